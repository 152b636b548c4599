!########################################################################
! Binding of libtlab_gpu.so (include/tlab_gpu.h) for the Fortran host of tlab.
!
! Not compiled in this repository's image (no Fortran compiler); it is the file a tlab maintainer adds to
! src/operators/ (see INTEGRATION.md).  The wrappers keep the argument lists of the procedures they replace:
!   OPR_Partial_X/Y/Z   src/operators/opr_partial.f90:31-377
!   OPR_Burgers_X/Y/Z   src/physics/opr_burgers.f90:190-431
!   OPR_Poisson         src/operators/opr_elliptic.f90:32-46,263-364 (procedure pointer)
!   TIME_RUNGEKUTTA     src/tools/dns/time.f90:185-333
! Arrays passed to the operator wrappers are DEVICE arrays: the host holds them as type(c_ptr) obtained from
! tlab_gpu_malloc and, where Fortran code must index them, as pointers set with c_f_pointer on managed memory.
!########################################################################
module TLab_GPU
    use, intrinsic :: iso_c_binding
    implicit none
    private

    integer(c_int), parameter, public :: TLAB_OPR_P1 = 1, TLAB_OPR_P2 = 2, TLAB_OPR_P2_P1 = 3
    integer(c_int), parameter, public :: TLAB_BCS_NN = 3

    type, bind(C), public :: tlab_dns_params
        integer(c_int) :: nx, ny, nz, nscal, rkm_mode, buoyancy_type, scal_limit
        integer(c_int) :: bcs_flow_jmin(3), bcs_flow_jmax(3)
        integer(c_int) :: bcs_scal_jmin(8), bcs_scal_jmax(8)
        real(c_double) :: visc
        real(c_double) :: schmidt(8)
        real(c_double) :: buoyancy_params(2)
        real(c_double) :: buoyancy_vector(3)
        real(c_double) :: scal_min(8), scal_max(8)
    end type tlab_dns_params

    interface
        integer(c_int) function tlab_gpu_init(device) bind(C, name='tlab_gpu_init')
            import :: c_int
            integer(c_int), value :: device
        end function
        integer(c_int) function tlab_gpu_malloc(ptr, bytes) bind(C, name='tlab_gpu_malloc')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr) :: ptr
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function tlab_gpu_upload(dst, src, bytes) bind(C, name='tlab_gpu_upload')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), value :: dst, src
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function tlab_gpu_download(dst, src, bytes) bind(C, name='tlab_gpu_download')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), value :: dst, src
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function tlab_fdm_plan_create(dir, n, nodes, periodic, uniform, mode1, mode2, plan) &
            bind(C, name='tlab_fdm_plan_create')
            import :: c_int, c_double, c_ptr
            integer(c_int), value :: dir, n, periodic, uniform, mode1, mode2
            real(c_double), intent(in) :: nodes(*)
            type(c_ptr) :: plan
        end function
        integer(c_int) function tlab_opr_partial(dir, itype, nx, ny, nz, bcs, plan, u, res, tmp1) &
            bind(C, name='tlab_opr_partial')
            import :: c_int, c_ptr
            integer(c_int), value :: dir, itype, nx, ny, nz
            integer(c_int), intent(in) :: bcs(4)
            type(c_ptr), value :: plan, u, res, tmp1
        end function
        integer(c_int) function tlab_opr_burgers_init(gx, gy, gz, visc, nscal, schmidt) bind(C, name='tlab_opr_burgers_init')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: gx, gy, gz
            real(c_double), value :: visc
            integer(c_int), value :: nscal
            real(c_double), intent(in) :: schmidt(*)
        end function
        integer(c_int) function tlab_opr_burgers(dir, ivel, is, nx, ny, nz, bcs, s, u, res, tmp1, u_t) &
            bind(C, name='tlab_opr_burgers')
            import :: c_int, c_ptr
            integer(c_int), value :: dir, ivel, is, nx, ny, nz
            integer(c_int), intent(in) :: bcs(4)
            type(c_ptr), value :: s, u, res, tmp1, u_t
        end function
        integer(c_int) function tlab_opr_elliptic_init(gx, gy, gz, kmax_local) bind(C, name='tlab_opr_elliptic_init')
            import :: c_int, c_ptr
            type(c_ptr), value :: gx, gy, gz
            integer(c_int), value :: kmax_local
        end function
        integer(c_int) function tlab_opr_poisson(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy) &
            bind(C, name='tlab_opr_poisson')
            import :: c_int, c_ptr
            integer(c_int), value :: nx, ny, nz, ibc
            type(c_ptr), value :: p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy
        end function
        integer(c_int) function tlab_mpi_get_unique_id(id) bind(C, name='tlab_mpi_get_unique_id')
            import :: c_int, c_char
            character(kind=c_char) :: id(128)
        end function
        integer(c_int) function tlab_mpi_init(rank, nranks, id) bind(C, name='tlab_mpi_init')
            import :: c_int, c_char
            integer(c_int), value :: rank, nranks
            character(kind=c_char), intent(in) :: id(128)
        end function
        integer(c_int) function tlab_dns_create(prm, gx, gy, gz, bbackground, dns) bind(C, name='tlab_dns_create')
            import :: c_int, c_double, c_ptr, tlab_dns_params
            type(tlab_dns_params), intent(in) :: prm
            type(c_ptr), value :: gx, gy, gz
            real(c_double), intent(in) :: bbackground(*)
            type(c_ptr) :: dns
        end function
        integer(c_int) function tlab_dns_upload_host(dns, name, src) bind(C, name='tlab_dns_upload_host')
            import :: c_int, c_double, c_ptr, c_char
            type(c_ptr), value :: dns
            character(kind=c_char), intent(in) :: name(*)
            real(c_double), intent(in) :: src(*)
        end function
        integer(c_int) function tlab_dns_download_host(dns, name, dst) bind(C, name='tlab_dns_download_host')
            import :: c_int, c_double, c_ptr, c_char
            type(c_ptr), value :: dns
            character(kind=c_char), intent(in) :: name(*)
            real(c_double) :: dst(*)
        end function
        integer(c_int) function tlab_time_rungekutta(dns, dtime) bind(C, name='tlab_time_rungekutta')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dtime
        end function
        ! TIME_SUBSTEP_INCOMPRESSIBLE_EXPLICIT + update of one stage (time.f90:559-670, 277-297)
        integer(c_int) function tlab_time_rungekutta_stage(dns, dtime, stage) bind(C, name='tlab_time_rungekutta_stage')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dtime
            integer(c_int), value :: stage
        end function
        ! same step from and to host arrays q(isize_field,3), s(isize_field,inb_scal)
        integer(c_int) function tlab_time_rungekutta_host(dns, dtime, q, s) bind(C, name='tlab_time_rungekutta_host')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dtime
            real(c_double) :: q(*), s(*)
        end function
        ! TIME_COURANT (time.f90:365-548): new dtime and the logged CFL / diffusion numbers
        integer(c_int) function tlab_time_courant(dns, cfla, cfld, prandtl, dtime, cfl_number, diffusion_number) &
            bind(C, name='tlab_time_courant')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: cfla, cfld, prandtl
            real(c_double) :: dtime, cfl_number, diffusion_number
        end function
        ! DNS_BOUNDS_CONTROL (dns_local.f90:94-234): extrema of the dilatation
        integer(c_int) function tlab_dns_bounds_control(dns, dil_min, dil_max) bind(C, name='tlab_dns_bounds_control')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double) :: dil_min, dil_max
        end function
        integer(c_int) function tlab_boundary_bcs_neumann_y(ibc, nx, ny, nz, g, u, bcs_hb, bcs_ht) &
            bind(C, name='tlab_boundary_bcs_neumann_y')
            import :: c_int, c_ptr
            integer(c_int), value :: ibc, nx, ny, nz
            type(c_ptr), value :: g, u, bcs_hb, bcs_ht
        end function
        integer(c_int) function tlab_gpu_set_tuning(key, value) bind(C, name='tlab_gpu_set_tuning')
            import :: c_int, c_char
            character(kind=c_char), intent(in) :: key(*)
            integer(c_int), value :: value
        end function
        integer(c_int) function tlab_mpi_finalize() bind(C, name='tlab_mpi_finalize')
            import :: c_int
        end function
        integer(c_int) function tlab_dns_destroy(dns) bind(C, name='tlab_dns_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: dns
        end function
        function tlab_gpu_last_error() bind(C, name='tlab_gpu_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
    end interface

    type(c_ptr), public :: plan_gpu(3) = c_null_ptr      ! device twins of FDM's g(1:3)
    type(c_ptr), public :: dns_gpu = c_null_ptr

    public :: tlab_gpu_init, tlab_gpu_malloc, tlab_gpu_upload, tlab_gpu_download
    public :: tlab_fdm_plan_create, tlab_opr_partial, tlab_opr_burgers_init, tlab_opr_burgers
    public :: tlab_opr_elliptic_init, tlab_opr_poisson, tlab_mpi_get_unique_id, tlab_mpi_init
    public :: tlab_dns_create, tlab_dns_upload_host, tlab_dns_download_host, tlab_time_rungekutta
    public :: tlab_time_rungekutta_stage, tlab_time_rungekutta_host, tlab_time_courant, tlab_dns_bounds_control
    public :: tlab_boundary_bcs_neumann_y, tlab_gpu_set_tuning, tlab_mpi_finalize, tlab_dns_destroy
    public :: OPR_Partial_GPU, OPR_Burgers_GPU, OPR_Poisson_GPU, TLab_GPU_Check

contains
    ! turn a non-zero return code into the reference's error path
    subroutine TLab_GPU_Check(ierr)
        use TLab_Constants, only: efile
        use TLab_WorkFlow, only: TLab_Write_ASCII, TLab_Stop
        integer(c_int), intent(in) :: ierr
        character(kind=c_char), pointer :: cmsg(:)
        character(len=256) :: msg
        integer :: i
        if (ierr == 0) return
        call c_f_pointer(tlab_gpu_last_error(), cmsg, [256])
        msg = ' '
        do i = 1, 256
            if (cmsg(i) == c_null_char) exit
            msg(i:i) = cmsg(i)
        end do
        call TLab_Write_ASCII(efile, 'TLab_GPU. '//trim(msg))
        call TLab_Stop(int(ierr))
    end subroutine

    ! OPR_Partial_X/Y/Z(type, nx, ny, nz, bcs, g, u, result, tmp1); dir replaces g (g%name -> 1, 2, 3)
    subroutine OPR_Partial_GPU(dir, type, nx, ny, nz, bcs, u, result, tmp1)
        integer, intent(in) :: dir, type, nx, ny, nz
        integer, intent(in) :: bcs(2, 2)
        type(c_ptr), intent(in) :: u, result
        type(c_ptr), intent(in), optional :: tmp1
        type(c_ptr) :: t1
        t1 = c_null_ptr
        if (present(tmp1)) t1 = tmp1
        call TLab_GPU_Check(tlab_opr_partial(int(dir, c_int), int(type, c_int), int(nx, c_int), int(ny, c_int), &
                                             int(nz, c_int), int(reshape(bcs, [4]), c_int), plan_gpu(dir), u, result, t1))
    end subroutine

    ! OPR_Burgers_X/Y/Z(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
    subroutine OPR_Burgers_GPU(dir, ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
        integer, intent(in) :: dir, ivel, is, nx, ny, nz
        integer, intent(in) :: bcs(2, 2)
        type(c_ptr), intent(in) :: s, u, result, tmp1
        type(c_ptr), intent(in), optional :: u_t
        type(c_ptr) :: ut
        ut = c_null_ptr
        if (present(u_t)) ut = u_t
        call TLab_GPU_Check(tlab_opr_burgers(int(dir, c_int), int(ivel, c_int), int(is, c_int), int(nx, c_int), &
                                             int(ny, c_int), int(nz, c_int), int(reshape(bcs, [4]), c_int), s, u, result, tmp1, ut))
    end subroutine

    ! target of the procedure pointer OPR_Poisson (device arrays)
    subroutine OPR_Poisson_GPU(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy)
        integer, intent(in) :: nx, ny, nz, ibc
        type(c_ptr), intent(in) :: p, tmp1, tmp2, bcs_hb, bcs_ht
        type(c_ptr), intent(in), optional :: dpdy
        type(c_ptr) :: d
        d = c_null_ptr
        if (present(dpdy)) d = dpdy
        call TLab_GPU_Check(tlab_opr_poisson(int(nx, c_int), int(ny, c_int), int(nz, c_int), int(ibc, c_int), &
                                             p, tmp1, tmp2, bcs_hb, bcs_ht, d))
    end subroutine
end module TLab_GPU
