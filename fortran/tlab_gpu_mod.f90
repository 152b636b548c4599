!########################################################################
! Binding of libtlab_gpu.so (include/tlab_gpu.h) for the Fortran host of tlab.
!
! Two layers:
!   1. module TLab_GPU_C: one interface per C entry point of include/tlab_gpu.h (all of them; tests/test_abi.py checks the
!      list against the header), iso_c_binding only, no dependence on tlab.
!   2. module TLab_GPU: procedures with EXACTLY the dummy-argument lists of the reference procedures they replace, so that a
!      maintainer switches a call site by renaming it (or by re-pointing a procedure pointer):
!        OPR_Partial_X/Y/Z(type, nx, ny, nz, bcs, g, u, result, tmp1)          src/operators/opr_partial.f90:31,155,265
!        OPR_Burgers_X/Y/Z(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t) src/physics/opr_burgers.f90:190,277,359
!        OPR_Poisson_interface (nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy)   src/operators/opr_elliptic.f90:32-48
!        BOUNDARY_BCS_NEUMANN_Y(ibc, nx, ny, nz, g, u, bcs_hb, bcs_ht, tmp1)   src/tools/dns/boundary_bcs.f90:368
!        FDM_Der1_Solve(nlines, ibc, g, lu1, u, result, wrk2d), FDM_Der2_Solve(nlines, g, lu, u, result, du, wrk2d)
!                                                                                src/fdm/fdm_derivative.f90:218,413
!        TRIDSS, TRIDPSS, PENTADSS, PENTADSS2                                   src/utils/linear3.f90:56,321, linear5.f90:76,209
!        TLab_Transpose(a, nra, nca, ma, b, mb)                                 src/utils/tlab_transpose.f90:14
!        OPR_Fourier_X_Forward/Backward(nx, ny, nz, in, out), OPR_Fourier_Z_Forward/Backward(in, out)
!                                                                                src/operators/opr_fourier.f90:219-433
!        TLabMPI_Trp_ExecK_Forward/Backward(a, b, trp_plan)                     src/base/tlab_mpi_transpose.f90:343-553
!        TIME_RUNGEKUTTA(), TIME_COURANT()                                      src/tools/dns/time.f90:185,365
!
! Memory model of layer 2.  The wrappers take ordinary real(wp) arrays, as the reference does, and pass c_loc(array) to the
! library; the arrays must therefore live in memory the GPU can address: unified memory from tlab_gpu_malloc_managed.  The one
! change on the host is in TLab_Allocate_Real (src/base/tlab_memory.f90:306-330): q, s, txc, wrk* become
! `real(wp), pointer, contiguous` associated with c_f_pointer on a managed allocation (TLab_GPU_Allocate below) instead of
! `allocatable` + allocate().  Host code keeps indexing them; kernels read and write the same addresses; pages migrate on
! first touch, or ahead of time with tlab_gpu_prefetch.  (The coarse seam -- the whole Runge-Kutta step on resident device
! fields, INTEGRATION.md section 2 -- needs none of this and is the fast path.)
!
! Not compiled in this repository's image (no Fortran compiler: gfortran, flang, nvfortran, ifort are all absent).  Syntax
! check on a machine with gfortran, using the stub modules of fortran/stubs/ in place of tlab's own:
!     gfortran -std=f2008 -fsyntax-only fortran/stubs/tlab_stubs.f90 fortran/tlab_gpu_mod.f90
! (INTEGRATION.md section 6).  tests/abi/abi_smoke.c runs the same entry points from plain C.
!########################################################################
module TLab_GPU_C
    use, intrinsic :: iso_c_binding
    implicit none
    public

    integer(c_int), parameter :: TLAB_OPR_P1 = 1, TLAB_OPR_P2 = 2, TLAB_OPR_P2_P1 = 3
    integer(c_int), parameter :: TLAB_BCS_DD = 0, TLAB_BCS_ND = 1, TLAB_BCS_DN = 2, TLAB_BCS_NN = 3
    integer(c_int), parameter :: TLAB_ERR_UNDEVELOP = 104, TLAB_ERR_CUDA = 200
    integer(c_int), parameter :: TLAB_MAX_SCAL = 8, TLAB_PROF_CLASSES = 11

    type, bind(C) :: tlab_dns_params
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
        ! ---- runtime ----------------------------------------------------------------------------
        integer(c_int) function tlab_gpu_init(device) bind(C, name='tlab_gpu_init')
            import :: c_int
            integer(c_int), value :: device
        end function
        integer(c_int) function tlab_gpu_finalize() bind(C, name='tlab_gpu_finalize')
            import :: c_int
        end function
        function tlab_gpu_last_error() bind(C, name='tlab_gpu_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        integer(c_int) function tlab_gpu_set_async(on) bind(C, name='tlab_gpu_set_async')
            import :: c_int
            integer(c_int), value :: on
        end function
        integer(c_int) function tlab_gpu_synchronize() bind(C, name='tlab_gpu_synchronize')
            import :: c_int
        end function
        integer(c_int) function tlab_gpu_malloc(ptr, bytes) bind(C, name='tlab_gpu_malloc')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr) :: ptr
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function tlab_gpu_free(ptr) bind(C, name='tlab_gpu_free')
            import :: c_int, c_ptr
            type(c_ptr), value :: ptr
        end function
        integer(c_int) function tlab_gpu_malloc_managed(ptr, bytes) bind(C, name='tlab_gpu_malloc_managed')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr) :: ptr
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function tlab_gpu_prefetch(ptr, bytes, to_device) bind(C, name='tlab_gpu_prefetch')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), value :: ptr
            integer(c_size_t), value :: bytes
            integer(c_int), value :: to_device
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
        integer(c_int) function tlab_gpu_copy(dst, src, bytes) bind(C, name='tlab_gpu_copy')
            import :: c_int, c_ptr, c_size_t
            type(c_ptr), value :: dst, src
            integer(c_size_t), value :: bytes
        end function
        integer(c_int) function tlab_gpu_set_tuning(key, value) bind(C, name='tlab_gpu_set_tuning')
            import :: c_int, c_char
            character(kind=c_char), intent(in) :: key(*)
            integer(c_int), value :: value
        end function
        integer(c_int) function tlab_gpu_get_counter(key, value) bind(C, name='tlab_gpu_get_counter')
            import :: c_int, c_char, c_long_long
            character(kind=c_char), intent(in) :: key(*)
            integer(c_long_long) :: value
        end function
        integer(c_int) function tlab_gpu_stream(stream) bind(C, name='tlab_gpu_stream')
            import :: c_int, c_ptr
            type(c_ptr) :: stream
        end function
        integer(c_int) function tlab_gpu_profile(on) bind(C, name='tlab_gpu_profile')
            import :: c_int
            integer(c_int), value :: on
        end function
        integer(c_int) function tlab_gpu_profile_report(ms_per_class, count_per_class, nclass) bind(C, name='tlab_gpu_profile_report')
            import :: c_int, c_double
            real(c_double) :: ms_per_class(*)
            integer(c_int) :: count_per_class(*)
            integer(c_int), value :: nclass
        end function
        ! ---- plans ------------------------------------------------------------------------------
        integer(c_int) function tlab_fdm_plan_create(dir, n, nodes, periodic, uniform, mode1, mode2, plan) &
            bind(C, name='tlab_fdm_plan_create')
            import :: c_int, c_double, c_ptr
            integer(c_int), value :: dir, n, periodic, uniform, mode1, mode2
            real(c_double), intent(in) :: nodes(*)
            type(c_ptr) :: plan
        end function
        integer(c_int) function tlab_fdm_plan_create_host(dir, n, nodes, periodic, uniform, mode1, mode2, plan) &
            bind(C, name='tlab_fdm_plan_create_host')
            import :: c_int, c_double, c_ptr
            integer(c_int), value :: dir, n, periodic, uniform, mode1, mode2
            real(c_double), intent(in) :: nodes(*)
            type(c_ptr) :: plan
        end function
        integer(c_int) function tlab_fdm_plan_destroy(plan) bind(C, name='tlab_fdm_plan_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: plan
        end function
        integer(c_int) function tlab_fdm_plan_get(plan, what, out_host, capacity, count) bind(C, name='tlab_fdm_plan_get')
            import :: c_int, c_double, c_ptr, c_char
            type(c_ptr), value :: plan
            character(kind=c_char), intent(in) :: what(*)
            real(c_double) :: out_host(*)
            integer(c_int), value :: capacity
            integer(c_int) :: count
        end function
        integer(c_int) function tlab_fdm_int1_system_host(plan, ibc, lambda, lhs, rhs, rhs_b, rhs_t) &
            bind(C, name='tlab_fdm_int1_system_host')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: plan
            integer(c_int), value :: ibc
            real(c_double), value :: lambda
            real(c_double) :: lhs(*), rhs(*), rhs_b(*), rhs_t(*)
        end function
        ! ---- operators --------------------------------------------------------------------------
        integer(c_int) function tlab_opr_partial(dir, itype, nx, ny, nz, bcs, plan, u, res, tmp1) bind(C, name='tlab_opr_partial')
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
        integer(c_int) function tlab_opr_burgers(dir, ivel, is, nx, ny, nz, bcs, s, u, res, tmp1, u_t) bind(C, name='tlab_opr_burgers')
            import :: c_int, c_ptr
            integer(c_int), value :: dir, ivel, is, nx, ny, nz
            integer(c_int), intent(in) :: bcs(4)
            type(c_ptr), value :: s, u, res, tmp1, u_t
        end function
        integer(c_int) function tlab_fdm_der1_solve(plan, nlines, ibc, u, res) bind(C, name='tlab_fdm_der1_solve')
            import :: c_int, c_ptr
            type(c_ptr), value :: plan, u, res
            integer(c_int), value :: nlines, ibc
        end function
        integer(c_int) function tlab_fdm_der2_solve(plan, nlines, is_or_minus1, u, du, res) bind(C, name='tlab_fdm_der2_solve')
            import :: c_int, c_ptr
            type(c_ptr), value :: plan, u, du, res
            integer(c_int), value :: nlines, is_or_minus1
        end function
        integer(c_int) function tlab_tridss(nmax, len, a, b, c, f) bind(C, name='tlab_tridss')
            import :: c_int, c_ptr
            integer(c_int), value :: nmax, len
            type(c_ptr), value :: a, b, c, f
        end function
        integer(c_int) function tlab_tridpss(nmax, len, a, b, c, d, e, f, wrk) bind(C, name='tlab_tridpss')
            import :: c_int, c_ptr
            integer(c_int), value :: nmax, len
            type(c_ptr), value :: a, b, c, d, e, f, wrk
        end function
        integer(c_int) function tlab_pentadss(nmax, len, a, b, c, d, e, f) bind(C, name='tlab_pentadss')
            import :: c_int, c_ptr
            integer(c_int), value :: nmax, len
            type(c_ptr), value :: a, b, c, d, e, f
        end function
        integer(c_int) function tlab_pentadss2(nmax, len, a, b, c, d, e, f) bind(C, name='tlab_pentadss2')
            import :: c_int, c_ptr
            integer(c_int), value :: nmax, len
            type(c_ptr), value :: a, b, c, d, e, f
        end function
        integer(c_int) function tlab_transpose(a, nra, nca, ma, b, mb) bind(C, name='tlab_transpose')
            import :: c_int, c_ptr
            type(c_ptr), value :: a, b
            integer(c_int), value :: nra, nca, ma, mb
        end function
        integer(c_int) function tlab_transpose_complex(a, nra, nca, ma, b, mb) bind(C, name='tlab_transpose_complex')
            import :: c_int, c_ptr
            type(c_ptr), value :: a, b
            integer(c_int), value :: nra, nca, ma, mb
        end function
        integer(c_int) function tlab_boundary_bcs_neumann_y(ibc, nx, ny, nz, g, u, bcs_hb, bcs_ht) &
            bind(C, name='tlab_boundary_bcs_neumann_y')
            import :: c_int, c_ptr
            integer(c_int), value :: ibc, nx, ny, nz
            type(c_ptr), value :: g, u, bcs_hb, bcs_ht
        end function
        integer(c_int) function tlab_opr_elliptic_init(gx, gy, gz, kmax_local) bind(C, name='tlab_opr_elliptic_init')
            import :: c_int, c_ptr
            type(c_ptr), value :: gx, gy, gz
            integer(c_int), value :: kmax_local
        end function
        integer(c_int) function tlab_opr_poisson(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy) bind(C, name='tlab_opr_poisson')
            import :: c_int, c_ptr
            integer(c_int), value :: nx, ny, nz, ibc
            type(c_ptr), value :: p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy
        end function
        integer(c_int) function tlab_opr_fourier_x_forward(nx, ny, nz, in, out) bind(C, name='tlab_opr_fourier_x_forward')
            import :: c_int, c_ptr
            integer(c_int), value :: nx, ny, nz
            type(c_ptr), value :: in, out
        end function
        integer(c_int) function tlab_opr_fourier_x_backward(nx, ny, nz, in, out) bind(C, name='tlab_opr_fourier_x_backward')
            import :: c_int, c_ptr
            integer(c_int), value :: nx, ny, nz
            type(c_ptr), value :: in, out
        end function
        integer(c_int) function tlab_opr_fourier_z_forward(nx, ny, nz, in, out) bind(C, name='tlab_opr_fourier_z_forward')
            import :: c_int, c_ptr
            integer(c_int), value :: nx, ny, nz
            type(c_ptr), value :: in, out
        end function
        integer(c_int) function tlab_opr_fourier_z_backward(nx, ny, nz, in, out) bind(C, name='tlab_opr_fourier_z_backward')
            import :: c_int, c_ptr
            integer(c_int), value :: nx, ny, nz
            type(c_ptr), value :: in, out
        end function
        ! ---- domain decomposition ---------------------------------------------------------------
        integer(c_int) function tlab_mpi_get_unique_id(id) bind(C, name='tlab_mpi_get_unique_id')
            import :: c_int, c_char
            character(kind=c_char) :: id(128)
        end function
        integer(c_int) function tlab_mpi_init(rank, nranks, id) bind(C, name='tlab_mpi_init')
            import :: c_int, c_char
            integer(c_int), value :: rank, nranks
            character(kind=c_char), intent(in) :: id(128)
        end function
        integer(c_int) function tlab_mpi_finalize() bind(C, name='tlab_mpi_finalize')
            import :: c_int
        end function
        integer(c_int) function tlab_mpi_rank(rank, nranks) bind(C, name='tlab_mpi_rank')
            import :: c_int
            integer(c_int) :: rank, nranks
        end function
        integer(c_int) function tlab_trp_exec_k_forward(a, b, nlines_total, kmax, is_complex) bind(C, name='tlab_trp_exec_k_forward')
            import :: c_int, c_ptr
            type(c_ptr), value :: a, b
            integer(c_int), value :: nlines_total, kmax, is_complex
        end function
        integer(c_int) function tlab_trp_exec_k_backward(b, a, nlines_total, kmax, is_complex) bind(C, name='tlab_trp_exec_k_backward')
            import :: c_int, c_ptr
            type(c_ptr), value :: b, a
            integer(c_int), value :: nlines_total, kmax, is_complex
        end function
        ! ---- time advance -----------------------------------------------------------------------
        integer(c_int) function tlab_dns_create(prm, gx, gy, gz, bbackground, dns) bind(C, name='tlab_dns_create')
            import :: c_int, c_ptr, tlab_dns_params
            type(tlab_dns_params), intent(in) :: prm
            type(c_ptr), value :: gx, gy, gz
            type(c_ptr), value :: bbackground          ! c_loc of a host array of ny doubles, or c_null_ptr
            type(c_ptr) :: dns
        end function
        integer(c_int) function tlab_dns_destroy(dns) bind(C, name='tlab_dns_destroy')
            import :: c_int, c_ptr
            type(c_ptr), value :: dns
        end function
        integer(c_int) function tlab_dns_field(dns, name, dev_ptr) bind(C, name='tlab_dns_field')
            import :: c_int, c_ptr, c_char
            type(c_ptr), value :: dns
            character(kind=c_char), intent(in) :: name(*)
            type(c_ptr) :: dev_ptr
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
        integer(c_int) function tlab_dns_launch_count(dns, count) bind(C, name='tlab_dns_launch_count')
            import :: c_int, c_ptr, c_long_long
            type(c_ptr), value :: dns
            integer(c_long_long) :: count
        end function
        integer(c_int) function tlab_time_rk_coefficients(rkm_mode, kdt, ktime, kco, nsub) bind(C, name='tlab_time_rk_coefficients')
            import :: c_int, c_double
            integer(c_int), value :: rkm_mode
            real(c_double) :: kdt(*), ktime(*), kco(*)
            integer(c_int) :: nsub
        end function
        integer(c_int) function tlab_rhs_global_incompressible_1(dns, dte) bind(C, name='tlab_rhs_global_incompressible_1')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dte
        end function
        integer(c_int) function tlab_time_substep(dns, dte, kco, scale_h) bind(C, name='tlab_time_substep')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dte, kco
            integer(c_int), value :: scale_h
        end function
        integer(c_int) function tlab_time_rungekutta_stage(dns, dtime, stage) bind(C, name='tlab_time_rungekutta_stage')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dtime
            integer(c_int), value :: stage
        end function
        integer(c_int) function tlab_time_rungekutta(dns, dtime) bind(C, name='tlab_time_rungekutta')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dtime
        end function
        integer(c_int) function tlab_time_courant(dns, cfla, cfld, prandtl, dtime, cfl_number, diffusion_number) &
            bind(C, name='tlab_time_courant')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: cfla, cfld, prandtl
            real(c_double) :: dtime, cfl_number, diffusion_number
        end function
        integer(c_int) function tlab_dns_bounds_control(dns, dil_min, dil_max) bind(C, name='tlab_dns_bounds_control')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double) :: dil_min, dil_max
        end function
        integer(c_int) function tlab_time_rungekutta_host(dns, dtime, q, s) bind(C, name='tlab_time_rungekutta_host')
            import :: c_int, c_double, c_ptr
            type(c_ptr), value :: dns
            real(c_double), value :: dtime
            real(c_double) :: q(*), s(*)
        end function
    end interface
end module TLab_GPU_C

!########################################################################
! Layer 2: the reference's own procedure interfaces on top of the C ABI.
!########################################################################
module TLab_GPU
    use, intrinsic :: iso_c_binding
    use TLab_GPU_C
    use TLab_Constants, only: wp, wi, efile
    use TLab_WorkFlow, only: TLab_Write_ASCII, TLab_Stop
    use FDM, only: fdm_dt
    use FDM_Derivative, only: fdm_derivative_dt
    implicit none
    private

    type(c_ptr), public :: plan_gpu(3) = c_null_ptr      ! device twins of FDM's g(1:3), created by TLab_GPU_CreatePlans
    type(c_ptr), public :: dns_gpu = c_null_ptr          ! state of the coarse seam (tlab_dns_create)
    integer(wi), public :: fourier_nx = 0, fourier_ny = 0, fourier_nz = 0   ! what OPR_Fourier_Initialize keeps in module variables

    public :: TLab_GPU_Check, TLab_GPU_CreatePlans, TLab_GPU_Allocate
    public :: OPR_Partial_X_GPU, OPR_Partial_Y_GPU, OPR_Partial_Z_GPU
    public :: OPR_Burgers_X_GPU, OPR_Burgers_Y_GPU, OPR_Burgers_Z_GPU
    public :: OPR_Poisson_GPU, BOUNDARY_BCS_NEUMANN_Y_GPU
    public :: FDM_Der1_Solve_GPU, FDM_Der2_Solve_GPU
    public :: TRIDSS_GPU, TRIDPSS_GPU, PENTADSS_GPU, PENTADSS2_GPU, TLab_Transpose_GPU
    public :: OPR_Fourier_X_Forward_GPU, OPR_Fourier_X_Backward_GPU, OPR_Fourier_Z_Forward_GPU, OPR_Fourier_Z_Backward_GPU
    public :: TLabMPI_Trp_ExecK_Forward_GPU, TLabMPI_Trp_ExecK_Backward_GPU
    public :: TIME_RUNGEKUTTA_GPU, TIME_COURANT_GPU

contains
    ! ###################################################################
    ! non-zero return code -> the reference's error path (TLab_Write_ASCII(efile, ...); TLab_Stop(code))
    subroutine TLab_GPU_Check(ierr)
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
    end subroutine TLab_GPU_Check

    ! ###################################################################
    ! after FDM_Initialize: one device plan per direction from the same node positions and scheme codes (FDM_CreatePlan)
    subroutine TLab_GPU_CreatePlans(g)
        type(fdm_dt), intent(in) :: g(3)
        integer :: id
        do id = 1, 3
            call TLab_GPU_Check(tlab_fdm_plan_create(int(id, c_int), int(g(id)%size, c_int), g(id)%nodes, &
                                                     merge(1_c_int, 0_c_int, g(id)%periodic), merge(1_c_int, 0_c_int, g(id)%uniform), &
                                                     int(g(id)%der1%mode_fdm, c_int), int(g(id)%der2%mode_fdm, c_int), plan_gpu(id)))
        end do
    end subroutine TLab_GPU_CreatePlans

    ! replacement of the allocate() inside TLab_Allocate_Real1/2 (tlab_memory.f90:306-330): unified memory
    subroutine TLab_GPU_Allocate(a, n1, n2)
        real(wp), pointer, contiguous, intent(out) :: a(:, :)
        integer(wi), intent(in) :: n1, n2
        type(c_ptr) :: p
        call TLab_GPU_Check(tlab_gpu_malloc_managed(p, int(n1, c_size_t)*int(n2, c_size_t)*c_sizeof(1.0_wp)))
        call c_f_pointer(p, a, [n1, n2])
        a = 0.0_wp
    end subroutine TLab_GPU_Allocate

    ! direction of a plan from its name ('x', 'y', 'z': fdm_dt%name, set by FDM_Initialize)
    integer function plan_dir(g)
        type(fdm_dt), intent(in) :: g
        select case (trim(adjustl(g%name)))
        case ('x')
            plan_dir = 1
        case ('y')
            plan_dir = 2
        case ('z')
            plan_dir = 3
        case default
            plan_dir = 0
            call TLab_Write_ASCII(efile, 'TLab_GPU. Plan name must be x, y or z.')
            call TLab_Stop(85)
        end select
    end function plan_dir

    ! ###################################################################
    ! OPR_Partial_X/Y/Z, opr_partial.f90:31-377
    subroutine OPR_Partial_GPU(dir, type, nx, ny, nz, bcs, u, result, tmp1)
        integer, intent(in) :: dir
        integer(wi), intent(in) :: type, nx, ny, nz
        integer(wi), intent(in) :: bcs(:, :)
        real(wp), intent(in), target :: u(nx*ny*nz)
        real(wp), intent(out), target :: result(nx*ny*nz)
        real(wp), intent(inout), optional, target :: tmp1(nx*ny*nz)
        type(c_ptr) :: t1
        integer(c_int) :: b(4)
        t1 = c_null_ptr
        if (present(tmp1)) t1 = c_loc(tmp1)
        b = int(reshape(bcs(1:2, 1:2), [4]), c_int)
        call TLab_GPU_Check(tlab_opr_partial(int(dir, c_int), int(type, c_int), int(nx, c_int), int(ny, c_int), int(nz, c_int), &
                                             b, plan_gpu(dir), c_loc(u), c_loc(result), t1))
    end subroutine OPR_Partial_GPU

    subroutine OPR_Partial_X_GPU(type, nx, ny, nz, bcs, g, u, result, tmp1)
        integer(wi), intent(in) :: type
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(:, :)
        type(fdm_dt), intent(in) :: g
        real(wp), intent(in) :: u(nx*ny*nz)
        real(wp), intent(out) :: result(nx*ny*nz)
        real(wp), intent(inout), optional :: tmp1(nx*ny*nz)
        call OPR_Partial_GPU(plan_dir(g), type, nx, ny, nz, bcs, u, result, tmp1)
    end subroutine OPR_Partial_X_GPU

    subroutine OPR_Partial_Y_GPU(type, nx, ny, nz, bcs, g, u, result, tmp1)
        integer(wi), intent(in) :: type
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(:, :)
        type(fdm_dt), intent(in) :: g
        real(wp), intent(in) :: u(nx*ny*nz)
        real(wp), intent(out) :: result(nx*ny*nz)
        real(wp), intent(inout), optional :: tmp1(nx*ny*nz)
        call OPR_Partial_GPU(plan_dir(g), type, nx, ny, nz, bcs, u, result, tmp1)
    end subroutine OPR_Partial_Y_GPU

    subroutine OPR_Partial_Z_GPU(type, nx, ny, nz, bcs, g, u, result, tmp1)
        integer(wi), intent(in) :: type
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(:, :)
        type(fdm_dt), intent(in) :: g
        real(wp), intent(in) :: u(nx*ny*nz)
        real(wp), intent(out) :: result(nx*ny*nz)
        real(wp), intent(inout), optional :: tmp1(nx*ny*nz)
        call OPR_Partial_GPU(plan_dir(g), type, nx, ny, nz, bcs, u, result, tmp1)
    end subroutine OPR_Partial_Z_GPU

    ! ###################################################################
    ! OPR_Burgers_X/Y/Z, opr_burgers.f90:190-431 (after tlab_opr_burgers_init in place of OPR_Burgers_Initialize).
    ! tmp1 and u_t carry the reference's transposed velocity; no transposed copies exist here, they are passed and ignored.
    subroutine OPR_Burgers_GPU(dir, ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
        integer, intent(in) :: dir, ivel, is
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(2, 2)
        real(wp), intent(in), target :: s(nx*ny*nz), u(nx*ny*nz)
        real(wp), intent(out), target :: result(nx*ny*nz)
        real(wp), intent(inout), target :: tmp1(nx*ny*nz)
        real(wp), intent(in), optional, target :: u_t(nx*ny*nz)
        type(c_ptr) :: ut
        integer(c_int) :: b(4)
        ut = c_null_ptr
        if (present(u_t)) ut = c_loc(u_t)
        b = int(reshape(bcs, [4]), c_int)
        call TLab_GPU_Check(tlab_opr_burgers(int(dir, c_int), int(ivel, c_int), int(is, c_int), int(nx, c_int), int(ny, c_int), &
                                             int(nz, c_int), b, c_loc(s), c_loc(u), c_loc(result), c_loc(tmp1), ut))
    end subroutine OPR_Burgers_GPU

    subroutine OPR_Burgers_X_GPU(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
        integer, intent(in) :: ivel
        integer, intent(in) :: is
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(2, 2)
        real(wp), intent(in) :: s(nx*ny*nz), u(nx*ny*nz)
        real(wp), intent(out) :: result(nx*ny*nz)
        real(wp), intent(inout) :: tmp1(nx*ny*nz)
        real(wp), intent(in), optional :: u_t(nx*ny*nz)
        call OPR_Burgers_GPU(1, ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
    end subroutine OPR_Burgers_X_GPU

    subroutine OPR_Burgers_Y_GPU(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
        integer, intent(in) :: ivel
        integer, intent(in) :: is
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(2, 2)
        real(wp), intent(in) :: s(nx*ny*nz), u(nx*ny*nz)
        real(wp), intent(out) :: result(nx*ny*nz)
        real(wp), intent(inout) :: tmp1(nx*ny*nz)
        real(wp), intent(in), optional :: u_t(nx*ny*nz)
        call OPR_Burgers_GPU(2, ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
    end subroutine OPR_Burgers_Y_GPU

    subroutine OPR_Burgers_Z_GPU(ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
        integer, intent(in) :: ivel
        integer, intent(in) :: is
        integer(wi), intent(in) :: nx, ny, nz
        integer(wi), intent(in) :: bcs(2, 2)
        real(wp), intent(in) :: s(nx*ny*nz), u(nx*ny*nz)
        real(wp), intent(out) :: result(nx*ny*nz)
        real(wp), intent(inout) :: tmp1(nx*ny*nz)
        real(wp), intent(in), optional :: u_t(nx*ny*nz)
        call OPR_Burgers_GPU(3, ivel, is, nx, ny, nz, bcs, s, u, result, tmp1, u_t)
    end subroutine OPR_Burgers_Z_GPU

    ! ###################################################################
    ! Conforms to OPR_Poisson_interface (opr_elliptic.f90:32-48): in OPR_Elliptic_Initialize,
    !     OPR_Poisson => OPR_Poisson_GPU            instead of      OPR_Poisson => OPR_Poisson_FourierXZ_Factorize  (:134)
    ! after  call TLab_GPU_Check(tlab_opr_elliptic_init(plan_gpu(1), plan_gpu(2), plan_gpu(3), int(kmax, c_int))).
    subroutine OPR_Poisson_GPU(nx, ny, nz, ibc, p, tmp1, tmp2, bcs_hb, bcs_ht, dpdy)
        integer(wi), intent(in) :: nx, ny, nz
        integer, intent(in) :: ibc
        real(wp), intent(inout) :: p(nx, ny, nz)
        real(wp), intent(inout), target :: tmp1(2*ny, nz, nx/2 + 1)
        real(wp), intent(inout), target :: tmp2(2*ny, nz, nx/2 + 1)
        real(wp), intent(in) :: bcs_hb(nx, nz), bcs_ht(nx, nz)
        real(wp), intent(out), optional :: dpdy(nx, ny, nz)
        target p, bcs_hb, bcs_ht, dpdy
        type(c_ptr) :: d
        d = c_null_ptr
        if (present(dpdy)) d = c_loc(dpdy)
        call TLab_GPU_Check(tlab_opr_poisson(int(nx, c_int), int(ny, c_int), int(nz, c_int), int(ibc, c_int), &
                                             c_loc(p), c_loc(tmp1), c_loc(tmp2), c_loc(bcs_hb), c_loc(bcs_ht), d))
    end subroutine OPR_Poisson_GPU

    ! ###################################################################
    ! BOUNDARY_BCS_NEUMANN_Y, boundary_bcs.f90:368-473 (tmp1 held the transposed field; unused here)
    subroutine BOUNDARY_BCS_NEUMANN_Y_GPU(ibc, nx, ny, nz, g, u, bcs_hb, bcs_ht, tmp1)
        integer(wi), intent(in) :: ibc
        integer(wi) nx, ny, nz
        type(fdm_dt), intent(in) :: g
        real(wp), intent(in) :: u(nx*nz, ny)
        real(wp), intent(inout) :: tmp1(nx*nz, ny)
        real(wp), intent(out) :: bcs_hb(nx*nz), bcs_ht(nx*nz)
        target u, bcs_hb, bcs_ht, tmp1
        call TLab_GPU_Check(tlab_boundary_bcs_neumann_y(int(ibc, c_int), int(nx, c_int), int(ny, c_int), int(nz, c_int), &
                                                        plan_gpu(plan_dir(g)), c_loc(u), c_loc(bcs_hb), c_loc(bcs_ht)))
    end subroutine BOUNDARY_BCS_NEUMANN_Y_GPU

    ! ###################################################################
    ! FDM_Der1_Solve / FDM_Der2_Solve, fdm_derivative.f90:218-278, 413-459.  g is the derivative plan of a direction; the
    ! library finds its twin through idir (the reference passes g(idir)%der1 and its lu, which the device plan already holds).
    subroutine FDM_Der1_Solve_GPU(nlines, ibc, g, lu1, u, result, wrk2d, idir)
        integer(wi), intent(in) :: nlines
        integer, intent(in) :: ibc
        type(fdm_derivative_dt), intent(in) :: g
        real(wp), intent(in) :: lu1(:, :)
        real(wp), intent(in), target :: u(nlines, g%size)
        real(wp), intent(out), target :: result(nlines, g%size)
        real(wp), intent(inout) :: wrk2d(*)
        integer, intent(in) :: idir
        call TLab_GPU_Check(tlab_fdm_der1_solve(plan_gpu(idir), int(nlines, c_int), int(ibc, c_int), c_loc(u), c_loc(result)))
    end subroutine FDM_Der1_Solve_GPU

    subroutine FDM_Der2_Solve_GPU(nlines, g, lu, u, result, du, wrk2d, idir)
        integer(wi), intent(in) :: nlines
        type(fdm_derivative_dt), intent(in) :: g
        real(wp), intent(in) :: lu(:, :)
        real(wp), intent(in), target :: u(nlines, g%size)
        real(wp), intent(in), target :: du(nlines, g%size)
        real(wp), intent(out), target :: result(nlines, g%size)
        real(wp), intent(out) :: wrk2d(*)
        integer, intent(in) :: idir
        call TLab_GPU_Check(tlab_fdm_der2_solve(plan_gpu(idir), int(nlines, c_int), -1_c_int, c_loc(u), c_loc(du), c_loc(result)))
    end subroutine FDM_Der2_Solve_GPU

    ! ###################################################################
    ! thomas3 / thomas5 substitution stages, linear3.f90:56-150, 321-442; linear5.f90:76-131, 209-244
    subroutine TRIDSS_GPU(nmax, len, a, b, c, f)
        integer(wi), intent(in) :: nmax, len
        real(wp), intent(in), target :: a(nmax), b(nmax), c(nmax)
        real(wp), intent(inout), target :: f(len, nmax)
        call TLab_GPU_Check(tlab_tridss(int(nmax, c_int), int(len, c_int), c_loc(a), c_loc(b), c_loc(c), c_loc(f)))
    end subroutine TRIDSS_GPU

    subroutine TRIDPSS_GPU(nmax, len, a, b, c, d, e, f, wrk)
        integer(wi), intent(in) :: nmax, len
        real(wp), intent(in), target :: a(nmax), b(nmax), c(nmax), d(nmax), e(nmax)
        real(wp), intent(inout), target :: f(len, nmax)
        real(wp), intent(inout), target :: wrk(len)
        call TLab_GPU_Check(tlab_tridpss(int(nmax, c_int), int(len, c_int), c_loc(a), c_loc(b), c_loc(c), c_loc(d), c_loc(e), &
                                         c_loc(f), c_loc(wrk)))
    end subroutine TRIDPSS_GPU

    subroutine PENTADSS_GPU(nmax, len, a, b, c, d, e, f)
        integer(wi) nmax, len
        real(wp), dimension(nmax), intent(in), target :: a, b, c, d, e
        real(wp), intent(inout), target :: f(len, nmax)
        call TLab_GPU_Check(tlab_pentadss(int(nmax, c_int), int(len, c_int), c_loc(a), c_loc(b), c_loc(c), c_loc(d), c_loc(e), c_loc(f)))
    end subroutine PENTADSS_GPU

    subroutine PENTADSS2_GPU(nmax, len, a, b, c, d, e, f)
        integer(wi) nmax, len
        real(wp), dimension(nmax), intent(in), target :: a, b, c, d, e
        real(wp), intent(inout), target :: f(len, nmax)
        call TLab_GPU_Check(tlab_pentadss2(int(nmax, c_int), int(len, c_int), c_loc(a), c_loc(b), c_loc(c), c_loc(d), c_loc(e), c_loc(f)))
    end subroutine PENTADSS2_GPU

    ! TLab_Transpose, tlab_transpose.f90:14-82
    subroutine TLab_Transpose_GPU(a, nra, nca, ma, b, mb)
        integer(wi), intent(in) :: nra, nca, ma, mb
        real(wp), intent(in), target :: a(ma, *)
        real(wp), intent(out), target :: b(mb, *)
        call TLab_GPU_Check(tlab_transpose(c_loc(a), int(nra, c_int), int(nca, c_int), int(ma, c_int), c_loc(b), int(mb, c_int)))
    end subroutine TLab_Transpose_GPU

    ! ###################################################################
    ! OPR_Fourier_*, opr_fourier.f90:219-433.  The Z transforms take their sizes from module variables in the reference
    ! (set by OPR_Fourier_Initialize); here from fourier_nx/ny/nz, which the X transforms record.
    subroutine OPR_Fourier_X_Forward_GPU(nx, ny, nz, in, out)
        integer(wi), intent(in) :: nx, ny, nz
        real(wp), intent(in), target :: in(nx*ny*nz)
        complex(wp), intent(out), target :: out(*)
        fourier_nx = nx; fourier_ny = ny; fourier_nz = nz
        call TLab_GPU_Check(tlab_opr_fourier_x_forward(int(nx, c_int), int(ny, c_int), int(nz, c_int), c_loc(in), c_loc(out)))
    end subroutine OPR_Fourier_X_Forward_GPU

    subroutine OPR_Fourier_X_Backward_GPU(nx, ny, nz, in, out)
        integer(wi) nx, ny, nz
        complex(wp), intent(in), target :: in(*)
        real(wp), intent(out), target :: out(nx*ny*nz)
        fourier_nx = nx; fourier_ny = ny; fourier_nz = nz
        call TLab_GPU_Check(tlab_opr_fourier_x_backward(int(nx, c_int), int(ny, c_int), int(nz, c_int), c_loc(in), c_loc(out)))
    end subroutine OPR_Fourier_X_Backward_GPU

    subroutine OPR_Fourier_Z_Forward_GPU(in, out)
        complex(wp), intent(inout), target :: in(*)
        complex(wp), intent(out), target :: out(*)
        call TLab_GPU_Check(tlab_opr_fourier_z_forward(int(fourier_nx, c_int), int(fourier_ny, c_int), int(fourier_nz, c_int), &
                                                       c_loc(in), c_loc(out)))
    end subroutine OPR_Fourier_Z_Forward_GPU

    subroutine OPR_Fourier_Z_Backward_GPU(in, out)
        complex(wp), intent(inout), target :: in(*)
        complex(wp), intent(out), target :: out(*)
        call TLab_GPU_Check(tlab_opr_fourier_z_backward(int(fourier_nx, c_int), int(fourier_ny, c_int), int(fourier_nz, c_int), &
                                                        c_loc(in), c_loc(out)))
    end subroutine OPR_Fourier_Z_Backward_GPU

    ! ###################################################################
    ! TLabMPI_Trp_ExecK_Forward/Backward_Real, tlab_mpi_transpose.f90:343-384, 403-440: slab a(nlines, kmax) <-> pencil
    ! b(nlines/P, kmax*P).  The reference's plan carries nlines and the MPI datatypes; only the extents matter here.
    subroutine TLabMPI_Trp_ExecK_Forward_GPU(a, b, nlines, kmax)
        real(wp), intent(in), target :: a(*)
        real(wp), intent(out), target :: b(*)
        integer(wi), intent(in) :: nlines, kmax           ! trp_plan%nlines * ims_npro_k, local slab thickness
        call TLab_GPU_Check(tlab_trp_exec_k_forward(c_loc(a), c_loc(b), int(nlines, c_int), int(kmax, c_int), 0_c_int))
    end subroutine TLabMPI_Trp_ExecK_Forward_GPU

    subroutine TLabMPI_Trp_ExecK_Backward_GPU(b, a, nlines, kmax)
        real(wp), intent(in), target :: b(*)
        real(wp), intent(out), target :: a(*)
        integer(wi), intent(in) :: nlines, kmax
        call TLab_GPU_Check(tlab_trp_exec_k_backward(c_loc(b), c_loc(a), int(nlines, c_int), int(kmax, c_int), 0_c_int))
    end subroutine TLabMPI_Trp_ExecK_Backward_GPU

    ! ###################################################################
    ! TIME_RUNGEKUTTA(), time.f90:185-333, and TIME_COURANT(), :365-548, on the device-resident state dns_gpu; both take no
    ! arguments in the reference and work on module variables (dtime, rtime, q, s): the caller keeps advancing rtime/itime.
    subroutine TIME_RUNGEKUTTA_GPU(dtime)
        real(wp), intent(in) :: dtime
        call TLab_GPU_Check(tlab_time_rungekutta(dns_gpu, real(dtime, c_double)))
    end subroutine TIME_RUNGEKUTTA_GPU

    subroutine TIME_COURANT_GPU(cfla, cfld, prandtl, dtime, cfl_number, diffusion_number)
        real(wp), intent(in) :: cfla, cfld, prandtl
        real(wp), intent(inout) :: dtime
        real(wp), intent(out) :: cfl_number, diffusion_number          ! logs_data(2:3) of the reference
        real(c_double) :: dt, c1, c2
        dt = dtime
        call TLab_GPU_Check(tlab_time_courant(dns_gpu, real(cfla, c_double), real(cfld, c_double), real(prandtl, c_double), dt, c1, c2))
        dtime = dt; cfl_number = c1; diffusion_number = c2
    end subroutine TIME_COURANT_GPU

end module TLab_GPU
