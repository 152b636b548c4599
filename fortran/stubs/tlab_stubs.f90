!########################################################################
! Minimal stand-ins for the tlab modules that fortran/tlab_gpu_mod.f90 uses, with the same names and the members it touches
! (TLab_Constants: src/base/tlab_constants.f90; TLab_WorkFlow: src/base/tlab_workflow.f90; FDM_Derivative:
! src/fdm/fdm_derivative.f90:14-49; FDM: src/fdm/fdm.f90:14-29).  Only for a stand-alone syntax / interface check of the binding:
!     gfortran -std=f2008 -fsyntax-only fortran/stubs/tlab_stubs.f90 fortran/tlab_gpu_mod.f90
! Inside tlab the real modules are used and this file is not compiled.
!########################################################################
module TLab_Constants
    implicit none
    integer, parameter :: wp = kind(1.0d0)
    integer, parameter :: wi = kind(1)
    character(len=*), parameter :: efile = 'dns.err'
end module TLab_Constants

module TLab_WorkFlow
    implicit none
contains
    subroutine TLab_Write_ASCII(file, lineloc)
        character(len=*), intent(in) :: file, lineloc
        write (*, *) trim(file)//': '//trim(lineloc)
    end subroutine TLab_Write_ASCII
    subroutine TLab_Stop(error_code)
        integer, intent(in) :: error_code
        if (error_code /= 0) error stop 1
        stop
    end subroutine TLab_Stop
end module TLab_WorkFlow

module FDM_Derivative
    use TLab_Constants, only: wp, wi
    implicit none
    type, public :: fdm_derivative_dt
        sequence
        integer mode_fdm
        integer(wi) size
        logical :: periodic = .false.
        logical :: need_1der = .false.
        real(wp), allocatable :: lhs(:, :), rhs(:, :), mwn(:), lu(:, :)
    end type fdm_derivative_dt
end module FDM_Derivative

module FDM
    use TLab_Constants, only: wp, wi
    use FDM_Derivative
    implicit none
    type, public :: fdm_dt
        sequence
        character*8 name
        integer(wi) size
        logical :: uniform = .false.
        logical :: periodic = .false.
        real(wp) scale
        real(wp), allocatable :: nodes(:)
        real(wp), allocatable :: jac(:, :)
        type(fdm_derivative_dt) :: der1
        type(fdm_derivative_dt) :: der2
    end type fdm_dt
    type(fdm_dt), public :: g(3)
end module FDM
