!!
!! NKA_TYPE (F95 procedural interface) over libnka_b200.so
!!
!! Drop-in for the public interface of src-F95/nka_type.F90:189-207 of nncarlson/nka:
!! same module name, same type name, same procedure names and argument orders.
!! The accelerator lives on the GPU; this module holds a handle and forwards.
!!
!!   call nka_init (this, vlen, mvec)          src-F95/nka_type.F90:211-225
!!   call nka_set_vec_tol (this, vtol)         :227-232
!!   call nka_accel_update (this, f [, dp])    :278-470   f: host array, overwritten in place
!!   call nka_accel_update (this, f_dev)       additive: f_dev = type(c_ptr) device address
!!   call nka_relax (this) / nka_restart (this)  :490-509 / :473-488
!!   call nka_delete (this)                    :266-275
!!   nka_num_vec, nka_max_vec, nka_vec_len, nka_vec_tol, nka_real_kind, nka_defined
!!
!! Differences, all forced by the device: a user dot product `dp` is used for what the
!! reference documents it for, the global sum of a parallel run -- each partial dot product p
!! of this process's portion becomes global as dp([p], [1.0]) (include/
!! nonlinear_krylov_accelerator.h); the default vtol (0.01, :194) is passed at creation.  Preconditions keep the reference's
!! ASSERT semantics (checked in the C library: message with file:line, then abort).
!!
!! NOT COMPILED in the build image (no Fortran compiler); see nka_b200_c.F90.
!!

module nka_type

  use, intrinsic :: iso_c_binding
  use nka_b200_c
  implicit none
  private

  integer, parameter :: r8 = selected_real_kind(15) ! 8-byte IEEE float

  type, public :: nka
    private
    type(c_ptr) :: handle = c_null_ptr
  end type nka

  public :: nka_init, nka_delete, nka_set_vec_tol, nka_defined
  public :: nka_vec_len, nka_num_vec, nka_max_vec, nka_vec_tol, nka_real_kind
  public :: nka_accel_update, nka_relax, nka_restart

  interface nka_accel_update
    module procedure nka_accel_update_host_array, nka_accel_update_device
  end interface

  abstract interface
    pure function dp_iface (x, y)
      import :: r8
      real(r8), intent(in) :: x(:), y(:)
      real(r8) :: dp_iface
    end function dp_iface
  end interface
  !! the dp of the nka_accel_update call in progress (the call is synchronous; not thread-safe,
  !! as little as the reference's module is)
  procedure(dp_iface), pointer, save :: current_dp => null()

contains

  subroutine nka_init (this, vlen, mvec)
    type(nka), intent(inout) :: this   ! intent(out) in the reference; inout so an old handle can be freed
    integer, intent(in) :: vlen
    integer, intent(in) :: mvec
    if (c_associated(this%handle)) call nka_delete_c (this%handle)
    this%handle = nka_init_ex (int(vlen, c_size_t), int(mvec, c_int), 0.01_c_double, -1_c_int, c_null_ptr)
  end subroutine nka_init

  subroutine nka_set_vec_tol (this, vtol)
    type(nka), intent(inout) :: this
    real(r8), intent(in) :: vtol
    call nka_set_vec_tol_c (this%handle, real(vtol, c_double))
  end subroutine nka_set_vec_tol

  integer function nka_num_vec (this)
    type(nka), intent(in) :: this
    nka_num_vec = nka_num_vec_c (this%handle)
  end function nka_num_vec

  integer function nka_max_vec (this)
    type(nka), intent(in) :: this
    nka_max_vec = nka_max_vec_c (this%handle)
  end function nka_max_vec

  integer function nka_vec_len (this)
    type(nka), intent(in) :: this
    nka_vec_len = nka_vec_len_c (this%handle)
  end function nka_vec_len

  real(r8) function nka_vec_tol (this)
    type(nka), intent(in) :: this
    nka_vec_tol = nka_vec_tol_c (this%handle)
  end function nka_vec_tol

  integer function nka_real_kind (this)
    type(nka), intent(in) :: this
    nka_real_kind = r8
  end function nka_real_kind

  logical function nka_defined (this)
    type(nka), intent(in) :: this
    nka_defined = .false.
    if (c_associated(this%handle)) nka_defined = (nka_defined_c (this%handle) /= 0)
  end function nka_defined

  subroutine nka_delete (this)
    type(nka), intent(inout) :: this
    if (c_associated(this%handle)) call nka_delete_c (this%handle)
    this%handle = c_null_ptr
  end subroutine nka_delete

  subroutine nka_accel_update_host_array (this, f, dp)
    type(nka), intent(inout) :: this
    real(r8),  intent(inout), target :: f(:)
    interface
      pure function dp (x, y)
        integer, parameter :: r8 = selected_real_kind(15)
        real(r8), intent(in) :: x(:), y(:)
        real(r8) :: dp
      end function dp
    end interface
    optional :: dp
    real(r8), allocatable :: tmp(:)
    if (present(dp)) then
      current_dp => dp
      call nka_set_dot_prod_ctx (this%handle, c_funloc(dp_trampoline), c_null_ptr)
    end if
    if (is_contiguous(f)) then
      call nka_accel_update_host (this%handle, f)
    else  ! strided section: stage through a contiguous copy
      tmp = f
      call nka_accel_update_host (this%handle, tmp)
      f = tmp
    end if
    if (present(dp)) then
      call nka_set_dot_prod_ctx (this%handle, c_null_funptr, c_null_ptr)
      current_dp => null()
    end if
  end subroutine nka_accel_update_host_array

  !! What the library calls (double (*)(int, double *, double *, void *)) while an update with a
  !! per-call dp is in progress.
  function dp_trampoline (n, x, y, ctx) bind(C) result(s)
    integer(c_int), value :: n
    real(c_double), intent(in) :: x(n), y(n)
    type(c_ptr), value :: ctx
    real(c_double) :: s
    s = current_dp (x, y)
  end function dp_trampoline

  subroutine nka_accel_update_device (this, f_dev)
    type(nka), intent(inout) :: this
    type(c_ptr), intent(in) :: f_dev   ! device address of vlen doubles; asynchronous on the handle's stream
    call nka_accel_update_dev (this%handle, f_dev)
  end subroutine nka_accel_update_device

  subroutine nka_restart (this)
    type(nka), intent(inout) :: this
    call nka_restart_c (this%handle)
  end subroutine nka_restart

  subroutine nka_relax (this)
    type(nka), intent(inout) :: this
    call nka_relax_c (this%handle)
  end subroutine nka_relax

end module nka_type
