!!
!! NKA_TYPE (F08 type-bound interface) over libnka_b200.so
!!
!! Drop-in for src-F08/nka_type.F90:148-181 of nncarlson/nka: same module, type and
!! binding names -- init, set_vec_tol, set_dot_prod, vec_len, num_vec, max_vec, vec_tol,
!! accel_update, relax, restart, defined -- plus delete and a finalizer (storage is
!! cudaMalloc'ed, so it cannot be released by automatic deallocation as in :380-383),
!! and accel_update overloaded for a device address.
!!
!!   call this%init(vlen, mvec)               src-F08/nka_type.F90:185-200
!!   call this%set_vec_tol(vtol)              :202-207
!!   call this%accel_update(f)                :249-419   f host array, overwritten in place
!!   call this%accel_update(f_dev)            additive: type(c_ptr) device address, asynchronous
!!   call this%relax() / this%restart()       :439-457 / :422-436
!!
!! set_dot_prod (:209-214) installs the caller's dot product as what the reference documents
!! it for, the GLOBAL sum of a parallel run: the device forms this process's partial dot
!! products and each partial p becomes global as dot_prod([p], [1.0]) -- so dot_prod must be a
!! Euclidean dot product followed by a sum over the processes (include/
!! nonlinear_krylov_accelerator.h).  comm_init is the built-in alternative (NVLink / NCCL).
!!
!! NOT COMPILED in the build image (no Fortran compiler); see nka_b200_c.F90.
!!

module nka_type

  use, intrinsic :: iso_fortran_env, only: r8 => real64, error_unit
  use, intrinsic :: iso_c_binding
  use nka_b200_c
  implicit none
  private

  type :: dp_holder
    procedure(dp), pointer, nopass :: f => null()
  end type dp_holder

  type, public :: nka
    private
    type(c_ptr) :: handle = c_null_ptr
    type(dp_holder), pointer :: hook => null()   ! heap cell: its address is the callback's context
  contains
    procedure :: init
    procedure :: set_vec_tol
    procedure :: set_dot_prod
    procedure :: vec_len
    procedure :: num_vec
    procedure :: max_vec
    procedure :: vec_tol
    generic   :: accel_update => accel_update_host_array, accel_update_device
    procedure, private :: accel_update_host_array, accel_update_device
    procedure :: relax
    procedure :: restart
    procedure :: defined
    procedure :: delete
    procedure :: comm_init
    final :: nka_final
  end type nka

  abstract interface
    real(r8) function dp(x, y)
      import :: r8
      real(r8), intent(in) :: x(:), y(:)
    end function
  end interface

contains

  subroutine init(this, vlen, mvec)
    class(nka), intent(inout) :: this
    integer, intent(in) :: vlen
    integer, intent(in) :: mvec
    call this%delete
    this%handle = nka_init_ex(int(vlen, c_size_t), int(mvec, c_int), 0.01_c_double, -1_c_int, c_null_ptr)
  end subroutine

  subroutine delete(this)
    class(nka), intent(inout) :: this
    if (c_associated(this%handle)) call nka_delete_c(this%handle)
    this%handle = c_null_ptr
    if (associated(this%hook)) deallocate(this%hook)
  end subroutine

  subroutine nka_final(this)
    type(nka), intent(inout) :: this
    if (c_associated(this%handle)) call nka_delete_c(this%handle)
    this%handle = c_null_ptr
    if (associated(this%hook)) deallocate(this%hook)
  end subroutine

  subroutine set_vec_tol(this, vtol)
    class(nka), intent(inout) :: this
    real(r8), intent(in) :: vtol
    call nka_set_vec_tol_c(this%handle, real(vtol, c_double))
  end subroutine

  subroutine set_dot_prod(this, dot_prod)
    class(nka), intent(inout) :: this
    procedure(dp), pointer :: dot_prod
    if (.not.associated(dot_prod)) then          ! ASSERT(associated(dot_prod)), src-F08/nka_type.F90:212
      write(error_unit,'(a)') 'nka%set_dot_prod: dot_prod is not associated'
      error stop 1
    end if
    if (.not.associated(this%hook)) allocate(this%hook)
    this%hook%f => dot_prod
    call nka_set_dot_prod_ctx(this%handle, c_funloc(dp_trampoline), c_loc(this%hook))
  end subroutine

  !! What the library calls (double (*)(int, double *, double *, void *)): finds the user's
  !! procedure pointer through ctx and hands it the two n-vectors as assumed-shape arrays.
  function dp_trampoline(n, x, y, ctx) bind(C) result(s)
    integer(c_int), value :: n
    real(c_double), intent(in) :: x(n), y(n)
    type(c_ptr), value :: ctx
    real(c_double) :: s
    type(dp_holder), pointer :: hook
    call c_f_pointer(ctx, hook)
    s = hook%f(x, y)
  end function

  !! Collective over the ranks that each own a slab of every vector (one process per GPU).
  subroutine comm_init(this, nranks, rank, id128)
    class(nka), intent(inout) :: this
    integer, intent(in) :: nranks, rank
    character(kind=c_char), intent(in) :: id128(128)
    if (nka_comm_init(this%handle, int(nranks, c_int), int(rank, c_int), id128) /= 0) then
      write(error_unit,'(a)') 'nka%comm_init: NCCL communicator creation failed'
      error stop 1
    end if
  end subroutine

  integer function num_vec(this)
    class(nka), intent(in) :: this
    num_vec = nka_num_vec_c(this%handle)
  end function

  integer function max_vec(this)
    class(nka), intent(in) :: this
    max_vec = nka_max_vec_c(this%handle)
  end function

  integer function vec_len(this)
    class(nka), intent(in) :: this
    vec_len = nka_vec_len_c(this%handle)
  end function

  real(r8) function vec_tol(this)
    class(nka), intent(in) :: this
    vec_tol = nka_vec_tol_c(this%handle)
  end function

  subroutine accel_update_host_array(this, f)
    class(nka), intent(inout) :: this
    real(r8),   intent(inout), contiguous :: f(:)
    call nka_accel_update_host(this%handle, f)
  end subroutine

  subroutine accel_update_device(this, f_dev)
    class(nka), intent(inout) :: this
    type(c_ptr), intent(in) :: f_dev
    call nka_accel_update_dev(this%handle, f_dev)
  end subroutine

  subroutine restart(this)
    class(nka), intent(inout) :: this
    call nka_restart_c(this%handle)
  end subroutine

  subroutine relax(this)
    class(nka), intent(inout) :: this
    call nka_relax_c(this%handle)
  end subroutine

  logical function defined(this)
    class(nka), intent(in) :: this
    defined = .false.
    if (c_associated(this%handle)) defined = (nka_defined_c(this%handle) /= 0)
  end function

end module nka_type
