!!
!! NKA_TYPE (abstract-vector interface) over libnka_b200.so
!!
!! Drop-in for src-F08-vector/nka_type.F90:148-171 of nncarlson/nka: init(vec, mvec),
!! set_vec_tol, num_vec, max_vec, vec_tol, accel_update(f), relax, restart, defined, for
!! vectors of the concrete class GPU_VECTOR.  Where the reference composes accel_update out
!! of 4M+6 virtual vector calls, each a full memory sweep (:237-238, 255-256, 262, 336, 347,
!! 374, 382), this forwards the whole update to the fused kernels: two sweeps and, across
!! GPUs, one all-reduce.  init clones nothing on the host: the accelerator adopts the
!! prototype's length, device, stream and communicator (nka_init_like).
!!
!! A vector of any other dynamic type stops with the base class's "incompatible arguments"
!! convention (vector_class.F90:157).
!!
!! NOT COMPILED in the build image (no Fortran compiler); see nka_b200_c.F90.
!!

module nka_type

  use, intrinsic :: iso_fortran_env, only: r8 => real64
  use, intrinsic :: iso_c_binding
  use vector_class
  use gpu_vector_type
  use nka_b200_c
  implicit none
  private

  type, public :: nka
    private
    type(c_ptr) :: handle = c_null_ptr
  contains
    procedure :: init
    procedure :: set_vec_tol
    procedure :: num_vec
    procedure :: max_vec
    procedure :: vec_tol
    procedure :: accel_update
    procedure :: relax
    procedure :: restart
    procedure :: defined
    procedure :: delete
    final :: nka_final
  end type nka

contains

  subroutine init(this, vec, mvec)
    class(nka), intent(inout) :: this
    class(vector), intent(in) :: vec
    integer, intent(in) :: mvec
    call this%delete
    select type (vec)
    class is (gpu_vector)
      this%handle = nka_init_like(vec%vec, int(mvec, c_int), 0.01_c_double)
    class default
      error stop 'incompatible arguments to NKA%INIT: this build accelerates GPU_VECTOR objects'
    end select
  end subroutine

  subroutine delete(this)
    class(nka), intent(inout) :: this
    if (c_associated(this%handle)) call nka_delete_c(this%handle)
    this%handle = c_null_ptr
  end subroutine

  subroutine nka_final(this)
    type(nka), intent(inout) :: this
    if (c_associated(this%handle)) call nka_delete_c(this%handle)
    this%handle = c_null_ptr
  end subroutine

  subroutine set_vec_tol(this, vtol)
    class(nka), intent(inout) :: this
    real(r8), intent(in) :: vtol
    call nka_set_vec_tol_c(this%handle, real(vtol, c_double))
  end subroutine

  integer function num_vec(this)
    class(nka), intent(in) :: this
    num_vec = nka_num_vec_c(this%handle)
  end function

  integer function max_vec(this)
    class(nka), intent(in) :: this
    max_vec = nka_max_vec_c(this%handle)
  end function

  real(r8) function vec_tol(this)
    class(nka), intent(in) :: this
    vec_tol = nka_vec_tol_c(this%handle)
  end function

  subroutine accel_update(this, f)
    class(nka), intent(inout) :: this
    class(vector), intent(inout) :: f
    select type (f)
    class is (gpu_vector)
      call nka_accel_update_vec(this%handle, f%vec)
    class default
      error stop 'incompatible arguments to NKA%ACCEL_UPDATE'
    end select
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
