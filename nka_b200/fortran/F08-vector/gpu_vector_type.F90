!!
!! GPU_VECTOR_TYPE -- a concrete extension of the reference's abstract VECTOR class
!! (src-F08-vector/vector_class.F90:90-109) whose data lives in GPU memory.
!!
!! Takes the place grid_vector_type.F90 has in the reference's example: every deferred
!! procedure (clone1, clone2, copy_, setval, scale, update1_..update4_, dot_, norm2) is one
!! CUDA kernel of libnka_b200.so (nka_vec.cu).  With a communicator (comm_init) each rank
!! holds a row slab and dot_/norm2 sum over all ranks, which is all the reference asks of a
!! parallel vector class (src-F08-vector/README.md:16-22).
!!
!! Build against the reference's own vector_class.F90 (it is not copied here).
!! NOT COMPILED in the build image (no Fortran compiler); see nka_b200_c.F90.
!!

module gpu_vector_type

  use, intrinsic :: iso_fortran_env, only: r8 => real64
  use, intrinsic :: iso_c_binding
  use vector_class
  use nka_b200_c
  implicit none
  private

  type, extends(vector), public :: gpu_vector
    type(c_ptr) :: vec = c_null_ptr     ! NKAVEC handle
  contains
    !! Deferred base class procedures
    procedure :: clone1
    procedure :: clone2
    procedure :: copy_
    procedure :: setval
    procedure :: scale
    procedure :: update1_
    procedure :: update2_
    procedure :: update3_
    procedure :: update4_
    procedure :: dot_
    procedure :: norm2 => norm2_
    !! Additional procedures specific to this type
    procedure :: init => init_len
    procedure :: length
    procedure :: device_ptr
    procedure :: set_from_host
    procedure :: get_to_host
    procedure :: comm_init
    procedure :: release
  end type

contains

  subroutine init_len(this, n)
    class(gpu_vector), intent(inout) :: this
    integer, intent(in) :: n
    call this%release
    this%vec = nka_vec_create(int(n, c_size_t), -1_c_int, c_null_ptr)
  end subroutine

  subroutine release(this)
    class(gpu_vector), intent(inout) :: this
    if (c_associated(this%vec)) call nka_vec_destroy(this%vec)
    this%vec = c_null_ptr
  end subroutine

  integer function length(this)
    class(gpu_vector), intent(in) :: this
    length = int(nka_vec_size(this%vec))
  end function

  type(c_ptr) function device_ptr(this)
    class(gpu_vector), intent(in) :: this
    device_ptr = nka_vec_data(this%vec)
  end function

  subroutine set_from_host(this, host)
    class(gpu_vector), intent(inout) :: this
    real(r8), intent(in), contiguous :: host(:)
    call nka_vec_set_host(this%vec, host)
  end subroutine

  subroutine get_to_host(this, host)
    class(gpu_vector), intent(in) :: this
    real(r8), intent(out), contiguous :: host(:)
    call nka_vec_get_host(this%vec, host)
  end subroutine

  subroutine comm_init(this, nranks, rank, id128)
    class(gpu_vector), intent(inout) :: this
    integer, intent(in) :: nranks, rank
    character(kind=c_char), intent(in) :: id128(128)
    if (nka_vec_comm_init(this%vec, int(nranks, c_int), int(rank, c_int), id128) /= 0) &
        error stop 'gpu_vector%comm_init: NCCL communicator creation failed'
  end subroutine

  !! allocate(clone, source=this) in grid_vector: a deep copy
  subroutine clone1(this, clone)
    class(gpu_vector), intent(in) :: this
    class(vector), allocatable, intent(out) :: clone
    type(gpu_vector), allocatable :: tmp
    allocate(tmp)
    tmp%vec = nka_vec_clone(this%vec)
    call move_alloc(tmp, clone)
  end subroutine

  subroutine clone2(this, clone, n)
    class(gpu_vector), intent(in) :: this
    class(vector), allocatable, intent(out) :: clone(:)
    integer, intent(in) :: n
    type(gpu_vector), allocatable :: tmp(:)
    integer :: j
    allocate(tmp(n))
    do j = 1, n
      tmp(j)%vec = nka_vec_clone(this%vec)
    end do
    call move_alloc(tmp, clone)
  end subroutine

  subroutine copy_(dest, src)
    class(gpu_vector), intent(inout) :: dest
    class(vector), intent(in) :: src
    select type (src)
    class is (gpu_vector)
      call nka_vec_copy(dest%vec, src%vec)
    end select
  end subroutine

  subroutine setval(this, val)
    class(gpu_vector), intent(inout) :: this
    real(r8), intent(in) :: val
    call nka_vec_setval(this%vec, real(val, c_double))
  end subroutine

  subroutine scale(this, a)
    class(gpu_vector), intent(inout) :: this
    real(r8), intent(in) :: a
    call nka_vec_scale(this%vec, real(a, c_double))
  end subroutine

  !! y <-- a*x + y
  subroutine update1_(this, a, x)
    class(gpu_vector), intent(inout) :: this
    class(vector), intent(in) :: x
    real(r8), intent(in) :: a
    select type (x)
    class is (gpu_vector)
      call nka_vec_update1(this%vec, real(a, c_double), x%vec)
    end select
  end subroutine

  !! y <-- a*x + b*y
  subroutine update2_(this, a, x, b)
    class(gpu_vector), intent(inout) :: this
    class(vector), intent(in) :: x
    real(r8), intent(in) :: a, b
    select type (x)
    class is (gpu_vector)
      call nka_vec_update2(this%vec, real(a, c_double), x%vec, real(b, c_double))
    end select
  end subroutine

  !! z <-- a*x + b*y + z
  subroutine update3_(this, a, x, b, y)
    class(gpu_vector), intent(inout) :: this
    class(vector), intent(in) :: x, y
    real(r8), intent(in) :: a, b
    select type (x)
    class is (gpu_vector)
      select type (y)
      class is (gpu_vector)
        call nka_vec_update3(this%vec, real(a, c_double), x%vec, real(b, c_double), y%vec)
      end select
    end select
  end subroutine

  !! z <-- a*x + b*y + c*z
  subroutine update4_(this, a, x, b, y, c)
    class(gpu_vector), intent(inout) :: this
    class(vector), intent(in) :: x, y
    real(r8), intent(in) :: a, b, c
    select type (x)
    class is (gpu_vector)
      select type (y)
      class is (gpu_vector)
        call nka_vec_update4(this%vec, real(a, c_double), x%vec, real(b, c_double), y%vec, real(c, c_double))
      end select
    end select
  end subroutine

  function dot_(x, y) result(dp)
    class(gpu_vector), intent(in) :: x
    class(vector), intent(in) :: y
    real(r8) :: dp
    dp = 0.0_r8
    select type (y)
    class is (gpu_vector)
      dp = nka_vec_dot(x%vec, y%vec)
    end select
  end function

  function norm2_(this)
    class(gpu_vector), intent(in) :: this
    real(r8) :: norm2_
    norm2_ = nka_vec_norm2(this%vec)
  end function

end module gpu_vector_type
