!!
!! NKA_B200_C -- ISO_C_BINDING interfaces to libnka_b200.so
!!
!! One interface block per C entry point of include/nonlinear_krylov_accelerator.h and
!! include/nka_b200.h.  Pure declarations: no executable code.  The three nka_type
!! modules in this directory (F95/, F08/, F08-vector/) are thin layers over these.
!!
!! NOT COMPILED in the build image (it has no Fortran compiler).  Standard F2003
!! interoperability only; every symbol below is exercised with the same argument
!! passing (scalars by value, arrays by address) by tests/test_fortran_abi.py.
!!

module nka_b200_c

  use, intrinsic :: iso_c_binding
  implicit none
  public

  interface

    !! NKA nka_init_ex (size_t vlen, int mvec, double vtol, int device, void *stream)
    function nka_init_ex(vlen, mvec, vtol, device, stream) bind(C, name='nka_init_ex') result(handle)
      import :: c_ptr, c_size_t, c_int, c_double
      integer(c_size_t), value :: vlen
      integer(c_int),    value :: mvec
      real(c_double),    value :: vtol
      integer(c_int),    value :: device
      type(c_ptr),       value :: stream
      type(c_ptr) :: handle
    end function

    !! void nka_delete (NKA)
    subroutine nka_delete_c(handle) bind(C, name='nka_delete')
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine

    !! void nka_accel_update (NKA, double *f) -- f host or device, detected by the library
    subroutine nka_accel_update_c(handle, f) bind(C, name='nka_accel_update')
      import :: c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(inout) :: f(*)
    end subroutine

    !! void nka_accel_update_host (NKA, double *f_host)
    subroutine nka_accel_update_host(handle, f) bind(C, name='nka_accel_update_host')
      import :: c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(inout) :: f(*)
    end subroutine

    !! void nka_accel_update_dev (NKA, double *f_dev) -- f is a device address held in a c_ptr
    subroutine nka_accel_update_dev(handle, f_dev) bind(C, name='nka_accel_update_dev')
      import :: c_ptr
      type(c_ptr), value :: handle
      type(c_ptr), value :: f_dev
    end subroutine

    subroutine nka_restart_c(handle) bind(C, name='nka_restart')
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine

    subroutine nka_relax_c(handle) bind(C, name='nka_relax')
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine

    function nka_num_vec_c(handle) bind(C, name='nka_num_vec') result(n)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: n
    end function

    function nka_max_vec_c(handle) bind(C, name='nka_max_vec') result(n)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: n
    end function

    function nka_vec_len_c(handle) bind(C, name='nka_vec_len') result(n)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: n
    end function

    function nka_vec_len64(handle) bind(C, name='nka_vec_len64') result(n)
      import :: c_ptr, c_size_t
      type(c_ptr), value :: handle
      integer(c_size_t) :: n
    end function

    function nka_vec_tol_c(handle) bind(C, name='nka_vec_tol') result(vtol)
      import :: c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double) :: vtol
    end function

    subroutine nka_set_vec_tol_c(handle, vtol) bind(C, name='nka_set_vec_tol')
      import :: c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: vtol
    end subroutine

    !! void nka_set_dot_prod_ctx (NKA, double (*dp)(int, double *, double *, void *), void *ctx)
    !! dp = c_funloc of a bind(C) trampoline, ctx = whatever it needs to find the user's procedure
    !! pointer; c_null_funptr restores the built-in reductions.
    subroutine nka_set_dot_prod_ctx(handle, dp, ctx) bind(C, name='nka_set_dot_prod_ctx')
      import :: c_ptr, c_funptr
      type(c_ptr), value :: handle
      type(c_funptr), value :: dp
      type(c_ptr), value :: ctx
    end subroutine

    function nka_defined_c(handle) bind(C, name='nka_defined') result(ok)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: ok
    end function

    subroutine nka_set_stream(handle, stream) bind(C, name='nka_set_stream')
      import :: c_ptr
      type(c_ptr), value :: handle, stream
    end subroutine

    subroutine nka_synchronize(handle) bind(C, name='nka_synchronize')
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine

    !! int nka_comm_unique_id (void *id128) ; int nka_comm_init (NKA, int nranks, int rank, const void *id128)
    function nka_comm_unique_id(id128) bind(C, name='nka_comm_unique_id') result(rc)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id128(128)
      integer(c_int) :: rc
    end function

    function nka_comm_init(handle, nranks, rank, id128) bind(C, name='nka_comm_init') result(rc)
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: handle
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id128(128)
      integer(c_int) :: rc
    end function

    !! ---- device vectors (gpu_vector) ----

    function nka_vec_create(n, device, stream) bind(C, name='nka_vec_create') result(v)
      import :: c_ptr, c_size_t, c_int
      integer(c_size_t), value :: n
      integer(c_int),    value :: device
      type(c_ptr),       value :: stream
      type(c_ptr) :: v
    end function

    function nka_vec_clone(src) bind(C, name='nka_vec_clone') result(v)
      import :: c_ptr
      type(c_ptr), value :: src
      type(c_ptr) :: v
    end function

    subroutine nka_vec_destroy(v) bind(C, name='nka_vec_destroy')
      import :: c_ptr
      type(c_ptr), value :: v
    end subroutine

    function nka_vec_size(v) bind(C, name='nka_vec_size') result(n)
      import :: c_ptr, c_size_t
      type(c_ptr), value :: v
      integer(c_size_t) :: n
    end function

    function nka_vec_data(v) bind(C, name='nka_vec_data') result(dptr)
      import :: c_ptr
      type(c_ptr), value :: v
      type(c_ptr) :: dptr
    end function

    subroutine nka_vec_set_host(v, host) bind(C, name='nka_vec_set_host')
      import :: c_ptr, c_double
      type(c_ptr), value :: v
      real(c_double), intent(in) :: host(*)
    end subroutine

    subroutine nka_vec_get_host(v, host) bind(C, name='nka_vec_get_host')
      import :: c_ptr, c_double
      type(c_ptr), value :: v
      real(c_double), intent(out) :: host(*)
    end subroutine

    subroutine nka_vec_copy(dst, src) bind(C, name='nka_vec_copy')
      import :: c_ptr
      type(c_ptr), value :: dst, src
    end subroutine

    subroutine nka_vec_setval(v, val) bind(C, name='nka_vec_setval')
      import :: c_ptr, c_double
      type(c_ptr), value :: v
      real(c_double), value :: val
    end subroutine

    subroutine nka_vec_scale(v, a) bind(C, name='nka_vec_scale')
      import :: c_ptr, c_double
      type(c_ptr), value :: v
      real(c_double), value :: a
    end subroutine

    subroutine nka_vec_update1(y, a, x) bind(C, name='nka_vec_update1')
      import :: c_ptr, c_double
      type(c_ptr), value :: y, x
      real(c_double), value :: a
    end subroutine

    subroutine nka_vec_update2(y, a, x, b) bind(C, name='nka_vec_update2')
      import :: c_ptr, c_double
      type(c_ptr), value :: y, x
      real(c_double), value :: a, b
    end subroutine

    subroutine nka_vec_update3(z, a, x, b, y) bind(C, name='nka_vec_update3')
      import :: c_ptr, c_double
      type(c_ptr), value :: z, x, y
      real(c_double), value :: a, b
    end subroutine

    subroutine nka_vec_update4(z, a, x, b, y, c) bind(C, name='nka_vec_update4')
      import :: c_ptr, c_double
      type(c_ptr), value :: z, x, y
      real(c_double), value :: a, b, c
    end subroutine

    function nka_vec_dot(x, y) bind(C, name='nka_vec_dot') result(dp)
      import :: c_ptr, c_double
      type(c_ptr), value :: x, y
      real(c_double) :: dp
    end function

    function nka_vec_norm2(x) bind(C, name='nka_vec_norm2') result(nrm)
      import :: c_ptr, c_double
      type(c_ptr), value :: x
      real(c_double) :: nrm
    end function

    function nka_vec_comm_init(v, nranks, rank, id128) bind(C, name='nka_vec_comm_init') result(rc)
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: v
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id128(128)
      integer(c_int) :: rc
    end function

    function nka_init_like(proto, mvec, vtol) bind(C, name='nka_init_like') result(handle)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: proto
      integer(c_int), value :: mvec
      real(c_double), value :: vtol
      type(c_ptr) :: handle
    end function

    subroutine nka_accel_update_vec(handle, f) bind(C, name='nka_accel_update_vec')
      import :: c_ptr
      type(c_ptr), value :: handle, f
    end subroutine

  end interface

end module nka_b200_c
