!!
!! NKA_EXAMPLE_C -- ISO_C_BINDING interfaces to the example's device-resident system and solver
!! (include/nka_example.h) and to nka_comm_mode (include/nka_b200.h).
!!
!! What a maintainer of the reference's nka_example.F90 binds to keep residual, pc_ssor and the
!! Picard loop on the GPU (system_type / solver_type, src-F08/nka_example.F90:67-258), on one
!! device or on row slabs across devices.  Pure declarations: no executable code.
!!
!! NOT COMPILED in the build image (it has no Fortran compiler).  Standard F2003 interoperability
!! only; every symbol below is checked from the C side by tests/test_fortran_abi.py.  Optional
!! output arrays that C accepts as NULL are passed as type(c_ptr), value (use c_loc / c_null_ptr).
!!

module nka_example_c

  use, intrinsic :: iso_c_binding
  implicit none
  public

  integer(c_int), parameter :: NKA_FIELD_U = 0, NKA_FIELD_R = 1, NKA_FIELD_Z = 2, &
                               NKA_FIELD_AXL = 3, NKA_FIELD_AYD = 4, NKA_FIELD_AC = 5

  interface

    !! NKASYS nka_system_init (int nx, int ny, double a, int scaling, int device, void *stream)
    function nka_system_init(nx, ny, a, scaling, device, stream) bind(C, name='nka_system_init') result(sys)
      import :: c_ptr, c_int, c_double
      integer(c_int), value :: nx
      integer(c_int), value :: ny
      real(c_double), value :: a
      integer(c_int), value :: scaling
      integer(c_int), value :: device
      type(c_ptr),    value :: stream
      type(c_ptr) :: sys
    end function

    !! NKASYS nka_system_init_slab (int nx, int ny_global, int k0, int k1, double a, int scaling, int device, void *stream)
    function nka_system_init_slab(nx, ny_global, k0, k1, a, scaling, device, stream) &
        bind(C, name='nka_system_init_slab') result(sys)
      import :: c_ptr, c_int, c_double
      integer(c_int), value :: nx
      integer(c_int), value :: ny_global
      integer(c_int), value :: k0
      integer(c_int), value :: k1
      real(c_double), value :: a
      integer(c_int), value :: scaling
      integer(c_int), value :: device
      type(c_ptr),    value :: stream
      type(c_ptr) :: sys
    end function

    !! int nka_system_comm_init (NKASYS, int nranks, int rank, const void *id128)
    function nka_system_comm_init(sys, nranks, rank, id128) bind(C, name='nka_system_comm_init') result(rc)
      import :: c_ptr, c_int, c_char
      type(c_ptr),    value :: sys
      integer(c_int), value :: nranks
      integer(c_int), value :: rank
      character(kind=c_char), intent(in) :: id128(128)
      integer(c_int) :: rc
    end function

    !! void nka_comm_share_system (NKA, NKASYS)
    subroutine nka_comm_share_system(handle, sys) bind(C, name='nka_comm_share_system')
      import :: c_ptr
      type(c_ptr), value :: handle
      type(c_ptr), value :: sys
    end subroutine

    !! int nka_comm_mode (NKA)
    function nka_comm_mode(handle) bind(C, name='nka_comm_mode') result(mode)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: mode
    end function

    !! void nka_system_delete (NKASYS)
    subroutine nka_system_delete(sys) bind(C, name='nka_system_delete')
      import :: c_ptr
      type(c_ptr), value :: sys
    end subroutine

    !! size_t nka_system_size (NKASYS)
    function nka_system_size(sys) bind(C, name='nka_system_size') result(n)
      import :: c_ptr, c_size_t
      type(c_ptr), value :: sys
      integer(c_size_t) :: n
    end function

    !! void *nka_system_stream (NKASYS)
    function nka_system_stream(sys) bind(C, name='nka_system_stream') result(stream)
      import :: c_ptr
      type(c_ptr), value :: sys
      type(c_ptr) :: stream
    end function

    !! double *nka_system_field (NKASYS, int field) -- a device address
    function nka_system_field(sys, field) bind(C, name='nka_system_field') result(dev)
      import :: c_ptr, c_int
      type(c_ptr),    value :: sys
      integer(c_int), value :: field
      type(c_ptr) :: dev
    end function

    !! size_t nka_system_index (NKASYS, int j, int k)
    function nka_system_index(sys, j, k) bind(C, name='nka_system_index') result(pos)
      import :: c_ptr, c_int, c_size_t
      type(c_ptr),    value :: sys
      integer(c_int), value :: j
      integer(c_int), value :: k
      integer(c_size_t) :: pos
    end function

    !! void nka_system_set_field (NKASYS, int field, const double *host)
    subroutine nka_system_set_field(sys, field, host) bind(C, name='nka_system_set_field')
      import :: c_ptr, c_int, c_double
      type(c_ptr),    value :: sys
      integer(c_int), value :: field
      real(c_double), intent(in) :: host(*)
    end subroutine

    !! void nka_system_get_field (NKASYS, int field, double *host)
    subroutine nka_system_get_field(sys, field, host) bind(C, name='nka_system_get_field')
      import :: c_ptr, c_int, c_double
      type(c_ptr),    value :: sys
      integer(c_int), value :: field
      real(c_double), intent(out) :: host(*)
    end subroutine

    !! double nka_system_residual (NKASYS, int subtract_z)     residual(uext, r): src-F08/nka_example.F90:103-145
    function nka_system_residual(sys, subtract_z) bind(C, name='nka_system_residual') result(rnorm)
      import :: c_ptr, c_int, c_double
      type(c_ptr),    value :: sys
      integer(c_int), value :: subtract_z
      real(c_double) :: rnorm
    end function

    !! int nka_system_pc_ssor (NKASYS, int nsweep, double omega)     pc_ssor: src-F08/nka_example.F90:147-179
    function nka_system_pc_ssor(sys, nsweep, omega) bind(C, name='nka_system_pc_ssor') result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr),    value :: sys
      integer(c_int), value :: nsweep
      real(c_double), value :: omega
      integer(c_int) :: rc
    end function

    !! int nka_example_solve (NKASYS, NKA acc, int nsweep, double omega, int maxitr, double tol,
    !!                        double *rnorm, int *nvec_seq)          solve: src-F08/nka_example.F90:226-256
    function nka_example_solve(sys, acc, nsweep, omega, maxitr, tol, rnorm, nvec_seq) &
        bind(C, name='nka_example_solve') result(iters)
      import :: c_ptr, c_int, c_double
      type(c_ptr),    value :: sys
      type(c_ptr),    value :: acc
      integer(c_int), value :: nsweep
      real(c_double), value :: omega
      integer(c_int), value :: maxitr
      real(c_double), value :: tol
      real(c_double), intent(out) :: rnorm(*)
      type(c_ptr),    value :: nvec_seq
      integer(c_int) :: iters
    end function

    !! int nka_system_ssor_trace (NKASYS, int on, unsigned long long *out)
    function nka_system_ssor_trace(sys, on, out) bind(C, name='nka_system_ssor_trace') result(nstrips)
      import :: c_ptr, c_int
      type(c_ptr),    value :: sys
      integer(c_int), value :: on
      type(c_ptr),    value :: out
      integer(c_int) :: nstrips
    end function

    !! unsigned long long nka_example_division_check (unsigned long long nsamples, unsigned long long seed, int device)
    function nka_example_division_check(nsamples, seed, device) bind(C, name='nka_example_division_check') result(nbad)
      import :: c_int, c_long_long
      integer(c_long_long), value :: nsamples
      integer(c_long_long), value :: seed
      integer(c_int),       value :: device
      integer(c_long_long) :: nbad
    end function

    !! void nka_system_timing_enable (NKASYS, int on)
    subroutine nka_system_timing_enable(sys, on) bind(C, name='nka_system_timing_enable')
      import :: c_ptr, c_int
      type(c_ptr),    value :: sys
      integer(c_int), value :: on
    end subroutine

    !! void nka_system_timing_read (NKASYS, double ms[2], unsigned long long count[2])
    subroutine nka_system_timing_read(sys, ms, count) bind(C, name='nka_system_timing_read')
      import :: c_ptr
      type(c_ptr), value :: sys
      type(c_ptr), value :: ms
      type(c_ptr), value :: count
    end subroutine

    !! unsigned long long nka_system_launch_count (NKASYS)
    function nka_system_launch_count(sys) bind(C, name='nka_system_launch_count') result(n)
      import :: c_ptr, c_long_long
      type(c_ptr), value :: sys
      integer(c_long_long) :: n
    end function

  end interface

end module nka_example_c
