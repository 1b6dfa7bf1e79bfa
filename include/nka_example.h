/* nka_example.h -- the reference example's discrete system and Picard solver on the device.
 *
 * C-ABI mirror of the two modules the reference's nka_example is made of, so that the
 * iterate never crosses PCIe between residual, preconditioner and accel_update:
 *   system_type  (init, residual, pc_ssor)   src-F08/nka_example.F90:67-181
 *                                            src-F95/nka_example.F90:139-222, src-C/nka_example.c:176-333
 *   solver_type  (init, solve)               src-F08/nka_example.F90:187-258, src-C/nka_example.c:109-173
 * The problem: -div((a+u) grad u) = q on the unit square, u = 0 on the boundary, nx x ny
 * cells, mimetic finite differences; Picard iteration preconditioned by SSOR sweeps in
 * lexicographic Gauss-Seidel order, accelerated by NKA.
 *
 * Grid functions live in device memory in WAVEFRONT-MAJOR order: anti-diagonal t = j + k
 * stored contiguously, diagonals packed back to back (nx*ny doubles, no padding).  Every
 * stencil neighbour of a cell on diagonal t sits at the same or adjacent position of
 * diagonal t-1 or t+1, so the exact-order Gauss-Seidel wavefront, the residual stencil and
 * accel_update (which is indifferent to element order) are all unit-stride.  The
 * set/get functions convert from/to the reference's natural order (x fastest).
 */
#ifndef NKA_EXAMPLE_H
#define NKA_EXAMPLE_H

#include <stddef.h>

#include "nonlinear_krylov_accelerator.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nka_system *NKASYS;

/* which grid function */
enum {
  NKA_FIELD_U = 0,    /* the solution u (interior cells)                                  */
  NKA_FIELD_R = 1,    /* the residual r of the last nka_system_residual                   */
  NKA_FIELD_Z = 2,    /* pc_ssor's output = the vector handed to accel_update = the update */
  NKA_FIELD_AXL = 3,  /* ax(j,k): coefficient of the face left of cell (j,k)              */
  NKA_FIELD_AYD = 4,  /* ay(j,k): coefficient of the face below cell (j,k)                */
  NKA_FIELD_AC = 5    /* ac(j,k): diagonal coefficient                                    */
};

/* system%init(a, nx, ny): src-F08/nka_example.F90:86-101.  u starts at 0.
 * scaling 0: the F95 / C flavour (q = hx*hy, face terms (hx/hy)*t: src-C/nka_example.c:97,
 * :227-245); scaling 1: the F08 flavours (q = 1, face terms t*hx**2: src-F08/nka_example.F90:100,
 * :131-135).  device < 0: current device.  stream: cudaStream_t or NULL. */
NKASYS nka_system_init (int nx, int ny, double a, int scaling, int device, void *stream);
/* Row slabs (one process per GPU; BASELINE.json configs[3]): this handle holds rows [k0, k1) of an
 * nx x ny_global grid (the reference's parallel recipe: each processing element passes its
 * portion of the vector, src-F08-vector/README.md:16-22).  nka_system_comm_init is collective;
 * rank r's slab lies directly above rank r-1's.  The residual then exchanges one row of u with
 * each neighbour and returns the GLOBAL norm; pc_ssor continues the exact lexicographic sweep
 * from the rank below (forward) / above (backward), strip by strip, through edge rows written
 * straight into the neighbour's memory (NVLink peer access), so results stay bit-identical to
 * the single-GPU and the CPU sweeps.  nka_comm_share_system makes an accelerator created with
 * vlen = nx*(k1-k0) sum its dot products over the same ranks. */
NKASYS nka_system_init_slab (int nx, int ny_global, int k0, int k1, double a, int scaling, int device, void *stream);
int nka_system_comm_init (NKASYS, int nranks, int rank, const void *id128);
void nka_comm_share_system (NKA, NKASYS);
void nka_system_delete (NKASYS);
size_t nka_system_size (NKASYS);                       /* nx*ny */
void *nka_system_stream (NKASYS);

/* Device pointer to a grid function (wavefront-major, nx*ny doubles).  U and R move between
 * two buffers: ask again after every nka_system_residual. */
double *nka_system_field (NKASYS, int field);
/* Position of cell (j,k), 0-based, in the wavefront-major arrays. */
size_t nka_system_index (NKASYS, int j, int k);
/* host (natural order: cell (j,k) at k*nx + j) <-> device; synchronous */
void nka_system_set_field (NKASYS, int field, const double *host);
void nka_system_get_field (NKASYS, int field, double *host);

/* residual(uext, r) (src-F08/nka_example.F90:103-145): if subtract_z, first u <- u - z
 * (:248); then rebuild the face coefficients from u and evaluate r.  Returns norm2(r)
 * (synchronises the stream). */
double nka_system_residual (NKASYS, int subtract_z);

/* pc_ssor(nsweep, omega, r) (src-F08/nka_example.F90:147-179): z <- nsweep symmetric sweeps
 * of SSOR on A z = r from z = 0, forward then backward, in the reference's lexicographic
 * order and with its operation order (bit-identical to the serial loops).  Asynchronous.
 * Returns 0, or nonzero if a PREVIOUS sweep on this system reported an internal error: the sweeps
 * are only queued here, so the status of THIS call's sweeps (1 = a strip waited longer than its
 * time limit for its upstream strip, 2 = for the neighbouring rank's edge row) is picked up by the
 * next nka_system_residual, which synchronises and reads the device's error word -- as
 * nka_example_solve does every iteration.  The sweep kernels wait on each other across CTAs: their
 * grid is sized to be co-resident (occupancy x SMs); a device shared with other work (another
 * process, MPS, a concurrent kernel of the caller's) can break that, and the symptom is this time-out
 * (~2 s on one GPU, ~2 min across GPUs), not a hang. */
int nka_system_pc_ssor (NKASYS, int nsweep, double omega);

/* solver%solve (src-F08/nka_example.F90:226-256): Picard iteration from the current u;
 * acc may be NULL (unaccelerated).  rnorm[0..maxitr] receives the residual norms,
 * nvec_seq[0..maxitr-1] (may be NULL) num_vec after each accel_update.  Stops when
 * rnorm < tol * rnorm[0].  Returns the number of iterations taken (<= maxitr).
 * acc must have been created with vlen = nx*ny on the system's device and stream. */
int nka_example_solve (NKASYS, NKA acc, int nsweep, double omega, int maxitr, double tol,
                       double *rnorm, int *nvec_seq);

/* Device timing with CUDA events on the system's stream: which 0 = pc_ssor (all sweeps of a
 * call), 1 = residual.  Same protocol as nka_timing_*. */
/* Tuning aid: per-strip timeline of the last SSOR sweep (see nka_example.cu). Returns nstrips. */
int nka_system_ssor_trace (NKASYS, int on, unsigned long long *out);
/* Self-check: the SSOR sweep forms x/ac as (reciprocal of ac, off the dependent chain) + three
 * chained operations; this compares that division with the device's IEEE division on nsamples
 * generated operand pairs (edge mantissas, out-of-range exponents included) and returns the
 * number of quotients that differ in any bit (must be 0). */
unsigned long long nka_example_division_check (unsigned long long nsamples, unsigned long long seed, int device);
void nka_system_timing_enable (NKASYS, int on);
void nka_system_timing_read (NKASYS, double ms[2], unsigned long long count[2]);
unsigned long long nka_system_launch_count (NKASYS);

#ifdef __cplusplus
}
#endif

#endif
