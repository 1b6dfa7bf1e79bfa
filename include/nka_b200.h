/* nka_b200.h -- additive C-ABI entry points of libnka_b200.so.
 *
 * Everything a host language needs beyond the reference's nine C functions
 * (include/nonlinear_krylov_accelerator.h): the parts of the Fortran API the C
 * header lacks, 64-bit lengths, device pointers and streams, the built-in
 * multi-GPU dot-product reduction, and read-only introspection for parity
 * tests and benchmarks.  Plain pointers and sizes only; no C++/torch types.
 *
 * These are the symbols the Fortran modules in nka_b200/fortran/ bind with
 * ISO_C_BINDING (see INTEGRATION.md for the interface blocks).
 */
#ifndef NKA_B200_H
#define NKA_B200_H

#include <stddef.h>

#include "nonlinear_krylov_accelerator.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NKA_B200_MAX_MVEC 32

/* ---- construction ------------------------------------------------------ */

/* nka_init with a 64-bit length (the reference's int arithmetic overflows at
 * (mvec+1)*vlen >= 2^31, src-C/nonlinear_krylov_accelerator.c:235,241).
 * device < 0: the current CUDA device.  stream: a cudaStream_t, or NULL for the
 * legacy default stream (work is then ordered after whatever the caller queued
 * there, as with the synchronous CPU interface this replaces).  Fortran: init, src-F08/nka_type.F90:185-200
 * (vtol defaults to 0.01 there, :160). */
NKA nka_init_ex (size_t vlen, int mvec, double vtol, int device, void *stream);

/* src-F08/nka_type.F90:202-207 (set_vec_tol), src-F95/nka_type.F90:227-232 */
void nka_set_vec_tol (NKA, double vtol);

/* src-F08/nka_type.F90:460-524 (defined), src-F95/nka_type.F90:511-578.
 * Copies the small device state to the host and checks every list invariant. */
int nka_defined (NKA);

size_t nka_vec_len64 (NKA);

/* src-F08/nka_type.F90:209-219 (set_dot_prod), and the per-call `dp` of src-F95/nka_type.F90:284-291:
 * install (or with dp == NULL remove) the global-sum callback after construction; see
 * nonlinear_krylov_accelerator.h for how dp is used.  The _ctx form passes a caller context
 * through (what a Fortran trampoline needs to find its procedure pointer). */
void nka_set_dot_prod (NKA, double (*dp)(int, double *, double *));
void nka_set_dot_prod_ctx (NKA, double (*dp)(int, double *, double *, void *), void *ctx);

/* ---- the hot path with explicit memory spaces -------------------------- */

/* f is a device pointer on the handle's device; asynchronous on the handle's
 * stream (kernels only, no host synchronisation). */
void nka_accel_update_dev (NKA, double *f_dev);
/* f is a host pointer; copies in, updates, copies out, returns when f holds
 * the result. */
void nka_accel_update_host (NKA, double *f_host);

void nka_set_stream (NKA, void *stream);
void *nka_get_stream (NKA);
/* Block until everything queued on the handle's stream has finished. */
void nka_synchronize (NKA);

/* ---- multi-GPU: one process per GPU, each holding a row slab ------------ */
/* The only exchange in accel_update is one fp64 sum-allreduce of the
 * 2*(mvec+1) partial dot products (the reference's `dp` hook contract:
 * src-C/...c:61-68, src-F08-vector/README.md:16-22).  vlen passed to
 * nka_init_ex is the LOCAL slab length.
 *   nka_comm_unique_id: rank 0 fills a 128-byte NCCL id; ship it to the other
 *   ranks with whatever the application has (MPI, torch.distributed, files).
 *   nka_comm_init: collective over all ranks.  Returns 0 on success. */
int nka_comm_unique_id (void *id128);
int nka_comm_init (NKA, int nranks, int rank, const void *id128);
/* Alternatively adopt an existing ncclComm_t (not owned). */
void nka_comm_adopt (NKA, void *nccl_comm, int nranks, int rank);
/* How the partial dot products are summed: 0 = single GPU (no exchange); 1 = one NCCL
 * all-reduce between pass A and the state kernel; 2 = fused into pass A through peer memory
 * (every rank maps every peer's exchange box with CUDA IPC; chosen automatically by
 * nka_comm_init / nka_comm_adopt when all ranks share one NVLink domain, unanimous across
 * ranks; NKA_PEER_REDUCE=0 in the environment of every rank forces 1). */
int nka_comm_mode (NKA);

/* ---- device vectors: the operations of the reference's abstract `vector` ---- */
/* What a concrete gpu_vector extension of src-F08-vector/vector_class.F90:90-109
 * needs.  Expression order follows grid_vector (grid_vector_type.F90:108-197).
 * dot/norm2 return the value to the host (they synchronise the vector's stream)
 * and, with a communicator, sum over all ranks' slabs. */
typedef struct nka_vec *NKAVEC;
NKAVEC nka_vec_create (size_t n, int device, void *stream);   /* contents undefined, like allocate() */
NKAVEC nka_vec_clone (NKAVEC src);                            /* allocate(clone, source=src): :86-97 */
void nka_vec_destroy (NKAVEC);
size_t nka_vec_size (NKAVEC);
double *nka_vec_data (NKAVEC);                                /* device pointer */
void nka_vec_set_host (NKAVEC, const double *host);           /* host -> device, synchronous */
void nka_vec_get_host (NKAVEC, double *host);                 /* device -> host, synchronous */
void nka_vec_copy (NKAVEC dst, NKAVEC src);                   /* copy_   :99-106 */
void nka_vec_setval (NKAVEC, double val);                     /* setval  :108-112 */
void nka_vec_scale (NKAVEC, double a);                        /* scale   :114-118 */
void nka_vec_update1 (NKAVEC y, double a, NKAVEC x);                       /* y = a*x + y        :121-129 */
void nka_vec_update2 (NKAVEC y, double a, NKAVEC x, double b);             /* y = a*x + b*y      :132-140 */
void nka_vec_update3 (NKAVEC z, double a, NKAVEC x, double b, NKAVEC y);   /* z = a*x + b*y + z  :143-154 */
void nka_vec_update4 (NKAVEC z, double a, NKAVEC x, double b, NKAVEC y, double c); /* z = a*x+b*y+c*z :157-168 */
double nka_vec_dot (NKAVEC x, NKAVEC y);                      /* dot_    :170-184 */
double nka_vec_norm2 (NKAVEC x);                              /* norm2   :186-197 */
int nka_vec_comm_init (NKAVEC, int nranks, int rank, const void *id128); /* clones share it */
/* The vector flavour's init(vec, mvec) and accel_update(f):
 * src-F08-vector/nka_type.F90:175-188, :219-222. */
NKA nka_init_like (NKAVEC proto, int mvec, double vtol);
void nka_accel_update_vec (NKA, NKAVEC f);

/* ---- introspection (tests, benchmarks; never on the hot path) ---------- */

typedef struct nka_state_view {
  int mvec, subspace, pending, first, last, free_slot;
  int next[NKA_B200_MAX_MVEC + 1], prev[NKA_B200_MAX_MVEC + 1];
  int chained[NKA_B200_MAX_MVEC + 1];
  int ndrop_last, evicted_last, relaxed_last, error;
  double vtol, min_margin, s_last;
  double c[NKA_B200_MAX_MVEC + 1];
  double s[NKA_B200_MAX_MVEC + 1];
  double h[(NKA_B200_MAX_MVEC + 1) * (NKA_B200_MAX_MVEC + 1)]; /* h[r*(mvec+1)+c], compacted */
  unsigned long long ncalls;
} nka_state_view;

void nka_get_state (NKA, nka_state_view *out);

/* Kernel launches issued by this handle so far (all of them ours). */
unsigned long long nka_launch_count (NKA);

/* Per-kernel device timing with CUDA events on the handle's stream.
 * which: 0 = pass A, 1 = state, 2 = materialise, 3 = pass B, 4 = allreduce.
 * nka_timing_read synchronises the stream and returns accumulated ms and
 * launch counts since the last nka_timing_reset. */
void nka_timing_enable (NKA, int on);
void nka_timing_reset (NKA);
void nka_timing_read (NKA, double ms[5], unsigned long long count[5]);

/* Grid/block geometry of the two streaming kernels for the next update;
 * *threads = 10000 * (pass A block size) + (pass B block size). */
void nka_launch_geometry (NKA, int *grid_a, int *grid_b, int *threads);

const char *nka_b200_version (void);

#ifdef __cplusplus
}
#endif

#endif
