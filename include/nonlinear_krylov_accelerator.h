/* nonlinear_krylov_accelerator.h -- the reference's C interface, served by
 * libnka_b200.so (hand-written sm_100a CUDA; no CPU compute path).
 *
 * Drop-in for /root/reference/src-C/nonlinear_krylov_accelerator.h:3-12: same
 * nine symbols, same signatures, same meaning
 * (documentation: src-C/nonlinear_krylov_accelerator.c:57-131).
 *
 * What changes behind the interface:
 *   - the subspace lives in device memory (HBM); nka_init allocates it on the
 *     current CUDA device, nka_delete frees it;
 *   - `f` passed to nka_accel_update may be a DEVICE pointer (zero copy, the
 *     intended use: iterates never cross PCIe) or a HOST pointer (staged
 *     host->device->host inside the call, synchronous, for drop-in use);
 *   - `dp` (src-C/...c:61-68, :227-231): NULL = the built-in reductions (one GPU, or
 *     all ranks of nka_comm_init in nka_b200.h).  A non-NULL dp is honoured as what the
 *     reference documents it for -- the GLOBAL sum in a parallel run: the device forms this
 *     process's partial dot products over its portion of the vectors, and each partial p is
 *     made global by calling dp(1, &p, &one) on the host (p times 1.0, summed over the
 *     processes by the caller's own reduction).  The n-long vectors are never handed to dp:
 *     it must be a plain Euclidean dot product followed by a sum over processes.  This path
 *     costs one device->host->device round trip per update (the 2*(mvec+1) scalars).
 *   - violated preconditions and CUDA failures print "file:line: message" on
 *     stderr and abort(), the reference's ASSERT convention
 *     (src-C/...c:217-219, src-F08/f90_assert.F90:37-47).
 */
#ifndef NONLINEAR_KRYLOV_ACCELERATOR_H
#define NONLINEAR_KRYLOV_ACCELERATOR_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nka_state *NKA;

/* src-C/nonlinear_krylov_accelerator.h:4 ; .c:211-258 */
extern NKA nka_init (int vlen, int mvec, double vtol, double (*dp)(int, double *, double *));
/* .h:5 ; .c:261-282 */
extern void nka_delete (NKA);
/* .h:6 ; .c:285-444 -- f (device or host pointer, vlen doubles) is overwritten
 * with the accelerated correction */
extern void nka_accel_update (NKA, double *f);
/* .h:7 ; .c:447-463 */
extern void nka_restart (NKA);
/* .h:8 ; .c:466-485 */
extern void nka_relax (NKA);
/* .h:9 ; .c:488-499 */
extern int nka_num_vec (NKA);
/* .h:10 ; .c:502 */
extern int nka_max_vec (NKA);
/* .h:11 ; .c:504 */
extern int nka_vec_len (NKA);
/* .h:12 ; .c:506 */
extern double nka_vec_tol (NKA);

#ifdef __cplusplus
}
#endif

#endif
