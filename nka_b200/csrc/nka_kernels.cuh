// nka_kernels.cuh -- the sm_100a kernels of accel_update.
//
//   nka_pass_a       one read-only sweep: every difference d_j of the subspace is
//                    formed in registers from the raw cached inputs and reduced
//                    against d_0 and f (2*ncol dot products), deterministic
//                    two-stage reduction, last CTA folds the per-CTA partials.
//                    Replaces src-C/nonlinear_krylov_accelerator.c:299-301,
//                    :323-324 and the dp() calls of :406.
//   nka_state_kernel one thread: nka_state_step() (nka_state.h).  :333-417
//   nka_materialise  rare: W[dst] -= W[sub] after a vtol drop / relax broke a chain.
//   nka_pass_b       one sweep: correction, new cached columns.  :397-398, :419-430
//
// All kernels are HBM-bandwidth bound (0.2-0.4 flop/byte, fp64); tensor cores
// do not apply.  Loads are 16-byte (double2) and coalesced; the grid is a
// multiple of the SM count.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "nka_state.h"

// Tunables (profiles/ records the sweep that picked the defaults).
#ifndef NKA_THREADS
#define NKA_THREADS 256
#endif
#ifndef NKA_MINB_A
#define NKA_MINB_A 1      // __launch_bounds__ min CTAs/SM for pass A (2 caps it at 128 registers)
#endif
#ifndef NKA_MINB_B
#define NKA_MINB_B 1
#endif
#define NKA_STATE_THREADS 128

// ---------------------------------------------------------------------------
// 16-byte / 8-byte element access with streaming cache hints.  V = 2 uses
// double2 (LDG.E.128 / STG.E.128), V = 1 is the fallback for a caller's f that
// is not 16-byte aligned.
// ---------------------------------------------------------------------------
template <int V> struct Vec;
template <> struct Vec<2> {
  double x, y;
  static __device__ __forceinline__ Vec ld(const double* p, size_t i) {
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p) + i);
    return {t.x, t.y};
  }
  static __device__ __forceinline__ Vec ld_keep(const double* p, size_t i) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p) + i);
    return {t.x, t.y};
  }
  __device__ __forceinline__ void st_stream(double* p, size_t i) const {
    __stcs(reinterpret_cast<double2*>(p) + i, make_double2(x, y));
  }
  __device__ __forceinline__ void st(double* p, size_t i) const {
    reinterpret_cast<double2*>(p)[i] = make_double2(x, y);
  }
  static __device__ __forceinline__ Vec zero() { return {0.0, 0.0}; }
  __device__ __forceinline__ Vec operator-(const Vec& o) const { return {x - o.x, y - o.y}; }
  __device__ __forceinline__ Vec operator+(const Vec& o) const { return {x + o.x, y + o.y}; }
  __device__ __forceinline__ Vec scaled(double a) const { return {a * x, a * y}; }
  __device__ __forceinline__ void fma_into(double a, Vec& acc) const { acc.x = fma(a, x, acc.x); acc.y = fma(a, y, acc.y); }
  __device__ __forceinline__ void dot_into(const Vec& o, double& acc) const { acc = fma(x, o.x, acc); acc = fma(y, o.y, acc); }
};
template <> struct Vec<1> {
  double x;
  static __device__ __forceinline__ Vec ld(const double* p, size_t i) { return {__ldcs(p + i)}; }
  static __device__ __forceinline__ Vec ld_keep(const double* p, size_t i) { return {__ldg(p + i)}; }
  __device__ __forceinline__ void st_stream(double* p, size_t i) const { __stcs(p + i, x); }
  __device__ __forceinline__ void st(double* p, size_t i) const { p[i] = x; }
  static __device__ __forceinline__ Vec zero() { return {0.0}; }
  __device__ __forceinline__ Vec operator-(const Vec& o) const { return {x - o.x}; }
  __device__ __forceinline__ Vec operator+(const Vec& o) const { return {x + o.x}; }
  __device__ __forceinline__ Vec scaled(double a) const { return {a * x}; }
  __device__ __forceinline__ void fma_into(double a, Vec& acc) const { acc.x = fma(a, x, acc.x); }
  __device__ __forceinline__ void dot_into(const Vec& o, double& acc) const { acc = fma(x, o.x, acc); }
};

__device__ __forceinline__ double nka_warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// Pass A.
//   acc[j]      += d_0 . d_j      acc[NC + j] += f . d_j       (j < ncol <= NC)
// One body, instantiated for V = 2 on the bulk and V = 1 on the odd tail.
// FULL = the plan streams exactly NC chained columns: no predicates at all.
// ---------------------------------------------------------------------------
template <int NC, int V, bool FULL>
__device__ __forceinline__ void nka_pass_a_elem(const double* __restrict__ f, const double* const (&wcol)[NC],
                                                size_t i, int ncol, unsigned submask, double (&acc)[2 * NC])
{
  using T = Vec<V>;
  const T x0 = T::ld_keep(f, i);          // f is read again by pass B: leave it in L2 if it fits
  T xs[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if (FULL || j < ncol) xs[j] = T::ld(wcol[j], i);
    else xs[j] = T::zero();
  }
  T prev = x0;
  T d0 = T::zero();
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    T d;
    if (FULL) d = xs[j] - prev;
    else d = ((submask >> j) & 1u) ? (xs[j] - prev) : xs[j];
    if (j == 0) d0 = d;
    d0.dot_into(d, acc[j]);
    x0.dot_into(d, acc[NC + j]);
    prev = xs[j];
  }
}

template <int NC, int V>
__global__ void __launch_bounds__(NKA_THREADS, NKA_MINB_A)
nka_pass_a(const double* __restrict__ f, const double* __restrict__ W, size_t ld, size_t n,
           const NkaDevState* __restrict__ S, double* __restrict__ partials, unsigned* __restrict__ ticket,
           double* __restrict__ dots)
{
  const int ncol = S->planA.ncol;
  const unsigned submask = S->planA.submask;
  const double* wcol[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) wcol[j] = W + (size_t)S->planA.col[j < ncol ? j : 0] * ld;

  double acc[2 * NC];
#pragma unroll
  for (int j = 0; j < 2 * NC; ++j) acc[j] = 0.0;

  const size_t nv = n / V;
  const size_t stride = (size_t)gridDim.x * NKA_THREADS;
  const size_t start = (size_t)blockIdx.x * NKA_THREADS + threadIdx.x;
  const bool full = (ncol == NC) && (submask == (NC >= 32 ? 0xffffffffu : ((1u << NC) - 1u)));
  if (full) {
    for (size_t i = start; i < nv; i += stride) nka_pass_a_elem<NC, V, true>(f, wcol, i, ncol, submask, acc);
  } else {
    for (size_t i = start; i < nv; i += stride) nka_pass_a_elem<NC, V, false>(f, wcol, i, ncol, submask, acc);
  }
  if (V == 2 && (n & 1) && start == 0) nka_pass_a_elem<NC, 1, false>(f, wcol, n - 1, ncol, submask, acc);

  // block reduction: shuffle tree inside each warp, fixed-order sum across warps
  __shared__ double red[NKA_THREADS / 32][2 * NC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 2 * NC; ++j) {
    const double v = nka_warp_sum(acc[j]);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2 * NC) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NKA_THREADS / 32; ++w) v += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * (2 * NC) + threadIdx.x] = v;
  }

  // last CTA to finish folds the per-CTA partials in a fixed order (run-to-run bit-stable)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int j = warp; j < 2 * NC; j += NKA_THREADS / 32) {
    double v = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(&partials[(size_t)b * (2 * NC) + j]);
    v = nka_warp_sum(v);
    if (lane == 0) dots[(j < NC) ? j : (NKA_MAXSLOT + (j - NC))] = v;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// ---------------------------------------------------------------------------
// State kernels.  The ~10 KB state is staged through shared memory: the scalar
// algorithm is a chain of dependent loads, and from global memory every one of
// them paid an L2 round trip (43 us at mvec = 10 on B200; see profiles/).
// ---------------------------------------------------------------------------
struct NkaStateStage {
  NkaDevState st;
  double dots[2 * NKA_MAXSLOT];
};

__device__ __forceinline__ void nka_stage_in(NkaStateStage& sm, const NkaDevState* S, const double* dots)
{
  static_assert(sizeof(NkaDevState) % 4 == 0, "state is copied in 4-byte words");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(S);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.st);
  for (unsigned i = threadIdx.x; i < sizeof(NkaDevState) / 4; i += blockDim.x) dst[i] = __ldcg(src + i);
  if (dots)
    for (unsigned i = threadIdx.x; i < 2 * NKA_MAXSLOT; i += blockDim.x) sm.dots[i] = __ldcg(dots + i);
  __syncthreads();
}

__device__ __forceinline__ void nka_stage_out(const NkaStateStage& sm, NkaDevState* S)
{
  __syncthreads();
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&sm.st);
  uint32_t* dst = reinterpret_cast<uint32_t*>(S);
  for (unsigned i = threadIdx.x; i < sizeof(NkaDevState) / 4; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(NKA_STATE_THREADS) nka_state_kernel(NkaDevState* S, const double* dots)
{
  __shared__ NkaStateStage sm;
  nka_stage_in(sm, S, dots);
  if (threadIdx.x == 0) nka_state_step(sm.st, sm.dots, NKA_MAXSLOT);
  nka_stage_out(sm, S);
}

__global__ void __launch_bounds__(NKA_STATE_THREADS) nka_relax_kernel(NkaDevState* S)
{
  __shared__ NkaStateStage sm;
  nka_stage_in(sm, S, nullptr);
  if (threadIdx.x == 0) nka_state_relax(sm.st);
  nka_stage_out(sm, S);
}

__global__ void __launch_bounds__(NKA_STATE_THREADS) nka_restart_kernel(NkaDevState* S)
{
  __shared__ NkaStateStage sm;
  nka_stage_in(sm, S, nullptr);
  if (threadIdx.x == 0) nka_state_restart(sm.st);
  nka_stage_out(sm, S);
}

__global__ void nka_init_kernel(NkaDevState* S, int mvec, double vtol)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) nka_state_init(*S, mvec, vtol);
}

__global__ void nka_set_vtol_kernel(NkaDevState* S, double vtol)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) S->vtol = vtol;
}

// ---------------------------------------------------------------------------
// Materialise: W[dst] -= W[sub] for each plan entry, oldest first, per element.
// Exits at once when the plan is empty (the common case).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NKA_THREADS)
nka_materialise(double* W, size_t ld, size_t n, const NkaDevState* __restrict__ S)
{
  const int m = S->planM.n;
  if (m == 0) return;
  const size_t stride = (size_t)gridDim.x * NKA_THREADS;
  for (size_t i = (size_t)blockIdx.x * NKA_THREADS + threadIdx.x; i < n; i += stride) {
    for (int e = 0; e < m; ++e) {
      double* dst = W + (size_t)S->planM.dst[e] * ld;
      const double* sub = W + (size_t)S->planM.sub[e] * ld;
      dst[i] = dst[i] - sub[i];
    }
  }
}

// ---------------------------------------------------------------------------
// Pass B.  NZ is the host's expectation of how many older Z columns the plan
// keeps (exact unless a vtol drop or the s == 0 guard fired); the FULL body is
// predicate-free.  Any other plan goes through the general body, which also
// handles nz > NZ with a run-time loop over the extra columns.
// ---------------------------------------------------------------------------
template <int NZ, int V, bool FULL>
__device__ __forceinline__ void nka_pass_b_elem(double* __restrict__ f, double* __restrict__ wnew, double* __restrict__ znew,
                                                double* __restrict__ zp, const double* const (&zcol)[NZ > 0 ? NZ : 1],
                                                const double (&coef)[NZ > 0 ? NZ : 1], double coef_p,
                                                int has_pair, int nz, int write_f, size_t i,
                                                const double* __restrict__ Z, size_t ld, const NkaPlanB* __restrict__ B)
{
  using T = Vec<V>;
  const T x0 = T::ld(f, i);
  T zs[NZ > 0 ? NZ : 1];
#pragma unroll
  for (int k = 0; k < NZ; ++k) {
    if (FULL || k < nz) zs[k] = T::ld(zcol[k], i);
    else zs[k] = T::zero();
  }
  T y = T::zero();
  if (FULL || has_pair) {
    const T zpv = T::ld(zp, i) + x0;        // Z'_p = Y_p + f
    zpv.st_stream(zp, i);
    zpv.fma_into(coef_p, y);
  }
#pragma unroll
  for (int k = 0; k < NZ; ++k) zs[k].fma_into(coef[k], y);   // coef is 0 beyond nz
  if (!FULL) {
    for (int k = NZ; k < nz; ++k) {
      const T z = T::ld(Z + (size_t)B->zcol[k] * ld, i);
      z.fma_into(B->coef[k], y);
    }
  }
  y.st_stream(znew, i);
  x0.st_stream(wnew, i);
  if (FULL || write_f) (x0 + y).st(f, i);
}

template <int NZ, int V>
__global__ void __launch_bounds__(NKA_THREADS, NKA_MINB_B)
nka_pass_b(double* __restrict__ f, double* __restrict__ W, double* __restrict__ Z, size_t ld, size_t n,
           const NkaDevState* __restrict__ S)
{
  constexpr int NZA = NZ > 0 ? NZ : 1;
  const NkaPlanB* B = &S->planB;
  const int nz = B->nz, has_pair = B->has_pair, write_f = B->write_f;
  double* wnew = W + (size_t)B->newslot * ld;
  double* znew = Z + (size_t)B->newslot * ld;
  double* zp = Z + (size_t)B->pslot * ld;
  const double coef_p = B->coef_p;
  const double* zcol[NZA];
  double coef[NZA];
#pragma unroll
  for (int k = 0; k < NZA; ++k) {
    const bool on = (k < nz) && (k < NZ);
    zcol[k] = Z + (size_t)(on ? B->zcol[k] : B->newslot) * ld;
    coef[k] = on ? B->coef[k] : 0.0;
  }
  const size_t nv = n / V;
  const size_t stride = (size_t)gridDim.x * NKA_THREADS;
  const size_t start = (size_t)blockIdx.x * NKA_THREADS + threadIdx.x;
  const bool full = (nz == NZ) && has_pair && write_f;
  if (full) {
    for (size_t i = start; i < nv; i += stride)
      nka_pass_b_elem<NZ, V, true>(f, wnew, znew, zp, zcol, coef, coef_p, has_pair, nz, write_f, i, Z, ld, B);
  } else {
    for (size_t i = start; i < nv; i += stride)
      nka_pass_b_elem<NZ, V, false>(f, wnew, znew, zp, zcol, coef, coef_p, has_pair, nz, write_f, i, Z, ld, B);
  }
  if (V == 2 && (n & 1) && start == 0)
    nka_pass_b_elem<NZ, 1, false>(f, wnew, znew, zp, zcol, coef, coef_p, has_pair, nz, write_f, n - 1, Z, ld, B);
}
