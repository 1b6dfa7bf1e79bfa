// nka_aux_kernels.cuh -- the small non-template kernels (state step, fix-up, relax/restart,
// materialise).  Included by nka_capi.cu only.
#pragma once

#include "nka_kernels.cuh"

// ---------------------------------------------------------------------------
// State step as its own kernel (first call: no pass A; multi-GPU: after NCCL).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NKA_STATE_THREADS) nka_state_kernel(NkaDevState* S, const double* dots, int have_last)
{
  __shared__ NkaStateStage sm;
  nka_pdl_wait();
  nka_stage_in(sm, S, dots);
  nka_run_state_step(sm, S, have_last);
}

// ---------------------------------------------------------------------------
// Fix-up for the lazily skipped oldest column: d_0 . d_last and f . d_last, then
// the state step again with every dot product present.  Every CTA leaves at
// once unless the first step asked for it.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NKA_THREADS)
nka_fixup_kernel(const double* __restrict__ f, const double* __restrict__ W, size_t ld, size_t n,
                 NkaDevState* __restrict__ S, double* __restrict__ partials, unsigned* __restrict__ ticket,
                 double* __restrict__ dots, NkaPeerCtx* __restrict__ peer)
{
  nka_pdl_wait();
  if (!S->need_fixup) return;          // identical on every rank: they ran the state step on the same bits
  const NkaPlanA& A = S->planA;
  const int jl = A.ncol - 1;
  const double* w0 = W + (size_t)A.col[0] * ld;
  const double* wl = W + (size_t)A.col[jl] * ld;
  const double* wp = W + (size_t)A.col[jl - 1] * ld;
  const bool sub = (A.submask >> jl) & 1ull;
  double acc[2] = {0.0, 0.0};
  const size_t stride = (size_t)gridDim.x * NKA_THREADS;
  for (size_t i = (size_t)blockIdx.x * NKA_THREADS + threadIdx.x; i < n; i += stride) {
    const double x0 = f[i];
    const double d0 = w0[i] - x0;
    const double dl = sub ? wl[i] - wp[i] : wl[i];
    acc[0] = fma(d0, dl, acc[0]);
    acc[1] = fma(x0, dl, acc[1]);
  }
  __shared__ NkaStateStage sm;
  __shared__ double xv[2];
  const bool last = nka_grid_reduce<2, NKA_THREADS>(acc, partials, ticket, partials, gridDim.x,
                                                  [&](int j, double v) { xv[j] = v; });
  if (last) {
    if (peer) nka_peer_allreduce(peer, xv, 2);
    if (threadIdx.x < 2) dots[(threadIdx.x == 0 ? 0 : NKA_MAXSLOT) + jl] = xv[threadIdx.x];
    __threadfence();
    __syncthreads();
    nka_stage_in(sm, S, dots);
    nka_run_state_step(sm, S, /*have_last=*/1);
  }
}

__global__ void __launch_bounds__(NKA_STATE_THREADS) nka_relax_kernel(NkaDevState* S)
{
  __shared__ NkaStateStage sm;
  nka_stage_in(sm, S, nullptr);
  if (threadIdx.x == 0) nka_state_relax(sm.st);
  __syncthreads();
  nka_stage_out(sm, S);
}

__global__ void __launch_bounds__(NKA_STATE_THREADS) nka_restart_kernel(NkaDevState* S)
{
  __shared__ NkaStateStage sm;
  nka_stage_in(sm, S, nullptr);
  if (threadIdx.x == 0) nka_state_restart(sm.st);
  __syncthreads();
  nka_stage_out(sm, S);
}

__global__ void nka_init_kernel(NkaDevState* S, int mvec, double vtol, int lazy_last)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    nka_state_init(*S, mvec, vtol);
    S->lazy_last = lazy_last;
    nka_build_plan_a(*S);
  }
}

__global__ void nka_set_vtol_kernel(NkaDevState* S, double vtol)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) S->vtol = vtol;
}

__global__ void nka_set_lazy_kernel(NkaDevState* S, int lazy_last)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) { S->lazy_last = lazy_last; nka_build_plan_a(*S); }
}

// ---------------------------------------------------------------------------
// Materialise (relax() only; inside accel_update pass B does it per element):
// W[dst] -= W[sub] for each plan entry, oldest first.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NKA_THREADS)
nka_materialise(double* W, size_t ld, size_t n, const NkaDevState* __restrict__ S)
{
  nka_pdl_wait();
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

