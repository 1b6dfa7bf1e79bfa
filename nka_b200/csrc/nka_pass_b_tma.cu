// nka_pass_b_tma.cu -- pass B with its operand tiles staged through shared memory by the TMA unit
// (cp.async.bulk, 1-D, completion on an mbarrier ring) instead of per-thread LDG.E.128.
//
// This is the recorded experiment DESIGN.md section 4 refers to: same arithmetic, same fma order
// and same stores as nka_pass_b<NZ,2>'s predicate-free body (nka_kernels.cuh), only the way the
// NZ + 1 input streams reach the SM differs.  Selected with NKA_PASS_B_TMA=1 for the steady-state
// plan (every streamed pair present, nothing to materialise, even n, 16-byte aligned f); any other
// plan runs the regular kernel.  Instantiated for the subspace sizes of the BASELINE configs only.
//
//   warp 8 (one elected lane) = producer: for each tile waits for the stage to be released, arms
//       the stage's `full` barrier with the byte count and issues NZ + 1 bulk copies (f tile and
//       one tile per Z column) of TILE doubles each;
//   warps 0-7 = consumers: wait on `full`, read their double2 of every column (LDS.128), release
//       the stage (arrive on `empty`), then do the fma chains and the three STG.E.128 stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "nka_dispatch.h"
#include "nka_kernels.cuh"

#define NKA_TMA_CONSUMERS 256
#define NKA_TMA_THREADS (NKA_TMA_CONSUMERS + 32)
#define NKA_TMA_TILE (2 * NKA_TMA_CONSUMERS)         // doubles per column per tile: one double2 per consumer thread

__device__ __forceinline__ uint32_t nka_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void nka_mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(nka_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void nka_mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(nka_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nka_mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(nka_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nka_mbar_wait(uint64_t* bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" :: "r"(nka_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void nka_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(nka_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(nka_smem_u32(bar)) : "memory");
}

template <int NZ> struct NkaTmaCfg {
  static constexpr int kStageBytes = (NZ + 1) * NKA_TMA_TILE * 8;
  static constexpr int kStages = (200 * 1024) / kStageBytes < 2 ? 2 : ((200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStages * 8 + 128;
};

template <int NZ>
__global__ void __launch_bounds__(NKA_TMA_THREADS, 1)
nka_pass_b_tma(double* __restrict__ f, double* W, double* Z, size_t ld, size_t n, const NkaDevState* __restrict__ S)
{
  using Cfg = NkaTmaCfg<NZ>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bar_empty = bar_full + Cfg::kStages;

  nka_pdl_wait();
  const NkaPlanB* B = &S->planB;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { nka_mbar_init(&bar_full[s], 1); nka_mbar_init(&bar_empty[s], NKA_TMA_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // anything but the steady-state plan: the regular kernel's general body, no staging
  const bool steady = (B->nz == NZ) && B->has_pair && B->write_f && (S->planM.n == 0);
  if (!steady) {
    if (threadIdx.x >= NKA_TMA_CONSUMERS) return;
    const int nz = B->nz, has_pair = B->has_pair, write_f = B->write_f;
    const double* zcol[NZ];
    double cN[NZ], cY[NZ];
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      const bool on = k < nz;
      zcol[k] = Z + (size_t)(on ? B->zcol[k] : B->newslot) * ld;
      cN[k] = on ? B->coefN[k] : 0.0;
      cY[k] = on ? B->coefY[k] : 0.0;
    }
    double* wnew_g = W + (size_t)B->newslot * ld;
    double* zp_g = Z + (size_t)B->pslot * ld;
    for (size_t i = (size_t)blockIdx.x * NKA_TMA_CONSUMERS + threadIdx.x; i < n / 2; i += (size_t)gridDim.x * NKA_TMA_CONSUMERS)
      nka_pass_b_elem<NZ, 2, false>(Vec<2>::ld_plain(f, i), f, wnew_g, zp_g, zcol, cN, cY, B->coef_p, has_pair, nz, write_f, i, W, Z, ld, S);
    return;
  }

  const size_t ntiles = (n + NKA_TMA_TILE - 1) / NKA_TMA_TILE;
  if (threadIdx.x >= NKA_TMA_CONSUMERS) {
    // ---- producer ----
    if (threadIdx.x == NKA_TMA_CONSUMERS) {
      const double* src[NZ + 1];
      src[0] = f;
#pragma unroll
      for (int k = 0; k < NZ; ++k) src[k + 1] = Z + (size_t)B->zcol[k] * ld;
      unsigned it = 0;
      for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % Cfg::kStages;
        const unsigned round = it / Cfg::kStages;
        if (round > 0) nka_mbar_wait(&bar_empty[s], (round - 1) & 1u);
        const size_t e0 = t * NKA_TMA_TILE;
        const unsigned cnt = (unsigned)((n - e0) < (size_t)NKA_TMA_TILE ? (n - e0) : (size_t)NKA_TMA_TILE);
        const unsigned bytes = cnt * 8u;                       // n is even: a multiple of 16
        nka_mbar_expect_tx(&bar_full[s], bytes * (NZ + 1));
        double* dst = tiles + (size_t)s * (NZ + 1) * NKA_TMA_TILE;
#pragma unroll
        for (int c = 0; c <= NZ; ++c) nka_bulk_g2s(dst + (size_t)c * NKA_TMA_TILE, src[c] + e0, bytes, &bar_full[s]);
      }
    }
    return;
  }

  // ---- consumers ----
  double* wnew = W + (size_t)B->newslot * ld;
  double* zp = Z + (size_t)B->pslot * ld;
  const double coef_p = B->coef_p;
  double coefN[NZ], coefY[NZ];
#pragma unroll
  for (int k = 0; k < NZ; ++k) { coefN[k] = B->coefN[k]; coefY[k] = B->coefY[k]; }
  using T = Vec<2>;
  unsigned it = 0;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % Cfg::kStages;
    nka_mbar_wait(&bar_full[s], (it / Cfg::kStages) & 1u);
    const double2* st = reinterpret_cast<const double2*>(tiles + (size_t)s * (NZ + 1) * NKA_TMA_TILE);
    const size_t i = t * (NKA_TMA_TILE / 2) + threadIdx.x;     // double2 index in the vector
    const bool live = 2 * i < n;
    T x0 = T::zero(), zs[NZ];
    if (live) {
      const double2 a = st[threadIdx.x];
      x0 = {a.x, a.y};
#pragma unroll
      for (int k = 0; k < NZ; ++k) {
        const double2 b = st[(size_t)(k + 1) * (NKA_TMA_TILE / 2) + threadIdx.x];
        zs[k] = {b.x, b.y};
      }
    }
    nka_mbar_arrive(&bar_empty[s]);                                 // the values are in registers: stage free
    if (!live) continue;
    T yprev = T::zero();                                        // same fma order as nka_pass_b_elem<NZ,2,true>
#pragma unroll
    for (int k = 0; k < NZ; ++k) zs[k].fma_into(coefY[k], yprev);
    const T zpv = yprev + x0;
    zpv.st_stream(zp, i);
    T y = T::zero();
    zpv.fma_into(coef_p, y);
#pragma unroll
    for (int k = 0; k < NZ; ++k) zs[k].fma_into(coefN[k], y);
    x0.st_stream(wnew, i);
    (x0 + y).st(f, i);
  }
}

// host side -----------------------------------------------------------------
template <int NZ> static void nka_tma_prepare()
{
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(nka_pass_b_tma<NZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, NkaTmaCfg<NZ>::kSmemBytes);
    done = true;
  }
}

// Returns the kernel for nz (nullptr: not instantiated) and its launch shape.
PassBFn nka_get_pass_b_tma(int nz, int* threads, int* smem_bytes)
{
  *threads = NKA_TMA_THREADS;
  switch (nz) {
    case 2: nka_tma_prepare<2>(); *smem_bytes = NkaTmaCfg<2>::kSmemBytes; return nka_pass_b_tma<2>;
    case 5: nka_tma_prepare<5>(); *smem_bytes = NkaTmaCfg<5>::kSmemBytes; return nka_pass_b_tma<5>;
    case 10: nka_tma_prepare<10>(); *smem_bytes = NkaTmaCfg<10>::kSmemBytes; return nka_pass_b_tma<10>;
    case 20: nka_tma_prepare<20>(); *smem_bytes = NkaTmaCfg<20>::kSmemBytes; return nka_pass_b_tma<20>;
    default: *smem_bytes = 0; return nullptr;
  }
}
