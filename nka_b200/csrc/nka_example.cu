// nka_example.cu -- the reference example's discrete system and Picard solver on the device
// (declarations and the reference lines being replaced: include/nka_example.h).
//
// Kernels (all sm_100a, fp64, none of them GEMM-shaped):
//   ex_residual_kernel   u <- u - z (optional), face coefficients ax/ay/ac from u, residual r,
//                        and sum r^2 in one stencil sweep (HBM bound; deterministic reduction).
//                        Replaces update_system + residual + `u = u - r` + norm2:
//                        src-F08/nka_example.F90:103-145, :248-250.
//   ex_ssor_sweep<DIR>   one Gauss-Seidel/SOR sweep in the reference's lexicographic order
//                        (DIR=+1 forward, -1 backward), src-F08/nka_example.F90:159-175, as a
//                        pipelined anti-diagonal wavefront: one thread per grid column, one warp
//                        (= one CTA) per strip of 32 columns, neighbouring strips hand over their edge values
//                        through a flag-in-data channel in L2.  Latency bound: the chain
//                        z(j-1,k) -> z(j,k) is serial by definition of the method.
//   ex_permute_kernel    natural order <-> wavefront-major (I/O only).
//
// Every cell is computed with the reference's operand order and without fma contraction
// (__dmul_rn/__dadd_rn/__ddiv_rn), from the same neighbour values the serial loops would use,
// so residual and SSOR are BIT-IDENTICAL to the CPU code (tests/test_gpu_example.py).

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/nka_b200.h"
#include "../../include/nka_example.h"
#include "nka_internal.h"
#include "nka_state.h"      // NKA_MAX_RANKS
#include "nka_res_items.h"

#if defined(__CUDACC__)
#define EX_HD __host__ __device__ __forceinline__
#else
#define EX_HD inline
#endif

// ---------------------------------------------------------------------------
// wavefront-major geometry: diagonal t = j + k holds cells j = jmin(t)..jmax(t)
// ---------------------------------------------------------------------------
EX_HD long long wf_off(int t, int nx, int ny)          // cells on diagonals < t
{
  const long long m = nx < ny ? nx : ny, M = nx < ny ? ny : nx;
  if (t <= m) return (long long)t * (t + 1) / 2;
  if (t <= M) return m * (m + 1) / 2 + ((long long)t - m) * m;
  const long long r = (long long)nx + ny - 1 - t;
  return (long long)nx * ny - r * (r + 1) / 2;
}
EX_HD long long wf_base(int t, int nx, int ny)         // index of cell (j, t-j) is wf_base(t) + j
{
  const int jm = t - (ny - 1);
  return wf_off(t, nx, ny) - (jm > 0 ? jm : 0);
}

// wf_base(t+1) - wf_base(t) for 0 <= t <= nx+ny-3: lets a walk along consecutive diagonals
// update its base with a few integer operations instead of re-evaluating wf_base
EX_HD int wf_step(int t, int nx, int ny)
{
  const int jmax = t < nx - 1 ? t : nx - 1;
  const int over = t - (ny - 1);                    // jmin(t) = max(0, over); jmin(t+1) - jmin(t) = (over >= 0)
  return jmax - (over > 0 ? over : 0) + (over >= 0 ? 0 : 1);
}

// ---------------------------------------------------------------------------
// residual
// ---------------------------------------------------------------------------
#define EX_RES_THREADS 256

struct ResParams {
  int nx, ny;
  const double* U;       // current solution
  const double* Zc;      // update to subtract first (nullptr: none)
  double* Unew;          // receives u - z when Zc != nullptr
  double *R, *AXL, *AYD, *AC, *AXR, *AYT;
  double a, fx, fy, q;
  double* partials;      // one per block
  unsigned* ticket;
  double* sumsq;         // result
  // row slabs (multi-GPU): u (already u - z) of the row below the slab's first row / above its last
  // row, received from the neighbouring ranks; nullptr at a physical boundary (u = 0 there)
  const double *HLO, *HHI;
};

__device__ __forceinline__ double ex_warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(EX_RES_THREADS) ex_residual_kernel(ResParams P)
{
  const int nx = P.nx, ny = P.ny;
  // Work items = (diagonal t, chunk of EX_RES_THREADS cells of it).  A fixed grid of resident CTAs
  // walks the items with a grid stride: one partial sum, one ticket and one block reduction per
  // CTA instead of per item (one CTA per item spent about half of the kernel on 131072 tickets,
  // fences and half-empty blocks at 4096^2).
  const int nchunk = ((nx < ny ? nx : ny) + EX_RES_THREADS - 1) / EX_RES_THREADS;
  const unsigned nitems = (unsigned)(nx + ny - 1) * (unsigned)nchunk;      // < 2^31: checked by nka_system_init_slab
  double rr = 0.0;
  for (unsigned item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int t = (int)(item / (unsigned)nchunk);
    const int jmin = t - (ny - 1) > 0 ? t - (ny - 1) : 0;
    const int jmax = t < nx - 1 ? t : nx - 1;
    const int p = (int)(item - (unsigned)t * (unsigned)nchunk) * EX_RES_THREADS + threadIdx.x;
    if (jmin + p > jmax) continue;
    const int j = jmin + p, k = t - j;
    const long long c = wf_base(t, nx, ny) + j;
    const long long bm = wf_base(t - 1, nx, ny), bp = wf_base(t + 1, nx, ny);
    const bool bot = k == 0, top = k + 1 >= ny;
    const bool hl = j > 0, hr = j + 1 < nx, hd = !bot || P.HLO, hu = !top || P.HHI;
    const double* __restrict__ U = P.U;
    const double* __restrict__ Zc = P.Zc;
    auto val = [&](long long i) {
      const double u = __ldg(U + i);
      return Zc ? __dsub_rn(u, __ldg(Zc + i)) : u;                 // u = u - r : F08 :248
    };
    const double uc = val(c);
    const double ul = hl ? val(bm + j - 1) : 0.0, ud = !bot ? val(bm + j) : (P.HLO ? __ldg(P.HLO + j) : 0.0);
    const double ur = hr ? val(bp + j + 1) : 0.0, uu = !top ? val(bp + j) : (P.HHI ? __ldg(P.HHI + j) : 0.0);
    // update_system (:122-145): t = 1/(a+u); each face sums the t*h^2 of its (one or two) cells
    // in cell order (left/lower cell first), then ax = 2/ax.
    const double tc = __ddiv_rn(1.0, __dadd_rn(P.a, uc));
    const double txc = __dmul_rn(tc, P.fx), tyc = __dmul_rn(tc, P.fy);
    double sxl = txc, sxr = txc, syd = tyc, syu = tyc;
    if (hl) sxl = __dadd_rn(__dmul_rn(__ddiv_rn(1.0, __dadd_rn(P.a, ul)), P.fx), txc);
    if (hr) sxr = __dadd_rn(txc, __dmul_rn(__ddiv_rn(1.0, __dadd_rn(P.a, ur)), P.fx));
    if (hd) syd = __dadd_rn(__dmul_rn(__ddiv_rn(1.0, __dadd_rn(P.a, ud)), P.fy), tyc);
    if (hu) syu = __dadd_rn(tyc, __dmul_rn(__ddiv_rn(1.0, __dadd_rn(P.a, uu)), P.fy));
    const double axl = __ddiv_rn(2.0, sxl), axr = __ddiv_rn(2.0, sxr);
    const double ayd = __ddiv_rn(2.0, syd), ayu = __ddiv_rn(2.0, syu);
    const double ac = __dadd_rn(__dadd_rn(__dadd_rn(axl, axr), ayd), ayu);                   // :142
    // residual (:115-117), left to right
    double r = __dmul_rn(ac, uc);
    r = __dsub_rn(r, __dmul_rn(axl, ul));
    r = __dsub_rn(r, __dmul_rn(axr, ur));
    r = __dsub_rn(r, __dmul_rn(ayd, ud));
    r = __dsub_rn(r, __dmul_rn(ayu, uu));
    r = __dsub_rn(r, P.q);
    P.R[c] = r; P.AXL[c] = axl; P.AYD[c] = ayd; P.AC[c] = ac;
    if (Zc) P.Unew[c] = uc;
    if (!hr) P.AXR[k] = axr;
    if (top) P.AYT[j] = ayu;
    rr = __dadd_rn(rr, __dmul_rn(r, r));
  }
  // deterministic two-stage sum of r^2: warp tree, fixed-order across warps, one partial per
  // block, the last block (ticket) folds all partials in a fixed order
  __shared__ double red[EX_RES_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  rr = ex_warp_sum(rr);
  if (lane == 0) red[warp] = rr;
  __syncthreads();
  const unsigned nblk = gridDim.x;
  const unsigned bid = blockIdx.x;
  if (threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < EX_RES_THREADS / 32; ++w) v += red[w];
    P.partials[bid] = v;
    __threadfence();
    is_last = (atomicAdd(P.ticket, 1u) == nblk - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double v = 0.0;
  for (unsigned b = threadIdx.x; b < nblk; b += EX_RES_THREADS) v += __ldcg(P.partials + b);
  v = ex_warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < EX_RES_THREADS / 32; ++w) s += red[w];
    *P.sumsq = s;
    *P.ticket = 0u;
  }
}

// ---------------------------------------------------------------------------
// residual, second formulation: every 1/(a+u) and every face coefficient formed ONCE
// ---------------------------------------------------------------------------
// ex_residual_kernel recomputes, in every cell, 1/(a+u) of the cell and its four neighbours and
// the 2/s of its four faces: nine IEEE divisions per cell, 403 instructions per cell, instruction
// bound at 27 % of the HBM rate (profiles/r1w_residual_ncu.txt).  But a face belongs to two cells
// and both form it from the same operands in the same order (left/lower cell's term first,
// src-F08/nka_example.F90:122-141), so computing it once and handing it to the neighbour gives
// the same bits: three divisions per cell (1/(a+u), the left face, the lower face).
//
// One WARP walks a strip of 30 grid columns (lanes 1..30; lanes 0 and 31 redo the neighbouring
// strips' edge columns, so no shared memory and no barrier is needed) down a band of consecutive
// anti-diagonals: lane l holds column j = 30 s - 1 + l, at step tau it sits on cell (j, tau - j),
// so the 32 lanes read and write 32 consecutive doubles of the wavefront-major arrays.  A lane
// keeps u, tx = t*fx, ty = t*fy and its two faces of the previous two diagonals in registers; the
// left neighbour's values come by shuffle.  At step tau the lane
//   ingests u(tau) (prefetched EX_RS_PF steps ahead), forms t, tx, ty                    1 division
//   forms its left face from lane-1's tx(tau-1) and its lower face from its own ty(tau-1)  2 divisions
//   finishes cell (j, tau-1-j): right face = lane+1's new left face, upper face = its own new
//   lower face; ac, r, r^2 exactly as the reference orders them (:115-117, :142).
// Rows -1 / ny (the neighbouring slabs' edge rows, or nothing at a physical boundary) and columns
// -1 / nx are walked like cells that are "absent": a face next to an absent cell sums one term.
// Work items = (band, strip), handed out in order through an atomic counter; each item leaves its
// own partial sum of r^2, folded in item order by the last CTA: run-to-run bit-stable.
#define EX_RS_COLS 30
#ifndef EX_RS_THREADS
#define EX_RS_THREADS 256
#endif
#ifndef EX_RS_PF
#define EX_RS_PF 2               // diagonals per load batch
#endif
#ifndef EX_RS_DEPTH
#define EX_RS_DEPTH 2            // load batches in flight ahead of the one being used (1: only the next batch)
#endif
#ifndef EX_RS_L2D
#define EX_RS_L2D 0              // diagonals by which prefetch.global.L2 runs ahead of the register loads; 0 = off: tried
                                 // with 4 / 8 / 16 and 3-6 % SLOWER (profiles/r2af_residual_l2_prefetch_ab.jsonl)
#endif
#ifndef EX_RS_MINB
#define EX_RS_MINB 3             // CTAs per SM the register budget is held to: 78 registers, NO spills.  With 4 (64 registers)
                                 // the spilled prefetch values made the warp wait for its loads at once: 57 % of all stall
                                 // samples sat on two spill instructions (profiles/r2ae_*); 3 is 6-8 % faster (r2ag)
#endif

template <bool V> struct ExTag { static constexpr bool value = V; };

struct ResStripParams {
  ResParams p;
  int nstrips, nbands, band;     // strips of 30 columns; bands of `band` diagonals per strip
  int absolute;                  // 1: a band is the same range of grid diagonals in every strip (neighbouring strips, handed out
                                 //    one after the other, then write the sectors they share at about the same time); 0: each
                                 //    strip's own diagonals cut into bands (equal work per item)
  unsigned nitems;               // absolute: the (band, strip) pairs that hold cells, numbered band by band
  const unsigned* first_item;    // absolute: [nbands + 1] number of the first item of each band
  const int* first_strip;        // absolute: [nbands] first strip with cells in each band
  unsigned* counter;             // next item
};

__global__ void __launch_bounds__(EX_RS_THREADS, EX_RS_MINB) ex_residual_strip_kernel(ResStripParams Q)
{
  const ResParams& P = Q.p;
  const int nx = P.nx, ny = P.ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nitems = Q.absolute ? Q.nitems : (unsigned)Q.nstrips * (unsigned)Q.nbands;
  const double* __restrict__ U = P.U;
  const double* __restrict__ Zc = P.Zc;
  const double* __restrict__ HLO = P.HLO;
  const double* __restrict__ HHI = P.HHI;
  const int tmax = nx + ny - 2;                       // last diagonal with cells
  const int kmin = HLO ? -1 : 0, kmax = HHI ? ny : ny - 1;          // rows that exist, edge rows of the neighbouring slabs included

  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(Q.counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    int b, s;
    if (Q.absolute) {
      // band of this item: the last b with first_item[b] <= item (binary search, warp-uniform)
      int lo = 0, hi = Q.nbands - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(Q.first_item + mid) <= item) lo = mid; else hi = mid - 1;
      }
      b = lo;
      s = __ldg(Q.first_strip + b) + (int)(item - __ldg(Q.first_item + b));
    } else {
      b = (int)(item / (unsigned)Q.nstrips);
      s = (int)(item - (unsigned)b * (unsigned)Q.nstrips);
    }
    const int j = s * EX_RS_COLS - 1 + lane;
    const bool cin = (unsigned)j < (unsigned)nx, cinl = (unsigned)(j - 1) < (unsigned)nx;
    const bool mine = lane >= 1 && lane <= EX_RS_COLS && cin;
    int t0 = Q.absolute ? b * Q.band : s * EX_RS_COLS + b * Q.band;   // first diagonal this item finishes
    int t1 = t0 + Q.band;                                             // one past the last
    const int tend = s * EX_RS_COLS + EX_RS_COLS - 1 + ny;            // one past the strip's last diagonal with cells
    if (t0 < s * EX_RS_COLS) t0 = s * EX_RS_COLS;                     // (absolute bands: the strip starts later / ends earlier)
    if (t1 > tend) t1 = tend;
    if (t1 > tmax + 1) t1 = tmax + 1;
    if (t0 >= t1) {                                                   // no cell of this strip in the band
      if (lane == 0) P.partials[item] = 0.0;
      continue;
    }

    // The walk, compiled twice.  IN = every cell the 32 lanes touch in this item (halo lanes and the two prologue /
    // one epilogue diagonals included) lies inside the grid with all four neighbours present -- no boundary row or
    // column, no neighbouring slab's edge row: all the predicates below are then true and drop out (the bulk of a
    // large grid's items; ~100 of the general walk's ~280 instructions per step are this bookkeeping).
    auto walk = [&](auto interior) -> double {
    constexpr bool IN = decltype(interior)::value;
    // what a lane sees at (j, k): in-grid value, a neighbouring slab's edge row, or nothing (zeros)
    auto fetch = [&](int tau, long long base, double& ur, double& zr) {
      if (IN) {                                                       // every lane's cell is inside the grid
        const bool live = tau <= t1;
        ur = live ? __ldg(U + (base + j)) : 0.0;
        zr = (live && Zc) ? __ldg(Zc + (base + j)) : 0.0;
        return;
      }
      const int k = tau - j;
      const bool live = cin && tau <= t1;                             // nothing beyond the band's last step
      const bool ing = live && (unsigned)k < (unsigned)ny;
      const double* pu = U + (base + j);
      bool pr = ing;
      if (live && k == -1 && HLO) { pu = HLO + j; pr = true; }
      if (live && k == ny && HHI) { pu = HHI + j; pr = true; }
      ur = pr ? __ldg(pu) : 0.0;
      zr = (ing && Zc) ? __ldg(Zc + (base + j)) : 0.0;
    };
    auto next_base = [&](int tau, long long base) -> long long {      // wf_base(tau + 1) from wf_base(tau)
      return (IN || (tau >= 0 && tau <= tmax - 1)) ? base + wf_step(tau, nx, ny) : 0;
    };

    // Operands are loaded a BATCH of EX_RS_PF diagonals at a time, one batch ahead of use, and handed
    // over between batches.  (Refilling one register slot per step, EX_RS_PF steps ahead, looks
    // equivalent but is not: the in-order warp has six load scoreboards, so waiting for the oldest
    // load also waits for the youngest that shares its scoreboard -- 57 % of the stall samples of
    // that version, profiles/r2e_residual_strip_ncu.txt.  A cp.async ring instead moved the stalls
    // to the memory-instruction queue, next to the step's eight shuffles: profiles/r2h_*.)
    double cu[EX_RS_PF], cz[EX_RS_PF], nu[EX_RS_PF], nz[EX_RS_PF];
    double mu[EX_RS_DEPTH > 1 ? EX_RS_PF : 1], mz[EX_RS_DEPTH > 1 ? EX_RS_PF : 1];   // the batch after next (EX_RS_DEPTH = 2)
    int tf = t0 - 1;
    long long basef = (tf >= 0 && tf <= tmax) ? wf_base(tf, nx, ny) : 0;
#pragma unroll
    for (int i = 0; i < EX_RS_PF; ++i) {
      fetch(tf, basef, cu[i], cz[i]);
      basef = next_base(tf, basef);
      ++tf;
    }
    if (EX_RS_DEPTH > 1) {
#pragma unroll
      for (int i = 0; i < EX_RS_PF; ++i) {
        fetch(tf, basef, nu[i], nz[i]);
        basef = next_base(tf, basef);
        ++tf;
      }
    }
    // Experiment (EX_RS_L2D > 0, off in the product build): the register loads run one batch ahead, which is less
    // than the DRAM latency under load (ncu: 67 % of the stall samples on the load scoreboard); an L2 prefetch
    // EX_RS_L2D diagonals further ahead was meant to turn the DRAM round trip into an L2 hit without holding
    // registers.  Measured 3-6 % slower at every distance: the ~14 extra instructions per step cost more than the
    // shorter wait gives back.
    int tl2 = tf + EX_RS_L2D;
    long long basel2 = (EX_RS_L2D > 0 && tl2 >= 0 && tl2 <= tmax) ? wf_base(tl2, nx, ny) : 0;
    auto l2_ahead = [&]() {
      if (EX_RS_L2D > 0) {
        const int k = tl2 - j;
        if (tl2 <= t1 && cin && (unsigned)k < (unsigned)ny) {
          asm volatile("prefetch.global.L2 [%0];" :: "l"(U + (basel2 + j)));
          if (Zc) asm volatile("prefetch.global.L2 [%0];" :: "l"(Zc + (basel2 + j)));
        }
        basel2 = next_base(tl2, basel2);
        ++tl2;
      }
    };

    double u_pp = 0.0, u_p = 0.0, tx_p = 0.0, ty_p = 0.0, fx_p = 1.0, fy_p = 1.0;
    bool pc_p = false;                                                 // was (j, k-1) there: the previous step's own cell
    long long base_c = (t0 - 1 >= 0 && t0 - 1 <= tmax) ? wf_base(t0 - 1, nx, ny) : 0, base_p = 0;
    double rr = 0.0;
    for (int tau0 = t0 - 1; tau0 <= t1; tau0 += EX_RS_PF) {
#pragma unroll
      for (int i = 0; i < EX_RS_PF; ++i) {                             // a batch ahead, all loads back to back
        if (EX_RS_DEPTH > 1) fetch(tf, basef, mu[i], mz[i]);
        else fetch(tf, basef, nu[i], nz[i]);
        basef = next_base(tf, basef);
        ++tf;
      }
#pragma unroll
      for (int i = 0; i < EX_RS_PF; ++i) l2_ahead();
#pragma unroll
      for (int i = 0; i < EX_RS_PF; ++i) {
        const int tau = tau0 + i;
        if (tau <= t1) {                                               // warp-uniform
          const int k = tau - j;
          const bool krange = IN || (k >= kmin && k <= kmax);
          const bool pc = IN || (cin && krange), pl = IN || (cinl && krange), pd = IN || pc_p;
          const double u_n = Zc ? __dsub_rn(cu[i], cz[i]) : cu[i];    // u = u - r : F08 :248 (z = 0 outside the grid)
          // update_system (:122-145)
          const double tc = __ddiv_rn(1.0, __dadd_rn(P.a, u_n));
          const double tx_n = __dmul_rn(tc, P.fx), ty_n = __dmul_rn(tc, P.fy);
          const double tx_l = __shfl_up_sync(0xffffffffu, tx_p, 1);    // (j-1, k) sits on diagonal tau-1
          const double sx = (pl && pc) ? __dadd_rn(tx_l, tx_n) : (pl ? tx_l : tx_n);
          const double sy = (pd && pc) ? __dadd_rn(ty_p, ty_n) : (pd ? ty_p : ty_n);
          const double fx_n = __ddiv_rn(2.0, sx);                      // left face of (j, k)
          const double fy_n = __ddiv_rn(2.0, sy);                      // lower face of (j, k)
          // finish cell (j, k-1) on diagonal tau-1
          const double u_l = __shfl_up_sync(0xffffffffu, u_pp, 1);     // (j-1, k-1): diagonal tau-2
          const double u_r = __shfl_down_sync(0xffffffffu, u_n, 1);    // (j+1, k-1): diagonal tau
          const double axr = __shfl_down_sync(0xffffffffu, fx_n, 1);   // left face of (j+1, k-1)
          const int kc = k - 1;
          if (mine && tau - 1 >= t0 && (IN || (unsigned)kc < (unsigned)ny)) {
            const double axl = fx_p, ayd = fy_p, ayu = fy_n;
            const double ac = __dadd_rn(__dadd_rn(__dadd_rn(axl, axr), ayd), ayu);        // :142
            double r = __dmul_rn(ac, u_p);                                                  // :115-117, left to right
            r = __dsub_rn(r, __dmul_rn(axl, u_l));
            r = __dsub_rn(r, __dmul_rn(axr, u_r));
            r = __dsub_rn(r, __dmul_rn(ayd, u_pp));
            r = __dsub_rn(r, __dmul_rn(ayu, u_n));
            r = __dsub_rn(r, P.q);
            const long long c = base_p + j;
            P.R[c] = r; P.AXL[c] = axl; P.AYD[c] = ayd; P.AC[c] = ac;
            if (Zc) P.Unew[c] = u_p;
            if (!IN && j == nx - 1) P.AXR[kc] = axr;
            if (!IN && kc == ny - 1) P.AYT[j] = ayu;
            rr = __dadd_rn(rr, __dmul_rn(r, r));
          }
          u_pp = u_p; u_p = u_n;
          tx_p = tx_n; ty_p = ty_n; fx_p = fx_n; fy_p = fy_n; pc_p = pc;
          base_p = base_c;
          base_c = next_base(tau, base_c);
        }
      }
#pragma unroll
      for (int i = 0; i < EX_RS_PF; ++i) {
        cu[i] = nu[i]; cz[i] = nz[i];
        if (EX_RS_DEPTH > 1) { nu[i] = mu[i]; nz[i] = mz[i]; }
      }
    }
    return rr;
    };
    const int js = s * EX_RS_COLS;
    const bool interior_item = s >= 1 && js + EX_RS_COLS <= nx - 1 && (t0 - 1) - (js + EX_RS_COLS) >= 1 && t1 - (js - 1) <= ny - 1;
    double rr = interior_item ? walk(ExTag<true>()) : walk(ExTag<false>());
    rr = ex_warp_sum(rr);
    if (lane == 0) P.partials[item] = rr;
  }

  // the last CTA folds the items' partial sums in item order
  __shared__ double red[EX_RS_THREADS / 32];
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double v = 0.0;
  for (unsigned i = threadIdx.x; i < nitems; i += EX_RS_THREADS) v += __ldcg(P.partials + i);
  v = ex_warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < EX_RS_THREADS / 32; ++w) sum += red[w];
    *P.sumsq = sum;
    *P.ticket = 0u;
    *Q.counter = 0u;
  }
}

// Row slabs: the slab's first and last rows of u (after u <- u - z, formed exactly as the residual
// kernel will form it) packed for the neighbouring ranks.
__global__ void ex_pack_rows_kernel(const double* __restrict__ U, const double* __restrict__ Zc, int nx, int ny,
                                    double* __restrict__ lo, double* __restrict__ hi)
{
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nx; j += gridDim.x * blockDim.x) {
    const long long c0 = wf_base(j, nx, ny) + j, c1 = wf_base(j + ny - 1, nx, ny) + j;
    lo[j] = Zc ? __dsub_rn(U[c0], Zc[c0]) : U[c0];
    hi[j] = Zc ? __dsub_rn(U[c1], Zc[c1]) : U[c1];
  }
}

// ---------------------------------------------------------------------------
// natural order <-> wavefront-major
// ---------------------------------------------------------------------------
__global__ void ex_permute_kernel(double* __restrict__ dst, const double* __restrict__ src, int nx, int ny, int to_wavefront)
{
  const size_t n = (size_t)nx * ny;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % nx), k = (int)(i / nx);
    const long long w = wf_base(j + k, nx, ny) + j;
    if (to_wavefront) dst[w] = src[i];
    else dst[i] = src[w];
  }
}

__global__ void ex_fill_u64(unsigned long long* p, size_t n, unsigned long long v)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------
// SSOR sweep
// ---------------------------------------------------------------------------
// One CTA = one warp = one strip of 32 grid columns; lane l walks column j0 + l down (forward)
// or up (backward) one row per step, so at any step the 32 lanes sit on ONE anti-diagonal t
// (warp-uniform: all index arithmetic on t is done once per warp, from blockIdx-derived values)
// and read/write 32 consecutive doubles of the wavefront-major arrays.
//
// The sweep is a chain of dependent cells, so what matters is the latency of one step, i.e. the
// number of instructions between receiving the upstream value and producing this cell's:
//   * operands are copied EX_SSOR_DIST steps ahead of use into a shared-memory ring (cp.async,
//     retired in order with wait_group: see ssor_issue);
//   * everything that does not depend on new values is formed one step ahead (ssor_cook;
//     products are rounded separately from the sums, so forming them early is bit-identical):
//     forward axr*z_old(j+1,k), ayu*z_old(j,k+1), (1-w)*z_old; backward r + axl*z_old(j-1,k)
//     (the first sum has no new operand), ayd*z_old(j,k-1), (1-w)*z_old;
//   * per step what is left is the reference's chain: 4 adds, 2-3 multiplies, one division.
#ifndef EX_SSOR_UNROLL
#define EX_SSOR_UNROLL 4
#endif
// "not written yet" mark of the edge channel: a signalling-NaN bit pattern arithmetic never produces
#define EX_SENT 0x7FF4DEADBEEF1234ULL
#define EX_SPIN_LIMIT (4000000000LL)   // cycles (~2 s): a stuck neighbour becomes an error, not a hang

struct SsorParams {
  int nx, ny, nstrips, zero_old;
  const double *R, *AC, *AXL, *AYD, *AXR, *AYT;
  double* Z;
  unsigned long long* bnd;       // [nstrips][ny]: edge column of strip s, all EX_SENT between sweeps
  double omega, om1;
  int* err;
  unsigned long long* trace;     // debugging aid (nka_system_ssor_trace): [nstrips][4] globaltimer stamps, or nullptr
  // Row slabs (multi-GPU): the sweep continues from the rank below (forward) / above (backward).
  // Edge rows travel as tagged 16-byte slots {lo32, tag, hi32, tag} written straight into the
  // neighbour's memory over NVLink (tag = sweep number: no flags to reset, no fences).
  const uint4* halo_lo;          // [nx] local: z of the row below this slab (written by the rank below), or nullptr
  const uint4* halo_hi;          // [nx] local: z of the row above this slab (written by the rank above), or nullptr
  uint4* peer_up_lo;             // the upper rank's halo_lo (forward: this slab's top row goes there), or nullptr
  uint4* peer_dn_hi;             // the lower rank's halo_hi (backward: this slab's bottom row goes there), or nullptr
  unsigned tag_cur, tag_prev;    // this sweep's number; the previous sweep's (whose edge row holds the "old" values)
  long long spin_limit;          // cycles a wait may take before it becomes an error (longer across GPUs)
};

__device__ __forceinline__ double tagged_wait(const uint4* slot, unsigned tag, int* err, long long limit)
{
  unsigned lo, t0, hi, t1;
  const long long c0 = clock64();
  for (unsigned it = 0;; ++it) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(slot) : "memory");
    if (t0 == tag && t1 == tag) break;
    if ((it & 255u) == 255u) {
      if (*(volatile int*)err) return 0.0;
      if (clock64() - c0 > limit) { *(volatile int*)err = 2; return 0.0; }
    }
  }
  return __hiloint2double((int)hi, (int)lo);
}

__device__ __forceinline__ void tagged_send(uint4* slot, double v, unsigned tag)
{
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
               :: "l"(slot), "r"((unsigned)__double2loint(v)), "r"(tag), "r"((unsigned)__double2hiint(v)), "r"(tag)
               : "memory");
}

// Raw operands of cell (j, tt - j): own r, ac, left/lower face, own old z; the right face (the left
// face of cell (j+1,k), or the boundary face); the old z of the downstream horizontal neighbour.
struct SsorEnt { double r, ac, axl, ayd, zo, zs, axr; };

// What a step needs that does not depend on new values, formed one step ahead of use from raw
// operands that were copied EX_SSOR_DIST steps ahead (so neither the loads nor these products sit on
// the dependent chain).  forward: a = r, b = axl, p0 = axr*z_old(j+1,k), p1 = ayu*z_old(j,k+1);
// backward: a = r + axl*z_old(j-1,k), b = axr, p0 = ayd*z_old(j,k-1), p1 unused.
struct SsorCooked { double a, b, p0, p1, po, ayd, ac; };

template <int DIR>
__device__ __forceinline__ SsorCooked ssor_cook(const SsorEnt& cur, const SsorEnt& nxt, bool edge, double ayt, double om1,
                                                double zold_edge)
{
  // edge: the cell's next row in travel direction lies outside this slab (forward: above the top
  // row; backward: below the bottom row).  Its old z is zold_edge: 0 at a physical boundary, the
  // neighbouring rank's edge row otherwise.
  SsorCooked c;
  c.po = __dmul_rn(om1, cur.zo);
  c.ayd = cur.ayd;
  c.ac = cur.ac;
  if (DIR > 0) {
    c.a = cur.r;
    c.b = cur.axl;
    c.p0 = __dmul_rn(cur.axr, cur.zs);
    c.p1 = __dmul_rn(edge ? ayt : nxt.ayd, edge ? zold_edge : nxt.zo);   // nxt = (j, k+1); beyond the top: ayt * edge z
  } else {
    c.a = __dadd_rn(cur.r, __dmul_rn(cur.axl, cur.zs));
    c.b = cur.axr;
    c.p0 = __dmul_rn(cur.ayd, edge ? zold_edge : nxt.zo);    // nxt = (j, k-1); below the bottom: ayd * edge z
    c.p1 = 0.0;
  }
  return c;
}

// Operand staging: cp.async (LDGSTS, 8 bytes per lane per array) into a shared-memory ring,
// EX_SSOR_DIST steps ahead of use, retired in order with cp.async.wait_group.  Register-ring
// prefetching with plain loads does not work here: the warp issues in order and has six load
// scoreboards for ~60 loads in flight, so a wait for an old load also waits for young ones that
// share its scoreboard (ncu: 36 % of the issue slots stalled on long_scoreboard, profiles/r1h_*).
#ifndef EX_SSOR_DIST
#define EX_SSOR_DIST 15
#endif
#define EX_SSOR_STAGES (EX_SSOR_DIST + 1)
enum { SS_R = 0, SS_AC, SS_AXL, SS_AYD, SS_AXR, SS_ZO, SS_ZS, SS_NF };

__device__ __forceinline__ unsigned long long ex_globaltimer()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// The edge channel lives in L2 and is only ever shared between SMs of this GPU: gpu-scope relaxed
// accesses (volatile would be system scope).
__device__ __forceinline__ unsigned long long ch_ld(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ch_st(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void ss_cp8(uint32_t dst, const double* src, bool on)
{
  const int nbytes = on ? 8 : 0;                  // 0: nothing is read, the 8 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(nbytes) : "memory");
}

// Issue the copies for cell (j, tt - j) into ring stage `stage`.  bc, bn, bp: wf_base of diagonals
// tt, tt+1, tt-1 (warp-uniform).  Cells outside the grid get zeros (z = 0 there).
template <int DIR>
__device__ __forceinline__ void ssor_issue(const SsorParams& P, uint32_t ring, int stage, int tt, long long bc,
                                           long long bn, long long bp, int j, bool jvalid, int lane)
{
  const int k = tt - j;
  const bool on = jvalid && k >= 0 && k < P.ny;
  const long long c = on ? bc + j : 0;
  const uint32_t dst = ring + (uint32_t)((stage * SS_NF * 32 + lane) * 8);
  ss_cp8(dst + SS_R * 256, P.R + c, on);
  ss_cp8(dst + SS_AC * 256, P.AC + c, on);
  ss_cp8(dst + SS_AXL * 256, P.AXL + c, on);
  ss_cp8(dst + SS_AYD * 256, P.AYD + c, on);
  const bool inner = j + 1 < P.nx;
  ss_cp8(dst + SS_AXR * 256, (on && inner) ? P.AXL + bn + j + 1 : P.AXR + (on ? k : 0), on);
  const bool zon = on && !P.zero_old;
  ss_cp8(dst + SS_ZO * 256, P.Z + c, zon);
  if (DIR > 0) ss_cp8(dst + SS_ZS * 256, P.Z + (zon && inner ? bn + j + 1 : 0), zon && inner);   // old z(j+1,k)
  else ss_cp8(dst + SS_ZS * 256, P.Z + (zon && j > 0 ? bp + j - 1 : 0), zon && j > 0);           // old z(j-1,k)
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ SsorEnt ssor_fetch(const double* ring, int stage, int lane)
{
  const double* p = ring + stage * SS_NF * 32 + lane;
  SsorEnt e;
  e.r = p[SS_R * 32]; e.ac = p[SS_AC * 32]; e.axl = p[SS_AXL * 32]; e.ayd = p[SS_AYD * 32];
  e.axr = p[SS_AXR * 32]; e.zo = p[SS_ZO * 32]; e.zs = p[SS_ZS * 32];
  return e;
}

// The edge values of the upstream strip reach the compute warp through a second, helper warp:
// polling L2 from the compute warp itself puts an L2 round trip (or, with polls issued ahead,
// the in-order warp's load-scoreboard aliasing: two of every four steps stalled ~450 ns,
// profiles/r1h_ssor_trace.txt) on the dependent chain.  The helper polls 32 rows at a time,
// re-arms each channel word for the next sweep, and drops the values into a 64-entry mailbox in
// shared memory; the compute warp's edge lane picks them up with a ~30-cycle shared-memory load.
#define EX_MBOX 64
__device__ __forceinline__ unsigned long long mbox_ld(const unsigned long long* p)
{
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void mbox_st(unsigned long long* p, unsigned long long v)
{
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

template <int DIR>
__device__ __forceinline__ void ssor_receiver(const SsorParams& P, unsigned long long* mbox, unsigned long long* cin,
                                              const int lane)
{
  const int ny = P.ny;
  for (int rb = 0; rb < ny; rb += 32) {
    const int ri = rb + lane;                      // ri-th row in travel order
    const int k = DIR > 0 ? ri : ny - 1 - ri;
    bool done = ri >= ny;
    unsigned long long v = EX_SENT;
    const long long t0 = clock64();
    unsigned it = 0;
    while (!__all_sync(0xffffffffu, done)) {
      if (!done) {
        if (v == EX_SENT) {
          v = ch_ld(cin + k);
          if (v != EX_SENT) ch_st(cin + k, EX_SENT);                     // ready for the next sweep
        }
        if (v != EX_SENT && mbox_ld(mbox + (ri & (EX_MBOX - 1))) == EX_SENT) {
          mbox_st(mbox + (ri & (EX_MBOX - 1)), v);
          done = true;
        }
      }
      if ((++it & 255u) == 0u) {
        if (*(volatile int*)P.err) return;
        if (clock64() - t0 > P.spin_limit) { *(volatile int*)P.err = 1; return; }
      }
    }
  }
}

__device__ __noinline__ unsigned long long ssor_mbox_wait(unsigned long long* slot, int* err)
{
  unsigned it = 0;
  for (;;) {
    const unsigned long long v = mbox_ld(slot);
    if (v != EX_SENT) return v;
    if ((++it & 1023u) == 0u && *(volatile int*)err) return 0ull;        // the receiver gave up: so do we
  }
}

template <int DIR>
__device__ __forceinline__ void ssor_strip(const SsorParams& P, double* ring_ptr, unsigned long long* mbox,
                                           const int strip, const int lane)
{
  constexpr int DIST = EX_SSOR_DIST, STAGES = EX_SSOR_STAGES, UNROLL = EX_SSOR_UNROLL;
  static_assert(DIST >= 3, "ring geometry");
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ring_ptr);
  const int nx = P.nx, ny = P.ny;
  const int j0 = strip * 32;
  const int j = j0 + lane;
  const bool jvalid = j < nx;
  const int jlast = j0 + 31 < nx - 1 ? j0 + 31 : nx - 1;
  const int nsteps = (jlast - j0) + ny;
  const int t_first = DIR > 0 ? j0 : jlast + ny - 1;
  const double ayt = jvalid ? __ldg(P.AYT + j) : 0.0;
  const double omega = P.omega;
  // forward: lane 31 hands z(j,k) to lane 0 of the next strip; backward: lane 0 to lane 31 of the previous one
  const bool is_prod = jvalid && (DIR > 0 ? (lane == 31 && j + 1 < nx) : (lane == 0 && strip > 0));
  const bool is_cons = jvalid && (DIR > 0 ? (lane == 0 && strip > 0) : (lane == 31 && j + 1 < nx));
  unsigned long long* cout = P.bnd + (size_t)strip * ny;

  // rolling window of diagonal bases around the prefetch diagonal tp: behind (tp - DIR), at, ahead (tp + DIR)
  int tp = t_first;
  long long b_behind = wf_base(tp - DIR, nx, ny), b_at = wf_base(tp, nx, ny), b_ahead = wf_base(tp + DIR, nx, ny);
  const long long b_first = b_at;
  int pstage = 0;                                   // ring stage of step number (tp - t_first) * DIR
  auto issue_next = [&]() {
    ssor_issue<DIR>(P, ring, pstage, tp, b_at, DIR > 0 ? b_ahead : b_behind, DIR > 0 ? b_behind : b_ahead, j, jvalid, lane);
    pstage = pstage + 1 == STAGES ? 0 : pstage + 1;
    tp += DIR;
    b_behind = b_at; b_at = b_ahead;
    b_ahead += DIR > 0 ? wf_step(tp, nx, ny) : -wf_step(tp - 1, nx, ny);      // now wf_base(tp + DIR)
  };
  __syncwarp();                                     // the previous strip's reads of the ring are done
  if (P.trace && lane == 0) P.trace[strip * 4 + 0] = ex_globaltimer();
  for (int i = 0; i < DIST; ++i) issue_next();      // steps 0 .. DIST-1 in flight

  // row slabs: the edge rows of the neighbouring ranks (new: where this sweep comes from; old: where it goes)
  double znew0 = 0.0, zold_edge = 0.0;
  if (jvalid) {
    const uint4* from = DIR > 0 ? P.halo_lo : P.halo_hi;
    const uint4* beyond = DIR > 0 ? P.halo_hi : P.halo_lo;
    if (from) znew0 = tagged_wait(from + j, P.tag_cur, P.err, P.spin_limit);
    if (beyond && !P.zero_old) zold_edge = tagged_wait(beyond + j, P.tag_prev, P.err, P.spin_limit);
  }
  uint4* const send_to = DIR > 0 ? P.peer_up_lo : P.peer_dn_hi;

  asm volatile("cp.async.wait_group %0;" :: "n"(DIST - 2) : "memory");          // steps 0 and 1 have landed
  SsorEnt e1 = ssor_fetch(ring_ptr, 1, lane);
  SsorCooked ck;
  {
    const SsorEnt e0 = ssor_fetch(ring_ptr, 0, lane);
    const int kf = t_first - j;
    ck = ssor_cook<DIR>(e0, e1, DIR > 0 ? kf + 1 >= ny : kf <= 0, ayt, P.om1, zold_edge);
  }
  // own result of the previous step = z(j, k-DIR), new; before the first row: the row the
  // neighbouring rank has just computed, or the boundary value 0
  double znew = znew0;
  double ayd_prev = 0.0;    // backward: lower face of the previous step's cell = upper face of this one
  long long b_cur = b_first;   // wf_base of the current step's diagonal (for the store)
  int fstage = 2;           // ring stage of step s + 2
  for (int sb = 0; sb < nsteps; sb += UNROLL) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int t = t_first + DIR * (sb + u);
      const int k = t - j;
      const bool act = jvalid && k >= 0 && k < ny;
      // upstream horizontal neighbour, new value: the adjacent lane's previous step
      double zh = DIR > 0 ? __shfl_up_sync(0xffffffffu, znew, 1) : __shfl_down_sync(0xffffffffu, znew, 1);
      if (DIR > 0 ? lane == 0 : lane == 31) zh = 0.0;                    // grid edge: boundary value 0
      if (is_cons && act) {
        // the edge lane's rows are its steps 0 .. ny-1, in travel order
        unsigned long long* slot = mbox + ((sb + u) & (EX_MBOX - 1));
        unsigned long long v = mbox_ld(slot);
        if (v == EX_SENT) { v = ssor_mbox_wait(slot, P.err); if (P.trace) P.trace[strip * 4 + 3] += 1; }
        mbox_st(slot, EX_SENT);                                          // free for the receiver
        zh = __longlong_as_double((long long)v);
        if (P.trace && sb + u == 0) P.trace[strip * 4 + 1] = ex_globaltimer();
      }
      // src-F08/nka_example.F90:163-165 (= :171-173): the reference's operation order, no fma
      //   z = (1-w) z + w (r + axl z(j-1,k) + axr z(j+1,k) + ayd z(j,k-1) + ayu z(j,k+1)) / ac
      double sm;
      if (DIR > 0) {
        sm = __dadd_rn(ck.a, __dmul_rn(ck.b, zh));                       // r + axl * new left
        sm = __dadd_rn(sm, ck.p0);                                       //   + axr * old right
        sm = __dadd_rn(sm, __dmul_rn(ck.ayd, znew));                     //   + ayd * new lower
        sm = __dadd_rn(sm, ck.p1);                                       //   + ayu * old upper
      } else {
        const double ayu = (k + 1 < ny) ? ayd_prev : ayt;
        sm = __dadd_rn(ck.a, __dmul_rn(ck.b, zh));                       // (r + axl * old left) + axr * new right
        sm = __dadd_rn(sm, ck.p0);                                       //   + ayd * old lower
        sm = __dadd_rn(sm, __dmul_rn(ayu, znew));                        //   + ayu * new upper
      }
      // (cells outside the grid divide 1 by 1: a zero numerator would take the division's slow path
      // for the whole warp during the 31 fill / drain steps of every strip)
      const double zc = __dadd_rn(ck.po, __ddiv_rn(act ? __dmul_rn(omega, sm) : 1.0, act ? ck.ac : 1.0));
      if (act) {
        znew = zc;
        ayd_prev = ck.ayd;
        P.Z[b_cur + j] = zc;
        if (is_prod) ch_st(cout + k, (unsigned long long)__double_as_longlong(zc));
        if (send_to && k == (DIR > 0 ? ny - 1 : 0)) tagged_send(send_to + j, zc, P.tag_cur);   // hand over to the next rank
      }
      if (P.trace && strip == P.nstrips / 2 && lane == 0 && sb + u < 256) P.trace[P.nstrips * 4 + sb + u] = ex_globaltimer();
      b_cur += DIR > 0 ? wf_step(t, nx, ny) : -wf_step(t - 1, nx, ny);
      // off the chain: issue the operands DIST steps ahead, retire the copies of step s + 2,
      // and form the next step's independent terms
      issue_next();
      asm volatile("cp.async.wait_group %0;" :: "n"(DIST - 2) : "memory");
      const SsorEnt e2 = ssor_fetch(ring_ptr, fstage, lane);
      fstage = fstage + 1 == STAGES ? 0 : fstage + 1;
      ck = ssor_cook<DIR>(e1, e2, DIR > 0 ? (k + DIR) + 1 >= ny : (k + DIR) <= 0, ayt, P.om1, zold_edge);
      e1 = e2;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (P.trace && lane == 0) P.trace[strip * 4 + 2] = ex_globaltimer();
}

template <int DIR>
__global__ void __launch_bounds__(64) ex_ssor_sweep(SsorParams P)
{
  __shared__ __align__(16) double ring[EX_SSOR_STAGES * SS_NF * 32];
  __shared__ unsigned long long mbox[EX_MBOX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // strips in dependency order: a strip only waits for one that an earlier CTA (or an earlier
  // round of this grid) owns, and every CTA of the grid is resident, so no wait can deadlock
  for (int i = blockIdx.x; i < P.nstrips; i += gridDim.x) {
    const int strip = DIR > 0 ? i : P.nstrips - 1 - i;
    const bool has_upstream = i > 0 && (DIR > 0 || (strip + 1) * 32 < P.nx + 0);
    if (threadIdx.x < EX_MBOX) mbox[threadIdx.x] = EX_SENT;
    __syncthreads();
    if (warp == 0) {
      ssor_strip<DIR>(P, ring, mbox, strip, lane);
    } else if (has_upstream) {
      // upstream strip: forward strip-1 (its lane 31 writes bnd[strip-1]); backward strip+1 (its lane 0 writes bnd[strip+1])
      ssor_receiver<DIR>(P, mbox, P.bnd + (size_t)(DIR > 0 ? strip - 1 : strip + 1) * P.ny, lane);
    }
    __syncthreads();
  }
}

#include "nka_ssor2.cuh"
#include "nka_ssor3.cuh"

// ex_ssor_sweep2 is compiled per direction x {trace off, on} x {one GPU, row slabs}
template <typename F>
static void ex2_for_each_variant(F f)
{
  f(ex_ssor_sweep2<1, false, false>); f(ex_ssor_sweep2<-1, false, false>);
  f(ex_ssor_sweep2<1, false, true>);  f(ex_ssor_sweep2<-1, false, true>);
  f(ex_ssor_sweep2<1, true, false>);  f(ex_ssor_sweep2<-1, true, false>);
  f(ex_ssor_sweep2<1, true, true>);   f(ex_ssor_sweep2<-1, true, true>);
}

template <int DIR>
static void ex2_launch(int grid, cudaStream_t stream, const SsorParams& P, bool slabs)
{
  if (P.trace) {
    if (slabs) ex_ssor_sweep2<DIR, true, true><<<grid, EX2_THREADS, EX2_SMEM_BYTES, stream>>>(P);
    else ex_ssor_sweep2<DIR, true, false><<<grid, EX2_THREADS, EX2_SMEM_BYTES, stream>>>(P);
  } else {
    if (slabs) ex_ssor_sweep2<DIR, false, true><<<grid, EX2_THREADS, EX2_SMEM_BYTES, stream>>>(P);
    else ex_ssor_sweep2<DIR, false, false><<<grid, EX2_THREADS, EX2_SMEM_BYTES, stream>>>(P);
  }
}

// ex_ssor_sweep3: the same variants
template <typename F>
static void ex3_for_each_variant(F f)
{
  f(ex_ssor_sweep3<1, false, false>); f(ex_ssor_sweep3<-1, false, false>);
  f(ex_ssor_sweep3<1, false, true>);  f(ex_ssor_sweep3<-1, false, true>);
  f(ex_ssor_sweep3<1, true, false>);  f(ex_ssor_sweep3<-1, true, false>);
  f(ex_ssor_sweep3<1, true, true>);   f(ex_ssor_sweep3<-1, true, true>);
}

template <int DIR>
static void ex3_launch(int grid, cudaStream_t stream, const SsorParams& P, bool slabs)
{
  if (P.trace) {
    if (slabs) ex_ssor_sweep3<DIR, true, true><<<grid, EX3_THREADS, EX3_SMEM_BYTES, stream>>>(P);
    else ex_ssor_sweep3<DIR, true, false><<<grid, EX3_THREADS, EX3_SMEM_BYTES, stream>>>(P);
  } else {
    if (slabs) ex_ssor_sweep3<DIR, false, true><<<grid, EX3_THREADS, EX3_SMEM_BYTES, stream>>>(P);
    else ex_ssor_sweep3<DIR, false, false><<<grid, EX3_THREADS, EX3_SMEM_BYTES, stream>>>(P);
  }
}

// ---------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------
struct ExSpan { cudaEvent_t beg, end; int kind; };

struct nka_system {
  int nx = 0, ny = 0, scaling = 1;
  double a = 0.0, hx = 0.0, hy = 0.0, fx = 0.0, fy = 0.0, q = 0.0;
  size_t n = 0;
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  double* U[2] = {nullptr, nullptr};
  int cur = 0;
  double *R = nullptr, *Z = nullptr, *AXL = nullptr, *AYD = nullptr, *AC = nullptr, *AXR = nullptr, *AYT = nullptr;
  unsigned long long* bnd = nullptr;
  unsigned long long* trace = nullptr;   // debugging aid, see nka_system_ssor_trace
  // row slabs (multi-GPU): this system holds rows [k0, k0 + ny) of an nx x ny_global grid
  int ny_global = 0, k0 = 0;
  bool has_lower = false, has_upper = false;
  NkaComm* comm = nullptr;
  double *halo_u_lo = nullptr, *halo_u_hi = nullptr;    // [nx] the neighbours' edge rows of u (received before each residual)
  double *send_lo = nullptr, *send_hi = nullptr;        // [nx] own edge rows of u, packed
  void* zbox = nullptr;                                 // IPC-exported: uint4 halo_lo[nx], halo_hi[nx] (tagged z rows)
  void* zbox_mapped[NKA_MAX_RANKS] = {};                // the neighbours' boxes as mapped here
  unsigned sweep_id = 0;
  int nstrips = 0;
  double* partials = nullptr;
  unsigned* ticket = nullptr;
  double* result = nullptr;        // device: [0] = sum r^2, [1] = error word of the SSOR kernels
  double* result_host = nullptr;   // pinned
  double* stage = nullptr;         // device scratch for the order conversions
  int res_grid = 0;                // residual kernel: resident CTAs walking (diagonal, chunk) items
  int res_kernel = 2;              // 2: ex_residual_strip_kernel (3 divisions per cell); 1: ex_residual_kernel (NKA_RESIDUAL_KERNEL=1, A/B)
  int rs_grid = 0, rs_strips = 0, rs_bands = 0, rs_band = 0, rs_absolute = 1;
  unsigned rs_items = 0;
  unsigned* rs_counter = nullptr;
  unsigned* rs_first_item = nullptr;   // device: [rs_bands + 1]
  int* rs_first_strip = nullptr;       // device: [rs_bands]
  int ssor_grid = 0;
  int ssor_kernel = 2;             // 2: ex_ssor_sweep2 (the chain on a warp of its own); 3: ex_ssor_sweep3 (two columns, two chains per lane); 1: ex_ssor_sweep (A/B timing)
  bool bnd_dirty = true;
  int error = 0;
  unsigned long long launches = 0;
  bool timing = false;
  std::vector<ExSpan> spans;
  double t_ms[2] = {0.0, 0.0};
  unsigned long long t_cnt[2] = {0, 0};
};

static void ex_fold(NKASYS sy)
{
  if (sy->spans.empty()) return;
  CUDA_CHECK(cudaStreamSynchronize(sy->stream));
  for (const ExSpan& sp : sy->spans) {
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, sp.beg, sp.end));
    sy->t_ms[sp.kind] += ms;
    sy->t_cnt[sp.kind] += 1;
    cudaEventDestroy(sp.beg);
    cudaEventDestroy(sp.end);
  }
  sy->spans.clear();
}

struct ExScope {
  NKASYS sy; int kind; cudaEvent_t beg = nullptr;
  ExScope(NKASYS s, int k) : sy(s), kind(k) {
    if (sy->timing) { CUDA_CHECK(cudaEventCreate(&beg)); CUDA_CHECK(cudaEventRecord(beg, sy->stream)); }
  }
  ~ExScope() {
    if (beg) {
      cudaEvent_t end;
      CUDA_CHECK(cudaEventCreate(&end));
      CUDA_CHECK(cudaEventRecord(end, sy->stream));
      sy->spans.push_back({beg, end, kind});
      if (sy->spans.size() >= 4096) ex_fold(sy);
    }
  }
};

extern "C" NKASYS nka_system_init(int nx, int ny, double a, int scaling, int device, void* stream)
{
  return nka_system_init_slab(nx, ny, 0, ny, a, scaling, device, stream);
}

extern "C" NKASYS nka_system_init_slab(int nx, int ny_global, int k0, int k1, double a, int scaling, int device, void* stream)
{
  NKA_REQUIRE(k0 >= 0 && k1 <= ny_global && k1 - k0 >= 3, "nka_system_init_slab: a slab needs at least 3 of the grid's rows");
  const int ny = k1 - k0;
  // preconditions: src-F08/nka_example.F90:90-92
  NKA_REQUIRE(a > 0.0, "nka_system_init: a must be > 0");
  NKA_REQUIRE(nx >= 3 && ny_global >= 3, "nka_system_init: nx, ny must be >= 3");
  NKA_REQUIRE(nx <= (1 << 20) && ny_global <= (1 << 20), "nka_system_init: grid too large");
  NKA_REQUIRE(scaling == 0 || scaling == 1, "nka_system_init: scaling must be 0 (F95/C) or 1 (F08)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    nka_fail(__FILE__, __LINE__, "no CUDA device: libnka_b200 has no CPU compute path");
  NKASYS sy = new nka_system();
  if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
  sy->device = device;
  DeviceGuard guard(device);
  CUDA_CHECK(cudaDeviceGetAttribute(&sy->num_sms, cudaDevAttrMultiProcessorCount, device));
  sy->stream = (cudaStream_t)stream;
  sy->nx = nx; sy->ny = ny; sy->a = a; sy->scaling = scaling;
  sy->n = (size_t)nx * ny;
  sy->ny_global = ny_global; sy->k0 = k0;
  sy->hx = 1.0 / nx; sy->hy = 1.0 / ny_global;
  if (scaling == 0) { sy->fx = sy->hx / sy->hy; sy->fy = sy->hy / sy->hx; sy->q = sy->hx * sy->hy; }   // src-C/nka_example.c:97,225-226
  else { sy->fx = sy->hx * sy->hx; sy->fy = sy->hy * sy->hy; sy->q = 1.0; }                           // F08 :100,:132-135
  const size_t bytes = sy->n * sizeof(double);
  double** grids[] = {&sy->U[0], &sy->U[1], &sy->R, &sy->Z, &sy->AXL, &sy->AYD, &sy->AC};
  for (double** g : grids) {
    if (cudaMalloc(g, bytes) != cudaSuccess) nka_fail(__FILE__, __LINE__, "nka_system_init: out of device memory");
    CUDA_CHECK(cudaMemsetAsync(*g, 0, bytes, sy->stream));
  }
  CUDA_CHECK(cudaMalloc(&sy->AXR, ny * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&sy->AYT, nx * sizeof(double)));
  const char* kv = getenv("NKA_SSOR_KERNEL");
  sy->ssor_kernel = kv ? atoi(kv) : 2;
  NKA_REQUIRE(sy->ssor_kernel >= 1 && sy->ssor_kernel <= 3, "NKA_SSOR_KERNEL must be 1, 2 or 3");
  const int strip_cols = sy->ssor_kernel == 3 ? EX3_W : 32;
  sy->nstrips = (nx + strip_cols - 1) / strip_cols;
  CUDA_CHECK(cudaMalloc(&sy->bnd, (size_t)sy->nstrips * ny * sizeof(unsigned long long)));
  {
    const size_t nitems = (size_t)(nx + ny - 1) * (((nx < ny ? nx : ny) + EX_RES_THREADS - 1) / EX_RES_THREADS);
    NKA_REQUIRE(nitems < ((size_t)1 << 31), "nka_system_init: grid too large");
    int occ_res = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_res, ex_residual_kernel, EX_RES_THREADS, 0));
    const char* wv = getenv("NKA_RES_WAVES");
    const size_t cap = (size_t)sy->num_sms * (occ_res > 0 ? occ_res : 1) * (wv ? atoi(wv) : 4);
    sy->res_grid = (int)(nitems < cap ? nitems : cap);
  }
  size_t npartials = (size_t)sy->res_grid;
  {
    // strip formulation: items = (band of diagonals, strip of 30 columns), a few per resident warp
    const char* rk = getenv("NKA_RESIDUAL_KERNEL");
    sy->res_kernel = rk ? atoi(rk) : 2;
    NKA_REQUIRE(sy->res_kernel == 1 || sy->res_kernel == 2, "NKA_RESIDUAL_KERNEL must be 1 or 2");
    int occ = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ex_residual_strip_kernel, EX_RS_THREADS, 0));
    if (occ < 1) occ = 1;
    sy->rs_strips = (nx + EX_RS_COLS - 1) / EX_RS_COLS;
    const int ndiag = ny + EX_RS_COLS - 1;                             // diagonals a strip has cells on
    const long long warps = (long long)sy->num_sms * occ * (EX_RS_THREADS / 32);
    sy->rs_absolute = 1;
    if (const char* e = getenv("NKA_RES_ABS_BANDS")) sy->rs_absolute = atoi(e) != 0;    // A/B switch
    int per_warp = 4;                                                  // per-strip bands: items per resident warp aimed at
    if (const char* e = getenv("NKA_RES_ITEMS_PER_WARP")) if (atoi(e) > 0) per_warp = atoi(e);
    long long bands = (warps * per_warp + sy->rs_strips - 1) / sy->rs_strips;
    int band = (int)((ndiag + bands - 1) / (bands > 0 ? bands : 1));
    if (band < 32) band = 32;                                          // two prologue diagonals per band: <= 6 % redone
    // absolute bands: 32 diagonals whatever the grid (0.248 ms at 4096^2, 0.927 at 8192^2, 1.87 at 32768 x 4096;
    // 24 / 48 / 64 are within a few per cent either way, larger bands lose: profiles/r2t_residual_band_sweep.jsonl)
    if (sy->rs_absolute) band = 32;
    if (const char* e = getenv("NKA_RES_BAND")) if (atoi(e) > 0) band = atoi(e);
    sy->rs_band = band;
    sy->rs_bands = ((sy->rs_absolute ? nx + ny - 1 : ndiag) + band - 1) / band;
    size_t items = (size_t)sy->rs_strips * sy->rs_bands;
    if (sy->rs_absolute) {
      // (band, strip) pairs that hold cells, numbered band by band: nka_res_items.h
      const NkaResItems ri = nka_res_items(nx, ny, EX_RS_COLS, band);
      NKA_REQUIRE(ri.nbands == sy->rs_bands, "nka_system_init: residual item geometry");
      const std::vector<unsigned>& first = ri.first;
      const std::vector<int>& s_lo = ri.s_lo;
      items = ri.count;
      sy->rs_items = (unsigned)ri.count;
      CUDA_CHECK(cudaMalloc(&sy->rs_first_item, first.size() * sizeof(unsigned)));
      CUDA_CHECK(cudaMalloc(&sy->rs_first_strip, s_lo.size() * sizeof(int)));
      CUDA_CHECK(cudaMemcpy(sy->rs_first_item, first.data(), first.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy(sy->rs_first_strip, s_lo.data(), s_lo.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    NKA_REQUIRE(items < ((size_t)1 << 31), "nka_system_init: grid too large");
    const size_t ctas = (items + EX_RS_THREADS / 32 - 1) / (EX_RS_THREADS / 32);
    const size_t cap = (size_t)sy->num_sms * occ;
    sy->rs_grid = (int)(ctas < cap ? ctas : cap);
    if (items > npartials) npartials = items;
    CUDA_CHECK(cudaMalloc(&sy->rs_counter, sizeof(unsigned)));
    CUDA_CHECK(cudaMemsetAsync(sy->rs_counter, 0, sizeof(unsigned), sy->stream));
  }
  CUDA_CHECK(cudaMalloc(&sy->partials, npartials * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&sy->ticket, sizeof(unsigned)));
  CUDA_CHECK(cudaMemsetAsync(sy->ticket, 0, sizeof(unsigned), sy->stream));
  CUDA_CHECK(cudaMalloc(&sy->result, 2 * sizeof(double)));
  CUDA_CHECK(cudaMemsetAsync(sy->result, 0, 2 * sizeof(double), sy->stream));
  CUDA_CHECK(cudaMallocHost(&sy->result_host, 2 * sizeof(double)));
  int occ = 0, occb = 0;
  if (sy->ssor_kernel == 3) {
    occ = occb = 1 << 20;
    ex3_for_each_variant([&occ](auto kernel) {
      CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EX3_SMEM_BYTES));
      int o = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, EX3_THREADS, EX3_SMEM_BYTES));
      if (o < occ) occ = o;
    });
  } else if (sy->ssor_kernel == 2) {
    // the grid must be co-resident whichever variant is launched: the smallest occupancy counts
    occ = occb = 1 << 20;
    ex2_for_each_variant([&occ](auto kernel) {
      CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EX2_SMEM_BYTES));
      int o = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, EX2_THREADS, EX2_SMEM_BYTES));
      if (o < occ) occ = o;
    });
  } else {
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ex_ssor_sweep<1>, 64, 0));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occb, ex_ssor_sweep<-1>, 64, 0));
  }
  if (occb < occ) occ = occb;
  NKA_REQUIRE(occ >= 1, "nka_system_init: the SSOR kernel does not fit on an SM");
  const int want = sy->nstrips;
  const int cap = occ * sy->num_sms;
  sy->ssor_grid = want < cap ? want : cap;
  return sy;
}

extern "C" void nka_system_delete(NKASYS sy)
{
  if (!sy) return;
  DeviceGuard guard(sy->device);
  cudaStreamSynchronize(sy->stream);
  for (const ExSpan& sp : sy->spans) { cudaEventDestroy(sp.beg); cudaEventDestroy(sp.end); }
  cudaFree(sy->U[0]); cudaFree(sy->U[1]); cudaFree(sy->R); cudaFree(sy->Z); cudaFree(sy->AXL);
  cudaFree(sy->AYD); cudaFree(sy->AC); cudaFree(sy->AXR); cudaFree(sy->AYT); cudaFree(sy->bnd);
  cudaFree(sy->trace);
  if (sy->comm) { nka_ipc_unmap(sy->comm, sy->zbox_mapped); nka_comm_release(sy->comm); }
  cudaFree(sy->zbox); cudaFree(sy->halo_u_lo); cudaFree(sy->halo_u_hi); cudaFree(sy->send_lo); cudaFree(sy->send_hi);
  cudaFree(sy->partials); cudaFree(sy->ticket); cudaFree(sy->result); cudaFree(sy->stage); cudaFree(sy->rs_counter);
  cudaFree(sy->rs_first_item); cudaFree(sy->rs_first_strip);
  cudaFreeHost(sy->result_host);
  delete sy;
}

// Collective.  Rank r must hold the slab directly above rank r-1's.  Creates the communicator
// (shared later with the accelerator: nka_comm_share_system), the receive buffers for the edge rows
// of u, and the tagged edge rows of z that the neighbouring ranks write through NVLink (CUDA IPC).
extern "C" int nka_system_comm_init(NKASYS sy, int nranks, int rank, const void* id128)
{
  NKA_REQUIRE(sy != NULL && id128 != NULL, "nka_system_comm_init: null argument");
  NKA_REQUIRE(nranks >= 1 && nranks <= NKA_MAX_RANKS && rank >= 0 && rank < nranks, "nka_system_comm_init: bad rank/nranks");
  NKA_REQUIRE(sy->comm == nullptr, "nka_system_comm_init: already attached");
  NKA_REQUIRE((rank == 0) == (sy->k0 == 0) && (rank == nranks - 1) == (sy->k0 + sy->ny == sy->ny_global),
              "nka_system_comm_init: slabs must be stacked in rank order");
  if (!nka_nccl_load() || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) return -1;
  DeviceGuard guard(sy->device);
  NkaId128 id;
  memcpy(id.bytes, id128, sizeof id.bytes);
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (rc != 0) return rc;
  NkaComm* c = new NkaComm();
  c->comm = comm; c->owned = true; c->nranks = nranks; c->rank = rank;
  sy->comm = c;
  sy->has_lower = rank > 0;
  sy->has_upper = rank + 1 < nranks;
  const size_t row = (size_t)sy->nx * sizeof(double);
  CUDA_CHECK(cudaMalloc(&sy->halo_u_lo, row)); CUDA_CHECK(cudaMalloc(&sy->halo_u_hi, row));
  CUDA_CHECK(cudaMalloc(&sy->send_lo, row)); CUDA_CHECK(cudaMalloc(&sy->send_hi, row));
  size_t zbytes = 2 * (size_t)sy->nx * sizeof(uint4);
  zbytes = (zbytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);          // whole 2 MiB granules: what CUDA IPC exports
  CUDA_CHECK(cudaMalloc(&sy->zbox, zbytes));
  CUDA_CHECK(cudaMemsetAsync(sy->zbox, 0, zbytes, sy->stream));          // tag 0 = never written (sweeps count from 1)
  CUDA_CHECK(cudaStreamSynchronize(sy->stream));
  bool want[NKA_MAX_RANKS];
  for (int r = 0; r < NKA_MAX_RANKS; ++r) want[r] = (r == rank - 1 || r == rank + 1);
  if (!nka_ipc_exchange(c, sy->stream, sy->zbox, sy->zbox_mapped, want))
    nka_fail(__FILE__, __LINE__, "nka_system_comm_init: the ranks cannot map each other's memory (CUDA IPC / NVLink peer "
                                 "access): the slab-pipelined SSOR sweep needs it");
  sy->zbox_mapped[rank] = nullptr;
  return 0;
}

// The accelerator of a slab system sums its dot products over the same ranks.
extern "C" void nka_comm_share_system(NKA acc, NKASYS sy)
{
  NKA_REQUIRE(acc != NULL && sy != NULL && sy->comm != NULL, "nka_comm_share_system: the system has no communicator");
  nka_attach_shared_comm(acc, sy->comm);
}

extern "C" size_t nka_system_size(NKASYS sy) { NKA_REQUIRE(sy != NULL, "nka_system_size: null handle"); return sy->n; }
extern "C" void* nka_system_stream(NKASYS sy) { NKA_REQUIRE(sy != NULL, "nka_system_stream: null handle"); return (void*)sy->stream; }

extern "C" double* nka_system_field(NKASYS sy, int field)
{
  NKA_REQUIRE(sy != NULL, "nka_system_field: null handle");
  switch (field) {
    case NKA_FIELD_U: return sy->U[sy->cur];
    case NKA_FIELD_R: return sy->R;
    case NKA_FIELD_Z: return sy->Z;
    case NKA_FIELD_AXL: return sy->AXL;
    case NKA_FIELD_AYD: return sy->AYD;
    case NKA_FIELD_AC: return sy->AC;
  }
  nka_fail(__FILE__, __LINE__, "nka_system_field: unknown field");
}

extern "C" size_t nka_system_index(NKASYS sy, int j, int k)
{
  NKA_REQUIRE(sy != NULL, "nka_system_index: null handle");
  NKA_REQUIRE(j >= 0 && j < sy->nx && k >= 0 && k < sy->ny, "nka_system_index: cell out of range");
  return (size_t)(wf_base(j + k, sy->nx, sy->ny) + j);
}

static int ex_grid(NKASYS sy, size_t work, int threads)
{
  size_t need = (work + threads - 1) / threads, cap = (size_t)sy->num_sms * 16;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

extern "C" void nka_system_set_field(NKASYS sy, int field, const double* host)
{
  double* dst = nka_system_field(sy, field);
  DeviceGuard guard(sy->device);
  if (!sy->stage) CUDA_CHECK(cudaMalloc(&sy->stage, sy->n * sizeof(double)));
  CUDA_CHECK(cudaMemcpyAsync(sy->stage, host, sy->n * sizeof(double), cudaMemcpyHostToDevice, sy->stream));
  ex_permute_kernel<<<ex_grid(sy, sy->n, 256), 256, 0, sy->stream>>>(dst, sy->stage, sy->nx, sy->ny, 1);
  CUDA_CHECK(cudaGetLastError());
  sy->launches += 1;
  CUDA_CHECK(cudaStreamSynchronize(sy->stream));
}

extern "C" void nka_system_get_field(NKASYS sy, int field, double* host)
{
  const double* src = nka_system_field(sy, field);
  DeviceGuard guard(sy->device);
  if (!sy->stage) CUDA_CHECK(cudaMalloc(&sy->stage, sy->n * sizeof(double)));
  ex_permute_kernel<<<ex_grid(sy, sy->n, 256), 256, 0, sy->stream>>>(sy->stage, src, sy->nx, sy->ny, 0);
  CUDA_CHECK(cudaGetLastError());
  sy->launches += 1;
  CUDA_CHECK(cudaMemcpyAsync(host, sy->stage, sy->n * sizeof(double), cudaMemcpyDeviceToHost, sy->stream));
  CUDA_CHECK(cudaStreamSynchronize(sy->stream));
}

extern "C" double nka_system_residual(NKASYS sy, int subtract_z)
{
  NKA_REQUIRE(sy != NULL, "nka_system_residual: null handle");
  DeviceGuard guard(sy->device);
  NkaRange nvtx("nka:example residual");
  ResParams P;
  P.nx = sy->nx; P.ny = sy->ny;
  P.U = sy->U[sy->cur];
  P.Zc = subtract_z ? sy->Z : nullptr;
  P.Unew = sy->U[sy->cur ^ 1];
  P.R = sy->R; P.AXL = sy->AXL; P.AYD = sy->AYD; P.AC = sy->AC; P.AXR = sy->AXR; P.AYT = sy->AYT;
  P.a = sy->a; P.fx = sy->fx; P.fy = sy->fy; P.q = sy->q;
  P.partials = sy->partials; P.ticket = sy->ticket; P.sumsq = sy->result;
  P.HLO = sy->has_lower ? sy->halo_u_lo : nullptr;
  P.HHI = sy->has_upper ? sy->halo_u_hi : nullptr;
  if (sy->comm && sy->comm->nranks > 1) {
    // edge rows of (u - z) to and from the neighbouring slabs: 2 x nx doubles each way
    ex_pack_rows_kernel<<<ex_grid(sy, sy->nx, 256), 256, 0, sy->stream>>>(P.U, P.Zc, sy->nx, sy->ny, sy->send_lo, sy->send_hi);
    CUDA_CHECK(cudaGetLastError());
    sy->launches += 1;
    const int me = sy->comm->rank;
    int rc = g_nccl.GroupStart();
    if (sy->has_lower) {
      rc |= g_nccl.Send(sy->send_lo, sy->nx, kNcclFloat64, me - 1, sy->comm->comm, sy->stream);
      rc |= g_nccl.Recv(sy->halo_u_lo, sy->nx, kNcclFloat64, me - 1, sy->comm->comm, sy->stream);
    }
    if (sy->has_upper) {
      rc |= g_nccl.Send(sy->send_hi, sy->nx, kNcclFloat64, me + 1, sy->comm->comm, sy->stream);
      rc |= g_nccl.Recv(sy->halo_u_hi, sy->nx, kNcclFloat64, me + 1, sy->comm->comm, sy->stream);
    }
    rc |= g_nccl.GroupEnd();
    if (rc != 0) nka_fail(__FILE__, __LINE__, "nka_system_residual: NCCL edge-row exchange failed");
  }
  {
    ExScope t(sy, 1);
    if (sy->res_kernel == 2) {
      ResStripParams Q;
      Q.p = P; Q.nstrips = sy->rs_strips; Q.nbands = sy->rs_bands; Q.band = sy->rs_band; Q.absolute = sy->rs_absolute; Q.counter = sy->rs_counter;
      Q.nitems = sy->rs_items; Q.first_item = sy->rs_first_item; Q.first_strip = sy->rs_first_strip;
      ex_residual_strip_kernel<<<sy->rs_grid, EX_RS_THREADS, 0, sy->stream>>>(Q);
    } else {
      ex_residual_kernel<<<sy->res_grid, EX_RES_THREADS, 0, sy->stream>>>(P);
    }
    CUDA_CHECK(cudaGetLastError());
    sy->launches += 1;
  }
  if (subtract_z) sy->cur ^= 1;
  if (sy->comm && sy->comm->nranks > 1) {
    // global norm: sum of the slabs' sums of squares, the same bits on every rank
    const int rc = g_nccl.AllReduce(sy->result, sy->result, 1, kNcclFloat64, kNcclSum, sy->comm->comm, sy->stream);
    if (rc != 0) nka_fail(__FILE__, __LINE__, "nka_system_residual: ncclAllReduce failed");
  }
  CUDA_CHECK(cudaMemcpyAsync(sy->result_host, sy->result, 2 * sizeof(double), cudaMemcpyDeviceToHost, sy->stream));
  CUDA_CHECK(cudaStreamSynchronize(sy->stream));
  int errw = 0;
  memcpy(&errw, &sy->result_host[1], sizeof errw);
  if (errw) { sy->error = errw; sy->bnd_dirty = true; }
  return sqrt(sy->result_host[0]);
}

extern "C" int nka_system_pc_ssor(NKASYS sy, int nsweep, double omega)
{
  NKA_REQUIRE(sy != NULL, "nka_system_pc_ssor: null handle");
  NKA_REQUIRE(nsweep >= 1, "nka_system_pc_ssor: nsweep must be >= 1");       // F08 :156-157
  NKA_REQUIRE(omega > 0.0, "nka_system_pc_ssor: omega must be > 0");
  DeviceGuard guard(sy->device);
  NkaRange nvtx("nka:example pc_ssor");
  if (sy->bnd_dirty) {
    const size_t nb = (size_t)sy->nstrips * sy->ny;
    ex_fill_u64<<<ex_grid(sy, nb, 256), 256, 0, sy->stream>>>(sy->bnd, nb, EX_SENT);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemsetAsync(sy->result + 1, 0, sizeof(double), sy->stream));
    sy->launches += 1;
    sy->bnd_dirty = false;
  }
  SsorParams P;
  P.nx = sy->nx; P.ny = sy->ny; P.nstrips = sy->nstrips;
  P.R = sy->R; P.AC = sy->AC; P.AXL = sy->AXL; P.AYD = sy->AYD; P.AXR = sy->AXR; P.AYT = sy->AYT;
  P.Z = sy->Z; P.bnd = sy->bnd;
  P.omega = omega; P.om1 = 1.0 - omega;
  P.err = reinterpret_cast<int*>(sy->result + 1);
  P.trace = sy->trace;
  const bool slabs = sy->comm && sy->comm->nranks > 1;
  uint4* zb = reinterpret_cast<uint4*>(sy->zbox);
  P.halo_lo = sy->has_lower ? zb : nullptr;
  P.halo_hi = sy->has_upper ? zb + sy->nx : nullptr;
  // the upper rank's halo_lo is the first row of its box, the lower rank's halo_hi the second row of its box
  P.peer_up_lo = sy->has_upper ? reinterpret_cast<uint4*>(sy->zbox_mapped[sy->comm->rank + 1]) : nullptr;
  P.peer_dn_hi = sy->has_lower ? reinterpret_cast<uint4*>(sy->zbox_mapped[sy->comm->rank - 1]) + sy->nx : nullptr;
  P.spin_limit = slabs ? 60 * EX_SPIN_LIMIT : EX_SPIN_LIMIT;          // ~2 s on one GPU, ~2 min when another GPU is upstream
  ExScope t(sy, 0);
  for (int i = 0; i < nsweep; ++i) {
    P.zero_old = (i == 0) ? 1 : 0;                       // z = 0 start (:158): nothing to read yet
    P.tag_prev = sy->sweep_id; P.tag_cur = ++sy->sweep_id;
    if (sy->ssor_kernel == 3) ex3_launch<1>(sy->ssor_grid, sy->stream, P, slabs);
    else if (sy->ssor_kernel == 2) ex2_launch<1>(sy->ssor_grid, sy->stream, P, slabs);
    else ex_ssor_sweep<1><<<sy->ssor_grid, 64, 0, sy->stream>>>(P);
    CUDA_CHECK(cudaGetLastError());
    P.zero_old = 0;
    P.tag_prev = sy->sweep_id; P.tag_cur = ++sy->sweep_id;
    if (sy->ssor_kernel == 3) ex3_launch<-1>(sy->ssor_grid, sy->stream, P, slabs);
    else if (sy->ssor_kernel == 2) ex2_launch<-1>(sy->ssor_grid, sy->stream, P, slabs);
    else ex_ssor_sweep<-1><<<sy->ssor_grid, 64, 0, sy->stream>>>(P);
    CUDA_CHECK(cudaGetLastError());
    sy->launches += 2;
  }
  return sy->error;
}

// Debugging / tuning aid (tools/ssor_trace.py; not part of the reference's interface): with on != 0 the
// next sweeps record per strip {start, first cell done, end} globaltimer stamps and the number of
// synchronous waits on the neighbour strip; out (may be NULL) receives nstrips*4 words of the last sweep.
extern "C" int nka_system_ssor_trace(NKASYS sy, int on, unsigned long long* out)
{
  NKA_REQUIRE(sy != NULL, "nka_system_ssor_trace: null handle");
  DeviceGuard guard(sy->device);
  const size_t bytes = ((size_t)sy->nstrips * 4 + 256) * sizeof(unsigned long long);
  CUDA_CHECK(cudaStreamSynchronize(sy->stream));
  if (out && sy->trace) CUDA_CHECK(cudaMemcpy(out, sy->trace, bytes, cudaMemcpyDeviceToHost));
  if (on && !sy->trace) CUDA_CHECK(cudaMalloc(&sy->trace, bytes));
  if (!on && sy->trace) { cudaFree(sy->trace); sy->trace = nullptr; }
  if (sy->trace) CUDA_CHECK(cudaMemset(sy->trace, 0, bytes));
  return sy->nstrips;
}

extern "C" int nka_example_solve(NKASYS sy, NKA acc, int nsweep, double omega, int maxitr, double tol,
                                 double* rnorm, int* nvec_seq)
{
  NKA_REQUIRE(sy != NULL && rnorm != NULL, "nka_example_solve: null argument");
  if (acc) {
    NKA_REQUIRE(nka_vec_len64(acc) == sy->n, "nka_example_solve: the accelerator's vlen must be nx*ny");
    NKA_REQUIRE(nka_get_stream(acc) == (void*)sy->stream, "nka_example_solve: accelerator and system must share a stream");
  }
  // src-F08/nka_example.F90:242-255 ; src-C/nka_example.c:131-166
  const double rnorm0 = nka_system_residual(sy, 0);
  rnorm[0] = rnorm0;
  int itr;
  for (itr = 1; itr <= maxitr; ++itr) {
    nka_system_pc_ssor(sy, nsweep, omega);
    if (acc) {
      nka_accel_update_dev(acc, sy->Z);
      if (nvec_seq) nvec_seq[itr - 1] = nka_num_vec(acc);
    }
    rnorm[itr] = nka_system_residual(sy, 1);
    if (sy->error) nka_fail(__FILE__, __LINE__, "nka_example_solve: an SSOR sweep timed out waiting for its neighbour strip");
    if (rnorm[itr] < tol * rnorm0) break;
  }
  return itr > maxitr ? maxitr : itr;
}

// ---------------------------------------------------------------------------
// self-check of the split division used by ex_ssor_sweep2 (ex2_rcp / ex2_div_fast / ex2_div_safe)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ex_mix64(unsigned long long z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Pairs (x, b): random signs and mantissas (every 8th with a mantissa of all ones, all zeros, or
// one bit off those: the cases where a quotient falls closest to a rounding boundary); exponents
// within +-60 of 1 for most, anywhere (denormals, infinities, NaN included) for one in 16, so the
// operand-range guard and the __ddiv_rn fallback are exercised too.
__global__ void ex_division_check_kernel(unsigned long long nsamples, unsigned long long seed, unsigned long long* mismatches)
{
  unsigned long long bad = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nsamples;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long h0 = ex_mix64(seed + 3 * i), h1 = ex_mix64(seed + 3 * i + 1), h2 = ex_mix64(seed + 3 * i + 2);
    unsigned long long mx = h0 & 0x000FFFFFFFFFFFFFull, mb = h1 & 0x000FFFFFFFFFFFFFull;
    if ((h2 & 7) == 0) {
      const unsigned long long pat[4] = {0x000FFFFFFFFFFFFFull, 0ull, 0x000FFFFFFFFFFFFEull, 1ull};
      mb = pat[(h2 >> 3) & 3];
      if (h2 & 32) mx = pat[(h2 >> 6) & 3];
    }
    unsigned long long ex, eb;
    if (((h2 >> 8) & 15) == 0) { ex = (h2 >> 12) & 0x7ff; eb = (h2 >> 23) & 0x7ff; }
    else { ex = 963 + ((h2 >> 12) % 121); eb = 963 + ((h2 >> 23) % 121); }
    const double x = __longlong_as_double((long long)(((h2 >> 40) & 1) << 63 | ex << 52 | mx));
    const double b = __longlong_as_double((long long)(((h2 >> 41) & 1) << 63 | eb << 52 | mb));
    const double want = __ddiv_rn(x, b);
    double got = ex2_div_fast(x, b, ex2_rcp(b));
    if (!(ex2_div_safe(b) && ex2_div_safe(x))) got = __ddiv_rn(x, b);
    if (__double_as_longlong(want) != __double_as_longlong(got)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

extern "C" unsigned long long nka_example_division_check(unsigned long long nsamples, unsigned long long seed, int device)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    nka_fail(__FILE__, __LINE__, "no CUDA device: libnka_b200 has no CPU compute path");
  if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
  DeviceGuard guard(device);
  unsigned long long* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof *d));
  CUDA_CHECK(cudaMemset(d, 0, sizeof *d));
  ex_division_check_kernel<<<148 * 8, 256>>>(nsamples, seed, d);
  CUDA_CHECK(cudaGetLastError());
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return h;
}

extern "C" void nka_system_timing_enable(NKASYS sy, int on)
{
  NKA_REQUIRE(sy != NULL, "nka_system_timing_enable: null handle");
  DeviceGuard guard(sy->device);
  if (!on) ex_fold(sy);
  else { sy->t_ms[0] = sy->t_ms[1] = 0.0; sy->t_cnt[0] = sy->t_cnt[1] = 0; }
  sy->timing = on != 0;
}

extern "C" void nka_system_timing_read(NKASYS sy, double ms[2], unsigned long long count[2])
{
  NKA_REQUIRE(sy != NULL, "nka_system_timing_read: null handle");
  DeviceGuard guard(sy->device);
  ex_fold(sy);
  for (int k = 0; k < 2; ++k) { ms[k] = sy->t_ms[k]; count[k] = sy->t_cnt[k]; }
}

extern "C" unsigned long long nka_system_launch_count(NKASYS sy)
{
  NKA_REQUIRE(sy != NULL, "nka_system_launch_count: null handle");
  return sy->launches;
}
