// nka_state.h -- the small, scalar half of accel_update: list bookkeeping,
// Gram row, Cholesky refactorisation, capacity eviction, vtol drop test,
// triangular solves, and the "plans" that tell the two streaming kernels which
// columns to touch.  Runs on the device in one thread of one warp
// (nka_state_kernel); written as plain C++ behind NKA_HD so the GPU-less build
// box can exercise the same logic in a test-only host model (tests/model/).
//
// Reference being replaced (same arithmetic order, so drop decisions agree):
//   src-C/nonlinear_krylov_accelerator.c:295-443   src-F08/nka_type.F90:263-417
//   relax  src-C/...c:466-485   restart  src-C/...c:447-463
//
// Storage model ("raw chain"):
//   W[k] (k = list slot) holds the RAW input f cached when slot k was created,
//   for as long as slot k is `chained`; the reference's normalised difference
//   is  w_k = (W[k] - W[prev[k]]) / s[k].  A slot whose newer neighbour leaves
//   the list is "materialised": W[k] <- W[k] - W[prev k], chained[k] = 0, and
//   from then on w_k = W[k] / s[k].
//   Z[k] holds, for a finished pair, Z'_k = Y_k + f_next, where Y_k = f_out - f_in
//   is the correction applied by the call that created slot k, so that
//   v_k - w_k = Z[k] / s[k]   (reference v, w: src-C/...c:316-320, :423).
//   Y of the pending slot is NOT stored: it is a linear combination of columns
//   pass B reads anyway, so the next pass B recomputes it from the saved
//   coefficients coefY[] (one column write less per update).
//   Lazy last column: when the list is full the oldest pair is evicted unless a
//   vtol drop makes room; pass A then skips its column (skip_last) and the rare
//   case that needs it after all runs a two-dot fix-up sweep (need_fixup).
#pragma once

#if defined(__CUDACC__)
#define NKA_HD __host__ __device__ __forceinline__
#else
#define NKA_HD inline
#endif

#ifndef NKA_MAXSLOT
#define NKA_MAXSLOT 33   // mvec <= 32
#endif
#define NKA_NIL (-1)

#if defined(__CUDA_ARCH__)
#define NKA_SQRT(x) sqrt(x)
#else
#include <cmath>
#define NKA_SQRT(x) std::sqrt(x)
#endif

// Which stored columns pass A streams, in list order (newest first), and how
// each difference is formed:  d_j = W[col[j]] - (bit j of submask ? prev : 0),
// prev = f for j == 0, W[col[j-1]] otherwise.
struct NkaPlanA {
  int ncol;          // list length on entry
  int skip_last;     // 1: pass A leaves out col[ncol-1] (it will be evicted unless a drop occurs)
  unsigned long long submask;   // 64-bit: the list holds up to NKA_MAXSLOT = 33 positions
  int col[NKA_MAXSLOT];
};

// In-place conversions W[dst] -= W[sub], executed oldest first, before pass B.
struct NkaPlanM {
  int n;
  int dst[NKA_MAXSLOT];
  int sub[NKA_MAXSLOT];
};

// What pass B does, with z_k = Z[zcol[k]] (the pairs on the list at entry, newest first):
//   if has_pair:  Y = sum_k coefY[k] z_k  (the previous call's correction, recomputed);
//                 zp = Y + f;  Z[pslot] <- zp
//   y = coef_p * zp + sum_k coefN[k] z_k      (coefN = 0 for pairs dropped this call)
//   W[newslot] <- f;  f <- f + y (if write_f).
struct NkaPlanB {
  int newslot;
  int has_pair;
  int pslot;
  int write_f;
  int nz;
  int zcol[NKA_MAXSLOT];
  double coef_p;
  double coefN[NKA_MAXSLOT];
  double coefY[NKA_MAXSLOT];
};

struct NkaDevState {
  // --- the reference's state (0-based slots, NKA_NIL terminates lists)
  int mvec;
  int subspace, pending;
  int first, last, free_;
  int next[NKA_MAXSLOT], prev[NKA_MAXSLOT];
  double vtol;
  double h[NKA_MAXSLOT * NKA_MAXSLOT];   // h[r*NKA_MAXSLOT + c]; raw Gram at (newer,older), factor at (older,newer)
  double c[NKA_MAXSLOT];
  // --- raw-chain bookkeeping
  double s[NKA_MAXSLOT];                 // norm of the raw difference of pair k
  double coefY[NKA_MAXSLOT];             // coefficient of Z[k] in the last correction y (0 if absent)
  int chained[NKA_MAXSLOT];
  int lazy_last;                         // feature switch: allow pass A to skip the doomed oldest column
  int need_fixup;                        // phase 1 found it needs the skipped column after all
  // --- plans
  NkaPlanA planA;
  NkaPlanM planM;
  NkaPlanB planB;
  // --- diagnostics of the last update (parity tests compare these with the oracle)
  int ndrop_last, evicted_last, relaxed_last;
  double min_margin;
  double s_last;
  unsigned long long ncalls;
  int error;                             // nonzero: an internal invariant failed
};

#define NKA_H(S, r, c_) ((S).h[(r) * NKA_MAXSLOT + (c_)])

// Multi-GPU, one process per GPU on one NVLink domain: the partial dot products are summed
// across ranks INSIDE pass A (nka_peer_allreduce, nka_kernels.cuh) through exchange boxes that
// every rank maps into every peer (CUDA IPC).  A box holds, for two alternating parities and
// each sending rank, NKA_PEER_K 16-byte slots {lo32, tag, hi32, tag}: value and tag travel in
// one 16-byte store whose two 8-byte halves carry their own tag, so no fence is needed.
#define NKA_MAX_RANKS 16
#define NKA_PEER_K (2 * NKA_MAXSLOT)
#define NKA_PEER_BOX_WORDS16 (2 * NKA_MAX_RANKS * NKA_PEER_K)      // 16-byte slots per box
struct NkaPeerCtx {
  void* box[NKA_MAX_RANKS];    // box[r]: rank r's exchange box as mapped in this process
  int nranks, rank;
  unsigned epoch;              // exchanges completed (identical on every rank)
  int timed_out;
  unsigned long long timeout_ns;   // a peer that never arrives: trap instead of hanging the GPU
};

NKA_HD void nka_build_plan_a(NkaDevState& S)
{
  NkaPlanA& A = S.planA;
  int j = 0;
  unsigned long long mask = 0;
  for (int k = S.first; k != NKA_NIL; k = S.next[k], ++j) {
    A.col[j] = k;
    if (j == 0 ? S.pending : S.chained[k]) mask |= (1ull << j);
  }
  A.ncol = j;
  A.submask = mask;
  A.skip_last = (S.lazy_last && S.pending && j == S.mvec + 1) ? 1 : 0;
}

NKA_HD void nka_state_restart(NkaDevState& S)
{
  // src-C/...c:447-463
  S.first = NKA_NIL;
  S.last = NKA_NIL;
  S.subspace = 0;
  S.pending = 0;
  S.free_ = 0;
  for (int k = 0; k < S.mvec; ++k) S.next[k] = k + 1;
  S.next[S.mvec] = NKA_NIL;
  for (int k = 0; k <= S.mvec; ++k) { S.prev[k] = NKA_NIL; S.chained[k] = 0; S.coefY[k] = 0.0; }
  S.planM.n = 0;
  S.need_fixup = 0;
  nka_build_plan_a(S);
}

NKA_HD void nka_state_init(NkaDevState& S, int mvec, double vtol)
{
  S.mvec = mvec;
  S.vtol = vtol;
  S.ncalls = 0;
  S.error = 0;
  S.ndrop_last = S.evicted_last = S.relaxed_last = 0;
  S.min_margin = 0.0;
  S.s_last = 0.0;
  for (int i = 0; i < NKA_MAXSLOT * NKA_MAXSLOT; ++i) S.h[i] = 0.0;
  for (int i = 0; i < NKA_MAXSLOT; ++i) { S.c[i] = 0.0; S.s[i] = 1.0; S.coefY[i] = 0.0; }
  S.lazy_last = 1;
  S.planB.newslot = 0; S.planB.has_pair = 0; S.planB.pslot = 0; S.planB.write_f = 0; S.planB.nz = 0;
  S.planB.coef_p = 0.0;
  nka_state_restart(S);
}

// List surgery of relax(): src-C/...c:470-484.  Returns the removed slot or NIL.
NKA_HD int nka_list_relax(NkaDevState& S)
{
  if (!S.pending) return NKA_NIL;
  const int head = S.first;
  S.first = S.next[head];
  if (S.first == NKA_NIL) S.last = NKA_NIL;
  else S.prev[S.first] = NKA_NIL;
  S.next[head] = S.free_;
  S.free_ = head;
  S.pending = 0;
  return head;
}

// Every surviving chained pair at a position after `jr` (positions refer to the
// list as pass A saw it: ord[0..L-1]) loses its newer neighbour's raw column,
// so it is converted in place, oldest first.  removed[slot] marks slots that
// left the list during this call.
NKA_HD void nka_plan_materialise(NkaDevState& S, const int* ord, int L, int jr, const bool* removed)
{
  NkaPlanM& M = S.planM;
  M.n = 0;
  for (int j = L - 1; j > jr; --j) {
    const int k = ord[j];
    if (S.chained[k]) {
      if (!removed[k]) {
        M.dst[M.n] = k;
        M.sub[M.n] = ord[j - 1];
        ++M.n;
      }
      S.chained[k] = 0;
    }
  }
}

// relax() as a host-requested operation between updates.
NKA_HD void nka_state_relax(NkaDevState& S)
{
  S.planM.n = 0;
  if (S.pending) {
    int ord[NKA_MAXSLOT];
    bool removed[NKA_MAXSLOT];
    const int L = S.planA.ncol;
    for (int j = 0; j < L; ++j) ord[j] = S.planA.col[j];
    for (int k = 0; k < NKA_MAXSLOT; ++k) removed[k] = false;
    const int head = nka_list_relax(S);
    removed[head] = true;
    nka_plan_materialise(S, ord, L, 0, removed);
  }
  nka_build_plan_a(S);
}

// How the step is executed: by one thread (the host model in tests/model/) or by the 32 lanes
// of a warp (the device).  Every entry of the Gram row, of the Cholesky factor and of the two
// triangular solves is computed with the reference's operations in the reference's order; the
// lanes only work on DIFFERENT entries at the same time (column-oriented: as soon as a vector is
// known to stay, every later row forms its entry against it), so results do not depend on the
// number of lanes.  The decisions (capacity eviction, vtol drop, relax guard) stay sequential.
struct NkaSerial {
  NKA_HD int lane() const { return 0; }
  NKA_HD int nlanes() const { return 1; }
  NKA_HD void sync() const {}
};

// Working storage of one step, shared by the lanes.
struct NkaStepScratch {
  int ord[NKA_MAXSLOT];          // the list as pass A saw it: slot at each position
  int removed[NKA_MAXSLOT];      // slot left the list during this call
  int act[NKA_MAXSLOT];          // the slots known to stay, in list order (p first)
  int lst[NKA_MAXSLOT];          // the list after the update
  double rhs[NKA_MAXSLOT];       // <f, w_k>, then the running right-hand side of the solves
  double hkk[NKA_MAXSLOT];       // 1 - sum of squares of the factor row being formed, per slot
  int na, nl, nvec, jr, has_pair, ret, stop, newcol, pending_at_entry, nw;
  double s;
};

// Row entries against the newly fixed column act[a], for every row at a position >= from that
// still awaits its decision:   h(k,j) = (G(j,k) - sum_{i before j} h(k,i) h(j,i)) / h(j,j)
// (src-C/...c:350-360), and the running  hkk -= h(k,j)^2.
template <class Par>
NKA_HD void nka_cholesky_column(NkaDevState& S, NkaStepScratch& X, int a, int from, int L, int missing, const Par& par)
{
  const int j = X.act[a];
  const double* hj = &NKA_H(S, j, 0);
  const double hjj = hj[j];
  for (int q = from + par.lane(); q < L; q += par.nlanes()) {
    if (q == missing) continue;
    const int k = X.ord[q];
    double* hk = &NKA_H(S, k, 0);
    double hkj = NKA_H(S, j, k);
    for (int ii = 0; ii < a; ++ii) hkj -= hk[X.act[ii]] * hj[X.act[ii]];
    hkj /= hjj;
    hk[j] = hkj;
    X.hkk[k] -= hkj * hkj;
  }
}

// The scalar part of one accel_update.  `dots` holds what pass A reduced, laid
// out as dd[j] = d_0 . d_j  (j < ncol)  followed by  fd[j] = f . d_j  at
// dots[stride + j]; ignored when the list was empty on entry.
//
// have_last = 0 means pass A skipped the oldest column (planA.skip_last): if the
// factorisation turns out to need it, nothing is committed and the function
// returns 1 with need_fixup set; the caller then computes the two missing dot
// products and calls again with have_last = 1.  Returns 0 when the step is done.
// Called by all par.nlanes() lanes together; S and X are shared by them.
template <class Par>
NKA_HD int nka_state_step(NkaDevState& S, NkaStepScratch& X, const double* dots, int stride, int have_last, const Par& par)
{
  const bool lead = par.lane() == 0;
  const int L = S.planA.ncol;
  const int missing = (S.planA.skip_last && !have_last) ? L - 1 : -1;   // position without dots
  const double* dd = dots;
  const double* fd = dots + stride;
  par.sync();
  if (lead) {
    for (int j = 0; j < L; ++j) X.ord[j] = S.planA.col[j];
    for (int k = 0; k < NKA_MAXSLOT; ++k) { X.removed[k] = 0; X.rhs[k] = 0.0; X.hkk[k] = 1.0; }
    X.pending_at_entry = S.pending;
    X.jr = L;            // first list position whose slot left the list this call
    X.has_pair = 0;
    X.ret = 0; X.stop = 0; X.newcol = 0;
    X.s = 0.0;
    S.ndrop_last = 0;
    S.evicted_last = 0;
    S.relaxed_last = 0;
    S.min_margin = 1.0e300;
    S.s_last = 0.0;
    S.planM.n = 0;
    S.need_fixup = 0;

    // Step A: norm of the new difference; zero guard.  src-C/...c:295-311
    if (S.pending) {
      X.s = NKA_SQRT(dd[0]);
      S.s_last = X.s;
      if (X.s == 0.0) {
        if (missing >= 0) { S.need_fixup = 1; X.ret = 1; }   // every old pair stays: its f.d is needed
        else {
          X.removed[nka_list_relax(S)] = 1;
          S.relaxed_last = 1;
          X.jr = 0;
        }
      }
    }
  }
  par.sync();
  if (X.ret) return 1;

  // Step B: Gram row, refactorisation, drops.  src-C/...c:313-385
  if (S.pending) {
    const int p = S.first;
    const double s = X.s;
    if (lead) {
      S.s[p] = s;
      S.chained[p] = 1;
      X.has_pair = 1;
      NKA_H(S, p, p) = 1.0;
      X.nvec = 1;
      X.na = 1;
      X.act[0] = p;
    }
    // <w_1, w_k> = (d_0 . d_k) / (s s_k); the reference normalises first (:317-324)
    for (int j = 1 + par.lane(); j < L; j += par.nlanes())
      if (j != missing) NKA_H(S, p, X.ord[j]) = (dd[j] / s) / S.s[X.ord[j]];
    par.sync();
    const double tol2 = S.vtol * S.vtol;
    nka_cholesky_column(S, X, 0, 1, L, missing, par);
    par.sync();
    for (int pos = 1; pos < L; ++pos) {            // the rows in list order: one decision each
      if (lead) {
        const int k = X.ord[pos];
        X.newcol = 0;
        if (++X.nvec > S.mvec) {                   // :339-347
          if (S.last != k) S.error = 1;
          S.next[S.last] = S.free_;
          S.free_ = k;
          S.last = S.prev[k];
          S.next[S.last] = NKA_NIL;
          X.removed[k] = 1;
          if (pos < X.jr) X.jr = pos;
          S.evicted_last = 1;
          X.stop = 1;
        } else if (pos == missing) {               // a drop made room: the skipped column matters
          S.need_fixup = 1;
          X.ret = 1;
        } else {
          const double hkk = X.hkk[k];
          if (hkk - tol2 < S.min_margin) S.min_margin = hkk - tol2;
          if (hkk > tol2) {                        // :362-363
            NKA_H(S, k, k) = NKA_SQRT(hkk);
            X.act[X.na++] = k;
            X.newcol = 1;
          } else {                                 // :364-379
            const int pk = S.prev[k], nk = S.next[k];
            S.next[pk] = nk;
            if (nk == NKA_NIL) S.last = pk;
            else S.prev[nk] = pk;
            S.next[k] = S.free_;
            S.free_ = k;
            X.removed[k] = 1;
            if (pos < X.jr) X.jr = pos;
            --X.nvec;
            ++S.ndrop_last;
          }
        }
      }
      par.sync();
      if (X.ret) return 1;
      if (X.stop) break;
      if (X.newcol && pos + 1 < L) nka_cholesky_column(S, X, X.na - 1, pos + 1, L, missing, par);
      par.sync();
    }
    if (lead) {
      S.subspace = 1;
      S.pending = 0;
    }
    par.sync();
  }

  if (lead) {
    // Pairs that lost their newer neighbour are converted before pass B overwrites anything.
    if (X.jr < L) {
      bool rem[NKA_MAXSLOT];
      for (int k = 0; k < NKA_MAXSLOT; ++k) rem[k] = X.removed[k] != 0;
      nka_plan_materialise(S, X.ord, L, X.jr, rem);
    }
    // Step C: storage for the new vectors.  src-C/...c:391-394
    if (S.free_ == NKA_NIL) { S.error = 2; X.ret = 2; }
    else {
      X.nw = S.free_;
      S.free_ = S.next[X.nw];
      X.nl = 0;
      for (int j = S.first; j != NKA_NIL; j = S.next[j]) X.lst[X.nl++] = j;   // the list after the update, newest first
    }
  }
  par.sync();
  if (X.ret) return 0;

  // Step D: projection.  src-C/...c:400-417
  NkaPlanB& B = S.planB;
  const int nl = X.nl;
  if (S.subspace) {
    for (int j = par.lane(); j < L; j += par.nlanes())
      if (!X.removed[X.ord[j]]) X.rhs[X.ord[j]] = fd[j] / S.s[X.ord[j]];      // <f, w_k>
    par.sync();
    // forward substitution (:405-411): c_j = (rhs_j - sum_{i before j} h(j,i) c_i) / h(j,j), the
    // subtractions applied to every later row as soon as c_i is known, i.e. in the reference's order
    for (int jj = 0; jj < nl; ++jj) {
      const int j = X.lst[jj];
      if (lead) S.c[j] = X.rhs[j] / NKA_H(S, j, j);
      par.sync();
      const double cj = S.c[j];
      for (int q = jj + 1 + par.lane(); q < nl; q += par.nlanes()) X.rhs[X.lst[q]] -= NKA_H(S, X.lst[q], j) * cj;
      par.sync();
    }
    // backward substitution (:412-417): c_j = (c_j - sum_{i after j, from the last} h(i,j) c_i) / h(j,j)
    for (int jj = nl - 1; jj >= 0; --jj) {
      const int j = X.lst[jj];
      if (lead) S.c[j] = S.c[j] / NKA_H(S, j, j);
      par.sync();
      const double cj = S.c[j];
      for (int q = jj - 1 - par.lane(); q >= 0; q -= par.nlanes()) S.c[X.lst[q]] -= NKA_H(S, j, X.lst[q]) * cj;
      par.sync();
    }
  }
  // Pass B streams the Z column of every pair that was on the list at entry: with the old
  // coefficient it rebuilds the pending correction Y, with the new one (0 if the pair left the
  // list) it forms the correction  f += sum_k c_k (v_k - w_k) = sum_k (c_k / s_k) Z[k]  (:419-424).
  const int z0 = X.pending_at_entry ? 1 : 0;
  for (int j = z0 + par.lane(); j < L; j += par.nlanes()) {
    const int k = X.ord[j];
    B.zcol[j - z0] = k;
    B.coefY[j - z0] = S.coefY[k];
    B.coefN[j - z0] = (S.subspace && !X.removed[k]) ? S.c[k] / S.s[k] : 0.0;
  }
  par.sync();
  if (lead) {
    const int nw = X.nw;
    B.newslot = nw;
    B.has_pair = X.has_pair;
    B.pslot = X.has_pair ? S.first : 0;
    B.nz = L > z0 ? L - z0 : 0;
    B.write_f = S.subspace ? 1 : 0;
    B.coef_p = X.has_pair ? S.c[S.first] / S.s[S.first] : 0.0;
    // remember how this call's correction y was assembled; the next pass B recomputes it
    for (int k = 0; k <= S.mvec; ++k) S.coefY[k] = 0.0;
    for (int i = 0; i < B.nz; ++i) S.coefY[B.zcol[i]] = B.coefN[i];
    if (X.has_pair) S.coefY[S.first] = B.coef_p;

    // Step E: push the new slot, mark pending.  src-C/...c:432-443
    S.prev[nw] = NKA_NIL;
    S.next[nw] = S.first;
    if (S.first == NKA_NIL) S.last = nw;
    else S.prev[S.first] = nw;
    S.first = nw;
    S.pending = 1;
    S.chained[nw] = 0;
    ++S.ncalls;

    nka_build_plan_a(S);
  }
  par.sync();
  return 0;
}

// Structural invariant check, src-F08/nka_type.F90:460-524 (0-based).
NKA_HD int nka_state_defined(const NkaDevState& S)
{
  if (S.mvec < 1 || S.mvec + 1 > NKA_MAXSLOT) return 0;
  if (!(S.vtol > 0.0)) return 0;
  if (S.error) return 0;
  const int n = S.mvec + 1;
  for (int k = 0; k < n; ++k)
    if (S.next[k] < NKA_NIL || S.next[k] >= n) return 0;
  if (S.first < NKA_NIL || S.first >= n) return 0;
  if (S.free_ < NKA_NIL || S.free_ >= n) return 0;
  bool tag[NKA_MAXSLOT];
  for (int k = 0; k < n; ++k) tag[k] = false;
  if (S.first == NKA_NIL) {
    if (S.last != NKA_NIL) return 0;
  } else {
    int k = S.first;
    if (S.prev[k] != NKA_NIL) return 0;
    tag[k] = true;
    while (S.next[k] != NKA_NIL) {
      if (S.prev[S.next[k]] != k) return 0;
      k = S.next[k];
      if (tag[k]) return 0;
      tag[k] = true;
    }
    if (S.last != k) return 0;
  }
  for (int k = S.free_; k != NKA_NIL; k = S.next[k]) {
    if (tag[k]) return 0;
    tag[k] = true;
  }
  for (int k = 0; k < n; ++k)
    if (!tag[k]) return 0;
  return 1;
}
