// nka_ssor3.cuh -- ex_ssor_sweep3<DIR>: the exact-order SSOR sweep with TWO independent dependent
// chains per lane and compact (instruction-cache resident) loops.  AN EXPERIMENT, selectable with
// NKA_SSOR_KERNEL=3, bit-identical, but SLOWER than ex_ssor_sweep2 (pc_ssor 6.86 vs 4.60 ms at
// 4096^2): measurements and the ncu stall breakdown in profiles/r2p_ssor3_two_chain_experiment.txt.
// Included by nka_example.cu after nka_ssor2.cuh (shares its division split, diagonal walk,
// predicated stores, mailbox helpers).  src-F08/nka_example.F90:159-175.
//
// Why two chains: ex_ssor_sweep2's consumer warp spends ~230 cycles on a step whose dependent chain
// is ~110 (shuffle + ten fp64 operations): one in-order warp with ONE chain cannot issue its ~55
// other instructions inside the chain's latency shadows.  Here a lane owns EX3_C = 2 adjacent grid
// columns.  On anti-diagonal t the lane's two cells (j, t-j) and (j+1, t-j-1) are independent of
// each other -- both depend only on values of diagonal t-1: the head cell on the adjacent lane's tail
// cell (one shuffle) and its own previous value, the tail cell on the lane's own two previous
// values -- so every step carries two chains that fill each other's latency, a strip is 64 columns
// wide (half as many strip-to-strip hand-overs), and the number of steps per sweep is unchanged
// (one per anti-diagonal: nx + ny - 1).
//
// Why compact loops: the first version of this kernel unrolled every role's loop over a block of 8
// steps, as ex_ssor_sweep2 does.  With twice the work per step the consumer's loop body was 14 KB
// and the six warps' loops together exceeded the SM's 32 KB L1.5 instruction cache: ncu attributed
// 53 % of the consumer's stall samples to instruction fetch (no_inst) and the sweep took 2.4 ms
// instead of 1.15.  Now the consumer's loop is two steps (register ping-pong, 4.3 KB: fits the ~6 KB
// L0), block hand-overs are predicated bar instructions instead of unrolled code, and the producers
// work in chunks of four steps: 1.7 ms per sweep.
//
// Why it still loses: the chains do fill each other's latency, but the step now has ~109 hot
// instructions (57 in ex_ssor_sweep2) and a lone in-order warp pays about one issue cycle plus one
// dependency cycle for each; with the shuffle and two taken branches the step is ~395 cycles where
// break-even against ex_ssor_sweep2 (230 cycles, twice as many strips) is 254.
//
// A CTA is six warps:
//   warp 0     consumer   both chains, the results' stores, the edge-channel store
//   warp 4     receiver   upstream strip's edge channel -> mailbox (shares warp 0's sub-partition: it
//                         only polls)
//   warps 1,5  copiers    one per sub-column: operands used unchanged, store offsets, 1/ac refinement
//   warps 2,3  cookers    one per sub-column: products of old values
// Ring, named-barrier hand-over and operand conventions are ex_ssor_sweep2's; the ring stores the two
// sub-columns of a field side by side, so the consumer fetches both with one LDS.128.
// Arithmetic per cell is unchanged (same operations, same order, same operands): bit-identical to
// the serial loops, tests/test_gpu_example.py runs every SSOR test with all three kernels.

#define EX3_C 2
#define EX3_W (32 * EX3_C)
#ifndef EX3_BLK
#define EX3_BLK 8                 // steps per hand-over block (even)
#endif
#ifndef EX3_NBLK
#define EX3_NBLK 4                // blocks in the consumer's ring (power of two)
#endif
#ifndef EX3_CH
#define EX3_CH 4                  // steps per unrolled chunk of the producers' loops
#endif
#define EX3_SLOTS (EX3_BLK * EX3_NBLK)
#define EX3_STAGES (4 * EX3_BLK)
#define EX3_THREADS 192
#define EX3_STEP_DOUBLES (CK_NF * 32 * EX3_C)
#define EX3_CK_DOUBLES (EX3_SLOTS * EX3_STEP_DOUBLES)
#define EX3_RAW_DOUBLES (EX3_STAGES * RW_NF * 32)          // per cooker
#define EX3_SMEM_BYTES ((EX3_CK_DOUBLES + EX3_C * EX3_RAW_DOUBLES) * 8 + EX3_SLOTS * 8 + (EX_MBOX + 2) * 8)
#define EX3_NPART (32 * (1 + 2 * EX3_C))                   // consumer + copiers + cookers
static_assert(EX3_BLK % 2 == 0 && EX3_BLK % EX3_CH == 0, "block geometry");
static_assert((EX3_NBLK & (EX3_NBLK - 1)) == 0 && (EX3_SLOTS & (EX3_SLOTS - 1)) == 0, "ring geometry: powers of two");

#define EX3_BAR_FULL(b) (1 + (b))
#define EX3_BAR_EMPTY(b) (1 + EX3_NBLK + (b))
__device__ __forceinline__ void ex3_bar_sync(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(EX3_NPART) : "memory"); }
__device__ __forceinline__ void ex3_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "n"(EX3_NPART) : "memory"); }
// warp-uniform predicates: the hand-over costs the consumer's loop an instruction, not a branch
__device__ __forceinline__ void ex3_bar_sync_if(int id, bool on)
{
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q bar.sync %0, %1; }" :: "r"(id), "n"(EX3_NPART), "r"((int)on) : "memory");
}
__device__ __forceinline__ void ex3_bar_arrive_if(int id, bool on)
{
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q bar.arrive %0, %1; }" :: "r"(id), "n"(EX3_NPART), "r"((int)on) : "memory");
}

// the division's slow path, out of line: keeps the cold block of the step a few instructions long
__device__ __noinline__ double ex3_div_slow(double x, double b) { return __ddiv_rn(x, b); }

struct Ex3Strip { int j0, jlast, nsteps, nblocks, t_first; };
template <int DIR>
__device__ __forceinline__ Ex3Strip ex3_strip(const SsorParams& P, int strip)
{
  Ex3Strip g;
  g.j0 = strip * EX3_W;
  g.jlast = g.j0 + EX3_W - 1 < P.nx - 1 ? g.j0 + EX3_W - 1 : P.nx - 1;
  g.nsteps = (g.jlast - g.j0) + P.ny;
  g.nblocks = (g.nsteps + EX3_BLK - 1) / EX3_BLK;
  g.t_first = DIR > 0 ? g.j0 : g.jlast + P.ny - 1;
  return g;
}

// ring addressing: [slot][field][lane][sub-column]
__device__ __forceinline__ int ex3_ck_index(int slot, int lane, int c) { return slot * EX3_STEP_DOUBLES + lane * EX3_C + c; }
#define EX3_FIELD (32 * EX3_C)          // doubles between consecutive fields of a slot

// ---------------------------------------------------------------------------
// copier of sub-column c: operands used unchanged, straight into the consumer's ring
// ---------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void ex3_copier(const SsorParams& P, double* ck_ptr, long long* sbase, const int strip,
                                           const int lane, const int c)
{
  const Ex3Strip g = ex3_strip<DIR>(P, strip);
  const uint32_t ck = (uint32_t)__cvta_generic_to_shared(ck_ptr);
  const int nx = P.nx, ny = P.ny, j = g.j0 + EX3_C * lane + c;
  const bool jvalid = j < nx;
  const bool inner = j + 1 < nx;
  Ex2Walk<DIR> w;
  w.start(g.t_first, nx, ny);
  auto add_reciprocals = [&](int bb) {
#pragma unroll
    for (int u = 0; u < EX3_BLK; ++u) {
      double* q = ck_ptr + ex3_ck_index(bb * EX3_BLK + u, lane, c);
      q[CK_Y * EX3_FIELD] = ex2_rcp(q[CK_AC * EX3_FIELD]);
    }
  };
  for (int i = 0; i < g.nblocks; ++i) {
    const int b = i % EX3_NBLK;
    if (i >= EX3_NBLK) ex3_bar_sync(EX3_BAR_EMPTY(b));             // the consumer has finished block i - EX3_NBLK
#pragma unroll 1
    for (int h = 0; h < EX3_BLK; h += EX3_CH) {
#pragma unroll
      for (int u = 0; u < EX3_CH; ++u) {
        const int slot = b * EX3_BLK + h + u;
        const int k = w.t - j;
        const bool on = jvalid && k >= 0 && k < ny;
        const long long cell = on ? w.b_at + j : 0;
        const uint32_t dst = ck + (uint32_t)(ex3_ck_index(slot, lane, c) * 8);
        ss_cp8(dst + CK_AC * EX3_FIELD * 8, on ? P.AC + cell : &ex2_one, true);
        if (DIR > 0) {
          ss_cp8(dst + CK_A * EX3_FIELD * 8, on ? P.R + cell : &ex2_one, true);
          ss_cp8(dst + CK_B * EX3_FIELD * 8, P.AXL + cell, on);
          ss_cp8(dst + CK_AY * EX3_FIELD * 8, P.AYD + cell, on);
        } else {
          // right face: the left face of cell (j+1,k) on diagonal t+1, or the boundary face of row k
          ss_cp8(dst + CK_B * EX3_FIELD * 8, (on && inner) ? P.AXL + w.b_plus() + j + 1 : P.AXR + (on ? k : 0), on);
          // upper face: the lower face of cell (j,k+1) on diagonal t+1, or the top face of column j
          ss_cp8(dst + CK_AY * EX3_FIELD * 8, (on && k + 1 < ny) ? P.AYD + w.b_plus() + j : P.AYT + (on ? j : 0), on);
        }
        if (c == 0 && lane == 0) sbase[slot] = w.b_at * 8;          // byte offset of this step's diagonal in a grid function
        w.advance();
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (i >= 1) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");         // block i-1 has landed
      add_reciprocals((i - 1) % EX3_NBLK);
      ex3_bar_arrive(EX3_BAR_FULL((i - 1) % EX3_NBLK));
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  add_reciprocals((g.nblocks - 1) % EX3_NBLK);
  ex3_bar_arrive(EX3_BAR_FULL((g.nblocks - 1) % EX3_NBLK));
  // pair off the consumer's last arrivals so every barrier is idle when the next strip starts
  for (int i = g.nblocks > EX3_NBLK ? g.nblocks : EX3_NBLK; i < g.nblocks + EX3_NBLK; ++i) ex3_bar_sync(EX3_BAR_EMPTY(i % EX3_NBLK));
}

// ---------------------------------------------------------------------------
// cooker of sub-column c: products of old values (raw fields as in nka_ssor2.cuh)
// ---------------------------------------------------------------------------
template <int DIR>
__device__ __forceinline__ void ex3_cooker(const SsorParams& P, double* ck, double* raw_ptr, const int strip, const int lane, const int c)
{
  static_assert((EX3_STAGES & (EX3_STAGES - 1)) == 0, "raw ring: power of two");
  static_assert(EX3_STAGES >= 3 * EX3_BLK + 1, "raw ring: three groups and a step in flight");
  const Ex3Strip g = ex3_strip<DIR>(P, strip);
  const uint32_t raw = (uint32_t)__cvta_generic_to_shared(raw_ptr);
  const int nx = P.nx, ny = P.ny, j = g.j0 + EX3_C * lane + c;
  const bool jvalid = j < nx;
  const double ayt = jvalid ? __ldg(P.AYT + j) : 0.0;
  const double om1 = P.om1;
  double zold_edge = 0.0;
  if (jvalid) {
    const uint4* beyond = DIR > 0 ? P.halo_hi : P.halo_lo;
    if (beyond && !P.zero_old) zold_edge = tagged_wait(beyond + j, P.tag_prev, P.err, P.spin_limit);
  }
  __syncwarp();
  // Copies are committed in groups of EX3_BLK steps, shifted by one step (group g = steps
  // g*BLK+1 .. g*BLK+BLK, step 0 rides with group 0): forming block i needs steps i*BLK .. i*BLK+BLK
  // (each cell also reads the cell one row further), i.e. groups <= i.
  Ex2Walk<DIR> w;                     // at the step being copied
  w.start(g.t_first, nx, ny);
  int sp = 0;
  auto issue_one = [&]() {
    ex2_raw_issue<DIR>(P, raw, sp & (EX3_STAGES - 1), w, j, jvalid, lane);
    ++sp;
    w.advance();
  };
  auto issue_group = [&]() {
#pragma unroll 1
    for (int h = 0; h < EX3_BLK; h += EX3_CH) {
#pragma unroll
      for (int u = 0; u < EX3_CH; ++u) issue_one();
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_one();
  issue_group();                      // group 0
  issue_group();                      // group 1
  int k = g.t_first - j;              // this lane's row at the step being formed
  for (int i = 0; i < g.nblocks; ++i) {
    const int b = i % EX3_NBLK;
    issue_group();                    // group i + 2
    asm volatile("cp.async.wait_group 2;" ::: "memory");            // groups <= i have landed
    if (i >= EX3_NBLK) ex3_bar_sync(EX3_BAR_EMPTY(b));
#pragma unroll 1
    for (int h = 0; h < EX3_BLK; h += EX3_CH) {
      Ex2Raw e[EX3_CH + 1];
#pragma unroll
      for (int u = 0; u <= EX3_CH; ++u) e[u] = ex2_raw_fetch<DIR>(raw_ptr, (i * EX3_BLK + h + u) & (EX3_STAGES - 1), lane);
#pragma unroll
      for (int u = 0; u < EX3_CH; ++u) {
        const Ex2Raw e1 = e[u], e2 = e[u + 1];                       // e2: the cell one row further in travel direction
        const bool on = jvalid && k >= 0 && k < ny;
        const bool edge = DIR > 0 ? k + 1 >= ny : k <= 0;            // the next row lies outside this slab
        double* dst = ck + ex3_ck_index(b * EX3_BLK + h + u, lane, c);
        dst[CK_PO * EX3_FIELD] = __dmul_rn(om1, e1.zo);              // (1-w) * z_old
        if (DIR > 0) {
          dst[CK_P0 * EX3_FIELD] = __dmul_rn(e1.f0, e1.zs);          // axr * old right
          const double p1 = __dmul_rn(edge ? ayt : e2.ayd, edge ? zold_edge : e2.zo);    // ayu * old upper
          dst[CK_P1 * EX3_FIELD] = on ? p1 : 0.0;
        } else {
          const double a = __dadd_rn(e1.f0, __dmul_rn(e1.axl, e1.zs));   // r + axl * old left
          dst[CK_A * EX3_FIELD] = on ? a : 1.0;
          const double p0 = __dmul_rn(e1.ayd, edge ? zold_edge : e2.zo);                 // ayd * old lower
          dst[CK_P0 * EX3_FIELD] = on ? p0 : 0.0;
        }
        k += DIR;
      }
    }
    ex3_bar_arrive(EX3_BAR_FULL(b));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  for (int i = g.nblocks > EX3_NBLK ? g.nblocks : EX3_NBLK; i < g.nblocks + EX3_NBLK; ++i) ex3_bar_sync(EX3_BAR_EMPTY(i % EX3_NBLK));
}

// ---------------------------------------------------------------------------
// warp 0: the two chains
// ---------------------------------------------------------------------------
struct Ex3Ops { double a[EX3_C], b[EX3_C], p0[EX3_C], p1[EX3_C], po[EX3_C], ay[EX3_C], ac[EX3_C], y[EX3_C]; long long sb; };

template <int DIR>
__device__ __forceinline__ void ex3_ops_load(Ex3Ops& o, const double* ck_lane, const long long* sbase, int slot)
{
  static_assert(EX3_C == 2, "one 16-byte load per field");
  const double2* p = reinterpret_cast<const double2*>(ck_lane + slot * EX3_STEP_DOUBLES);
  auto ld = [&](int field, double* out) { const double2 v = p[field * 32]; out[0] = v.x; out[1] = v.y; };
  ld(CK_A, o.a); ld(CK_B, o.b); ld(CK_P0, o.p0);
  if (DIR > 0) ld(CK_P1, o.p1); else { o.p1[0] = 0.0; o.p1[1] = 0.0; }
  ld(CK_PO, o.po); ld(CK_AY, o.ay); ld(CK_AC, o.ac); ld(CK_Y, o.y);
  o.sb = sbase[slot];
}

template <bool V> struct Ex3Tag { static constexpr bool value = V; };

template <int DIR, bool TRACE, bool SLAB>
__device__ __forceinline__ void ex3_consumer(const SsorParams& P, const double* ck, const long long* sbase,
                                             unsigned long long* mbox, const int strip, const int lane)
{
  constexpr int C = EX3_C;
  constexpr int CH = DIR > 0 ? 0 : C - 1;        // head sub-column: first in travel order, its upstream neighbour is in the adjacent lane / strip
  constexpr int CT = DIR > 0 ? C - 1 : 0;        // tail sub-column: its value goes to the adjacent lane / strip
  const Ex3Strip g = ex3_strip<DIR>(P, strip);
  const int ny = P.ny, jb = g.j0 + C * lane;
  const double omega = P.omega;
  bool jv[C];
#pragma unroll
  for (int c = 0; c < C; ++c) jv[c] = jb + c < P.nx;
  // forward: lane 31's tail column hands z(j,k) to lane 0 of the next strip; backward: lane 0's to lane 31 of the previous one
  const bool is_prod = jv[CT] && (DIR > 0 ? (lane == 31 && jb + CT + 1 < P.nx) : (lane == 0 && strip > 0));
  const bool is_cons = jv[CH] && (DIR > 0 ? (lane == 0 && strip > 0) : (lane == 31 && jb + CH + 1 < P.nx));
  const bool first_lane = DIR > 0 ? lane == 0 : lane == 31;
  unsigned long long* const cout = P.bnd + (size_t)strip * ny;
  uint32_t mbox0 = (uint32_t)__cvta_generic_to_shared(mbox);
  unsigned ny_act[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    ny_act[c] = jv[c] ? (unsigned)ny : 0u;                           // rows this lane owns in column c (none: beyond nx)
    asm volatile("mov.u32 %0, %0;" : "+r"(ny_act[c]));
  }
  unsigned ny_cons = is_cons ? (unsigned)ny : 0u;                    // steps in which this lane reads the mailbox
  unsigned long long sent = EX_SENT;
  asm volatile("mov.u32 %0, %0;" : "+r"(mbox0));
  asm volatile("mov.u32 %0, %0;" : "+r"(ny_cons));
  asm volatile("mov.u64 %0, %0;" : "+l"(sent));
  const uint32_t zero_slot = mbox0 + EX_MBOX * 8;
  uint4* const send_to = SLAB ? (DIR > 0 ? P.peer_up_lo : P.peer_dn_hi) : nullptr;
  uint4* const send_slot = send_to ? send_to + jb : nullptr;
  const int k_send = send_to ? (DIR > 0 ? ny - 1 : 0) : -1;
  char* const zcol = reinterpret_cast<char*>(P.Z + jb);
  const double* const ck_lane = ck + lane * C;
  if (TRACE && lane == 0) P.trace[strip * 4 + 0] = ex_globaltimer();

  // own results of the previous step, per sub-column: z(j, k-DIR), new; before the first row the row the
  // neighbouring rank has just computed, or the boundary value 0
  double znew[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    znew[c] = 0.0;
    if (SLAB && jv[c]) {
      const uint4* from = DIR > 0 ? P.halo_lo : P.halo_hi;
      if (from) znew[c] = tagged_wait(from + jb + c, P.tag_cur, P.err, P.spin_limit);
    }
  }
  uint32_t mslot = is_cons ? mbox0 : zero_slot;
  unsigned long long ext = ex2_mbox_ld(mslot);
  if (ext == EX_SENT) ext = ex2_mbox_wait(mslot, P.err);
  int k[C];
#pragma unroll
  for (int c = 0; c < C; ++c) k[c] = g.t_first - (jb + c);
  int s = 0;
  const int nsteps_ring = g.nblocks * EX3_BLK;                       // steps the producers fill (whole blocks)
  unsigned long long* cptr = cout + k[CT];                           // this step's word of the edge channel
  double pend_z[C]; bool pend_act[C]; int pend_k[C]; long long pend_sb = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) { pend_z[c] = 0.0; pend_act[c] = false; pend_k[c] = 0; }

  // One step.  `o`: this step's operands (loaded during the previous step), `on_`: the next step's,
  // loaded here.  BOUNDARY: this step may be the last of its hand-over block (odd steps only, EX3_BLK
  // is even): the next block is waited for before its first operands are read, a step early, and
  // this block is released at the end.
  auto step = [&](auto boundary, const Ex3Ops& o, Ex3Ops& on_) {
    constexpr bool BOUNDARY = decltype(boundary)::value;
    double zh[C];
    zh[CH] = DIR > 0 ? __shfl_up_sync(0xffffffffu, znew[CT], 1) : __shfl_down_sync(0xffffffffu, znew[CT], 1);
    if (first_lane) zh[CH] = __longlong_as_double((long long)ext);
#pragma unroll
    for (int c = 0; c < C; ++c) if (c != CH) zh[c] = znew[c - DIR];
    // ---- off the chains ----
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ex2_st_f64(reinterpret_cast<double*>(zcol + pend_sb) + c, pend_z[c], pend_act[c]);
      if (SLAB) ex2_st_tagged(send_slot + c, pend_z[c], P.tag_cur, pend_act[c] && pend_k[c] == k_send);
    }
    bool last = false;
    if (BOUNDARY) {
      last = (s & (EX3_BLK - 1)) == EX3_BLK - 1;
      ex3_bar_sync_if(EX3_BAR_FULL(((s + 1) / EX3_BLK) & (EX3_NBLK - 1)), last && s + 1 < nsteps_ring);
    }
    ex3_ops_load<DIR>(on_, ck_lane, sbase, (s + 1) & (EX3_SLOTS - 1));   // (past the last block: stale values, unused)
    ex2_mbox_free(mslot, sent, (unsigned)s < ny_cons);
    mslot = ((unsigned)(s + 1) < ny_cons) ? mbox0 + ((s + 1) & (EX_MBOX - 1)) * 8 : zero_slot;
    const unsigned long long ext_next = ex2_mbox_ld(mslot);
    bool act[C], safe_b[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { act[c] = (unsigned)k[c] < ny_act[c]; safe_b[c] = ex2_div_safe(o.ac[c]); }
    // ---- the chains.  src-F08/nka_example.F90:163-165 (= :171-173): the reference's operation order, no fma
    //   z = (1-w) z + w (r + axl z(j-1,k) + axr z(j+1,k) + ayd z(j,k-1) + ayu z(j,k+1)) / ac
    double x[C], q[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double sm = __dadd_rn(o.a[c], __dmul_rn(o.b[c], zh[c]));
      sm = __dadd_rn(sm, o.p0[c]);
      sm = __dadd_rn(sm, __dmul_rn(o.ay[c], znew[c]));
      if (DIR > 0) sm = __dadd_rn(sm, o.p1[c]);
      x[c] = __dmul_rn(omega, sm);
      q[c] = ex2_div_fast(x[c], o.ac[c], o.y[c]);
    }
    ext = ext_next;
    bool unsafe_any = false, unsafe[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { unsafe[c] = !(safe_b[c] && ex2_div_safe(x[c])); unsafe_any = unsafe_any || unsafe[c]; }
    if (__builtin_expect(unsafe_any || ext_next == sent, 0)) {       // rare, one branch for all of it
#pragma unroll
      for (int c = 0; c < C; ++c) if (unsafe[c]) q[c] = ex3_div_slow(x[c], o.ac[c]);
      if (ext_next == sent) ext = ex2_mbox_wait_counted(mslot, P.err, TRACE ? P.trace + strip * 4 + 3 : nullptr);
    }
    double zc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) zc[c] = __dadd_rn(o.po[c], q[c]);
    ex2_st_ch(cptr, zc[CT], act[CT] && is_prod);                     // the downstream strip is waiting for this one: not deferred
#pragma unroll
    for (int c = 0; c < C; ++c) {
      znew[c] = act[c] ? zc[c] : znew[c];
      pend_z[c] = zc[c]; pend_act[c] = act[c]; pend_k[c] = k[c];
      k[c] += DIR;
    }
    pend_sb = o.sb;
    if (TRACE) {
      if (s == 0 && is_cons) P.trace[strip * 4 + 1] = ex_globaltimer();
      if (strip == P.nstrips / 2 && lane == 0 && s < 256) P.trace[P.nstrips * 4 + s] = ex_globaltimer();
    }
    if (BOUNDARY) ex3_bar_arrive_if(EX3_BAR_EMPTY((s / EX3_BLK) & (EX3_NBLK - 1)), last);
    cptr += DIR;
    ++s;
  };

  ex3_bar_sync(EX3_BAR_FULL(0));
  Ex3Ops oa, ob;
  ex3_ops_load<DIR>(oa, ck_lane, sbase, 0);
#pragma unroll 1
  for (int v = 0; v < nsteps_ring; v += 2) {
    step(Ex3Tag<false>(), oa, ob);
    step(Ex3Tag<true>(), ob, oa);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    ex2_st_f64(reinterpret_cast<double*>(zcol + pend_sb) + c, pend_z[c], pend_act[c]);
    if (SLAB) ex2_st_tagged(send_slot + c, pend_z[c], P.tag_cur, pend_act[c] && pend_k[c] == k_send);
  }
  if (TRACE && lane == 0) P.trace[strip * 4 + 2] = ex_globaltimer();
}

template <int DIR, bool TRACE, bool SLAB>
__global__ void __launch_bounds__(EX3_THREADS) ex_ssor_sweep3(SsorParams P)
{
  extern __shared__ __align__(16) unsigned char ex3_smem[];
  double* ck = reinterpret_cast<double*>(ex3_smem);                                // [EX3_SLOTS][CK_NF][32][EX3_C]
  double* raw = ck + EX3_CK_DOUBLES;                                               // [EX3_C][EX3_STAGES][RW_NF][32]
  long long* sbase = reinterpret_cast<long long*>(raw + EX3_C * EX3_RAW_DOUBLES);  // [EX3_SLOTS]
  unsigned long long* mbox = reinterpret_cast<unsigned long long*>(sbase + EX3_SLOTS);   // [EX_MBOX] + the zero slot
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // strips in dependency order, every CTA of the grid resident (as ex_ssor_sweep2)
  for (int i = blockIdx.x; i < P.nstrips; i += gridDim.x) {
    const int strip = DIR > 0 ? i : P.nstrips - 1 - i;
    const bool has_upstream = i > 0;
    if (threadIdx.x < EX_MBOX) mbox[threadIdx.x] = EX_SENT;
    if (threadIdx.x == EX_MBOX) mbox[EX_MBOX] = 0ull;
    __syncthreads();
    if (warp == 0) ex3_consumer<DIR, TRACE, SLAB>(P, ck, sbase, mbox, strip, lane);
    else if (warp == 4) {
      if (has_upstream) ssor_receiver<DIR>(P, mbox, P.bnd + (size_t)(DIR > 0 ? strip - 1 : strip + 1) * P.ny, lane);
    } else if (warp == 1 || warp == 5) ex3_copier<DIR>(P, ck, sbase, strip, lane, warp == 1 ? 0 : 1);
    else ex3_cooker<DIR>(P, ck, raw + (warp == 2 ? 0 : EX3_RAW_DOUBLES), strip, lane, warp == 2 ? 0 : 1);
    __syncthreads();
  }
}
