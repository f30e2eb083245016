// frontier_quad.cu — packed-lanes variant of the frontier pass: one warp evaluates
// R = 32 / Wq StoryRuns per trip (Wq = words per run rounded up to a power of two; R = 4 at
// S = 256, 16 at S = 64, 1 at S = 1024).
//
// Lane l = g * Wq + w holds word w (32 steps) of run g of the trip, so every bit-plane stage
// (sync rewrite, buildStateMaps classification, fail-fast / compensation marking, group
// selection, cond masking, result words) runs on all 32 lanes and its instructions are
// shared by R runs; per-run reductions use sub-warp redux / ballots.  The R state records
// of a trip are adjacent in HBM and arrive with ONE bulk copy; lanes 0..R-1 are the
// producers of the R topology records (each runs its own slot-id -> slot-entry prefetch
// chain and TMA copy; all arrive on the stage's mbarrier).  The dependency walk (one step
// per lane over a CSR row) is inherently per run and visits the candidate words of all R
// runs one after another.
//
// Shared memory is addressed through 32-bit shared-window addresses throughout (ld.shared / st.shared, no generic
// pointers); the walk reads a 16-byte per-group entry {col_idx base, row_ptr base, status base, max_deg} written by
// the group's first lane, so an item (run, word) costs one extra ld.shared over the one-run-per-warp walk.
//
// Used for single-pass evaluation of batches whose topologies have no `parallel` steps;
// the general kernel (frontier_kernel.cu) covers fixpoint mode and parallel joins.
// Semantics are identical (same stages, same citations); parity tests run both.
#include "kernel_common.cuh"

namespace bf {

extern __shared__ __align__(128) uint8_t smem_q[];

template <bool FIXUP>
DI void walk_quad(uint32_t lane, uint32_t CAND, uint32_t wq_log2, uint32_t col_off, uint32_t meta, uint32_t tr0_off,
                  uint32_t topo_buf, uint32_t st0_off, const uint32_t* mFAIL, uint32_t fclass, uint32_t& met_w,
                  uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  const uint32_t wmask = (1u << wq_log2) - 1u;
  uint32_t todo = __ballot_sync(FULL, CAND != 0);  // (run, word) pairs with at least one candidate step
  while (todo) {
    const uint32_t L = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t g = L >> wq_log2, j = L & wmask;
    const uint32_t candw = __shfl_sync(FULL, CAND, L);
    const uint32_t colo = __shfl_sync(FULL, col_off, L);
    const uint32_t mt = __shfl_sync(FULL, meta, L);  // zidx | max_deg << 16
    const uint32_t zidx = mt & 0xFFFFu, max_deg = mt >> 16;
    const uint16_t* row_ptr = reinterpret_cast<const uint16_t*>(smem_q + tr0_off + g * topo_buf + sizeof(TopoHeader));
    const uint8_t* st = smem_q + st0_off + (g << (5 + wq_log2));
    const bool cand = (candw >> lane) & 1u;
    const uint32_t i = j * 32 + lane;
    uint32_t e0 = 0, n = 0;
    if (cand) {
      e0 = row_ptr[i];
      n = row_ptr[i + 1] - e0;
    }
    const uint16_t* cp = reinterpret_cast<const uint16_t*>(smem_q + colo) + e0;
    bool unmet, fdp;
    if (!FIXUP) {
      const uint32_t x0 = cp[0], x1 = cp[1], x2 = cp[2], x3 = cp[3];  // may run past the row: masked below
      const uint32_t s0 = st[min(x0, zidx)], s1 = st[min(x1, zidx)], s2 = st[min(x2, zidx)], s3 = st[min(x3, zidx)];
      uint32_t w = ((s3 * 256u + s2) * 256u + s1) * 256u + s0;
      w &= bmsk_clamp(0u, n * 8u);
      if (max_deg > 4) {  // warp-uniform
        for (uint32_t e = 4; e < n; ++e) w |= st[cp[e]];
      }
      unmet = (w & 0x01010101u) != 0;
      fdp = (w & 0x02020202u) != 0;
    } else {
      const uint32_t fc = __shfl_sync(FULL, fclass, L);
      const uint32_t* mf = mFAIL + (g << wq_log2);
      uint32_t acc = 0;
      for (uint32_t e = 0; e < n; ++e) {
        const uint32_t d = cp[e];
        uint32_t sb = st[d];
        if (d < i && ((mf[d >> 5] >> (d & 31u)) & 1u)) sb = fc;
        acc |= sb;
      }
      unmet = (acc & 1u) != 0;
      fdp = (acc & 2u) != 0;
    }
    const uint32_t fdb = __ballot_sync(FULL, fdp);
    const uint32_t metb = __ballot_sync(FULL, cand && !unmet);
    if (lane == L) {
      fd_w = fdb;
      met_w = metb;
    }
  }
}

// Lean walk over the (run, word) items of a trip: one step per lane, first four deps branch-free and clamp-free
// (col_idx carries zero padding, device_record.h).  tab_a: per-group entries {col_a, rp_a, st_a, max_deg}.
template <bool NEED_FD, bool LONG_ROWS>
DI void walk_items(uint32_t lane, uint32_t CAND, uint32_t lg, uint32_t tab_a, uint32_t& met_w, uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  const uint32_t wmask = (1u << lg) - 1u;
  const uint32_t lane2 = lane * 2u;
  uint32_t todo = __ballot_sync(FULL, CAND != 0);  // (run, word) pairs with at least one candidate step
  while (todo) {
    const uint32_t L = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t candw = __shfl_sync(FULL, CAND, L);
    const uint4 t = lds_v4(tab_a + (L >> lg) * 16u);  // col_a, rp_a, st_a, max_deg of the item's run
    const bool cand = (candw >> lane) & 1u;
    uint32_t e0 = 0, n = 0;
    if (cand) {
      const uint32_t a = t.y + (L & wmask) * 64u + lane2;
      e0 = lds_u16(a);
      n = lds_u16(a + 2u) - e0;
    }
    const uint32_t cpa = t.x + e0 * 2u;
    uint32_t x0, x1, x2, x3;
    asm volatile("ld.shared.u16 %0, [%4];\n\tld.shared.u16 %1, [%4+2];\n\tld.shared.u16 %2, [%4+4];\n\tld.shared.u16 %3, [%4+6];"
                 : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(cpa));  // may run past the row: valid indices, masked below
    const uint32_t s0 = lds_u8(t.z + x0), s1 = lds_u8(t.z + x1), s2 = lds_u8(t.z + x2), s3 = lds_u8(t.z + x3);
    uint32_t w = ((s3 * 256u + s2) * 256u + s1) * 256u + s0;
    w &= bmsk_clamp(0u, n * 8u);
    if (LONG_ROWS) {
      for (uint32_t e = 4; e < n; ++e) w |= lds_u8(t.z + lds_u16(cpa + e * 2u));
    }
    const uint32_t metb = __ballot_sync(FULL, cand && (w & 0x01010101u) == 0);
    if (NEED_FD) {
      const uint32_t fdb = __ballot_sync(FULL, (w & 0x02020202u) != 0);
      if (lane == L) fd_w = fdb;
    }
    if (lane == L) met_w = metb;
  }
}

// CD: cond and/or decision codes present   XO: any of fail/needs_cond/skip_dep/phase_out requested
template <bool CD, bool XO>
__global__ void __launch_bounds__(512) frontier_quad_kernel(const KParams P) {
  const uint32_t lane = pin(threadIdx.x & 31u);  // pinned: otherwise rematerialised from S2R inside the loop
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t ST = P.stages;
  const uint32_t Wq = P.wq, lg = P.wq_log2, R = 32u >> lg;
  const uint32_t g = pin(lane >> lg), w = pin(lane & (Wq - 1u));
  const uint32_t gmask = (Wq == 32u ? FULL : ((1u << Wq) - 1u)) << (g << lg);

  // ---- shared memory carve-up: [block counters 128 B][warp regions] ----
  unsigned long long* blk_counts = reinterpret_cast<unsigned long long*>(smem_q);
  const uint32_t run_bytes = P.state_stride + P.topo_buf_bytes;
  const uint32_t stage_bytes = R * run_bytes;
  const uint32_t ring_bytes = ST * stage_bytes;
  const uint32_t per_warp = ring_bytes + P.work_bytes + 64;
  const uint32_t wbase_off = 128 + warp * per_warp;
  uint8_t* const wbase = smem_q + wbase_off;
  const uint32_t smem_base = smem_u32(smem_q);
  const uint32_t ring = pin(smem_base + wbase_off);                         // shared-window address of the warp's ring
  const uint32_t bars = ring + ring_bytes + P.work_bytes;                   // mbarriers: behind the scratch, away from the TMA destinations
  // scratch: [fix-up fail words 128 B][status bytes 1024 + 16][per-group walk entries 16 x 16 B]
  uint32_t* const mFAIL = reinterpret_cast<uint32_t*>(wbase + ring_bytes);  // [32] (fix-up path only)
  const uint32_t st0_off = wbase_off + ring_bytes + 128;
  const uint32_t st0_a = ring + ring_bytes + 128u;
  const uint32_t tab_a = st0_a + 1040u;

  if (threadIdx.x < 4) blk_counts[threadIdx.x] = 0ull;
  if (lane == 0) {
    for (uint32_t s = 0; s < ST; ++s) mbar_init(bars + 8 * s, R);
    fence_barrier_init();
  }
  __syncthreads();

  const uint32_t gw = blockIdx.x * P.warps_per_block + warp;
  const uint32_t G = gridDim.x * P.warps_per_block;
  const uint32_t N = P.n_runs;
  const uint32_t n_trips = (N + R - 1) >> (5 - lg);
  const uint32_t my_trips = gw < n_trips ? (n_trips - gw + G - 1) / G : 0;
  const size_t trip_state_step = (size_t)G * R * P.state_stride;

  // ---- producers: lane g' < R owns run g' of every trip (slot id -> slot entry -> TMA) ----
  const bool producer = lane < R;
  const uint8_t* src_state = P.state + ((size_t)gw * R + lane) * P.state_stride;  // my run's record (producers)
  uint32_t run_i = gw * R + lane;       // global run index of the producer's next issue
  uint32_t ni = 0, is = 0;
  uint64_t ent_addr = 0;
  uint32_t ent_bytes = 0;
  uint32_t sid_q = 0xFFFFFFFFu;
  uint32_t ok_bits = 0, defer_bits = 0;
  bool ent_defer = false;  // the topology has parallel steps: this run goes to the general kernel

  auto load_ent = [&](uint32_t sid) {
    ent_addr = 0;
    ent_bytes = 0;
    ent_defer = false;
    if (sid < P.n_slots) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.slots + sid));
      ent_addr = (uint64_t)v.x | ((uint64_t)v.y << 32);
      ent_bytes = v.z;
      ent_defer = (v.w >> 16) != 0 && P.defer_list != nullptr;
    }
  };
  auto issue = [&]() {  // executed by the R producer lanes
    if (ni < my_trips) {
      const uint32_t buf = ring + is * stage_bytes;
      const uint32_t bar = bars + 8 * is;
      const bool valid = run_i < N;
      const bool dfr = valid && ent_defer;
      const bool ok = valid && !dfr && ent_addr != 0 && ent_bytes <= P.topo_buf_bytes;
      uint32_t tx = ok ? ent_bytes : 0u;
      ok_bits = ok ? (ok_bits | (1u << is)) : (ok_bits & ~(1u << is));
      defer_bits = dfr ? (defer_bits | (1u << is)) : (defer_bits & ~(1u << is));
      uint32_t sbytes = 0;
      if (lane == 0) {  // the trip's state records are adjacent: one copy
        const uint32_t first = run_i, nvalid = min(R, N - first);
        sbytes = nvalid * P.state_stride;
      }
      mbar_expect_tx(bar, tx + sbytes);
      if (lane == 0) bulk_g2s(buf, src_state, sbytes, bar);
      if (ok) bulk_g2s(buf + R * P.state_stride + lane * P.topo_buf_bytes, reinterpret_cast<const void*>(ent_addr), tx, bar);
    }
    load_ent(sid_q);
    const uint32_t run_n2 = run_i + 2 * G * R;
    sid_q = (ni + 2 < my_trips && run_n2 < N) ? __ldg(reinterpret_cast<const uint32_t*>(src_state + 2 * trip_state_step)) : 0xFFFFFFFFu;
    src_state += trip_state_step;
    run_i += G * R;
    ++ni;
    is = (is + 1 == ST) ? 0 : is + 1;
  };
  if (producer && my_trips != 0) {
    load_ent(run_i < N ? __ldg(reinterpret_cast<const uint32_t*>(src_state)) : 0xFFFFFFFFu);
    sid_q = (my_trips > 1 && run_i + G * R < N) ? __ldg(reinterpret_cast<const uint32_t*>(src_state + trip_state_step)) : 0xFFFFFFFFu;
    for (uint32_t s = 0; s < ST; ++s) issue();
  }

  const uint32_t Wmax = P.words;
  const bool has_cond = CD && P.off_cond != BF_OFF_NONE;
  const bool has_dec = CD && P.off_decision != BF_OFF_NONE;

  uint32_t lane_ready = 0, lane_skip = 0, lane_evals = 0;  // per-lane running totals, reduced once at the end
  uint32_t cs = 0, cpar = 0;
  uint32_t r = gw * R + g;  // my group's run index
  uint8_t* rr = P.result + (size_t)r * P.result_stride;
  const size_t trip_result_step = (size_t)G * R * P.result_stride;

  for (uint32_t k = 0; k < my_trips; ++k, r += G * R, rr += trip_result_step) {
    mbar_wait(bars + 8 * cs, cpar);
    const uint32_t stage_off = wbase_off + cs * stage_bytes;
    const uint32_t tr0_off = stage_off + R * P.state_stride;
    const uint32_t stage_a = ring + cs * stage_bytes;
    const uint32_t sr_a = stage_a + g * P.state_stride;                                 // my run's state record
    const uint32_t tr_a = stage_a + R * P.state_stride + g * P.topo_buf_bytes;          // my run's topology record
    const bool staged = (__shfl_sync(FULL, ok_bits, g) >> cs) & 1u;
    const bool deferred = (__shfl_sync(FULL, defer_bits, g) >> cs) & 1u;
    cs = (cs + 1 == ST) ? 0 : cs + 1;
    cpar ^= (cs == 0);

    const bool in_batch = r < N;
    const uint4 h0 = lds_v4(tr_a);
    const uint4 h1 = lds_v4(tr_a + 16);
    uint32_t S = h0.x & 0xFFFFu, Wt = h0.x >> 16;
    const uint32_t max_deg = h0.y & 0xFFFFu;
    uint32_t n_main = h0.z & 0xFFFFu, n_comp = h0.z >> 16, n_final = h0.w & 0xFFFFu;
    const bool live = in_batch && staged && Wt <= Wmax;  // group-uniform
    if (!live) { S = 0; Wt = 0; n_main = n_comp = n_final = 0; }
    const uint32_t rflags = live ? lds_u8(sr_a + 4) : (uint32_t)BF_RF_HOST_GROUP | (BF_GROUP_DONE << BF_RF_HOST_GROUP_SHIFT);

    // ---------------- planes of my word ----------------
    const bool act = w < Wt;
    uint32_t AF = 0, TS = 0, HASIF = 0, G1 = 0, G2 = 0, VALID = 0, SYNC_T = 0;
    uint32_t c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    if (act) {
      const uint32_t sp = tr_a + h1.y + w * 4u, ps = Wt * 4u;   // static planes: my word, plane stride
      AF = lds_u32(sp + PL_AF * ps);
      G1 = lds_u32(sp + PL_G1 * ps); G2 = lds_u32(sp + PL_G2 * ps);
      if (XO) HASIF = lds_u32(sp + PL_HASIF * ps);
      VALID = bmsk_clamp(0u, S - w * 32u);
      const uint32_t pw = sr_a + P.off_phase + w * 4u, ds = Wmax * 4u;  // dynamic planes: stride of the layout
      p0 = lds_u32(pw); p1 = lds_u32(pw + ds); p2 = lds_u32(pw + 2u * ds); p3 = lds_u32(pw + 3u * ds);
      const uint32_t keep = VALID & ~(p0 & p1 & p2 & p3);  // steps >= S and the reserved code 15 read as 0
      p0 &= keep; p1 &= keep; p2 &= keep; p3 &= keep;
      if (CD) {
        if (has_cond) { const uint32_t cw = sr_a + P.off_cond + w * 4u; c0 = lds_u32(cw); c1 = lds_u32(cw + ds); }
        if (has_dec) {
          const uint32_t dw = sr_a + P.off_decision + w * 4u; d0 = lds_u32(dw); d1 = lds_u32(dw + ds);
          const uint32_t t0 = lds_u32(sp + PL_T0 * ps), t1 = lds_u32(sp + PL_T1 * ps), t2 = lds_u32(sp + PL_T2 * ps);
          TS = lds_u32(sp + PL_TS * ps);
          SYNC_T = t0 & (t1 | t2);  // sleep(3) | wait(5) | gate(7)
        }
      }
    }
    const uint32_t q0 = p0, q1 = p1, q2 = p2, q3 = p3;
    const uint32_t GM = VALID & ~G1 & ~G2;
    const bool fail_fast = rflags & BF_RF_FAIL_FAST;
    const bool realtime = rflags & BF_RF_REALTIME;
    const bool topo_term = rflags & BF_RF_TOPOLOGY_TERMINATED;
    const bool host_group = rflags & BF_RF_HOST_GROUP;
    bool marked = false;  // group-uniform

    // ---------------- stage G: gate / sleep / wait sync (dag.go:1469-1533, 1235-1277, 1327-1437) ----------------
    if (CD) {
      if (has_dec) {
        const uint32_t syn = SYNC_T & plut<BF_LUT_RUNNING>(p0, p1, p2, p3);
        const uint32_t n0 = d0;
        const uint32_t n1 = d0 & ~(d1 & TS);
        const uint32_t n2 = d1 & (~d0 | TS);
        const uint32_t n3 = ~(d0 ^ d1);
        p0 = (p0 & ~syn) | (n0 & syn);
        p1 = (p1 & ~syn) | (n1 & syn);
        p2 = (p2 & ~syn) | (n2 & syn);
        p3 = (p3 & ~syn) | (n3 & syn);
      }
    }

    // ---------------- stage B: classification (dag.go:3377-3388, 2020-2033) ----------------
    uint32_t TERM = plut<BF_LUT_TERMINAL>(p0, p1, p2, p3);
    uint32_t COMPL = plut<BF_LUT_COMPLETED0>(p0, p1, p2, p3) | (TERM & AF);
    uint32_t RUNQ = plut<BF_LUT_RUNNING_Q>(p0, p1, p2, p3);
    uint32_t FAILED = TERM & ~COMPL;
    uint32_t group, sum = 0;
    // ---------------- stage I (dag.go:422-495): per-run reductions = one sub-warp redux.or ----------------
    {
      uint32_t RUN = plut<BF_LUT_RUNNING>(p0, p1, p2, p3);
      const uint32_t DONE = COMPL | FAILED;
      const uint32_t mark_ff = GM & ~COMPL & ~RUNQ & ~TERM;           // markFailFastSkipped candidates (:3289-3312)
      const uint32_t mark_cs = G1 & ~COMPL & ~RUN & ~FAILED & ~TERM;   // markCompensationsSkipped candidates (:3314-3342)
      uint32_t bits = ((FAILED & GM) != 0 ? 1u : 0u) | ((GM & ~DONE) != 0 ? 2u : 0u) | ((GM & ~DONE & ~mark_ff) != 0 ? 4u : 0u) |
                      ((G1 & ~DONE) != 0 ? 8u : 0u) | ((G1 & ~DONE & ~mark_cs) != 0 ? 16u : 0u) | ((G2 & ~DONE) != 0 ? 32u : 0u) |
                      ((FAILED & G1) != 0 ? 64u : 0u) | ((FAILED & G2) != 0 ? 128u : 0u) | (mark_ff != 0 ? 256u : 0u) |
                      (mark_cs != 0 ? 512u : 0u);
      bits = __reduce_or_sync(gmask, bits);
      // no collective may sit under a per-run branch: host-group runs (tier K1) take part in every
      // vote below and simply never qualify for a marking
      const bool auto_group = !host_group;
      bool amf = bits & 1u;
      const bool do_ff = auto_group && fail_fast && amf;
      if (do_ff) {
        marked = marked || (bits & 256u);
        pset<BF_PHASE_SKIPPED>(mark_ff, p0, p1, p2, p3);
        TERM |= mark_ff; COMPL |= mark_ff; RUN &= ~mark_ff;
      }
      bool main_done = n_main == 0 || !(bits & (do_ff ? 4u : 2u));
      const bool acf = bits & 64u;
      // dag.go:436-464 (rare)
      const bool tt = auto_group && !main_done && realtime && topo_term;
      if (__any_sync(FULL, tt)) {
        const uint32_t mtt = tt ? (GM & (p0 | p1 | p2 | p3) & ~TERM) : 0u;
        pset<BF_PHASE_FAILED>(mtt, p0, p1, p2, p3);
        TERM |= mtt; COMPL |= mtt & AF; FAILED |= mtt & ~AF; RUN &= ~mtt; RUNQ &= ~mtt;
        const uint32_t b2 = __reduce_or_sync(gmask, (mtt != 0 ? 1u : 0u) | ((FAILED & GM) != 0 ? 2u : 0u));
        if (tt) {
          marked = marked || (b2 & 1u);
          main_done = true;
          amf = b2 & 2u;
        }
      }
      bool comp_done;
      if (auto_group && main_done && !amf && n_comp != 0) {
        marked = marked || (bits & 512u);
        pset<BF_PHASE_SKIPPED>(mark_cs, p0, p1, p2, p3);
        TERM |= mark_cs; COMPL |= mark_cs;
        comp_done = !(bits & 16u);
      } else {
        comp_done = n_comp == 0 || !(bits & 8u);
      }
      const bool final_done = n_final == 0 || !(bits & 32u);
      const bool aff = bits & 128u;
      if (!main_done) group = BF_GROUP_MAIN;
      else if (amf && n_comp != 0 && !comp_done) group = BF_GROUP_COMPENSATION;
      else if (n_final != 0 && !final_done) group = BF_GROUP_FINALLY;
      else group = BF_GROUP_DONE;
      sum = (main_done ? BF_SUM_MAIN_DONE : 0u) | (amf ? BF_SUM_MAIN_FAILED : 0u) | (comp_done ? BF_SUM_COMP_DONE : 0u) |
            (final_done ? BF_SUM_FINAL_DONE : 0u) | (acf ? BF_SUM_COMP_FAILED : 0u) | (aff ? BF_SUM_FINAL_FAILED : 0u);
      if (host_group) {
        group = (rflags >> BF_RF_HOST_GROUP_SHIFT) & 3u;
        sum = 0;
      }
    }
    uint32_t summary = sum | group;

    // ------------- D-prep: dependency classes under this pass's policy (dag.go:499-502) -------------
    const bool evaluate = group != BF_GROUP_DONE;
    const bool allow_failed = group != BF_GROUP_MAIN;
    const bool skip_on_failed = group == BF_GROUP_MAIN && !fail_fast;
    const uint32_t GSEL = group == BF_GROUP_MAIN ? GM : (group == BF_GROUP_COMPENSATION ? G1 : G2);
    const uint32_t SAT = COMPL | (realtime ? plut<BF_LUT_RT_SAT>(p0, p1, p2, p3) : 0u) | (allow_failed ? TERM : 0u);
    const uint32_t U = ~SAT;
    const uint32_t FD = skip_on_failed ? (TERM & ~SAT) : 0u;
    const uint32_t CAND = evaluate ? (GSEL & ~COMPL & ~RUNQ & ~TERM) : 0u;
    // ------------- stage C: one status byte per step (bit0 unmet, bit1 failed-dep), all R runs -------------
    const bool any_fd = __any_sync(FULL, skip_on_failed);              // some run of the trip has a failed-dependency class
    const bool any_long = __any_sync(FULL, live && max_deg > 4);       // some run has rows longer than the straight-line four
    __syncwarp();
#pragma unroll
    for (uint32_t t = 0; t < 4; ++t) {
      const uint32_t item = t * 32 + lane;  // 8 steps: byte (item & 3) of the word held by lane item >> 2
      const uint32_t sh = (item & 3u) * 8u;
      const uint32_t ub = __shfl_sync(FULL, U, item >> 2) >> sh;
      uint32_t vx = bits4_to_bytes(ub & 0xFu), vy = bits4_to_bytes((ub >> 4) & 0xFu);
      if (any_fd) {  // warp-uniform
        const uint32_t fb = __shfl_sync(FULL, FD, item >> 2) >> sh;
        vx |= bits4_to_bytes(fb & 0xFu) << 1;
        vy |= bits4_to_bytes((fb >> 4) & 0xFu) << 1;
      }
      sts_v2(st0_a + item * 8u, vx, vy);
    }
    // per-group walk entry: CSR bases, status base, longest row
    const uint32_t col_a = tr_a + h1.x;
    if (w == 0) sts_v4(tab_a + g * 16u, col_a, tr_a + (uint32_t)sizeof(TopoHeader), st0_a + (g << (5u + lg)), max_deg);
    __syncwarp();
    // ------------- stage D: walk the needs rows (dag.go:2711-2733) -------------
    const uint32_t col_off = col_a - smem_base;             // fix-up path (generic pointers)
    const uint32_t meta = (32u * Wt) | (max_deg << 16);
    uint32_t met_w, fd_w;
    if (any_long) {
      if (any_fd) walk_items<true, true>(lane, CAND, lg, tab_a, met_w, fd_w);
      else walk_items<false, true>(lane, CAND, lg, tab_a, met_w, fd_w);
    } else {
      if (any_fd) walk_items<true, false>(lane, CAND, lg, tab_a, met_w, fd_w);
      else walk_items<false, false>(lane, CAND, lg, tab_a, met_w, fd_w);
    }
    uint32_t ready_w = met_w, skipc_w = 0, fail_w = 0;
    if (CD) {
      ready_w = met_w & ~c0 & ~c1;   // BF_COND_PASS
      skipc_w = met_w & c0 & ~c1;    // BF_COND_SKIP
      fail_w = met_w & c0 & c1;      // BF_COND_FAIL
      if (__any_sync(FULL, fail_w != 0)) {
        // a step set Failed inside the loop is visible to LATER steps of the list only (dag.go:2745, :497)
        const uint32_t fclass = allow_failed ? 0u : (skip_on_failed ? 3u : 1u);
        for (uint32_t round = 0; round <= 32u * Wmax; ++round) {
          __syncwarp();
          mFAIL[lane] = fail_w;
          __syncwarp();
          walk_quad<true>(lane, CAND, lg, col_off, meta, tr0_off, P.topo_buf_bytes, st0_off, mFAIL, fclass, met_w, fd_w);
          const uint32_t nf = met_w & c0 & c1;
          const bool same = !__any_sync(FULL, nf != fail_w);
          fail_w = nf;
          if (same) break;
        }
        ready_w = met_w & ~c0 & ~c1;
        skipc_w = met_w & c0 & ~c1;
      }
      pset<BF_PHASE_FAILED>(fail_w, p0, p1, p2, p3);  // dag.go:2745-2747, 2810-2812
    }
    const uint32_t acc_ready = ready_w, acc_skip = fd_w | skipc_w;

    // ---------------- stage E: result records ----------------
    const uint32_t cnt = __reduce_add_sync(gmask, (uint32_t)__popc(acc_ready) | ((uint32_t)__popc(acc_skip) << 16));
    bool changed = marked;
    if (CD) changed = (__ballot_sync(FULL, ((p0 ^ q0) | (p1 ^ q1) | (p2 ^ q2) | (p3 ^ q3)) != 0) & gmask) != 0;
    if (in_batch && deferred) {
      if (w == 0) P.defer_list[atomicAdd(P.defer_count, 1u)] = r;  // the general kernel writes this run's record
    } else if (in_batch) {
      if (w == 0) {
        summary = live ? (summary | (changed ? BF_SUM_PHASE_CHANGED : 0u) | (1u << BF_SUM_ITER_SHIFT)) : 0xFFFFFFFFu;
        *reinterpret_cast<uint4*>(rr) = make_uint4(summary, cnt & 0xFFFFu, cnt >> 16, 0u);
        if (P.exp_counts) P.exp_counts[r] = 0;
      }
      if (w < Wmax) {
        reinterpret_cast<uint32_t*>(rr + P.off_ready)[w] = acc_ready;
        reinterpret_cast<uint32_t*>(rr + P.off_skip)[w] = acc_skip;
        if (XO) {
          if (P.off_fail != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_fail)[w] = fail_w;
          if (P.off_needs_cond != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_needs_cond)[w] = realtime ? 0u : (met_w & HASIF);
          if (P.off_skip_dep != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_skip_dep)[w] = fd_w;
          if (P.off_phase_out != BF_OFF_NONE) {
            uint32_t* po = reinterpret_cast<uint32_t*>(rr + P.off_phase_out) + w;
            po[0] = p0; po[Wmax] = p1; po[2 * Wmax] = p2; po[3 * Wmax] = p3;
          }
        }
      }
      if (P.result_tail != P.result_stride)
        for (uint32_t x = P.result_tail / 4 + w; x < P.result_stride / 4; x += Wq) reinterpret_cast<uint32_t*>(rr)[x] = 0u;
    }
    lane_ready += (uint32_t)__popc(acc_ready);
    lane_skip += (uint32_t)__popc(acc_skip);
    lane_evals += (w == 0 && live) ? S : 0u;

    __syncwarp();  // every lane is done with this stage's buffers
    if (producer) issue();
  }

  // ---- counters: lane -> warp (redux) -> block (shared atomics) -> one global atomic per block ----
  if (P.counts) {
    const uint32_t wr = redux_add(lane_ready), ws = redux_add(lane_skip), we = redux_add(lane_evals);
    if (lane == 0 && my_trips != 0) {
      atomicAdd(&blk_counts[0], (unsigned long long)wr);
      atomicAdd(&blk_counts[1], (unsigned long long)ws);
      atomicAdd(&blk_counts[3], (unsigned long long)we);
    }
    __syncthreads();
    if (threadIdx.x < 4 && blk_counts[threadIdx.x] != 0ull) atomicAdd(&P.counts[threadIdx.x], blk_counts[threadIdx.x]);
  }
}

typedef void (*QuadFn)(const KParams);
static QuadFn pick_quad(const KParams& P) {
  const bool cd = P.off_cond != BF_OFF_NONE || P.off_decision != BF_OFF_NONE;
  const bool xo = P.off_fail != BF_OFF_NONE || P.off_needs_cond != BF_OFF_NONE || P.off_skip_dep != BF_OFF_NONE ||
                  P.off_phase_out != BF_OFF_NONE;
  if (cd) return xo ? frontier_quad_kernel<true, true> : frontier_quad_kernel<true, false>;
  return xo ? frontier_quad_kernel<false, true> : frontier_quad_kernel<false, false>;
}

cudaError_t launch_frontier_quad(const KParams& P, uint32_t grid, uint32_t smem_bytes, cudaStream_t stream) {
  QuadFn fn = pick_quad(P);
  static QuadFn configured[8][4] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  bool known = false;
  if (dev >= 0 && dev < 8)
    for (int i = 0; i < 4; ++i) known = known || configured[dev][i] == fn;
  if (!known) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 8)
      for (int i = 0; i < 4; ++i)
        if (configured[dev][i] == nullptr) { configured[dev][i] = fn; break; }
  }
  fn<<<grid, P.warps_per_block * 32, smem_bytes, stream>>>(P);
  return cudaGetLastError();
}

int frontier_quad_max_blocks_per_sm(const KParams& P, uint32_t threads, uint32_t smem_bytes) {
  int n = 0;
  QuadFn fn = pick_quad(P);
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, (int)threads, smem_bytes) != cudaSuccess) return 1;
  return n;
}

}  // namespace bf
