// frontier_split.cu — two-phase single-pass evaluation.
//
// Most of a StoryRun's bytes are its needs-adjacency (CSR), but the adjacency is only consulted for
// CANDIDATE steps (findReadySteps `continue`s past completed / running / terminal steps before it ever
// looks at `deps`, dag.go:2649-2708) — and a run whose group selection ends in "finalize", or whose
// steps are all running or done, has no candidate at all.  So:
//
//   phase 1  classify_kernel   every run: state record + the topology's static bit planes (~430 B at
//                              S = 256).  Packed lanes (R = 32 / Wq runs per warp trip): sync rewrite,
//                              buildStateMaps classification, fail-fast / compensation marking, group
//                              selection (dag.go:409-511) — and the complete result record of every run
//                              that has nothing to walk.  Runs with candidates get a compact hand-over
//                              entry (candidate / unmet / failed-dep masks + CSR address) appended to a list.
//   phase 2  walk_kernel       listed runs only: TMA-stages the CSR block (row_ptr + col_idx) and the
//                              entry, walks the needs rows (dag.go:2711-2733), applies the cond codes,
//                              patches ready / skip (and the optional masks) into the result record.
//
// Runs whose topology has `parallel` steps are deferred to the general kernel (join + expansion live
// there); fixpoint mode always uses the general kernel.  Same stages, same citations, same results:
// the parity tests run this path and the general kernel against the oracle.
#include "kernel_common.cuh"

namespace bf {

// ============================================================================ phase 1
template <bool CD, bool XO>
__global__ void __launch_bounds__(256) classify_kernel(const KParams P) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t Wq = P.wq, lg = P.wq_log2, R = 32u >> lg;
  const uint32_t g = lane >> lg, w = lane & (Wq - 1u);
  const uint32_t gmask = (Wq == 32u ? FULL : ((1u << Wq) - 1u)) << (g << lg);
  const uint32_t N = P.n_runs, Wmax = P.words;
  const uint32_t n_trips = (N + R - 1) >> (5 - lg);
  const bool has_cond = CD && P.off_cond != BF_OFF_NONE;
  const bool has_dec = CD && P.off_decision != BF_OFF_NONE;
  uint32_t lane_evals = 0;
  const uint32_t warp = threadIdx.x >> 5;
  __shared__ uint32_t s_cnt[8];
  __shared__ uint32_t s_base;
  __shared__ unsigned long long s_evals;
  if (threadIdx.x == 0) s_evals = 0ull;

  // every warp of a block runs the same number of loop trips (the list reservation below is block-wide)
  for (uint32_t trip0 = blockIdx.x * 8u; trip0 < n_trips; trip0 += gridDim.x * 8u) {
    const uint32_t trip = trip0 + warp;
    const uint32_t r = trip * R + g;
    const bool in_batch = trip < n_trips && r < N;
    const uint8_t* srec = P.state + (size_t)(in_batch ? r : 0) * P.state_stride;
    const uint2 hdr = *reinterpret_cast<const uint2*>(srec);  // topo_slot, run_flags
    const uint32_t slot = hdr.x;
    SlotInfo si = {0, 0, 0, 0, 0, 0, 0};
    if (in_batch && slot < P.n_slots) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(P.slot_info + slot));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(P.slot_info + slot) + 1);
      si.addr = (uint64_t)a.x | ((uint64_t)a.y << 32); si.csr_bytes = a.z; si.off_planes = a.w;
      si.s_w = b.x; si.deg_p = b.y; si.main_comp = b.z; si.n_final = b.w;
    }
    uint32_t S = si.s_w & 0xFFFFu, Wt = si.s_w >> 16;
    const uint32_t max_deg = si.deg_p & 0xFFFFu, nP = si.deg_p >> 16;
    uint32_t n_main = si.main_comp & 0xFFFFu, n_comp = si.main_comp >> 16, n_final = si.n_final;
    const bool deferred = in_batch && si.addr != 0 && nP != 0 && P.defer_list != nullptr;
    const bool live = in_batch && si.addr != 0 && Wt <= Wmax && !deferred;  // group-uniform
    if (!live) { S = 0; Wt = 0; n_main = n_comp = n_final = 0; }
    const uint32_t rflags = live ? (hdr.y & 0xFFu) : ((uint32_t)BF_RF_HOST_GROUP | (BF_GROUP_DONE << BF_RF_HOST_GROUP_SHIFT));

    const bool act = w < Wt;
    uint32_t AF = 0, TS = 0, HASIF = 0, G1 = 0, G2 = 0, VALID = 0, SYNC_T = 0;
    uint32_t c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    if (act) {
      const uint32_t* sp = reinterpret_cast<const uint32_t*>(si.addr + si.off_planes) + w;
      AF = __ldg(sp + PL_AF * Wt);
      G1 = __ldg(sp + PL_G1 * Wt); G2 = __ldg(sp + PL_G2 * Wt);
      if (XO) HASIF = __ldg(sp + PL_HASIF * Wt);
      const uint32_t rem = S - w * 32;
      VALID = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
      const uint32_t* pw = reinterpret_cast<const uint32_t*>(srec + P.off_phase) + w;
      p0 = pw[0]; p1 = pw[Wmax]; p2 = pw[2 * Wmax]; p3 = pw[3 * Wmax];
      const uint32_t keep = VALID & ~(p0 & p1 & p2 & p3);  // steps >= S and the reserved code 15 read as 0
      p0 &= keep; p1 &= keep; p2 &= keep; p3 &= keep;
      if (CD) {
        if (has_cond) { const uint32_t* cw = reinterpret_cast<const uint32_t*>(srec + P.off_cond) + w; c0 = cw[0]; c1 = cw[Wmax]; }
        if (has_dec) {
          const uint32_t* dw = reinterpret_cast<const uint32_t*>(srec + P.off_decision) + w; d0 = dw[0]; d1 = dw[Wmax];
          const uint32_t t0 = __ldg(sp + PL_T0 * Wt), t1 = __ldg(sp + PL_T1 * Wt), t2 = __ldg(sp + PL_T2 * Wt);
          TS = __ldg(sp + PL_TS * Wt);
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
    bool marked = false;

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
      const bool auto_group = !host_group;  // no collective sits under a per-run branch
      bool amf = bits & 1u;
      const bool do_ff = auto_group && fail_fast && amf;
      if (do_ff) {
        marked = marked || (bits & 256u);
        pset<BF_PHASE_SKIPPED>(mark_ff, p0, p1, p2, p3);
        TERM |= mark_ff; COMPL |= mark_ff; RUN &= ~mark_ff;
      }
      bool main_done = n_main == 0 || !(bits & (do_ff ? 4u : 2u));
      const bool acf = bits & 64u;
      const bool tt = auto_group && !main_done && realtime && topo_term;  // dag.go:436-464 (rare)
      if (__any_sync(FULL, tt)) {
        const uint32_t mtt = tt ? (GM & (p0 | p1 | p2 | p3) & ~TERM) : 0u;
        pset<BF_PHASE_FAILED>(mtt, p0, p1, p2, p3);
        TERM |= mtt; COMPL |= mtt & AF; FAILED |= mtt & ~AF; RUN &= ~mtt; RUNQ &= ~mtt;
        const uint32_t b2 = __reduce_or_sync(gmask, (mtt != 0 ? 1u : 0u) | ((FAILED & GM) != 0 ? 2u : 0u));
        if (tt) { marked = marked || (b2 & 1u); main_done = true; amf = b2 & 2u; }
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
      if (host_group) { group = (rflags >> BF_RF_HOST_GROUP_SHIFT) & 3u; sum = 0; }
    }
    // ------------- D-prep: dependency classes under this pass's policy (dag.go:499-502) -------------
    const bool evaluate = group != BF_GROUP_DONE;
    const bool allow_failed = group != BF_GROUP_MAIN;
    const bool skip_on_failed = group == BF_GROUP_MAIN && !fail_fast;
    const uint32_t GSEL = group == BF_GROUP_MAIN ? GM : (group == BF_GROUP_COMPENSATION ? G1 : G2);
    const uint32_t SAT = COMPL | (realtime ? plut<BF_LUT_RT_SAT>(p0, p1, p2, p3) : 0u) | (allow_failed ? TERM : 0u);
    const uint32_t U = ~SAT;
    const uint32_t FD = skip_on_failed ? (TERM & ~SAT) : 0u;
    const uint32_t CAND = evaluate ? (GSEL & ~COMPL & ~RUNQ & ~TERM) : 0u;
    const uint32_t votes = __ballot_sync(FULL, CAND != 0);
    const bool has_cand = (votes & gmask) != 0;
    bool changed = marked;
    if (CD) changed = (__ballot_sync(FULL, ((p0 ^ q0) | (p1 ^ q1) | (p2 ^ q2) | (p3 ^ q3)) != 0) & gmask) != 0;
    uint32_t summary = sum | group | (changed ? BF_SUM_PHASE_CHANGED : 0u) | (1u << BF_SUM_ITER_SHIFT);

    // ---------------- result record of every run (walked runs: phase 2 patches ready / skip / counts) ----------------
    uint8_t* rr = P.result + (size_t)(in_batch ? r : 0) * P.result_stride;
    if (deferred) {
      if (w == 0) P.defer_list[atomicAdd(P.defer_count, 1u)] = r;  // the general kernel writes this run's record
    } else if (in_batch) {
      if (w == 0) {
        *reinterpret_cast<uint4*>(rr) = make_uint4(live ? summary : 0xFFFFFFFFu, 0u, 0u, 0u);
        if (P.exp_counts) P.exp_counts[r] = 0;
      }
      if (w < Wmax) {
        reinterpret_cast<uint32_t*>(rr + P.off_ready)[w] = 0u;
        reinterpret_cast<uint32_t*>(rr + P.off_skip)[w] = 0u;
        if (XO) {
          if (P.off_fail != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_fail)[w] = 0u;
          if (P.off_needs_cond != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_needs_cond)[w] = 0u;
          if (P.off_skip_dep != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_skip_dep)[w] = 0u;
          if (P.off_phase_out != BF_OFF_NONE) {
            uint32_t* po = reinterpret_cast<uint32_t*>(rr + P.off_phase_out) + w;
            po[0] = p0; po[Wmax] = p1; po[2 * Wmax] = p2; po[3 * Wmax] = p3;
          }
        }
      }
      if (P.result_tail != P.result_stride)
        for (uint32_t x = P.result_tail / 4 + w; x < P.result_stride / 4; x += Wq) reinterpret_cast<uint32_t*>(rr)[x] = 0u;
    }
    // ---------------- hand-over entry for runs with candidates ----------------
    // one global atomic per BLOCK: 100k same-address atomics would serialise in L2 for longer than the kernel runs
    const uint32_t heads = __ballot_sync(FULL, has_cand && w == 0);   // bit (g << lg) set: run g of this trip walks
    if (lane == 0) s_cnt[warp] = (uint32_t)__popc(heads);
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (uint32_t x = 0; x < 8; ++x) { const uint32_t c = s_cnt[x]; s_cnt[x] = tot; tot += c; }
      s_base = tot ? atomicAdd(P.walk_count, tot) : 0u;
    }
    __syncthreads();
    const uint32_t idx = s_base + s_cnt[warp] + (uint32_t)__popc(heads & ((1u << (g << lg)) - 1u));
    __syncthreads();  // s_cnt / s_base are rewritten by the next trip
    if (has_cand) {
      uint8_t* ent = P.walk_entries + (size_t)idx * P.walk_entry_bytes;
      if (w == 0) {
        const uint32_t fclass = allow_failed ? 0u : (skip_on_failed ? 3u : 1u);
        const uint64_t csr = si.addr + sizeof(TopoHeader);
        reinterpret_cast<uint4*>(ent)[0] = make_uint4((uint32_t)csr, (uint32_t)(csr >> 32), si.csr_bytes, r);
        reinterpret_cast<uint4*>(ent)[1] = make_uint4(Wt | (max_deg << 16), summary, fclass,
                                                      (2u * (S + 1u) + 15u) & ~15u);  // col_idx offset inside the CSR block
      }
      if (w < Wmax) {
        uint32_t* aw = reinterpret_cast<uint32_t*>(ent + sizeof(WalkEntry)) + w;
        aw[0] = CAND; aw[Wmax] = U; aw[2 * Wmax] = FD;
        uint32_t k = 3;
        if (CD) { aw[k * Wmax] = c0; aw[(k + 1) * Wmax] = c1; k += 2; }
        if (XO) { aw[k * Wmax] = realtime ? 0u : HASIF; k += 1; }
        if (CD && XO) { aw[k * Wmax] = p0; aw[(k + 1) * Wmax] = p1; aw[(k + 2) * Wmax] = p2; aw[(k + 3) * Wmax] = p3; }
      }
    }
    lane_evals += (w == 0 && live) ? S : 0u;
  }
  if (P.counts) {
    const uint32_t we = redux_add(lane_evals);
    if (lane == 0 && we != 0) atomicAdd(&s_evals, (unsigned long long)we);
    __syncthreads();
    if (threadIdx.x == 0 && s_evals != 0ull) atomicAdd(&P.counts[3], s_evals);
  }
}

// ============================================================================ phase 2
extern __shared__ __align__(128) uint8_t smem_w[];

template <bool CD, bool XO>
__global__ void __launch_bounds__(512) walk_kernel(const KParams P) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t ST = P.stages;
  const uint32_t Wmax = P.words;
  unsigned long long* blk_counts = reinterpret_cast<unsigned long long*>(smem_w);
  const uint32_t EB = P.walk_entry_bytes;
  const uint32_t stage_bytes = EB + P.topo_buf_bytes;         // entry + CSR block
  const uint32_t ring_bytes = ST * stage_bytes;
  const uint32_t per_warp = ring_bytes + P.work_bytes + 64;
  uint8_t* const wbase = smem_w + 128 + warp * per_warp;
  const uint32_t bars = smem_u32(wbase + ring_bytes + P.work_bytes);
  const uint32_t ring = smem_u32(wbase);
  uint32_t* const mFAIL = reinterpret_cast<uint32_t*>(wbase + ring_bytes);
  uint8_t* const st = wbase + ring_bytes + ((4u * Wmax + 15u) & ~15u);
  const uint32_t st_addr = pin(smem_u32(st));

  if (threadIdx.x < 4) blk_counts[threadIdx.x] = 0ull;
  if (lane == 0) {
    for (uint32_t s = 0; s < ST; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const uint32_t gw = blockIdx.x * P.warps_per_block + warp;
  const uint32_t G = gridDim.x * P.warps_per_block;
  const uint32_t N = min(*reinterpret_cast<const volatile uint32_t*>(P.walk_count), P.n_runs);
  const uint32_t my_runs = gw < N ? (N - gw + G - 1) / G : 0;
  const size_t ent_step = (size_t)G * EB;

  // ---- producer (lane 0): entry header prefetch one issue ahead -> two TMA copies ----
  const uint8_t* src_ent = P.walk_entries + (size_t)gw * EB;
  uint32_t ni = 0, is = 0;
  uint64_t csr_addr = 0;
  uint32_t csr_bytes = 0;
  auto load_hdr = [&](const uint8_t* e) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(e));  // written by phase 1 (the previous kernel)
    csr_addr = (uint64_t)v.x | ((uint64_t)v.y << 32);
    csr_bytes = v.z;
  };
  auto issue = [&]() {
    if (ni < my_runs) {
      const uint32_t buf = ring + is * stage_bytes;
      const uint32_t bar = bars + 8 * is;
      const uint32_t cb = csr_bytes <= P.topo_buf_bytes ? csr_bytes : 0u;
      mbar_expect_tx(bar, EB + cb);
      bulk_g2s(buf, src_ent, EB, bar);
      if (cb) bulk_g2s(buf + EB, reinterpret_cast<const void*>(csr_addr), cb, bar);
    }
    src_ent += ent_step;
    ++ni;
    if (ni < my_runs) load_hdr(src_ent);
    is = (is + 1 == ST) ? 0 : is + 1;
  };
  if (lane == 0 && my_runs != 0) {
    load_hdr(src_ent);
    for (uint32_t s = 0; s < ST; ++s) issue();
  }

  uint32_t lane_ready = 0, lane_skip = 0;
  uint32_t cs = 0, cpar = 0;
  for (uint32_t k = 0; k < my_runs; ++k) {
    mbar_wait(bars + 8 * cs, cpar);
    const uint8_t* ent = wbase + cs * stage_bytes;
    const uint8_t* csr = ent + EB;
    cs = (cs + 1 == ST) ? 0 : cs + 1;
    cpar ^= (cs == 0);

    const uint4 e0 = *reinterpret_cast<const uint4*>(ent);
    const uint4 e1 = *reinterpret_cast<const uint4*>(ent + 16);
    const uint32_t r = e0.w;
    const uint32_t Wt = e1.x & 0xFFFFu, max_deg = e1.x >> 16;
    const uint32_t summary = e1.y, fclass = e1.z;
    const uint16_t* row_ptr = reinterpret_cast<const uint16_t*>(csr);
    const uint16_t* col = reinterpret_cast<const uint16_t*>(csr + e1.w);
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(ent + sizeof(WalkEntry));
    const bool act = lane < Wt;
    uint32_t CAND = 0, c0 = 0, c1 = 0, HIF = 0, p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    if (act) {
      CAND = aw[lane];
      uint32_t kk = 3;
      if (CD) { c0 = aw[kk * Wmax + lane]; c1 = aw[(kk + 1) * Wmax + lane]; kk += 2; }
      if (XO) { HIF = aw[kk * Wmax + lane]; kk += 1; }
      if (CD && XO) { p0 = aw[kk * Wmax + lane]; p1 = aw[(kk + 1) * Wmax + lane]; p2 = aw[(kk + 2) * Wmax + lane]; p3 = aw[(kk + 3) * Wmax + lane]; }
    }
    // ------------- stage C: masks -> one status byte per step (bit0 unmet, bit1 failed-dep) -------------
    {
      const uint8_t* ub8 = reinterpret_cast<const uint8_t*>(aw + Wmax);
      const uint8_t* fb8 = reinterpret_cast<const uint8_t*>(aw + 2 * Wmax);
      for (uint32_t m = lane; m < 4 * Wt; m += 32) {
        const uint32_t ub = ub8[m], fb = fb8[m];
        uint2 v;
        v.x = bits4_to_bytes(ub & 0xFu) | (bits4_to_bytes(fb & 0xFu) << 1);
        v.y = bits4_to_bytes(ub >> 4) | (bits4_to_bytes(fb >> 4) << 1);
        reinterpret_cast<uint2*>(st)[m] = v;
      }
    }
    __syncwarp();
    // ------------- stage D: walk the needs rows (dag.go:2711-2733) -------------
    const uint32_t zidx = 32 * Wt;
    uint32_t met_w, fd_w;
    walk_rows_s<true, true>(lane, CAND, smem_u32(row_ptr), smem_u32(col), st_addr, met_w, fd_w);
    uint32_t ready_w = met_w, skipc_w = 0, fail_w = 0;
    if (CD) {
      ready_w = met_w & ~c0 & ~c1;   // BF_COND_PASS
      skipc_w = met_w & c0 & ~c1;    // BF_COND_SKIP
      fail_w = met_w & c0 & c1;      // BF_COND_FAIL
      if (__any_sync(FULL, fail_w != 0)) {
        // a step set Failed inside the loop is visible to LATER steps of the list only (dag.go:2745, :497)
        for (uint32_t round = 0; round <= 32u * Wmax; ++round) {
          __syncwarp();
          if (act) mFAIL[lane] = fail_w;
          __syncwarp();
          walk_rows<true>(lane, CAND, zidx, max_deg, row_ptr, col, st, mFAIL, fclass, met_w, fd_w);
          const uint32_t nf = met_w & c0 & c1;
          const bool same = !__any_sync(FULL, nf != fail_w);
          fail_w = nf;
          if (same) break;
        }
        ready_w = met_w & ~c0 & ~c1;
        skipc_w = met_w & c0 & ~c1;
      }
    }
    const uint32_t skip_w = fd_w | skipc_w;
    // ------------- patch the result record -------------
    const uint32_t cnt = redux_add((uint32_t)__popc(ready_w) | ((uint32_t)__popc(skip_w) << 16));
    uint8_t* rr = P.result + (size_t)r * P.result_stride;
    uint32_t sum2 = summary;
    if (CD) { if (__any_sync(FULL, fail_w != 0)) sum2 |= BF_SUM_PHASE_CHANGED; }
    if (lane == 0) *reinterpret_cast<uint4*>(rr) = make_uint4(sum2, cnt & 0xFFFFu, cnt >> 16, 0u);
    if (act) {
      reinterpret_cast<uint32_t*>(rr + P.off_ready)[lane] = ready_w;
      reinterpret_cast<uint32_t*>(rr + P.off_skip)[lane] = skip_w;
      if (XO) {
        if (CD && P.off_fail != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_fail)[lane] = fail_w;
        if (P.off_needs_cond != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_needs_cond)[lane] = met_w & HIF;
        if (P.off_skip_dep != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_skip_dep)[lane] = fd_w;
        if (CD && P.off_phase_out != BF_OFF_NONE && fail_w != 0) {
          pset<BF_PHASE_FAILED>(fail_w, p0, p1, p2, p3);  // dag.go:2745-2747, 2810-2812
          uint32_t* po = reinterpret_cast<uint32_t*>(rr + P.off_phase_out) + lane;
          po[0] = p0; po[Wmax] = p1; po[2 * Wmax] = p2; po[3 * Wmax] = p3;
        }
      }
    }
    lane_ready += (uint32_t)__popc(ready_w);
    lane_skip += (uint32_t)__popc(skip_w);
    __syncwarp();  // every lane is done with this stage's buffers
    if (lane == 0) issue();
  }
  if (P.counts) {
    const uint32_t wr = redux_add(lane_ready), ws = redux_add(lane_skip);
    if (lane == 0 && my_runs != 0) {
      atomicAdd(&blk_counts[0], (unsigned long long)wr);
      atomicAdd(&blk_counts[1], (unsigned long long)ws);
    }
    __syncthreads();
    if (threadIdx.x < 2 && blk_counts[threadIdx.x] != 0ull) atomicAdd(&P.counts[threadIdx.x], blk_counts[threadIdx.x]);
  }
}

// ============================================================================ host side
typedef void (*SplitFn)(const KParams);
static void variant(const KParams& P, bool& cd, bool& xo) {
  cd = P.off_cond != BF_OFF_NONE || P.off_decision != BF_OFF_NONE;
  xo = P.off_fail != BF_OFF_NONE || P.off_needs_cond != BF_OFF_NONE || P.off_skip_dep != BF_OFF_NONE || P.off_phase_out != BF_OFF_NONE;
}
static SplitFn pick_classify(const KParams& P) {
  bool cd, xo;
  variant(P, cd, xo);
  if (cd) return xo ? classify_kernel<true, true> : classify_kernel<true, false>;
  return xo ? classify_kernel<false, true> : classify_kernel<false, false>;
}
static SplitFn pick_walk(const KParams& P) {
  bool cd, xo;
  variant(P, cd, xo);
  if (cd) return xo ? walk_kernel<true, true> : walk_kernel<true, false>;
  return xo ? walk_kernel<false, true> : walk_kernel<false, false>;
}

uint32_t split_entry_bytes(const KParams& P) {
  bool cd, xo;
  variant(P, cd, xo);
  const uint32_t nw = 3u + (cd ? 2u : 0u) + (xo ? 1u : 0u) + ((cd && xo) ? 4u : 0u);
  return ((uint32_t)sizeof(WalkEntry) + nw * 4u * P.words + 15u) & ~15u;
}

int walk_max_blocks_per_sm(const KParams& P, uint32_t threads, uint32_t smem_bytes) {
  int n = 0;
  SplitFn fn = pick_walk(P);
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, (int)threads, smem_bytes) != cudaSuccess) return 1;
  return n;
}

cudaError_t launch_classify(const KParams& P, uint32_t sm_count, cudaStream_t stream) {
  const uint32_t R = 32u >> P.wq_log2;
  const uint32_t trips = (P.n_runs + R - 1) / R;
  uint32_t blocks = (trips + 7) / 8;            // 8 warps per block, one trip per warp per loop iteration
  const uint32_t cap = sm_count * 8u;            // resident blocks; each loops over its share of the trips
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  pick_classify(P)<<<blocks, 256, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_walk(const KParams& P, uint32_t grid, uint32_t smem_bytes, cudaStream_t stream) {
  SplitFn fn = pick_walk(P);
  static SplitFn configured[8][4] = {};
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

}  // namespace bf
