// frontier_kernel.cu — the StoryRun ready-frontier pass for sm_100a.
//
// One WARP evaluates one StoryRun per trip of a persistent loop.  For every run the
// warp's elected lane issues two cp.async.bulk (TMA) copies — the run's dynamic state
// record and its Story topology record (CSR + bit-sliced static flags) — into a
// per-warp, multi-stage shared-memory ring guarded by mbarriers; the copies of the
// next runs are in flight while the current one is evaluated, so HBM streams at full
// rate with no register staging.
//
// Evaluation is BIT-SLICED: the 4-bit phase codes of a run are transposed into four
// bit planes (one u32 word = 32 steps), after which every classification of
// buildStateMaps (dag.go:3358-3391), the gate/sleep/wait rewrite (dag.go:1455-1547),
// fail-fast / compensation marking and group selection (dag.go:422-511) are a
// handful of LOP3s per 32 steps, held by lanes 0..W-1.  The dependency walk of
// findReadySteps (dag.go:2711-2733) then runs one step per lane over the CSR row in
// shared memory against a per-step status byte, and __ballot_sync folds the 32
// per-step verdicts of a trip straight into one word of the ready / skip bit masks.
//
// Integer only; no tensor cores.  HBM-bound: ~3 KB in, 80 B out per run at the
// BASELINE configuration.  See DESIGN.md for the roofline accounting.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bobrafrontier.h"
#include "device_record.h"

namespace bf {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

template <int IMM>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
  return r;
}
// 16-entry boolean table over a bit-sliced 4-bit code: 3 LOP3 for 32 steps.
template <uint32_t T16>
__device__ __forceinline__ uint32_t plut(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
  const uint32_t lo = lop3<(T16 & 0xFF)>(p2, p1, p0);
  const uint32_t hi = lop3<((T16 >> 8) & 0xFF)>(p2, p1, p0);
  return lop3<0xCA>(p3, hi, lo);  // p3 ? hi : lo
}
// set the code of the steps in mask m to the constant CODE
template <int CODE>
__device__ __forceinline__ void pset(uint32_t m, uint32_t& p0, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  p0 = (CODE & 1) ? (p0 | m) : (p0 & ~m);
  p1 = (CODE & 2) ? (p1 | m) : (p1 & ~m);
  p2 = (CODE & 4) ? (p2 | m) : (p2 & ~m);
  p3 = (CODE & 8) ? (p3 | m) : (p3 & ~m);
}

// gather bits 0,4,8,..,28 of x into the low byte
__device__ __forceinline__ uint32_t squeeze4(uint32_t x) {
  x = (x | (x >> 3)) & 0x03030303u;
  x = (x | (x >> 6)) & 0x000F000Fu;
  x = (x | (x >> 12)) & 0xFFu;
  return x;
}
// gather bits 0,2,4,..,30 of x into the low half
__device__ __forceinline__ uint32_t squeeze2(uint32_t x) {
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0xFFFFu;
  return x;
}
// inverse of squeeze4: byte -> bits 0,4,..,28
__device__ __forceinline__ uint32_t spread4(uint32_t x) {
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}
// low nibble -> one 0/1 byte per bit
__device__ __forceinline__ uint32_t bits4_to_bytes(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ uint32_t get_nibble(const uint8_t* base, uint32_t i) {
  uint32_t v = (base[i >> 1] >> ((i & 1u) * 4u)) & 0xFu;
  return v == 15u ? 0u : v;
}

struct RunCtx {
  uint32_t lane;
  uint32_t Wt;      // words of this topology
  uint32_t S;
  const uint8_t* sr;  // state record (smem)
  const uint8_t* tr;  // topology record (smem)
  const uint16_t* row_ptr;
  const uint16_t* col;
  uint8_t* st;       // status bytes [32*Wt]
};

// Stage D: one step per lane per trip; returns via smem words.  FIXUP adds the
// "Failed earlier in this same loop" visibility rule (dag.go:2744/2810 mutate
// stepStates while `completed` stays as built at :497).
template <bool FIXUP>
__device__ __forceinline__ void walk_deps(const RunCtx& c, const uint32_t* mCAND, const uint32_t* mFAIL,
                                          uint32_t failed_class, uint32_t& met_w, uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  for (uint32_t j = 0; j < c.Wt; ++j) {
    const uint32_t candw = mCAND[j];
    if (candw == 0) continue;  // warp-uniform
    const uint32_t i = j * 32 + c.lane;
    const bool cand = (candw >> c.lane) & 1u;
    uint32_t acc = 0;
    if (cand) {
      uint32_t e = c.row_ptr[i];
      const uint32_t e1 = c.row_ptr[i + 1];
      if (!FIXUP) {
        for (; e + 4 <= e1; e += 4) {
          const uint32_t a = c.col[e], b = c.col[e + 1], cc = c.col[e + 2], d = c.col[e + 3];
          acc |= c.st[a] | c.st[b] | c.st[cc] | c.st[d];
        }
        for (; e < e1; ++e) acc |= c.st[c.col[e]];
      } else {
        for (; e < e1; ++e) {
          const uint32_t d = c.col[e];
          uint32_t s = c.st[d];
          if (d < i && ((mFAIL[d >> 5] >> (d & 31u)) & 1u)) s = failed_class;
          acc |= s;
        }
      }
    }
    const uint32_t fdb = __ballot_sync(0xffffffffu, cand && (acc & 2u));
    const uint32_t metb = __ballot_sync(0xffffffffu, cand && acc == 0u);
    if (c.lane == j) {
      fd_w = fdb;
      met_w = metb;
    }
  }
}

extern __shared__ __align__(128) uint8_t smem_raw[];

__global__ void __launch_bounds__(512) frontier_kernel(const KParams P) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t WPB = P.warps_per_block;
  const uint32_t ST = P.stages;
  const uint32_t FULL = 0xffffffffu;

  // ---- shared memory carve-up: [block counters 128 B][warp regions] ----
  unsigned long long* blk_counts = reinterpret_cast<unsigned long long*>(smem_raw);
  const uint32_t ring_bytes = ST * P.stage_bytes;
  const uint32_t per_warp = ring_bytes + P.work_bytes + 64;  // + mbarriers (<= 8 stages)
  uint8_t* wbase = smem_raw + 128 + warp * per_warp;
  uint32_t* work = reinterpret_cast<uint32_t*>(wbase + ring_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + ring_bytes + P.work_bytes);

  if (threadIdx.x < 4) blk_counts[threadIdx.x] = 0ull;
  if (lane == 0) {
    for (uint32_t s = 0; s < ST; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();

  const uint32_t gw = blockIdx.x * WPB + warp;
  const uint32_t G = gridDim.x * WPB;
  const uint32_t N = P.n_runs;
  const uint32_t my_runs = gw < N ? (N - gw + G - 1) / G : 0;

  const bool has_cond = P.off_cond != BF_OFF_NONE;
  const bool has_dec = P.off_decision != BF_OFF_NONE;
  const bool has_child = P.off_child != BF_OFF_NONE;
  const bool fixpoint = (P.flags & BF_EVAL_FIXPOINT) != 0;

  // ---- producer state (lane 0): two-level prefetch of slot id -> slot entry ----
  uint32_t ni = 0;          // next issue index
  Slot ent_q = {0, 0, 0};   // entry for issue index ni
  uint32_t sid_q = 0;       // slot id for issue index ni+1
  uint32_t ok_bits = 0;     // bit s: stage s holds a staged topology record
  auto load_sid = [&](uint32_t n) -> uint32_t {
    if (n >= my_runs) return 0xFFFFFFFFu;
    const uint8_t* hdr = P.state + (size_t)(gw + n * G) * P.state_stride;
    return __ldg(reinterpret_cast<const uint32_t*>(hdr));
  };
  auto load_ent = [&](uint32_t sid) -> Slot {
    Slot e = {0, 0, 0};
    if (sid < P.n_slots) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.slots + sid));
      e.addr = (uint64_t)v.x | ((uint64_t)v.y << 32);
      e.bytes = v.z;
      e.S = v.w;
    }
    return e;
  };
  auto issue = [&]() {
    // copies for issue index ni into stage ni % ST
    if (ni < my_runs) {
      const uint32_t s = ni % ST;
      uint8_t* buf = wbase + s * P.stage_bytes;
      const uint32_t bar = smem_u32(&bars[s]);
      const uint8_t* src_state = P.state + (size_t)(gw + ni * G) * P.state_stride;
      const bool ok = ent_q.addr != 0 && ent_q.bytes <= P.topo_buf_bytes;
      const uint32_t tb = ok ? ent_q.bytes : 0u;
      ok_bits = ok ? (ok_bits | (1u << s)) : (ok_bits & ~(1u << s));
      mbar_expect_tx(bar, P.state_stride + tb);
      bulk_g2s(smem_u32(buf), src_state, P.state_stride, bar);
      if (ok) bulk_g2s(smem_u32(buf + P.state_stride), reinterpret_cast<const void*>(ent_q.addr), tb, bar);
    }
    ent_q = load_ent(sid_q);
    sid_q = load_sid(ni + 2);
    ++ni;
  };
  if (lane == 0) {
    ent_q = load_ent(load_sid(0));
    sid_q = load_sid(1);
    for (uint32_t s = 0; s < ST; ++s) issue();
  }

  // per-warp running totals (lane-uniform)
  uint32_t tot_ready = 0, tot_skip = 0, tot_exp = 0, tot_evals = 0;

  for (uint32_t k = 0; k < my_runs; ++k) {
    const uint32_t s = k % ST;
    const uint32_t r = gw + k * G;
    mbar_wait(smem_u32(&bars[s]), (k / ST) & 1u);

    const uint8_t* sr = wbase + s * P.stage_bytes;
    const uint8_t* tr = sr + P.state_stride;
    const uint32_t rflags = sr[4];
    const uint64_t registered = *reinterpret_cast<const uint64_t*>(sr + 8);
    // Was a topology staged for this run?  (the producer lane knows)
    bool topo_ok = (__shfl_sync(FULL, ok_bits, 0) >> s) & 1u;
    if (topo_ok) topo_ok = reinterpret_cast<const TopoHeader*>(tr)->W <= P.words;
    uint8_t* rr = P.result + (size_t)r * P.result_stride;

    if (!topo_ok) {  // dead / out-of-range slot: empty result, summary all-ones
      for (uint32_t x = lane; x < P.result_stride / 4; x += 32)
        reinterpret_cast<uint32_t*>(rr)[x] = x == 0 ? 0xFFFFFFFFu : 0u;
      if (P.exp_counts && lane == 0) P.exp_counts[r] = 0;
      __syncwarp();
      if (lane == 0) issue();
      continue;
    }

    const TopoHeader th = *reinterpret_cast<const TopoHeader*>(tr);
    const uint32_t S = th.S, Wt = th.W;
    const uint32_t* splanes = reinterpret_cast<const uint32_t*>(tr + th.off_planes);

    uint32_t* pl = work;                 // [4][Wt] phase planes
    uint32_t* cpl = pl + 4 * P.words;    // [2][Wt] cond planes
    uint32_t* dpl = cpl + 2 * P.words;   // [2][Wt] decision planes
    uint32_t* mU = dpl + 2 * P.words;
    uint32_t* mFD = mU + P.words;
    uint32_t* mCAND = mFD + P.words;
    uint32_t* mREADY = mCAND + P.words;
    uint32_t* mFAIL = mREADY + P.words;
    uint8_t* st = reinterpret_cast<uint8_t*>(work) + ((52u * P.words + 15u) & ~15u);  // 13 word arrays above

    // ---------------- stage A: transpose packed codes into bit planes ----------------
    {
      const uint32_t* pw = reinterpret_cast<const uint32_t*>(sr + P.off_phase);
      for (uint32_t m = lane; m < 4 * Wt; m += 32) {
        uint32_t w = pw[m];
        uint32_t f = w & (w >> 1);
        f = f & (f >> 2) & 0x11111111u;  // nibble == 15 (reserved) -> 0
        w &= ~(f * 15u);
#pragma unroll
        for (int b = 0; b < 4; ++b)
          reinterpret_cast<uint8_t*>(pl + b * Wt)[m] = (uint8_t)squeeze4((w >> b) & 0x11111111u);
      }
      if (has_cond) {
        const uint32_t* cw = reinterpret_cast<const uint32_t*>(sr + P.off_cond);
        for (uint32_t m = lane; m < 2 * Wt; m += 32) {
          const uint32_t w = cw[m];
          reinterpret_cast<uint16_t*>(cpl)[m] = (uint16_t)squeeze2(w & 0x55555555u);
          reinterpret_cast<uint16_t*>(cpl + Wt)[m] = (uint16_t)squeeze2((w >> 1) & 0x55555555u);
        }
      }
      if (has_dec) {
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(sr + P.off_decision);
        for (uint32_t m = lane; m < 2 * Wt; m += 32) {
          const uint32_t w = dw[m];
          reinterpret_cast<uint16_t*>(dpl)[m] = (uint16_t)squeeze2(w & 0x55555555u);
          reinterpret_cast<uint16_t*>(dpl + Wt)[m] = (uint16_t)squeeze2((w >> 1) & 0x55555555u);
        }
      }
    }
    __syncwarp();

    // static planes + validity for word `lane`
    const bool act = lane < Wt;
    uint32_t t0 = 0, t1 = 0, t2 = 0, AF = 0, TS = 0, HASIF = 0, G1 = 0, G2 = 0, VALID = 0;
    uint32_t c0 = 0, c1 = 0, d0 = 0, d1 = 0;
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;  // input planes (for the "changed" flag)
    if (act) {
      t0 = splanes[PL_T0 * Wt + lane]; t1 = splanes[PL_T1 * Wt + lane]; t2 = splanes[PL_T2 * Wt + lane];
      AF = splanes[PL_AF * Wt + lane]; TS = splanes[PL_TS * Wt + lane]; HASIF = splanes[PL_HASIF * Wt + lane];
      G1 = splanes[PL_G1 * Wt + lane]; G2 = splanes[PL_G2 * Wt + lane];
      const uint32_t rem = S - lane * 32;
      VALID = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
      if (has_cond) { c0 = cpl[lane]; c1 = cpl[Wt + lane]; }
      if (has_dec) { d0 = dpl[lane]; d1 = dpl[Wt + lane]; }
      q0 = pl[lane] & VALID; q1 = pl[Wt + lane] & VALID; q2 = pl[2 * Wt + lane] & VALID; q3 = pl[3 * Wt + lane] & VALID;
      pl[lane] = q0; pl[Wt + lane] = q1; pl[2 * Wt + lane] = q2; pl[3 * Wt + lane] = q3;
    }
    const uint32_t GM = VALID & ~G1 & ~G2;
    const uint32_t SYNC_T = t0 & (t1 | t2);            // sleep(3) | wait(5) | gate(7)
    const uint32_t T_COND = t0 & ~t1 & ~t2;            // 1
    const uint32_t T_PAR = ~t0 & t1 & ~t2 & VALID;     // 2
    const uint32_t T_STOP = ~t0 & ~t1 & t2 & VALID;    // 4

    const bool fail_fast = rflags & BF_RF_FAIL_FAST;
    const bool realtime = rflags & BF_RF_REALTIME;
    const bool topo_term = rflags & BF_RF_TOPOLOGY_TERMINATED;
    const bool host_group = rflags & BF_RF_HOST_GROUP;

    uint32_t acc_ready = 0, acc_skip = 0, acc_fail = 0, acc_needs = 0, acc_skipdep = 0;
    uint32_t summary = 0, iters = 0;
    const uint32_t cap = fixpoint ? (P.max_iter ? P.max_iter : S + 1) : 1u;

    RunCtx rc;
    rc.lane = lane; rc.Wt = Wt; rc.S = S; rc.sr = sr; rc.tr = tr;
    rc.row_ptr = reinterpret_cast<const uint16_t*>(tr + sizeof(TopoHeader));
    rc.col = reinterpret_cast<const uint16_t*>(tr + th.off_col);
    rc.st = st;

    for (uint32_t it = 0; it < cap; ++it) {
      ++iters;
      // ---------------- stage H: parallel join (dag.go:1131-1198) ----------------
      if (has_child && th.P != 0) {
        __syncwarp();
        const ParDesc* pd = reinterpret_cast<const ParDesc*>(tr + th.off_par);
        const uint8_t* child = sr + P.off_child;
        for (uint32_t q = 0; q < th.P; ++q) {
          const ParDesc d = pd[q];
          if (!((registered >> q) & 1ull)) continue;
          const uint32_t wj = d.step >> 5, wb = d.step & 31u;
          const uint32_t ph = ((pl[wj] >> wb) & 1u) | (((pl[Wt + wj] >> wb) & 1u) << 1) |
                              (((pl[2 * Wt + wj] >> wb) & 1u) << 2) | (((pl[3 * Wt + wj] >> wb) & 1u) << 3);
          if (ph == 0 || ((BF_LUT_TERMINAL >> ph) & 1u)) continue;
          const uint32_t* allow = reinterpret_cast<const uint32_t*>(tr + d.allow_off);
          bool all_done = true, any_failed = false;
          for (uint32_t b = lane; b < d.branches; b += 32) {
            const uint32_t cp = get_nibble(child, d.child_first + b);
            const bool done = cp != 0 && ((BF_LUT_TERMINAL >> cp) & 1u);
            const bool okc = cp == BF_PHASE_SUCCEEDED || cp == BF_PHASE_SKIPPED || ((allow[b >> 5] >> (b & 31u)) & 1u);
            all_done = all_done && done;
            any_failed = any_failed || (done && !okc);
          }
          all_done = __all_sync(FULL, all_done);
          any_failed = __any_sync(FULL, any_failed);
          if (all_done && lane == 0) {
            uint32_t a = pl[wj], b = pl[Wt + wj], c = pl[2 * Wt + wj], e = pl[3 * Wt + wj];
            const uint32_t m = 1u << wb;
            if (any_failed) pset<BF_PHASE_FAILED>(m, a, b, c, e); else pset<BF_PHASE_SUCCEEDED>(m, a, b, c, e);
            pl[wj] = a; pl[Wt + wj] = b; pl[2 * Wt + wj] = c; pl[3 * Wt + wj] = e;
          }
          __syncwarp();
        }
      }

      // ---------------- stage B: classification on planes (lanes < Wt) ----------------
      uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
      if (act) { p0 = pl[lane]; p1 = pl[Wt + lane]; p2 = pl[2 * Wt + lane]; p3 = pl[3 * Wt + lane]; }

      if (has_dec) {  // stage G: gate / sleep / wait sync (dag.go:1469-1533, 1235-1277, 1327-1437)
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

      uint32_t TERM = plut<BF_LUT_TERMINAL>(p0, p1, p2, p3);
      uint32_t COMPL = plut<BF_LUT_COMPLETED0>(p0, p1, p2, p3) | (TERM & AF);
      uint32_t RUN = plut<BF_LUT_RUNNING>(p0, p1, p2, p3);
      uint32_t RUNQ = plut<BF_LUT_RUNNING_Q>(p0, p1, p2, p3);
      uint32_t FAILED = TERM & ~COMPL;
      uint32_t group;
      uint32_t sum = 0;
      if (host_group) {
        group = (rflags >> BF_RF_HOST_GROUP_SHIFT) & 3u;
      } else {  // stage I: dag.go:422-495
        bool amf = __any_sync(FULL, (FAILED & GM) != 0);
        if (fail_fast && amf) {  // markFailFastSkipped dag.go:3289-3312
          const uint32_t m = GM & ~COMPL & ~RUNQ & ~TERM;
          pset<BF_PHASE_SKIPPED>(m, p0, p1, p2, p3);
          TERM |= m; COMPL |= m; RUN &= ~m;
        }
        bool main_done = th.n_main == 0 || !__any_sync(FULL, (GM & ~(COMPL | FAILED)) != 0);
        if (!main_done && realtime && topo_term) {  // dag.go:436-464
          const uint32_t m = GM & (p0 | p1 | p2 | p3) & ~TERM;
          pset<BF_PHASE_FAILED>(m, p0, p1, p2, p3);
          TERM |= m; COMPL |= m & AF; FAILED |= m & ~AF; RUN &= ~m; RUNQ &= ~m;
          main_done = true;
          amf = __any_sync(FULL, (FAILED & GM) != 0);
        }
        if (main_done && !amf && th.n_comp != 0) {  // markCompensationsSkipped dag.go:3314-3342
          const uint32_t m = G1 & ~COMPL & ~RUN & ~FAILED & ~TERM;
          pset<BF_PHASE_SKIPPED>(m, p0, p1, p2, p3);
          TERM |= m; COMPL |= m;
        }
        const bool comp_done = th.n_comp == 0 || !__any_sync(FULL, (G1 & ~(COMPL | FAILED)) != 0);
        const bool final_done = th.n_final == 0 || !__any_sync(FULL, (G2 & ~(COMPL | FAILED)) != 0);
        const bool acf = __any_sync(FULL, (FAILED & G1) != 0);
        const bool aff = __any_sync(FULL, (FAILED & G2) != 0);
        if (!main_done) group = BF_GROUP_MAIN;
        else if (amf && th.n_comp != 0 && !comp_done) group = BF_GROUP_COMPENSATION;
        else if (th.n_final != 0 && !final_done) group = BF_GROUP_FINALLY;
        else group = BF_GROUP_DONE;
        sum = (main_done ? BF_SUM_MAIN_DONE : 0u) | (amf ? BF_SUM_MAIN_FAILED : 0u) |
              (comp_done ? BF_SUM_COMP_DONE : 0u) | (final_done ? BF_SUM_FINAL_DONE : 0u) |
              (acf ? BF_SUM_COMP_FAILED : 0u) | (aff ? BF_SUM_FINAL_FAILED : 0u);
      }
      summary = sum | group;

      uint32_t it_ready = 0, it_skip = 0, it_fail = 0;
      if (group != BF_GROUP_DONE) {
        // ------------- D-prep: dependency classes under this pass's policy (dag.go:499-502) -------------
        const bool allow_failed = group != BF_GROUP_MAIN;
        const bool skip_on_failed = group == BF_GROUP_MAIN && !fail_fast;
        const uint32_t GSEL = group == BF_GROUP_MAIN ? GM : (group == BF_GROUP_COMPENSATION ? G1 : G2);
        const uint32_t SAT = COMPL | (realtime ? plut<BF_LUT_RT_SAT>(p0, p1, p2, p3) : 0u) | (allow_failed ? TERM : 0u);
        const uint32_t FD = skip_on_failed ? (TERM & ~SAT) : 0u;
        const uint32_t CAND = GSEL & ~COMPL & ~RUNQ & ~TERM;
        if (act) { mU[lane] = ~SAT; mFD[lane] = FD; mCAND[lane] = CAND; }
        __syncwarp();
        // ------------- stage C: masks -> one status byte per step (bit0 unmet, bit1 failed-dep) -------------
        for (uint32_t m = lane; m < 4 * Wt; m += 32) {
          const uint32_t ub = reinterpret_cast<const uint8_t*>(mU)[m];
          const uint32_t fb = reinterpret_cast<const uint8_t*>(mFD)[m];
          uint2 v;
          v.x = bits4_to_bytes(ub & 0xFu) | (bits4_to_bytes(fb & 0xFu) << 1);
          v.y = bits4_to_bytes(ub >> 4) | (bits4_to_bytes(fb >> 4) << 1);
          reinterpret_cast<uint2*>(st)[m] = v;
        }
        __syncwarp();
        // ------------- stage D: walk the needs rows (dag.go:2711-2733) -------------
        uint32_t met_w, fd_w;
        walk_deps<false>(rc, mCAND, mFAIL, 0u, met_w, fd_w);
        uint32_t ready_w = met_w & ~c0 & ~c1;   // BF_COND_PASS
        uint32_t skipc_w = met_w & c0 & ~c1;    // BF_COND_SKIP
        uint32_t fail_w = met_w & c0 & c1;      // BF_COND_FAIL (HOLD = c1 & ~c0: nothing)
        if (__any_sync(FULL, fail_w != 0)) {
          // A step set Failed inside the loop is visible to LATER steps of the list only; iterate
          // to the unique fixed point (at most one extra round per chained failure).
          const uint32_t fclass = allow_failed ? 0u : (skip_on_failed ? 3u : 1u);
          for (uint32_t round = 0; round <= S; ++round) {
            __syncwarp();
            if (act) mFAIL[lane] = fail_w;
            __syncwarp();
            walk_deps<true>(rc, mCAND, mFAIL, fclass, met_w, fd_w);
            const uint32_t nf = met_w & c0 & c1;
            const bool same = !__any_sync(FULL, nf != fail_w);
            fail_w = nf;
            if (same) break;
          }
          ready_w = met_w & ~c0 & ~c1;
          skipc_w = met_w & c0 & ~c1;
        }
        it_ready = ready_w;
        it_skip = fd_w | skipc_w;
        it_fail = fail_w;
        acc_needs |= realtime ? 0u : (met_w & HASIF);
        acc_skipdep |= fd_w;
        pset<BF_PHASE_FAILED>(fail_w, p0, p1, p2, p3);  // dag.go:2745-2747, 2810-2812
      }
      acc_ready |= it_ready; acc_skip |= it_skip; acc_fail |= it_fail;

      bool more = false;
      if (fixpoint && group != BF_GROUP_DONE) {
        // ------------- launch effects (dag.go:1735-1775, step_executor.go:132-185) -------------
        pset<BF_PHASE_SKIPPED>(it_skip, p0, p1, p2, p3);
        const uint32_t none_or_q = ~(p0 | p1 | p2 | p3) | (~p0 & p1 & p2 & p3);  // code 0 or 14
        const uint32_t rd = it_ready;
        pset<BF_PHASE_SUCCEEDED>(rd & T_COND, p0, p1, p2, p3);
        pset<BF_PHASE_PAUSED>(rd & SYNC_T, p0, p1, p2, p3);
        pset<BF_PHASE_RUNNING>(rd & T_PAR, p0, p1, p2, p3);
        pset<BF_PHASE_RUNNING>(rd & ~T_COND & ~SYNC_T & ~T_PAR & ~T_STOP & none_or_q, p0, p1, p2, p3);
        const bool progress = __any_sync(FULL, (it_ready | it_skip) != 0);
        const bool stop_ready = __any_sync(FULL, (it_ready & T_STOP) != 0);
        more = progress && !stop_ready;
      }
      __syncwarp();
      if (act) { pl[lane] = p0; pl[Wt + lane] = p1; pl[2 * Wt + lane] = p2; pl[3 * Wt + lane] = p3; }
      __syncwarp();
      if (!more) break;
    }

    // ---------------- stage E: result record ----------------
    uint32_t fp0 = 0, fp1 = 0, fp2 = 0, fp3 = 0;
    if (act) { fp0 = pl[lane]; fp1 = pl[Wt + lane]; fp2 = pl[2 * Wt + lane]; fp3 = pl[3 * Wt + lane]; }
    const bool changed = __any_sync(FULL, ((fp0 ^ q0) | (fp1 ^ q1) | (fp2 ^ q2) | (fp3 ^ q3)) != 0);
    const uint32_t n_ready = __reduce_add_sync(FULL, (uint32_t)__popc(acc_ready));
    const uint32_t n_skip = __reduce_add_sync(FULL, (uint32_t)__popc(acc_skip));
    uint32_t n_exp = 0;
    if (th.P != 0) {
      if (act) mREADY[lane] = acc_ready;
      __syncwarp();
      const ParDesc* pd = reinterpret_cast<const ParDesc*>(tr + th.off_par);
      uint32_t mine = 0;
      for (uint32_t q = lane; q < th.P; q += 32) {
        const uint32_t stp = pd[q].step;
        if ((mREADY[stp >> 5] >> (stp & 31u)) & 1u) mine += pd[q].branches;
      }
      n_exp = __reduce_add_sync(FULL, mine);
    }
    summary |= (changed ? BF_SUM_PHASE_CHANGED : 0u) | (iters << BF_SUM_ITER_SHIFT);
    if (lane == 0) {
      *reinterpret_cast<uint4*>(rr) = make_uint4(summary, n_ready, n_skip, n_exp);
      if (P.exp_counts) P.exp_counts[r] = n_exp;
    }
    if (lane < P.words) {
      reinterpret_cast<uint32_t*>(rr + P.off_ready)[lane] = acc_ready;
      reinterpret_cast<uint32_t*>(rr + P.off_skip)[lane] = acc_skip;
      if (P.off_fail != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_fail)[lane] = acc_fail;
      if (P.off_needs_cond != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_needs_cond)[lane] = acc_needs;
      if (P.off_skip_dep != BF_OFF_NONE) reinterpret_cast<uint32_t*>(rr + P.off_skip_dep)[lane] = acc_skipdep;
    }
    if (P.off_phase_out != BF_OFF_NONE) {
      uint32_t* po = reinterpret_cast<uint32_t*>(rr + P.off_phase_out);
      for (uint32_t m = lane; m < 4 * P.words; m += 32) {
        uint32_t w = 0;
        if (m < 4 * Wt) {
#pragma unroll
          for (int b = 0; b < 4; ++b) w |= spread4(reinterpret_cast<const uint8_t*>(pl + b * Wt)[m]) << b;
        }
        po[m] = w;
      }
    }
    for (uint32_t x = P.result_tail / 4 + lane; x < P.result_stride / 4; x += 32) reinterpret_cast<uint32_t*>(rr)[x] = 0u;
    tot_ready += n_ready; tot_skip += n_skip; tot_exp += n_exp; tot_evals += S;

    __syncwarp();  // every lane is done with this stage's buffers
    if (lane == 0) issue();
  }

  // ---- counters: warp -> block (shared atomics) -> one global atomic per block ----
  if (P.counts) {
    if (lane == 0 && my_runs != 0) {
      atomicAdd(&blk_counts[0], (unsigned long long)tot_ready);
      atomicAdd(&blk_counts[1], (unsigned long long)tot_skip);
      atomicAdd(&blk_counts[2], (unsigned long long)tot_exp);
      atomicAdd(&blk_counts[3], (unsigned long long)tot_evals);
    }
    __syncthreads();
    if (threadIdx.x < 4 && blk_counts[threadIdx.x] != 0ull) atomicAdd(&P.counts[threadIdx.x], blk_counts[threadIdx.x]);
  }
}

// Host-side launcher (called from abi.cu).
cudaError_t launch_frontier(const KParams& P, uint32_t grid, uint32_t smem_bytes, cudaStream_t stream) {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    e = cudaFuncSetAttribute(frontier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  frontier_kernel<<<grid, P.warps_per_block * 32, smem_bytes, stream>>>(P);
  return cudaGetLastError();
}


int frontier_max_blocks_per_sm(uint32_t threads, uint32_t smem_bytes) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaFuncSetAttribute(frontier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, frontier_kernel, (int)threads, smem_bytes) != cudaSuccess) return 1;
  return n;
}

}  // namespace bf
