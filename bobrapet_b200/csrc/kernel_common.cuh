// kernel_common.cuh — PTX / bit-plane helpers shared by the frontier kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bobrafrontier.h"
#include "device_record.h"

namespace bf {

#define DI __device__ __forceinline__
constexpr uint32_t FULL = 0xffffffffu;

// ------------------------------------------------------------------ PTX helpers
DI uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
DI void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
DI void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
DI uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
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
DI void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
DI void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
DI void sts_zero16(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory"); }
DI void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <int IMM>
DI uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
  return r;
}
// 16-entry boolean table over a bit-sliced 4-bit code: 3 LOP3 for 32 steps.
template <uint32_t T16>
DI uint32_t plut(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
  const uint32_t lo = lop3<(T16 & 0xFF)>(p2, p1, p0);
  const uint32_t hi = lop3<((T16 >> 8) & 0xFF)>(p2, p1, p0);
  return lop3<0xCA>(p3, hi, lo);  // p3 ? hi : lo
}
// set the code of the steps in mask m to the constant CODE
template <int CODE>
DI void pset(uint32_t m, uint32_t& p0, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  p0 = (CODE & 1) ? (p0 | m) : (p0 & ~m);
  p1 = (CODE & 2) ? (p1 | m) : (p1 & ~m);
  p2 = (CODE & 4) ? (p2 | m) : (p2 & ~m);
  p3 = (CODE & 8) ? (p3 | m) : (p3 & ~m);
}
DI uint32_t squeeze4(uint32_t x) {  // bits 0,4,..,28 -> low byte
  x = (x | (x >> 3)) & 0x03030303u;
  x = (x | (x >> 6)) & 0x000F000Fu;
  x = (x | (x >> 12)) & 0xFFu;
  return x;
}
DI uint32_t squeeze2(uint32_t x) {  // bits 0,2,..,30 -> low half
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0xFFFFu;
  return x;
}
DI uint32_t spread4(uint32_t x) {  // inverse of squeeze4
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}
DI uint32_t bits4_to_bytes(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }  // 4 bits -> 4 0/1 bytes
DI uint32_t get_nibble(const uint8_t* base, uint32_t i) {
  const uint32_t v = (base[i >> 1] >> ((i & 1u) * 4u)) & 0xFu;
  return v == 15u ? 0u : v;
}
DI uint32_t redux_or(uint32_t v) { return __reduce_or_sync(FULL, v); }
DI uint32_t redux_add(uint32_t v) { return __reduce_add_sync(FULL, v); }

DI uint32_t bmsk_clamp(uint32_t pos, uint32_t width) {
  uint32_t r;
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(pos), "r"(width));
  return r;
}

// ---- shared-space accessors on 32-bit addresses (keeps the walk's address math in 32 bits) ----
DI uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
DI uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
DI uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
DI uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
DI void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
DI void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
DI void sts_v2(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
// generic pointer of a shared-window address (rare paths only: parallel join, in-loop failure fix-up)
DI const uint8_t* gptr(uint32_t a) { return static_cast<const uint8_t*>(__cvta_shared_to_generic(a)); }
// keep a loop-invariant value in a register instead of letting the compiler rematerialise it from constants
DI uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }

// ------------------------------------------------------------------ stage D: the dependency walk (one run per warp)
// One step per lane per trip over the CSR rows of the candidate steps, visiting only the 32-step words that hold
// a candidate.  Status byte of a dependency: bit0 = not satisfied, bit1 = failed dependency.  The first four deps of
// a row are fetched branch-free (index clamped into the status array, verdict masked by the row length); longer rows
// exist only when the topology header says so.  FIXUP adds the "set Failed earlier in this same loop" visibility
// rule (dag.go:2744/2810 mutate stepStates while `completed` stays as built at :497).
template <bool FIXUP>
DI void walk_rows(uint32_t lane, uint32_t CAND, uint32_t zidx, uint32_t max_deg, const uint16_t* __restrict__ row_ptr,
                  const uint16_t* __restrict__ col, const uint8_t* __restrict__ st, const uint32_t* mFAIL,
                  uint32_t failed_class, uint32_t& met_w, uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  uint32_t todo = __ballot_sync(FULL, CAND != 0);  // words with at least one candidate step
  while (todo) {
    const uint32_t j = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t candw = __shfl_sync(FULL, CAND, j);
    const bool cand = (candw >> lane) & 1u;
    const uint32_t i = j * 32 + lane;
    uint32_t e0 = 0, n = 0;
    if (cand) {
      e0 = row_ptr[i];
      n = row_ptr[i + 1] - e0;
    }
    const uint16_t* cp = col + e0;
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
      uint32_t acc = 0;
      for (uint32_t e = 0; e < n; ++e) {
        const uint32_t d = cp[e];
        uint32_t sb = st[d];
        if (d < i && ((mFAIL[d >> 5] >> (d & 31u)) & 1u)) sb = failed_class;
        acc |= sb;
      }
      unmet = (acc & 1u) != 0;
      fdp = (acc & 2u) != 0;
    }
    const uint32_t fdb = __ballot_sync(FULL, fdp);
    const uint32_t metb = __ballot_sync(FULL, cand && !unmet);
    if (lane == j) {
      fd_w = fdb;
      met_w = metb;
    }
  }
}


// Same walk with explicit 32-bit shared addresses (ld.shared), for kernels whose CSR / status arrays are given
// as shared-window addresses: keeps every address computation a single 32-bit add.
// NEED_FD = false: the policy has no failed-dependency class (status bit 1 is never set), fd_w stays 0.
// LONG_ROWS = false: no row of this topology has more than 4 dependencies (TopoHeader.max_deg), the tail loop is
// compiled out instead of being tested per lane.
template <bool NEED_FD, bool LONG_ROWS>
DI void walk_rows_s(uint32_t lane, uint32_t CAND, uint32_t rp_addr, uint32_t col_addr,
                    uint32_t st_addr, uint32_t& met_w, uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  uint32_t todo = __ballot_sync(FULL, CAND != 0);
  const uint32_t rp_lane = rp_addr + lane * 2u;
  while (todo) {
    const uint32_t j = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t candw = __shfl_sync(FULL, CAND, j);
    const bool cand = (candw >> lane) & 1u;
    uint32_t e0 = 0, n = 0;
    if (cand) {
      const uint32_t a = rp_lane + j * 64u;
      e0 = lds_u16(a);
      n = lds_u16(a + 2u) - e0;
    }
    const uint32_t cpa = col_addr + e0 * 2u;
    // The four fetches may run past the row (and, for the last rows, past E): every u16 there is another row's
    // entry or the record's zero padding (device_record.h), i.e. a valid step index; the verdict is masked below.
    uint32_t x0, x1, x2, x3;
    asm volatile("ld.shared.u16 %0, [%4];\n\tld.shared.u16 %1, [%4+2];\n\tld.shared.u16 %2, [%4+4];\n\tld.shared.u16 %3, [%4+6];"
                 : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(cpa));
    const uint32_t s0 = lds_u8(st_addr + x0), s1 = lds_u8(st_addr + x1), s2 = lds_u8(st_addr + x2), s3 = lds_u8(st_addr + x3);
    uint32_t w = ((s3 * 256u + s2) * 256u + s1) * 256u + s0;
    w &= bmsk_clamp(0u, n * 8u);
    if (LONG_ROWS) {
      for (uint32_t e = 4; e < n; ++e) w |= lds_u8(st_addr + lds_u16(cpa + e * 2u));
    }
    const uint32_t metb = __ballot_sync(FULL, cand && (w & 0x01010101u) == 0);
    if (NEED_FD) {
      const uint32_t fdb = __ballot_sync(FULL, (w & 0x02020202u) != 0);
      if (lane == j) fd_w = fdb;
    }
    if (lane == j) met_w = metb;
  }
}


// K candidate words at once: K independent load chains in flight (ILP).  All K words exist (the driver below
// only calls it with at least K left), so there is no wasted half trip.
template <int K, bool NEED_FD, bool LONG_ROWS>
DI void walk_group(uint32_t lane, uint32_t CAND, uint32_t& todo, uint32_t rp_lane, uint32_t col_addr, uint32_t st_addr,
                   uint32_t& met_w, uint32_t& fd_w) {
  uint32_t j[K], e[K], n[K], w[K];
  bool c[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    j[k] = __ffs(todo) - 1;
    todo &= todo - 1;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t cw = __shfl_sync(FULL, CAND, j[k]);
    c[k] = (cw >> lane) & 1u;
    e[k] = 0; n[k] = 0;
    if (c[k]) { const uint32_t a = rp_lane + j[k] * 64u; e[k] = lds_u16(a); n[k] = lds_u16(a + 2u) - e[k]; }
  }
  uint32_t x[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // may run past the row (and, for the last rows, past E): every u16 there is another row's entry or the record's
    // zero padding (device_record.h), i.e. a valid step index; the verdict is masked below
    const uint32_t p = col_addr + e[k] * 2u;
    x[k][0] = lds_u16(p); x[k][1] = lds_u16(p + 2u); x[k][2] = lds_u16(p + 4u); x[k][3] = lds_u16(p + 6u);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t s0 = lds_u8(st_addr + x[k][0]), s1 = lds_u8(st_addr + x[k][1]), s2 = lds_u8(st_addr + x[k][2]), s3 = lds_u8(st_addr + x[k][3]);
    w[k] = (((s3 * 256u + s2) * 256u + s1) * 256u + s0) & bmsk_clamp(0u, n[k] * 8u);
    if (LONG_ROWS) {
      const uint32_t p = col_addr + e[k] * 2u;
      for (uint32_t q = 4; q < n[k]; ++q) w[k] |= lds_u8(st_addr + lds_u16(p + q * 2u));
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t m = __ballot_sync(FULL, c[k] && (w[k] & 0x01010101u) == 0);
    if (lane == j[k]) met_w = m;
    if (NEED_FD) {
      const uint32_t f = __ballot_sync(FULL, (w[k] & 0x02020202u) != 0);
      if (lane == j[k]) fd_w = f;
    }
  }
}

// Stage D driver: the candidate words of a run, KMAX at a time while that many are left, then 2, then 1.
// KMAX = 1 is the plain walk; 2 pays at two CTAs per SM (+2 %), 4 where one CTA per SM leaves little else to hide
// latency (S = 1024: +10 %).  Measured on B200, see DESIGN.md section 5.
template <int KMAX, bool NEED_FD, bool LONG_ROWS>
DI void walk_words(uint32_t lane, uint32_t CAND, uint32_t rp_addr, uint32_t col_addr, uint32_t st_addr, uint32_t& met_w,
                   uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  uint32_t todo = __ballot_sync(FULL, CAND != 0);  // words with at least one candidate step
  const uint32_t rp_lane = rp_addr + lane * 2u;
  if (KMAX >= 4)
    while (__popc(todo) >= 4) walk_group<4, NEED_FD, LONG_ROWS>(lane, CAND, todo, rp_lane, col_addr, st_addr, met_w, fd_w);
  if (KMAX >= 2) {
    if (KMAX >= 4) {
      if (__popc(todo) >= 2) walk_group<2, NEED_FD, LONG_ROWS>(lane, CAND, todo, rp_lane, col_addr, st_addr, met_w, fd_w);
    } else {
      while (__popc(todo) >= 2) walk_group<2, NEED_FD, LONG_ROWS>(lane, CAND, todo, rp_lane, col_addr, st_addr, met_w, fd_w);
    }
  }
  if (KMAX >= 2) {
    if (todo) walk_group<1, NEED_FD, LONG_ROWS>(lane, CAND, todo, rp_lane, col_addr, st_addr, met_w, fd_w);
  } else {
    while (todo) walk_group<1, NEED_FD, LONG_ROWS>(lane, CAND, todo, rp_lane, col_addr, st_addr, met_w, fd_w);
  }
}

// ------------------------------------------------------------------ the same walk for packed lanes (R runs per trip)
// One step per lane over the CSR rows of the candidate steps (findReadySteps, dag.go:2711-2733), visiting only
// the 32-step words that hold a candidate, TWO words per loop trip so two independent load chains overlap.
// Status byte of a dependency: bit0 = not satisfied, bit1 = failed dependency.  The first four deps of a row
// are fetched branch-free (index clamped into the status array, verdict masked by the row length); longer rows
// exist only when the topology header says so.
//
// Lane l of the CAND mask = group g = l >> lg (one StoryRun of the trip), word j = l & (2^lg - 1):
//   row_ptr of group g at rp0 + g*topo_buf (u16 entries), status bytes of group g at st0 + (g << (5+lg)),
//   col_addr / meta are per-lane (group-uniform) values fetched from the word's owner lane.
struct WalkOne {
  bool cand, unmet, fdp;
};
DI WalkOne walk_word(uint32_t lane, uint32_t L, bool enable, uint32_t CAND, uint32_t lg, uint32_t col_addr, uint32_t meta,
                     uint32_t rp0, uint32_t topo_buf, uint32_t st0) {
  const uint32_t g = L >> lg, j = L & ((1u << lg) - 1u);
  const uint32_t candw = __shfl_sync(FULL, CAND, L);
  const uint32_t cola = __shfl_sync(FULL, col_addr, L);
  const uint32_t mt = __shfl_sync(FULL, meta, L);  // zidx | max_deg << 16
  const uint32_t zidx = mt & 0xFFFFu;
  const uint32_t rpa = rp0 + g * topo_buf + (j * 32u + lane) * 2u;
  const uint32_t sta = st0 + (g << (5u + lg));
  WalkOne o;
  o.cand = enable && ((candw >> lane) & 1u);
  uint32_t e0 = 0, n = 0;
  if (o.cand) {
    e0 = lds_u16(rpa);
    n = lds_u16(rpa + 2u) - e0;
  }
  const uint32_t cpa = cola + e0 * 2u;
  const uint32_t x0 = lds_u16(cpa), x1 = lds_u16(cpa + 2u), x2 = lds_u16(cpa + 4u), x3 = lds_u16(cpa + 6u);  // may run past the row
  const uint32_t s0 = lds_u8(sta + min(x0, zidx)), s1 = lds_u8(sta + min(x1, zidx)), s2 = lds_u8(sta + min(x2, zidx)),
                 s3 = lds_u8(sta + min(x3, zidx));
  uint32_t w = ((s3 * 256u + s2) * 256u + s1) * 256u + s0;
  w &= bmsk_clamp(0u, n * 8u);
  if ((mt >> 16) > 4u) {  // warp-uniform: some row of this topology is longer than 4
    for (uint32_t e = 4; e < n; ++e) w |= lds_u8(sta + lds_u16(cpa + e * 2u));
  }
  o.unmet = (w & 0x01010101u) != 0;
  o.fdp = (w & 0x02020202u) != 0;
  return o;
}

DI void walk_deps2(uint32_t lane, uint32_t CAND, uint32_t lg, uint32_t col_addr, uint32_t meta, uint32_t rp0, uint32_t topo_buf,
                   uint32_t st0, uint32_t& met_w, uint32_t& fd_w) {
  met_w = 0;
  fd_w = 0;
  uint32_t todo = __ballot_sync(FULL, CAND != 0);  // (run, word) pairs with at least one candidate step
  while (todo) {
    const uint32_t L1 = __ffs(todo) - 1;
    todo &= todo - 1;
    const bool two = todo != 0;
    const uint32_t L2 = two ? __ffs(todo) - 1 : L1;
    todo &= todo - 1;  // (0 & -1 == 0 when there was no second word)
    const WalkOne a = walk_word(lane, L1, true, CAND, lg, col_addr, meta, rp0, topo_buf, st0);
    const WalkOne b = walk_word(lane, L2, two, CAND, lg, col_addr, meta, rp0, topo_buf, st0);
    const uint32_t fd1 = __ballot_sync(FULL, a.fdp), met1 = __ballot_sync(FULL, a.cand && !a.unmet);
    const uint32_t fd2 = __ballot_sync(FULL, b.fdp), met2 = __ballot_sync(FULL, b.cand && !b.unmet);
    if (lane == L1) { fd_w = fd1; met_w = met1; }
    if (two && lane == L2) { fd_w = fd2; met_w = met2; }
  }
}

}  // namespace bf
