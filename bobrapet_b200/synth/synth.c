/*
 * synth.c — seeded synthetic Stories / StoryRuns for the five BASELINE.json configurations
 * (SURVEY.md section 8(d)).  Data generation only: no frontier semantics live here.
 *
 * PRNG: splitmix64, seed = 0xB0B2A9E7 ^ (cfg << 56) ^ (stream << 48) ^ run_id, so any
 * sub-range of runs can be generated independently (ranks generate their own shard).
 *
 *  cfg 2: S=64  lattice 8 layers x 8 wide, step (l,w) needs (l-1,w),(l-1,(w+1)%8); E=112
 *  cfg 3: S=256 step i needs min(i,4) distinct uniform picks from [max(0,i-32), i); E=1014
 *  cfg 4: cfg 3 + 25% of steps carry an `if` (cond PASS 60 / SKIP 25 / HOLD 15) and 25% are
 *         `gate` steps (decision pending 40 / approved 50 / rejected 8 / timed-out 2)
 *  cfg 5: S=1024, same rule; 8 `parallel` steps at 128k+64 with 128 branches each
 *  any other cfg value: the cfg-3 rule at the caller's S (used by tests for odd sizes)
 *
 * State (cfg 2-5): progress p ~ U[0,S]; i<p: Succeeded 88 / Skipped 3 / Failed 2 / Running 5 /
 * Paused 2 (%); i>=p: none 97 / Pending-queued 3; allowFailure on 10% of steps;
 * failFast false on 50% of runs.
 *
 * Runs are independent streams, so the loops over runs are OpenMP-parallel (synth_set_threads).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC synth.c -o libsynth.so
 */
#include <stdint.h>
#include <string.h>

#include "../../include/bobrafrontier.h"

static int g_threads = 1;
void synth_set_threads(int t) { g_threads = t < 1 ? 1 : t; }

static inline uint64_t sm64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint32_t below(uint64_t* s, uint32_t n) { return (uint32_t)(((sm64(s) >> 32) * (uint64_t)n) >> 32); }
static inline uint64_t seed_of(uint32_t cfg, uint32_t stream, uint64_t run) {
  return 0xB0B2A9E7ull ^ ((uint64_t)cfg << 56) ^ ((uint64_t)stream << 48) ^ run;
}

uint32_t synth_edges(uint32_t cfg, uint32_t S) {
  if (cfg == 2) return S >= 8 ? 2 * (S - 8) : 0;
  uint32_t e = 0;
  for (uint32_t i = 0; i < S; ++i) e += i < 4 ? i : 4;
  return e;
}

/* One topology.  row_ptr[S+1] u32, col_idx[E] u16 (ascending inside a row), flags[S]. */
static void topo_one(uint32_t cfg, uint64_t run, uint32_t S, uint32_t* row_ptr, uint16_t* col, uint8_t* flags) {
  uint64_t st = seed_of(cfg, 1, run);
  uint32_t e = 0;
  for (uint32_t i = 0; i < S; ++i) {
    row_ptr[i] = e;
    if (cfg == 2) {
      const uint32_t l = i / 8, w = i % 8;
      if (l > 0) {
        uint32_t a = (l - 1) * 8 + w, b = (l - 1) * 8 + (w + 1) % 8;
        if (a > b) { uint32_t t = a; a = b; b = t; }
        col[e++] = (uint16_t)a;
        col[e++] = (uint16_t)b;
      }
    } else {
      const uint32_t k = i < 4 ? i : 4, lo = i > 32 ? i - 32 : 0, span = i - lo;
      uint32_t picks[4], n = 0;
      while (n < k) {
        const uint32_t c = lo + below(&st, span);
        int dup = 0;
        for (uint32_t x = 0; x < n; ++x) dup |= picks[x] == c;
        if (!dup) picks[n++] = c;
      }
      for (uint32_t a = 1; a < n; ++a) { /* insertion sort */
        uint32_t v = picks[a], b = a;
        while (b > 0 && picks[b - 1] > v) { picks[b] = picks[b - 1]; --b; }
        picks[b] = v;
      }
      for (uint32_t x = 0; x < n; ++x) col[e++] = (uint16_t)picks[x];
    }
  }
  row_ptr[S] = e;
  for (uint32_t i = 0; i < S; ++i) {
    uint8_t f = BF_STEP_ENGRAM;
    if (below(&st, 100) < 10) f |= BF_SF_ALLOW_FAILURE;
    if (cfg == 4) {
      const uint32_t r = below(&st, 100);
      if (r < 25) f |= BF_SF_HAS_IF;
      else if (r < 50) { f = (uint8_t)((f & ~BF_SF_TYPE_MASK) | BF_STEP_GATE); if (below(&st, 2)) f |= BF_SF_ON_TIMEOUT_SKIP; }
    }
    if (cfg == 5 && i % 128 == 64) f = (uint8_t)((f & ~BF_SF_TYPE_MASK) | BF_STEP_PARALLEL);
    flags[i] = f;
  }
}

/* Topologies for runs [run_lo, run_lo+n): arrays are concatenated, fixed S per call.
 * par (n*P descs) and allow_bits (n*P*B/8 bytes) are filled only for cfg 5 (P=8, B=128). */
void synth_topologies(uint32_t cfg, uint64_t run_lo, uint32_t n, uint32_t S, uint32_t* row_ptr, uint16_t* col,
                      uint8_t* flags, bf_parallel_desc* par, uint8_t* allow_bits) {
  const uint32_t E = synth_edges(cfg, S);
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (uint32_t r = 0; r < n; ++r) {
    topo_one(cfg, run_lo + r, S, row_ptr + (size_t)r * (S + 1), col + (size_t)r * E, flags + (size_t)r * S);
    if (cfg == 5 && par) {
      uint64_t st = seed_of(cfg, 3, run_lo + r);
      const uint32_t P = S / 128, B = 128;
      for (uint32_t q = 0; q < P; ++q) {
        bf_parallel_desc* d = &par[(size_t)r * P + q];
        d->step = (uint16_t)(128 * q + 64);
        d->branches = (uint16_t)B;
        d->allow_first = (uint32_t)(((size_t)r * P + q) * B);
        if (allow_bits) {
          uint8_t* ab = allow_bits + ((size_t)r * P + q) * (B / 8);
          memset(ab, 0, B / 8);
          for (uint32_t b = 0; b < B; ++b)
            if (below(&st, 100) < 5) ab[b >> 3] |= (uint8_t)(1u << (b & 7u));
        }
      }
    }
  }
}

static inline void put_nib(uint8_t* a, uint32_t i, uint32_t v) { a[i >> 1] |= (uint8_t)(v << ((i & 1u) * 4u)); }
/* bit-sliced code: plane b at base + b*W words */
static inline void put_code(uint8_t* base, uint32_t W, int nbits, uint32_t i, uint32_t v) {
  uint32_t* w = (uint32_t*)base;
  for (int b = 0; b < nbits; ++b)
    if ((v >> b) & 1u) w[(uint32_t)b * W + (i >> 5)] |= 1u << (i & 31u);
}

/* State records for runs [run_lo, run_lo+n) into `state` (n * L->state_stride bytes, zeroed here).
 * slots[r] is written to the header; flags[] are the topology's step flags (n*S) (for gate/if/parallel
 * placement); child_first[P] gives the nibble offsets of the parallel descs (cfg 5). */
void synth_state(uint32_t cfg, uint64_t run_lo, uint32_t n, uint32_t S, const bf_layout* L, const uint32_t* slots,
                 const uint8_t* flags, const uint32_t* child_first, uint32_t P, uint32_t B, uint8_t* state) {
  memset(state, 0, (size_t)n * L->state_stride);
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (uint32_t r = 0; r < n; ++r) {
    uint8_t* rec = state + (size_t)r * L->state_stride;
    bf_run_header* h = (bf_run_header*)rec;
    uint64_t st = seed_of(cfg, 2, run_lo + r);
    const uint8_t* fl = flags + (size_t)r * S;
    h->topo_slot = slots[r];
    h->run_flags = below(&st, 2) ? BF_RF_FAIL_FAST : 0;
    const uint32_t p = below(&st, S + 1);
    uint8_t* ph = rec + L->off_phase;
    uint8_t* cd = L->off_cond != BF_OFF_NONE ? rec + L->off_cond : 0;
    uint8_t* dc = L->off_decision != BF_OFF_NONE ? rec + L->off_decision : 0;
    uint8_t* ch = L->off_child != BF_OFF_NONE ? rec + L->off_child : 0;
    uint32_t q = 0;
    for (uint32_t i = 0; i < S; ++i) {
      uint32_t code;
      const uint32_t u = below(&st, 100);
      if (i < p) code = u < 88 ? BF_PHASE_SUCCEEDED : u < 91 ? BF_PHASE_SKIPPED : u < 93 ? BF_PHASE_FAILED : u < 98 ? BF_PHASE_RUNNING : BF_PHASE_PAUSED;
      else code = u < 97 ? BF_PHASE_NONE : BF_PHASE_PENDING_QUEUED;
      const uint32_t ty = fl[i] & BF_SF_TYPE_MASK;
      if (cd) {
        uint32_t c = BF_COND_PASS;
        if (fl[i] & BF_SF_HAS_IF) { const uint32_t v = below(&st, 100); c = v < 60 ? BF_COND_PASS : v < 85 ? BF_COND_SKIP : BF_COND_HOLD; }
        put_code(cd, L->words, 2, i, c);
      }
      if (dc) {
        uint32_t d = BF_DEC_PENDING;
        if (ty == BF_STEP_GATE) { const uint32_t v = below(&st, 100); d = v < 40 ? BF_DEC_PENDING : v < 90 ? BF_DEC_SUCCEED : v < 98 ? BF_DEC_FAIL : BF_DEC_TIMED_OUT; }
        put_code(dc, L->words, 2, i, d);
      }
      if (ty == BF_STEP_PARALLEL && q < P) {
        if (i < p && code == BF_PHASE_SUCCEEDED && below(&st, 100) < 30) code = BF_PHASE_RUNNING;
        if (code != BF_PHASE_NONE && code != BF_PHASE_PENDING_QUEUED) {
          h->children_registered |= 1ull << q;
          if (ch) {
            const uint32_t mode = below(&st, 100);
            for (uint32_t b = 0; b < B; ++b) {
              uint32_t cc;
              const uint32_t v = below(&st, 1000);
              if (mode < 50 || code == BF_PHASE_SUCCEEDED) cc = v < 970 ? BF_PHASE_SUCCEEDED : BF_PHASE_FAILED;
              else cc = v < 776 ? BF_PHASE_SUCCEEDED : v < 800 ? BF_PHASE_FAILED : v < 950 ? BF_PHASE_RUNNING : BF_PHASE_NONE;
              put_nib(ch, child_first[q] + b, cc);
            }
          }
        }
        q++;
      }
      put_code(ph, L->words, 4, i, code);
    }
  }
}
