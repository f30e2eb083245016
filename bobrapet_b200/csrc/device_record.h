// device_record.h — HBM layouts shared by the host packer (abi.cu) and the kernels.
//
// A topology (one Story generation: needs-adjacency + static step flags) is one
// contiguous, 16-byte aligned, 16-byte padded record in the arena so that a warp
// stages it into shared memory with ONE cp.async.bulk (TMA) copy.  Two adjacency formats,
// chosen per topology when it is uploaded (plan_record, abi.cu):
//
//   CSR (ell = 0)
//   +0            TopoHeader (32 B)
//   +32           row_ptr   u16[S+1]          CSR over allStorySteps (dag.go:3270), pad 16
//   +off_col      col_idx   u16[E]            dependency step indices, then >= 4 zero entries, pad 16
//
//   fixed-width rows (ell = K in {2, 4}; taken when no step has more than K needs and S*K entries are not
//   larger than the CSR block they replace; S > 256): row_ptr is implicit (row i starts at entry i*K)
//   +32 = off_col col_idx   u16[32*W][K]      the needs of step i, unused entries (and the rows past S) = PAD
//                                             PAD = 32*W: the status byte just past the last step word, always 0
//                                             ("satisfied, not failed"), so a padded entry never changes a verdict
//
//   fixed-width rows with BYTE entries (ell = 0x100 | K; taken for every topology of at most 256 steps that qualifies
//   for fixed-width rows): a step index fits a byte, so the adjacency is HALF the size again
//   +32 = off_col col_idx   u8[32*W][K]       the needs of step i; a row shorter than K repeats its first entry (the OR
//                                             over a row's status bytes is idempotent); there is no PAD index (32*W = 256
//                                             does not fit a byte at S = 256), so a row WITHOUT needs holds its own index
//                                             and its step is flagged in a ninth static plane, NODEP: flagged steps never
//                                             enter the walk, their dependencies are trivially met
//
//   fixed-width rows with 10-BIT entries (ell = 0x200 | 4; taken for topologies of more than 512 steps — the ones only the
//   one-run-per-warp kernel stages — whose rows have at most 4 needs): a step index below 1024 fits 10 bits
//   +32 = off_col lo        u32[32*W]         entries 0, 1, 2 in bits 0-9, 10-19, 20-29; the low 2 bits of entry 3 in bits 30-31
//   +32 + 128*W   hi        u8[32*W]          the high 8 bits of entry 3          (5 bytes per row against 8)
//                                             short rows, rows without needs and NODEP as for byte entries
//
//   all
//   +off_planes   planes    u32[8 or 9][W]    static step flags, BIT-SLICED (W = ceil(S/32)):
//                                             t0,t1,t2 (type), AF, TS, HAS_IF, G1 (comp), G2 (finally)
//                                             = S bytes, same size as the canonical u8 step_flags[S];
//                                             byte-entry and 10-bit rows: + NODEP
//   +off_par      ParDesc[P] (16 B each) followed by the branch allowFailure bit words
//
// Canonical ("algorithmic") bytes per topology, SURVEY.md section 8(d):
//   2*(S+1) + 2*E + S  (u16 CSR + u8 flags).  The CSR record adds the header and the 16-byte paddings
//   (~1.5% at cfg3); the fixed-width record drops row_ptr, so it is SMALLER than the canonical figure
//   (cfg3: 2 336 B with u16 entries, 1 344 B with byte entries, against 2 798 B).
#pragma once
#include <stdint.h>

namespace bf {

struct TopoHeader {
  uint16_t S;          // steps
  uint16_t W;          // ceil(S/32)
  uint16_t max_deg;    // largest in-degree (the kernel's straight-line walk covers <= 4; E = row_ptr[S])
  uint16_t P;          // parallel descs
  uint16_t n_main, n_comp, n_final;
  uint16_t child_nibbles;  // total child nibbles of all descs
  uint16_t off_col;    // byte offset of col_idx
  uint16_t ell;        // 0 = CSR (row_ptr at +32), K = 2 / 4: fixed-width rows of K u16 entries, no row_ptr; 0x100 | K: byte entries; 0x204: 10-bit entries
  uint32_t off_planes;
  uint32_t off_par;
  uint32_t rec_bytes;  // multiple of 16
};
static_assert(sizeof(TopoHeader) == 32, "TopoHeader must be 32 bytes");

enum Plane { PL_T0 = 0, PL_T1, PL_T2, PL_AF, PL_TS, PL_HASIF, PL_G1, PL_G2, PL_COUNT, PL_NODEP = PL_COUNT };
constexpr uint32_t ELL_BYTE = 0x100u;                                                   // TopoHeader::ell flag: byte entries
constexpr uint32_t ELL_PACK10 = 0x200u;                                                 // TopoHeader::ell flag: 10-bit entries (K = 4)
__host__ __device__ inline uint32_t ell_k(uint32_t ell) { return ell & 0xFFu; }                      // entries per row
__host__ __device__ inline uint32_t ell_row_bytes(uint32_t ell) {
  return (ell & ELL_PACK10) ? 5u : ((ell & ELL_BYTE) ? (ell & 0xFFu) : 2u * ell);
}
__host__ __device__ inline bool ell_has_nodep(uint32_t ell) { return (ell & (ELL_BYTE | ELL_PACK10)) != 0; }
__host__ __device__ inline uint32_t plane_count(uint32_t ell) { return ell_has_nodep(ell) ? PL_COUNT + 1u : PL_COUNT; }
// entry e of row i of a fixed-width block at `col` (any of the three entry widths); W = words of the topology
__host__ __device__ inline uint32_t ell_entry(const uint8_t* col, uint32_t ell, uint32_t W, uint32_t i, uint32_t e) {
  if (ell & ELL_PACK10) {
    const uint32_t lo = reinterpret_cast<const uint32_t*>(col)[i];
    return e < 3u ? (lo >> (10u * e)) & 0x3FFu : (lo >> 30) | ((uint32_t)col[128u * W + i] << 2);
  }
  if (ell & ELL_BYTE) return col[i * (ell & 0xFFu) + e];
  return reinterpret_cast<const uint16_t*>(col)[i * ell + e];
}

struct ParDesc {
  uint16_t step;
  uint16_t branches;
  uint32_t child_first;  // nibble offset in the run's child area
  uint32_t allow_off;    // byte offset (from record start) of this desc's allowFailure bit words
  uint32_t reserved;
};
static_assert(sizeof(ParDesc) == 16, "ParDesc must be 16 bytes");

struct Slot {          // device slot table entry
  uint64_t addr;       // device address of the record (0 = dead slot)
  uint32_t bytes;      // rec_bytes
  uint32_t meta;       // S | P << 16 (P = number of parallel descs)
};
static_assert(sizeof(Slot) == 16, "Slot must be 16 bytes");

struct KParams {
  const uint8_t* state;
  uint8_t* result;
  const Slot* slots;
  unsigned long long* counts;   // bf_counts (4 x u64) or nullptr
  // compact ticks whose pass is the packed-lanes kernel alone: the pass itself leaves every run's head word (summary | listed |
  // event count) and adds the counts to the per-512-run totals the emit kernel scans — no separate heads kernel
  uint32_t* head;               // [n_runs] or nullptr
  unsigned long long* head_sums;   // [ceil(n_runs / 512)], zero on entry
  unsigned long long* acc;      // BF_EVAL_COUNTS_SET, packed-lanes kernel alone: ctx scratch {4 totals, CTA ticket}; the last CTA
                                // out copies the totals to `counts` and clears the scratch (nullptr: add to `counts`)
  uint32_t* exp_counts;         // [n_runs] compact per-run expansion counts or nullptr
  // two-tier dispatch: the packed-lanes kernel appends the runs it cannot take (topologies with
  // parallel steps) to defer_list; the general kernel then runs over run_list[0..*run_list_count)
  uint32_t* defer_list;
  uint32_t* defer_count;
  const uint32_t* run_list;
  const uint32_t* run_list_count;
  uint32_t n_slots;
  uint32_t n_runs;
  uint32_t flags;               // BF_EVAL_*
  uint32_t max_iter;
  uint32_t any_parallel;        // some live topology has `parallel` steps (selects the CH kernel variants)
  // layout (bf_layout)
  uint32_t words;
  uint32_t state_stride, off_phase, off_cond, off_decision, off_child;
  uint32_t result_tail;         // first byte after the last result field (padding up to the stride is zeroed)
  uint32_t result_stride, off_ready, off_skip, off_fail, off_needs_cond, off_skip_dep, off_phase_out;
  // shared-memory plan
  uint32_t stages;              // general kernel: ring depth per warp
  uint32_t topo_buf_bytes;      // per-run room for a topology record
  uint32_t stage_bytes;         // general kernel: state_stride + topo_buf_bytes (multiple of 16)
  uint32_t work_bytes;          // per-warp scratch
  uint32_t warps_per_block;
  uint32_t occ2;                // general kernel: launch the build compiled for two resident CTAs per SM
  uint32_t run_blocked;         // general kernel: 1: a warp takes a contiguous block of runs, 0: every G-th run
  // packed-lanes kernel (frontier_pack.cu): words per run rounded up to a power of two, slot groups per CTA
  uint32_t wq, wq_log2;
  uint32_t slot_groups;
};

// compact.cu: result records -> one head word per run + 16-bit (step | kind << 10) events
struct CompactParams {
  const uint8_t* result;
  const uint8_t* prev_result;        // previous tick's records (changed-only mode) or nullptr: every run is listed
  uint32_t* head;                    // [n_runs]
  uint32_t* host_head;               // the caller's pinned head buffer (device-visible address) or nullptr: heads are also posted
                                     // straight to it, no separate download
  uint16_t* events;                  // [cap]
  unsigned long long cap;
  unsigned long long* block_sums;    // scratch: ceil(n_runs / 512)
  unsigned long long* zero_sums;     // the OTHER totals buffer: the emit kernel leaves it zeroed for the next tick's pass to add into
  uint32_t zero_len;                 // entries of zero_sums
  uint32_t heads_done;               // the pass wrote head words and block totals itself (every run listed): no compact_heads
  unsigned long long* total;         // out: [0] events of the batch, [1] listed runs
  // the pass's small results, written straight to pinned host memory by the last block (no separate small D2H copies):
  // host_tail[0] = events, [1..4] = bf_counts, [5] = rejected deltas, [6] = listed runs
  unsigned long long* host_tail;     // device-visible address of the pinned block, or nullptr
  const unsigned long long* counts;  // device bf_counts of the pass
  const uint32_t* rejected;          // device counter of rejected deltas, or nullptr
  uint32_t n_runs, words, result_stride, result_tail, off_ready, off_skip, off_fail, off_needs_cond, off_skip_dep;
};

// resident.cu (row f2)
struct DeltaParams {
  uint8_t* state;
  const bf_delta* deltas;
  uint32_t n, n_runs;
  uint32_t words, state_stride, off_phase, off_cond, off_decision, off_child, child_nibbles;
  uint32_t* rejected;  // count of deltas outside the record (bad run / index / absent field)
};

// limiters.cu (rows a9 / f4)
struct SchedParams {
  const uint8_t* state;
  const uint8_t* result;
  const Slot* slots;
  const bf_sched_run* runs;
  uint8_t* records;            // [n_runs][stride]
  uint32_t* story_running;     // totals (pre-loaded with the base counts)
  uint32_t* queue_running;
  int32_t* queue_maxprio;      // pre-loaded with INT32_MIN
  uint32_t* global_running;
  const int32_t* story_limit;
  const int32_t* queue_limit;
  const int32_t* queue_aging;
  const uint32_t* story_base;  // StepRuns the batch does not hold, or nullptr
  const uint32_t* queue_base;
  const int32_t* queue_maxprio_base;  // highest effective priority among runs with demand OUTSIDE the batch, or nullptr
  uint32_t global_base;
  int32_t global_limit;
  uint32_t n_stories, n_queues, n_slots, n_runs;
  uint32_t words, state_stride, off_phase, off_child, child_nibbles, result_stride, off_ready, stride;
};

}  // namespace bf
