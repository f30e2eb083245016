// validate.cu — batched dependency-graph validation on the device.
//
// validateRuntimeDependencyGraph (internal/controller/runs/dag.go:3076-3146) runs Kahn's algorithm per
// reconcile on string maps (O(S^2 log S)); the webhook does the same at admission
// (internal/webhook/v1alpha1/story_webhook.go:1501-1547).  Here one warp peels one topology record
// level-synchronously straight from the arena: a step joins level L when all its `needs` are in
// levels < L.  Steps that never join are on (or behind) a cycle.  Unknown dependencies (col_idx >= S)
// are rejected on the host before a record is even built.
//
// status word per topology: bit 0 = cycle detected, bits 8.. = number of levels (longest chain).
#include "kernel_common.cuh"

namespace bf {

constexpr int VAL_WARPS = 8;

__global__ void __launch_bounds__(VAL_WARPS * 32) validate_kernel(const Slot* slots, const uint32_t* slot_ids, uint32_t n,
                                                                  uint32_t n_slots, uint32_t* status) {
  __shared__ uint32_t done_s[VAL_WARPS][BF_MAX_STEPS / 32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t t = blockIdx.x * VAL_WARPS + warp;
  if (t >= n) return;
  const uint32_t sid = slot_ids[t];
  if (sid >= n_slots || slots[sid].addr == 0) {
    if (lane == 0) status[t] = 0xFFFFFFFFu;
    return;
  }
  const uint8_t* rec = reinterpret_cast<const uint8_t*>(slots[sid].addr);
  const TopoHeader* th = reinterpret_cast<const TopoHeader*>(rec);
  const uint32_t S = th->S, W = th->W, ell = th->ell;  // ell: fixed-width rows, no row_ptr (device_record.h)
  const uint16_t* row_ptr = reinterpret_cast<const uint16_t*>(rec + sizeof(TopoHeader));
  const uint16_t* col = reinterpret_cast<const uint16_t*>(rec + th->off_col);
  const uint8_t* colb = rec + th->off_col;                                                // byte-entry rows (device_record.h)
  const bool nodep_fmt = ell_has_nodep(ell);
  const uint32_t K = ell_k(ell);
  const uint32_t* nodep = reinterpret_cast<const uint32_t*>(rec + th->off_planes) + PL_NODEP * W;
  uint32_t* done = done_s[warp];
  for (uint32_t w = lane; w < W; w += 32) done[w] = 0;
  __syncwarp();
  uint32_t levels = 0, n_done = 0;
  for (;;) {
    // level-synchronous: decide from the masks of the previous level, publish after the sweep
    uint32_t newly = 0;  // my lane's newly-done steps, bit k = step lane + 32k
    for (uint32_t i = lane, k = 0; i < S; i += 32, ++k) {
      if ((done[i >> 5] >> (i & 31u)) & 1u) continue;
      bool ok = true;
      uint32_t e0 = ell ? i * K : row_ptr[i], e1 = ell ? e0 + K : row_ptr[i + 1];
      if (nodep_fmt && ((nodep[i >> 5] >> (i & 31u)) & 1u)) e1 = e0;  // a row without needs holds its own index
      for (uint32_t e = e0; e < e1 && ok; ++e) {
        const uint32_t d = ell ? ell_entry(colb, ell, W, i, e - e0) : col[e];
        if (d >= S) continue;  // unused entry of a fixed-width row
        ok = (done[d >> 5] >> (d & 31u)) & 1u;
      }
      if (ok) newly |= 1u << k;
    }
    __syncwarp();
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < W; ++k) {  // step lane + 32k lives in word k, bit lane
      const uint32_t word = __ballot_sync(FULL, (newly >> k) & 1u);
      cnt += __popc(word);
      if (lane == 0) done[k] |= word;
    }
    __syncwarp();
    if (cnt == 0) break;
    n_done += cnt;
    ++levels;
    if (n_done == S) break;
  }
  if (lane == 0) status[t] = (n_done != S ? 1u : 0u) | (levels << 8);
}

// Redrive closure (resolveRedriveFromStepSet, internal/controller/runs/storyrun_controller.go:535-558): the steps to
// reset when a run is re-driven from `start` = start plus everything downstream of it through the dependents
// edges of buildDependencyGraphs over the start step's OWN group (findStepGroup :560-577).  One warp per query;
// monotone fixed point over the CSR rows instead of the reference's queue (same closure, order-free).
__global__ void __launch_bounds__(VAL_WARPS * 32) closure_kernel(const Slot* slots, const uint32_t* slot_ids, const uint32_t* starts,
                                                                 uint32_t n, uint32_t n_slots, uint32_t words_out, uint32_t* masks) {
  __shared__ uint32_t sel_s[VAL_WARPS][BF_MAX_STEPS / 32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t t = blockIdx.x * VAL_WARPS + warp;
  if (t >= n) return;
  uint32_t* out = masks + (size_t)t * words_out;
  for (uint32_t w = lane; w < words_out; w += 32) out[w] = 0u;
  const uint32_t sid = slot_ids[t];
  if (sid >= n_slots || slots[sid].addr == 0) return;
  const uint8_t* rec = reinterpret_cast<const uint8_t*>(slots[sid].addr);
  const TopoHeader* th = reinterpret_cast<const TopoHeader*>(rec);
  const uint32_t S = th->S, W = th->W, start = starts[t], ell = th->ell;
  if (start >= S || W > words_out) return;
  const uint16_t* row_ptr = reinterpret_cast<const uint16_t*>(rec + sizeof(TopoHeader));
  const uint16_t* col = reinterpret_cast<const uint16_t*>(rec + th->off_col);
  const uint8_t* colb = rec + th->off_col;
  const uint32_t K = ell_k(ell);
  const uint32_t* planes = reinterpret_cast<const uint32_t*>(rec + th->off_planes);
  auto group_of = [&](uint32_t i) -> uint32_t {
    const uint32_t w = i >> 5, b = i & 31u;
    return ((planes[PL_G1 * W + w] >> b) & 1u) | (((planes[PL_G2 * W + w] >> b) & 1u) << 1);
  };
  const uint32_t g0 = group_of(start);
  uint32_t* sel = sel_s[warp];
  for (uint32_t w = lane; w < W; w += 32) sel[w] = 0;
  __syncwarp();
  if (lane == 0) sel[start >> 5] = 1u << (start & 31u);
  __syncwarp();
  for (;;) {
    bool changed = false;
    for (uint32_t i = lane; i < S; i += 32) {
      if ((sel[i >> 5] >> (i & 31u)) & 1u) continue;
      if (group_of(i) != g0) continue;
      const uint32_t e0 = ell ? i * K : row_ptr[i], e1 = ell ? e0 + K : row_ptr[i + 1];
      for (uint32_t e = e0; e < e1; ++e) {
        const uint32_t d = ell ? ell_entry(colb, ell, W, i, e - e0) : col[e];   // (a NODEP row points at its own, unselected, step: no effect)
        if (d < S && ((sel[d >> 5] >> (d & 31u)) & 1u)) {
          atomicOr(&sel[i >> 5], 1u << (i & 31u));
          changed = true;
          break;
        }
      }
    }
    __syncwarp();
    if (!__any_sync(FULL, changed)) break;
  }
  for (uint32_t w = lane; w < W; w += 32) out[w] = sel[w];
}

cudaError_t launch_closure(const Slot* slots, const uint32_t* slot_ids, const uint32_t* starts, uint32_t n, uint32_t n_slots,
                           uint32_t words_out, uint32_t* masks, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  closure_kernel<<<(n + VAL_WARPS - 1) / VAL_WARPS, VAL_WARPS * 32, 0, stream>>>(slots, slot_ids, starts, n, n_slots, words_out, masks);
  return cudaGetLastError();
}

cudaError_t launch_validate(const Slot* slots, const uint32_t* slot_ids, uint32_t n, uint32_t n_slots, uint32_t* status,
                            cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  validate_kernel<<<(n + VAL_WARPS - 1) / VAL_WARPS, VAL_WARPS * 32, 0, stream>>>(slots, slot_ids, n, n_slots, status);
  return cudaGetLastError();
}

}  // namespace bf
