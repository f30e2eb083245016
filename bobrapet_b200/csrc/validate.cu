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
  const uint32_t S = th->S, W = th->W;
  const uint16_t* row_ptr = reinterpret_cast<const uint16_t*>(rec + sizeof(TopoHeader));
  const uint16_t* col = reinterpret_cast<const uint16_t*>(rec + th->off_col);
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
      for (uint32_t e = row_ptr[i]; e < row_ptr[i + 1] && ok; ++e) {
        const uint32_t d = col[e];
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

cudaError_t launch_validate(const Slot* slots, const uint32_t* slot_ids, uint32_t n, uint32_t n_slots, uint32_t* status,
                            cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  validate_kernel<<<(n + VAL_WARPS - 1) / VAL_WARPS, VAL_WARPS * 32, 0, stream>>>(slots, slot_ids, n, n_slots, status);
  return cudaGetLastError();
}

}  // namespace bf
