// compact.cu — result records -> compact lists, on the device (SURVEY.md rows a9 / d).
//
// The consumer of a pass (findAndLaunchReadySteps, dag.go:1735-1775) wants LISTS: the ready steps to hand to
// StepExecutor.Execute and the skipped steps to mark.  A frontier pass over 100k runs x 256 steps leaves ~3 such steps
// per run, yet the dense result records are 80 bytes per run (16-byte header + two 256-bit masks), 8 MB per pass over
// PCIe.  Here the masks are turned into 8-byte (run, step, kind) events, run-major and step-ascending — the order of
// the reference's lists — plus one summary word per run, so a tick ships ~0.4 MB + 8 bytes per event.  Two launches:
// per-block event totals, then every block sums the totals before it, scans its own runs and emits (deterministic order,
// so the list is bit-exact against the oracle's masks).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/bobrafrontier.h"
#include "device_record.h"

namespace bf {

constexpr int CB = 512;  // runs per block (one thread per run)
constexpr uint32_t STAGE_EVENTS = 4096;  // events of a block staged in shared memory (32 KB)

__device__ __forceinline__ uint32_t union_word(const CompactParams& P, const uint8_t* rr, uint32_t w) {
  uint32_t u = reinterpret_cast<const uint32_t*>(rr + P.off_ready)[w] | reinterpret_cast<const uint32_t*>(rr + P.off_skip)[w];
  if (P.off_fail != BF_OFF_NONE) u |= reinterpret_cast<const uint32_t*>(rr + P.off_fail)[w];
  if (P.off_needs_cond != BF_OFF_NONE) u |= reinterpret_cast<const uint32_t*>(rr + P.off_needs_cond)[w];
  if (P.off_skip_dep != BF_OFF_NONE) u |= reinterpret_cast<const uint32_t*>(rr + P.off_skip_dep)[w];
  return u;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* sh) {  // all threads get the block total
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  v = __reduce_add_sync(0xffffffffu, v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  uint32_t t = lane < CB / 32 ? sh[lane] : 0u;
  t = __reduce_add_sync(0xffffffffu, t);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(CB) compact_count(const CompactParams P) {
  __shared__ uint32_t sh[32];
  const uint32_t r = blockIdx.x * CB + threadIdx.x;
  uint32_t c = 0;
  if (r < P.n_runs) {
    const uint8_t* rr = P.result + (size_t)r * P.result_stride;
    const uint32_t summary = *reinterpret_cast<const uint32_t*>(rr);
    if (P.summary) P.summary[r] = summary;
    if (summary != 0xFFFFFFFFu)  // a dead slot's record is empty by contract
      for (uint32_t w = 0; w < P.words; ++w) c += __popc(union_word(P, rr, w));
  }
  const uint32_t tot = block_sum(c, sh);
  if (threadIdx.x == 0) P.block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(CB) compact_emit(const CompactParams P) {
  __shared__ uint32_t sh[32];
  __shared__ unsigned long long base_s;
  // events of the blocks before mine
  unsigned long long part = 0;
  for (uint32_t b = threadIdx.x; b < blockIdx.x; b += CB) part += P.block_sums[b];
  {
    __shared__ unsigned long long sh64[CB / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
    if (lane == 0) sh64[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (int k = 0; k < CB / 32; ++k) t += sh64[k];
      base_s = t;
    }
    __syncthreads();
  }
  const uint32_t r = blockIdx.x * CB + threadIdx.x;
  const uint8_t* rr = P.result + (size_t)r * P.result_stride;
  uint32_t c = 0;
  const bool alive = r < P.n_runs && *reinterpret_cast<const uint32_t*>(rr) != 0xFFFFFFFFu;
  if (alive)
    for (uint32_t w = 0; w < P.words; ++w) c += __popc(union_word(P, rr, w));
  // exclusive scan of c over the block
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += n;
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = lane < CB / 32 ? sh[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= (uint32_t)o) v += n;
    }
    sh[lane] = v;
  }
  __syncthreads();
  const uint32_t before = inc - c + (warp ? sh[warp - 1] : 0u);
  unsigned long long pos = base_s + before;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == CB - 1) {
    *P.total = pos + c;  // the batch's event count
    if (P.host_tail) {   // posted writes to pinned host memory; visible to the host once the stream has been synchronised
      P.host_tail[0] = pos + c;
      for (int k = 0; k < 4; ++k) P.host_tail[1 + k] = P.counts ? P.counts[k] : 0ull;
      P.host_tail[5] = P.rejected ? (unsigned long long)*P.rejected : 0ull;
    }
  }
  // Small blocks of the list are staged in shared memory and written out by consecutive threads: the destination may be
  // PINNED HOST memory (zero-copy results: no D2H copy command at all), where scattered 8-byte stores would each be a
  // PCIe write of their own; staged, a warp writes 256 contiguous bytes per instruction.
  __shared__ bf_step_event stage[STAGE_EVENTS];
  const uint32_t btot = sh[CB / 32 - 1];
  const bool staged = btot <= STAGE_EVENTS;
  if (staged) {
    if (alive && c != 0) {
      uint32_t at = before;
      for (uint32_t w = 0; w < P.words; ++w) {
        const uint32_t rd = reinterpret_cast<const uint32_t*>(rr + P.off_ready)[w], sk = reinterpret_cast<const uint32_t*>(rr + P.off_skip)[w];
        const uint32_t fl = P.off_fail != BF_OFF_NONE ? reinterpret_cast<const uint32_t*>(rr + P.off_fail)[w] : 0u;
        const uint32_t nc = P.off_needs_cond != BF_OFF_NONE ? reinterpret_cast<const uint32_t*>(rr + P.off_needs_cond)[w] : 0u;
        const uint32_t sd = P.off_skip_dep != BF_OFF_NONE ? reinterpret_cast<const uint32_t*>(rr + P.off_skip_dep)[w] : 0u;
        for (uint32_t u = rd | sk | fl | nc | sd; u; u &= u - 1) {
          const uint32_t b = __ffs(u) - 1;
          bf_step_event e;
          e.run = r;
          e.step = (uint16_t)(w * 32u + b);
          e.kind = (uint16_t)(((rd >> b) & 1u) * BF_EVT_READY | ((sk >> b) & 1u) * BF_EVT_SKIP | ((fl >> b) & 1u) * BF_EVT_FAIL |
                              ((nc >> b) & 1u) * BF_EVT_NEEDS_COND | ((sd >> b) & 1u) * BF_EVT_SKIP_DEP);
          stage[at++] = e;
        }
      }
    }
    __syncthreads();
    const unsigned long long base = base_s;
    for (uint32_t i = threadIdx.x; i < btot; i += CB)
      if (base + i < P.cap) P.events[base + i] = stage[i];
    return;
  }
  if (!alive || c == 0) return;
  for (uint32_t w = 0; w < P.words; ++w) {
    const uint32_t rd = reinterpret_cast<const uint32_t*>(rr + P.off_ready)[w], sk = reinterpret_cast<const uint32_t*>(rr + P.off_skip)[w];
    const uint32_t fl = P.off_fail != BF_OFF_NONE ? reinterpret_cast<const uint32_t*>(rr + P.off_fail)[w] : 0u;
    const uint32_t nc = P.off_needs_cond != BF_OFF_NONE ? reinterpret_cast<const uint32_t*>(rr + P.off_needs_cond)[w] : 0u;
    const uint32_t sd = P.off_skip_dep != BF_OFF_NONE ? reinterpret_cast<const uint32_t*>(rr + P.off_skip_dep)[w] : 0u;
    for (uint32_t u = rd | sk | fl | nc | sd; u; u &= u - 1) {
      const uint32_t b = __ffs(u) - 1;
      if (pos < P.cap) {
        bf_step_event e;
        e.run = r;
        e.step = (uint16_t)(w * 32u + b);
        e.kind = (uint16_t)(((rd >> b) & 1u) * BF_EVT_READY | ((sk >> b) & 1u) * BF_EVT_SKIP | ((fl >> b) & 1u) * BF_EVT_FAIL |
                            ((nc >> b) & 1u) * BF_EVT_NEEDS_COND | ((sd >> b) & 1u) * BF_EVT_SKIP_DEP);
        P.events[pos] = e;
      }
      ++pos;
    }
  }
}

// scratch: block_sums needs ceil(n / 512) u64
cudaError_t launch_compact(const CompactParams& P, cudaStream_t stream) {
  if (P.n_runs == 0) {
    if (P.host_tail) memset(P.host_tail, 0, 6 * sizeof(unsigned long long));   // pinned host memory: plain store
    return cudaMemsetAsync(P.total, 0, sizeof(unsigned long long), stream);
  }
  const uint32_t nb = (P.n_runs + CB - 1) / CB;
  compact_count<<<nb, CB, 0, stream>>>(P);
  compact_emit<<<nb, CB, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace bf
