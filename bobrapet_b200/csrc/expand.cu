// expand.cu — `parallel` fan-out expansion (executeParallelStep, step_executor.go:740-811).
//
// For every ready `parallel` step the host must create one child StepRun per branch of
// with.steps, in branch order.  The frontier kernel leaves a per-run expansion count;
// here an exclusive scan turns the counts into offsets and a second kernel writes the
// (run, step, branch) tuples in (run, step, branch) order — deterministic, so the
// tuple list is bit-exact against the oracle.  Two launches: per-block totals, then scan + emit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bobrafrontier.h"
#include "device_record.h"

namespace bf {

constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, uint32_t lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (uint32_t)o) v += n;
  }
  return v;
}

// block-wide inclusive scan of one value per thread (1024 threads)
__device__ __forceinline__ unsigned long long block_incl_scan(unsigned long long v, unsigned long long* total) {
  __shared__ unsigned long long wsum[32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  v = warp_incl_scan(v, lane);
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = wsum[lane];
    w = warp_incl_scan(w, lane);
    wsum[lane] = w;
  }
  __syncthreads();
  if (warp > 0) v += wsum[warp - 1];
  if (total) *total = wsum[31];
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) exp_block_sums(const uint32_t* counts, uint32_t n, unsigned long long* block_sums) {
  const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  unsigned long long tot;
  block_incl_scan(i < n ? counts[i] : 0u, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// Second launch: block b sums the totals of the blocks before it, scans its own 1024 runs in shared memory and its warps
// emit the tuples: one warp per group of 32 runs, runs with a non-zero count are expanded cooperatively.
__global__ void __launch_bounds__(SCAN_BLOCK) exp_scan_emit(const KParams P, const unsigned long long* block_sums, bf_expansion* out,
                                                            unsigned long long cap) {
  __shared__ unsigned long long off_s[SCAN_BLOCK];
  __shared__ unsigned long long part_s[32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned long long part = 0;
  for (uint32_t b = threadIdx.x; b < blockIdx.x; b += SCAN_BLOCK) part += block_sums[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if (lane == 0) part_s[warp] = part;
  __syncthreads();
  unsigned long long base = 0;
  for (int k = 0; k < 32; ++k) base += part_s[k];
  __syncthreads();
  const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const uint32_t c = i < P.n_runs ? P.exp_counts[i] : 0u;
  const unsigned long long inc = block_incl_scan(c, nullptr);
  off_s[threadIdx.x] = base + inc - c;
  __syncthreads();
  const uint32_t r0 = blockIdx.x * SCAN_BLOCK + warp * 32;
  uint32_t todo = __ballot_sync(0xffffffffu, c != 0);
  while (todo) {
    const uint32_t l = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t r = r0 + l;
    const uint8_t* srec = P.state + (size_t)r * P.state_stride;
    const uint32_t slot = *reinterpret_cast<const uint32_t*>(srec);
    if (slot >= P.n_slots) continue;
    const Slot se = P.slots[slot];
    if (se.addr == 0) continue;
    const uint8_t* tr = reinterpret_cast<const uint8_t*>(se.addr);
    const TopoHeader* th = reinterpret_cast<const TopoHeader*>(tr);
    const ParDesc* pd = reinterpret_cast<const ParDesc*>(tr + th->off_par);
    const uint32_t* ready = reinterpret_cast<const uint32_t*>(P.result + (size_t)r * P.result_stride + P.off_ready);
    unsigned long long pos = off_s[warp * 32 + l];
    for (uint32_t q = 0; q < th->P; ++q) {
      const uint32_t stp = pd[q].step, B = pd[q].branches;
      if (!((ready[stp >> 5] >> (stp & 31u)) & 1u)) continue;
      for (uint32_t b = lane; b < B; b += 32) {
        if (pos + b < cap) {
          bf_expansion t;
          t.run = r; t.step = (uint16_t)stp; t.branch = (uint16_t)b;
          out[pos + b] = t;
        }
      }
      pos += B;
    }
  }
}

// scratch: block_sums needs ceil(n/1024) u64
cudaError_t launch_expansion(const KParams& P, unsigned long long* block_sums, unsigned long long* /*offsets: unused*/,
                             bf_expansion* out, unsigned long long cap, cudaStream_t stream, uint32_t* launches) {
  const uint32_t n = P.n_runs;
  if (n == 0) return cudaSuccess;
  const uint32_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  exp_block_sums<<<nb, SCAN_BLOCK, 0, stream>>>(P.exp_counts, n, block_sums);
  exp_scan_emit<<<nb, SCAN_BLOCK, 0, stream>>>(P, block_sums, out, cap);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

}  // namespace bf
