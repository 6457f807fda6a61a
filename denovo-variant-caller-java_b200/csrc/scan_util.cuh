// scan_util.cuh — small order-preserving compaction / scan building blocks
// shared by the READ pipeline (calls in input order) and the VARIANT filter
// (per-person queues and PositionCalls in stream order).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dnv {

constexpr int kScanThreads = 256;  // bit-words per compaction block

// block_sums[b] = popcount of the b-th block of kScanThreads words
static __global__ void __launch_bounds__(kScanThreads)
popc_blocks_kernel(const uint32_t* __restrict__ bits, int64_t n_words, uint32_t* __restrict__ block_sums) {
  const int64_t w = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
  int v = (w < n_words) ? __popc(bits[w]) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ int s[kScanThreads / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; i++) t += s[i];
    block_sums[blockIdx.x] = (uint32_t)t;
  }
}

// One 1024-thread block: exclusive scan of vals[0..n) in place; returns the
// total (valid in every thread).
__device__ __forceinline__ uint32_t block_exclusive_scan_inplace(uint32_t* vals, int n) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < n) ? vals[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t wsum = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, wsum, o);
        if (threadIdx.x >= o) wsum += y;
      }
      s_warp[threadIdx.x] = wsum;
    }
    __syncthreads();
    const uint32_t warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t carry = s_carry;
    if (i < n) vals[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + x;
    __syncthreads();
  }
  return s_carry;
}

static __global__ void __launch_bounds__(1024)
scan_blocks_generic_kernel(uint32_t* __restrict__ block_sums, int n_blocks, unsigned long long* __restrict__ total) {
  const uint32_t t = block_exclusive_scan_inplace(block_sums, n_blocks);
  if (threadIdx.x == 0) *total = t;
}

// Rank (within the whole mask) of the first set bit of this thread's word.
// Must be called by all kScanThreads threads of the block.
__device__ __forceinline__ int64_t word_rank(uint32_t word, const uint32_t* __restrict__ block_offsets) {
  const int cnt = __popc(word);
  int x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  __shared__ int s[kScanThreads / 32];
  if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
  __syncthreads();
  int warp_off = 0;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) warp_off += s[i];
  __syncthreads();
  return (int64_t)block_offsets[blockIdx.x] + warp_off + x - cnt;
}

}  // namespace dnv
