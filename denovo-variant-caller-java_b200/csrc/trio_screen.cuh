// trio_screen.cuh — exact integer screen in front of the fused posterior kernel.
//
// For BAYES / LRT a site can only be called when Lden > T * Lmen.  Let a*, b*, e* be the
// largest genotype likelihoods of dad, mom and child.  Every one of the 750 de novo triples
// carries the factor P(d) P(m) mu / nDen <= mu / 384 (w <= 2 per parent, nDen >= 6), so
//     Lden <= (mu / 384) * (sum_d a_d)(sum_m b_m)(sum_c e_c) <= (mu / 384) * 1000 * a* b* e*.
// If the three most likely genotypes (g_d, g_m, g_c) form a MENDELIAN-consistent triple, that
// single triple already gives
//     Lmen >= P(g_d) P(g_m) T(g_d, g_m, g_c) a b e >= ((1 - mu) / 256) * a b e
// (a, b, e = the likelihoods of the chosen genotypes, each within a factor 2 of the maximum).
// Hence Lden / Lmen <= 8000 * 256 / 384 * mu / (1 - mu) = 5334 mu / (1 - mu): whenever that is
// below the decision threshold (mu < 1.8e-4 for BAYES) the site is provably NOT de novo — also
// under the reference's raw-exp underflow semantics (terms of Lden are then below those of
// Lmen, so Lden flushes to zero first; 0/0 = NaN compares false).  The most likely genotype of
// a person follows from integer counts alone: the best homozygote is the most frequent allele
// (k_hom reads), the best heterozygote the two most frequent ones (k_het reads), and
// r_hom^k_hom vs r_het^k_het is compared as k * log2(r) in 16.16 fixed point.
// So ~99 % of sites are decided with integer instructions and no floating point at all; the
// kernel streams the counts at HBM speed and appends the remaining sites to a list that the
// fused FP64 kernel (trio_fast.cuh) evaluates.
//
// Two levels, so the common case costs ~40 instructions:
//   level 1  the same allele X dominates all three pileups so strongly that XX beats every
//            heterozygote:  n_X * log2(r_hom) >= N * log2(r_het)  (N = depth; implies n_X > N/2,
//            so XX is each person's most likely genotype) -> the triple (XX, XX, XX) is Mendelian.
//   level 2  the general test above (top two alleles per person).  Sites failing level 1 are
//            parked in a warp-private shared-memory queue and level 2 runs on 32 of them at a
//            time, so the ~10 % that need it do not make whole warps diverge.
#pragma once
#include <algorithm>

#include "dnv_internal.cuh"
#include "trio_fast.cuh"

namespace dnv {

constexpr int kScreenThreads = 256;
constexpr int kScreenWarps = kScreenThreads / 32;
constexpr int kScreenBufs = 3;  // tile buffers per warp: 2 tiles in flight ahead of the one being consumed

struct ScreenConsts {
  uint32_t c_hom, c_het;  // log2(r_hom), log2(r_het) in 16.16 fixed point
};

// One person: keys are (count << 2) | allele, so integer max/min order counts and carry the
// allele along (equal counts -> any maximiser is fine).  Returns the allele mask of the most
// likely genotype in bits 0-3 and its two alleles in bits 4-5 / 6-7; bit 8 set = too deep.
__device__ __forceinline__ int screen_person(int k0, int k1, int k2, int k3, const ScreenConsts& k) {
  const int m01 = max(k0, k1), l01 = min(k0, k1), m23 = max(k2, k3), l23 = min(k2, k3);
  const int top = max(m01, m23);                          // most frequent allele
  const int sec = max(min(m01, m23), max(l01, l23));      // second most frequent
  const int k_hom = top >> 2, k_het = k_hom + (sec >> 2);
  const int x = top & 3, y = sec & 3;
  const bool het = (uint32_t)k_het * k.c_het > (uint32_t)k_hom * k.c_hom;  // r_het^k_het > r_hom^k_hom
  const int y2 = het ? y : x;
  return ((1 << x) | (1 << y2)) | (x << 4) | (y2 << 6) | (k_het >= kPowTable ? 256 : 0);
}

// true = provably not de novo (see the header of this file); w = the site's three uint2 words
template <bool kU8>
__device__ __forceinline__ bool screen_site(const uint32_t* w, const ScreenConsts& k) {
  int g[3];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    int k0, k1, k2, k3;
    if (kU8) {
      const uint32_t v = w[p];
      k0 = (int)((v & 0xFFu) << 2);
      k1 = (int)(((v >> 6) & 0x3FCu) | 1u);
      k2 = (int)(((v >> 14) & 0x3FCu) | 2u);
      k3 = (int)(((v >> 22) & 0x3FCu) | 3u);
    } else {
      const uint32_t lo = w[2 * p], hi = w[2 * p + 1];
      k0 = (int)((lo & 0xFFFFu) << 2);
      k1 = (int)(((lo >> 14) & 0x3FFFCu) | 1u);
      k2 = (int)(((hi & 0xFFFFu) << 2) | 2u);
      k3 = (int)(((hi >> 14) & 0x3FFFCu) | 3u);
    }
    g[p] = screen_person(k0, k1, k2, k3, k);
  }
  // isDenovoMap, M/DenovoUtil.java:41-50, on the chosen triple (0 dad, 1 mom, 2 child)
  const int c1 = (g[2] >> 4) & 3, c2 = (g[2] >> 6) & 3;
  const int mendel = ((g[1] >> c1) & (g[0] >> c2) & 1) | ((g[1] >> c2) & (g[0] >> c1) & 1);
  return mendel && !((g[0] | g[1] | g[2]) & 256);
}

// level 1 (see the header): w = the site's packed count words
template <bool kU8>
__device__ __forceinline__ bool screen_level1(const uint32_t* w, const ScreenConsts& k) {
  // candidate X = dad's most frequent allele
  int k0, k1, k2, k3;
  if (kU8) {
    const uint32_t v = w[0];
    k0 = (int)((v & 0xFFu) << 2);
    k1 = (int)(((v >> 6) & 0x3FCu) | 1u);
    k2 = (int)(((v >> 14) & 0x3FCu) | 2u);
    k3 = (int)(((v >> 22) & 0x3FCu) | 3u);
  } else {
    const uint32_t lo = w[0], hi = w[1];
    k0 = (int)((lo & 0xFFFFu) << 2);
    k1 = (int)(((lo >> 14) & 0x3FFFCu) | 1u);
    k2 = (int)(((hi & 0xFFFFu) << 2) | 2u);
    k3 = (int)(((hi >> 14) & 0x3FFFCu) | 3u);
  }
  const int x = max(max(k0, k1), max(k2, k3)) & 3;
  bool ok = true;
#pragma unroll
  for (int p = 0; p < 3; p++) {
    uint32_t nx, depth;
    if (kU8) {
      nx = (w[p] >> (8 * x)) & 0xFFu;
      depth = __dp4a(w[p], 0x01010101u, 0u);
    } else {
      nx = ((x & 2 ? w[2 * p + 1] : w[2 * p]) >> (16 * (x & 1))) & 0xFFFFu;
      depth = __dp2a_lo(w[2 * p], 0x0101u, __dp2a_lo(w[2 * p + 1], 0x0101u, 0u));
    }
    ok &= depth < (uint32_t)kPowTable && nx * k.c_hom >= depth * k.c_het;
  }
  return ok;
}

// Each warp walks tiles of 128 sites (four per lane, 3072 / 1536 bytes); TMA bulk copies keep two
// tiles in flight ahead of the one being consumed, in warp-private buffers.
constexpr int kScreenSites = 128;
constexpr int kScreenPerLane = kScreenSites / 32;
constexpr int kScreenQueue = 64;  // level-2 queue entries (site indices) per warp: at most 31 + 32 parked

template <bool kU8>
__global__ void __launch_bounds__(kScreenThreads)
trio_screen_kernel(const unsigned char* __restrict__ counts, int64_t n, const ScreenConsts k,
                   uint32_t* __restrict__ list, unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kBps = kU8 ? 12 : 24;
  constexpr int kTile = kScreenSites * kBps;
  constexpr int kWords = kBps / 4;  // 32-bit words per site
  unsigned char* tiles = smem_raw;  // [kScreenWarps][kScreenBufs][kTile]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kScreenWarps * kScreenBufs * kTile);
  uint32_t* queues = reinterpret_cast<uint32_t*>(bars + kScreenWarps * kScreenBufs);  // [kScreenWarps][kScreenQueue]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* my_tiles = tiles + (size_t)warp * kScreenBufs * kTile;
  unsigned long long* my_bars = bars + warp * kScreenBufs;
  uint32_t* my_queue = queues + warp * kScreenQueue;
  const int64_t n_tiles = (n + kScreenSites - 1) / kScreenSites;
  const bool tma_ok = (reinterpret_cast<uintptr_t>(counts) & 15) == 0;
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kScreenBufs; b++) mbar_init(&my_bars[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int64_t t, int buf) {
    if (lane == 0 && tma_ok && t < n_tiles && (t + 1) * kScreenSites <= n) {
      mbar_expect_tx(&my_bars[buf], kTile);
      tma_bulk_g2s(my_tiles + buf * kTile, counts + t * kTile, kTile, &my_bars[buf]);
    }
  };
  // level 2 on up to 32 parked sites (their counts are re-read through L2, where the tile that
  // just streamed in still sits); what it cannot decide goes to the FP64 kernel's list
  int qcount = 0;
  auto drain = [&](int take) {
    __syncwarp();
    bool fail = false;
    uint32_t site = 0;
    if (lane < take) {
      site = my_queue[qcount - take + lane];
      uint32_t w[kWords];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(counts + (size_t)site * kBps);
#pragma unroll
      for (int q = 0; q < kWords; q++) w[q] = __ldg(src + q);
      fail = !screen_site<kU8>(w, k);
    }
    const uint32_t fw = __ballot_sync(0xffffffffu, fail);
    if (fw) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&counters[1], (unsigned long long)__popc(fw));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (fail) list[base + __popc(fw & ((1u << lane) - 1))] = site;
    }
    qcount -= take;
    __syncwarp();
  };
  const int64_t warps_total = (int64_t)gridDim.x * kScreenWarps;
  const int64_t t0 = (int64_t)blockIdx.x * kScreenWarps + warp;
#pragma unroll
  for (int b = 0; b < kScreenBufs - 1; b++) issue(t0 + b * warps_total, b);
  uint32_t phases = 0;
  int buf = 0;
  for (int64_t t = t0; t < n_tiles; t += warps_total, buf = (buf + 1) % kScreenBufs) {
    uint32_t w[kScreenPerLane][kWords];
    const int64_t site0 = t * kScreenSites + lane;
    if (tma_ok && (t + 1) * kScreenSites <= n) {
      mbar_wait(&my_bars[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(my_tiles + buf * kTile);
#pragma unroll
      for (int h = 0; h < kScreenPerLane; h++)
#pragma unroll
        for (int q = 0; q < kWords; q++) w[h][q] = src[(h * 32 + lane) * kWords + q];
    } else {
#pragma unroll
      for (int h = 0; h < kScreenPerLane; h++) {
        const int64_t site = site0 + h * 32;
#pragma unroll
        for (int q = 0; q < kWords; q++)
          w[h][q] = site < n ? __ldg(reinterpret_cast<const uint32_t*>(counts + site * kBps) + q) : 0u;
      }
    }
    __syncwarp();  // the buffer consumed in the previous trip may now be refilled
    issue(t + (kScreenBufs - 1) * warps_total, (buf + kScreenBufs - 1) % kScreenBufs);
#pragma unroll
    for (int h = 0; h < kScreenPerLane; h++) {
      const int64_t site = site0 + h * 32;
      const bool park = site < n && !screen_level1<kU8>(w[h], k);
      const uint32_t pw = __ballot_sync(0xffffffffu, park);
      if (pw) {
        if (park) my_queue[qcount + __popc(pw & ((1u << lane) - 1))] = (uint32_t)site;
        qcount += __popc(pw);
        if (qcount >= 32) drain(32);
      }
    }
  }
  if (qcount > 0) drain(qcount);
}

template <bool kU8>
static int launch_screen(dnv_ctx* ctx, const void* d_counts, int64_t n, const ScreenConsts& sk) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  constexpr int kTile = kScreenSites * (kU8 ? 12 : 24);
  const size_t smem = (size_t)kScreenWarps * kScreenBufs * kTile + kScreenWarps * kScreenBufs * 8 +
                      (size_t)kScreenWarps * kScreenQueue * 4;
  auto kern = trio_screen_kernel<kU8>;
  DNV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  DNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kScreenThreads, smem));
  const int64_t want = ((n + kScreenSites - 1) / kScreenSites + kScreenWarps - 1) / kScreenWarps;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(want, 1), (int64_t)sms * std::max(per_sm, 1));
  kern<<<grid, kScreenThreads, smem, ctx->stream>>>((const unsigned char*)d_counts, n, sk, ctx->d_list, ctx->d_counters);
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

}  // namespace dnv
