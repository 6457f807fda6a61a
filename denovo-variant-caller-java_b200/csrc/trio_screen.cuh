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
// So ~99 % of sites are decided with ~100 integer instructions and no floating point at all;
// the kernel streams the counts at HBM speed and appends the remaining sites to a list that the
// fused FP64 kernel (trio_fast.cuh) evaluates.
#pragma once
#include "dnv_internal.cuh"
#include "trio_fast.cuh"

namespace dnv {

constexpr int kScreenThreads = 256;
constexpr int kScreenWarps = kScreenThreads / 32;
constexpr int kScreenBufs = 4;  // tiles in flight per warp: 3 ahead of the one being consumed

struct ScreenConsts {
  uint32_t c_hom, c_het;  // log2(r_hom), log2(r_het) in 16.16 fixed point
};

struct ScreenShared {
  alignas(16) unsigned char tile[kScreenWarps][kScreenBufs][kWarpTileBytes];
  alignas(8) unsigned long long bar[kScreenWarps][kScreenBufs];
};

// most frequent allele (index) and its count; ties -> lowest index (any maximiser is fine)
__device__ __forceinline__ void argmax4(int n0, int n1, int n2, int n3, int& x, int& v) {
  const int x01 = n1 > n0 ? 1 : 0, m01 = max(n0, n1);
  const int x23 = n3 > n2 ? 3 : 2, m23 = max(n2, n3);
  x = m23 > m01 ? x23 : x01;
  v = max(m01, m23);
}

// true = provably not de novo (see the header of this file)
__device__ __forceinline__ bool screen_site(const int* c, const ScreenConsts& k) {
  int a1[3], a2[3], mask[3];
  bool ok = true;
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int* n = c + 4 * p;
    int x, k_hom, y, second;
    argmax4(n[0], n[1], n[2], n[3], x, k_hom);
    argmax4(x == 0 ? -1 : n[0], x == 1 ? -1 : n[1], x == 2 ? -1 : n[2], x == 3 ? -1 : n[3], y, second);
    const int k_het = k_hom + second;
    ok &= k_het < kPowTable;  // deeper than the tables: the fused kernel hands it to the exact path
    const bool het = (uint32_t)k_het * k.c_het > (uint32_t)k_hom * k.c_hom;
    a1[p] = x;
    a2[p] = het ? y : x;
    mask[p] = (1 << x) | (het ? (1 << y) : 0);
  }
  // isDenovoMap, M/DenovoUtil.java:41-50, on the chosen triple (0 dad, 1 mom, 2 child)
  const int c1 = a1[2], c2 = a2[2];
  const int mendel = ((mask[1] >> c1) & (mask[0] >> c2) & 1) | ((mask[1] >> c2) & (mask[0] >> c1) & 1);
  return ok && mendel;
}

template <bool kU8>
__global__ void __launch_bounds__(kScreenThreads)
trio_screen_kernel(const unsigned char* __restrict__ counts, int64_t n, const ScreenConsts k,
                   uint32_t* __restrict__ list, unsigned long long* __restrict__ counters) {
  __shared__ ScreenShared s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kBps = kU8 ? 12 : 24;
  constexpr int kTile = 32 * kBps;
  const int64_t n_tiles = (n + 31) / 32;
  const bool tma_ok = (reinterpret_cast<uintptr_t>(counts) & 15) == 0;
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kScreenBufs; b++) mbar_init(&s.bar[warp][b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int64_t t, int buf) {
    if (lane == 0 && tma_ok && t < n_tiles && (t + 1) * 32 <= n) {
      mbar_expect_tx(&s.bar[warp][buf], kTile);
      tma_bulk_g2s(s.tile[warp][buf], counts + t * kTile, kTile, &s.bar[warp][buf]);
    }
  };
  const int64_t warps_total = (int64_t)gridDim.x * kScreenWarps;
  const int64_t t0 = (int64_t)blockIdx.x * kScreenWarps + warp;
#pragma unroll
  for (int b = 0; b < kScreenBufs - 1; b++) issue(t0 + b * warps_total, b);
  uint32_t phases = 0;
  int buf = 0;
  for (int64_t t = t0; t < n_tiles; t += warps_total, buf = (buf + 1) % kScreenBufs) {
    const int64_t site = t * 32 + lane;
    int c[12];
    if (tma_ok && (t + 1) * 32 <= n) {
      mbar_wait(&s.bar[warp][buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      if (kU8) unpack_counts8(reinterpret_cast<const uint32_t*>(s.tile[warp][buf] + lane * 12), c);
      else unpack_counts(reinterpret_cast<const uint2*>(s.tile[warp][buf] + lane * 24), c);
    } else if (site < n) {
      if (kU8) unpack_counts8(reinterpret_cast<const uint32_t*>(counts + site * 12), c);
      else unpack_counts(reinterpret_cast<const uint2*>(counts + site * 24), c);
    } else {
#pragma unroll
      for (int q = 0; q < 12; q++) c[q] = 0;
    }
    __syncwarp();  // the buffer consumed in the previous trip may now be refilled
    issue(t + (kScreenBufs - 1) * warps_total, (buf + kScreenBufs - 1) % kScreenBufs);
    const bool fail = site < n && !screen_site(c, k);
    const uint32_t fw = __ballot_sync(0xffffffffu, fail);
    if (fw) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&counters[1], (unsigned long long)__popc(fw));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (fail) list[base + __popc(fw & ((1u << lane) - 1))] = (uint32_t)site;
    }
  }
}

}  // namespace dnv
