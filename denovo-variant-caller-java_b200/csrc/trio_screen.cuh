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
// below the decision threshold the site is provably NOT de novo — also under the reference's
// raw-exp underflow semantics: every term of Lden is below that one term M of Lmen by the factor
// 5.33 mu, so Lden flushes to zero first (0/0 = NaN compares false); and when M is subnormal each
// of the <= 750 terms of Lden that does not flush may round up by half an ulp, at most doubling it,
// so 2 * 5334 mu / (1 - mu) < T still proves Lden <= T * Lmen.  tables.cpp enables the screen
// with a factor 4 (mu < 4.7e-5 for BAYES; the default is 1e-8).  The most likely genotype of
// a person follows from integer counts alone: the best homozygote is the most frequent allele
// (k_hom reads), the best heterozygote the two most frequent ones (k_het reads), and
// r_hom^k_hom vs r_het^k_het is compared as k * log2(r) in 16.16 fixed point.
// So ~99 % of sites are decided with integer instructions and no floating point at all; the
// kernel streams the counts at HBM speed and hands the remaining sites, in ascending order, to
// the fused FP64 list kernel (trio_list.cuh).
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

struct ScreenConsts {
  uint32_t c_hom, c_het;  // log2(r_hom), log2(r_het) in 16.16 fixed point
  uint32_t w[4];          // level 1 weights per candidate allele (Tables::scr_w)
  int32_t l3_lambda;      // level 3: log2 threshold (16.16), 0 = level 3 off
  int32_t l3_max_depth;   // level 3: largest total depth it may decide
  uint32_t shift32;       // 32 (a run-time value the assembler cannot fold: x >> 32 == 0 orders the tile refill)
  unsigned long long* ktrace;  // DNV_KTRACE time line of the lane, or null
  const unsigned int* ktrace_seq;
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

// Level 3 — the sites level 2 leaves are almost all of ONE kind: both parents clearly homozygous (alleles x_d, x_m), and
// a child whose most likely genotype c* they cannot have transmitted (the de novo pattern).  Most of them are still
// not called: one more read in a parent would have made "the parent is heterozygous and we missed it" the better
// explanation.  With d* = x_d x_d, m* = x_m x_m and every person's best genotype at least 13 bits ahead of its
// second best (so that the nine other genotypes add at most 9 * 2^-13 to a person's likelihood sum):
//     Lden <= P(d*) P(m*) (mu / 9) a* b* E + (mu / 384) (A B - a* b*) E  <=  1.0143 * (mu / 2304) a* b* e*
// (nDen = 9 for two homozygous parents; every other parental pair carries a factor 2^-13), while each Mendelian
// ALTERNATIVE is a single term of Lmen with the same weight (1 - mu) / 256:
//     child x_d x_m instead of c*                                  (1/16)(1/16)(1 - mu)        a*  b*  e_alt
//     dad x_d u instead of x_d x_d, if c* = {u, x_m}               (1/8)(1/16)((1 - mu) / 2)   a_alt b* e*
//     mom x_m u instead of x_m x_m, if c* = {u, x_d}               (1/16)(1/8)((1 - mu) / 2)   a*  b_alt e*
// so Lden < T * Lmen as soon as the likelihood ratio best / alternative of ONE of them is below
// 9 T (1 - mu) / (1.0143 mu) — compared as k * log2(r) in 16.16 fixed point like level 2 (tables.cpp folds the
// constant, the fixed-point error and a guard for FP64 rounding into l3_lambda).  The site's total depth is
// limited so that both sides stay normal doubles in the reference's raw-exp sums (no flush / subnormal effects).
// Checked against the oracle in tests/test_screen_theory.py.  n = counts [person][allele].
__device__ __forceinline__ int pick4(int n0, int n1, int n2, int n3, int a) {  // n[a] without dynamic indexing
  return a == 0 ? n0 : a == 1 ? n1 : a == 2 ? n2 : n3;
}
// top allele (first maximiser) and the largest count among the others
__device__ __forceinline__ void top_two(int n0, int n1, int n2, int n3, int& xa, int& k1, int& k2) {
  xa = 0; k1 = n0;
  if (n1 > k1) { k1 = n1; xa = 1; }
  if (n2 > k1) { k1 = n2; xa = 2; }
  if (n3 > k1) { k1 = n3; xa = 3; }
  k2 = max(max(xa == 0 ? 0 : n0, xa == 1 ? 0 : n1), max(xa == 2 ? 0 : n2, xa == 3 ? 0 : n3));
}
__device__ __forceinline__ bool screen_level3(const int (&n)[3][4], const ScreenConsts& k) {
  const int c_hom = (int)k.c_hom, c_het = (int)k.c_het;
  constexpr int kGap = 13 << 16;
  const int total = ((n[0][0] + n[0][1]) + (n[0][2] + n[0][3])) + ((n[1][0] + n[1][1]) + (n[1][2] + n[1][3])) +
                    ((n[2][0] + n[2][1]) + (n[2][2] + n[2][3]));
  if (total > k.l3_max_depth) return false;
  // parents: the homozygote of the top allele, 13 bits ahead of everything else
  int xd, xm, k1, k2;
  top_two(n[0][0], n[0][1], n[0][2], n[0][3], xd, k1, k2);
  if (k1 * c_hom < max(k2 * c_hom, (k1 + k2) * c_het) + kGap) return false;
  const int nd_x = k1;
  top_two(n[1][0], n[1][1], n[1][2], n[1][3], xm, k1, k2);
  if (k1 * c_hom < max(k2 * c_hom, (k1 + k2) * c_het) + kGap) return false;
  const int nm_x = k1;
  // child: top three alleles, best genotype (hom of the top one or het of the top two), 13 bits ahead of its second best
  int t1;
  top_two(n[2][0], n[2][1], n[2][2], n[2][3], t1, k1, k2);
  int t2 = -1, kk = -1;  // first allele other than t1 holding k2 reads
  if (t1 != 0 && n[2][0] > kk) { kk = n[2][0]; t2 = 0; }
  if (t1 != 1 && n[2][1] > kk) { kk = n[2][1]; t2 = 1; }
  if (t1 != 2 && n[2][2] > kk) { kk = n[2][2]; t2 = 2; }
  if (t1 != 3 && n[2][3] > kk) { kk = n[2][3]; t2 = 3; }
  const int k3 = max(max((t1 == 0 || t2 == 0) ? 0 : n[2][0], (t1 == 1 || t2 == 1) ? 0 : n[2][1]),
                     max((t1 == 2 || t2 == 2) ? 0 : n[2][2], (t1 == 3 || t2 == 3) ? 0 : n[2][3]));
  const int hom_c = k1 * c_hom, het_c = (k1 + k2) * c_het;
  const bool het = het_c > hom_c;
  const int e_star = het ? het_c : hom_c;
  const int second = het ? max(hom_c, (k1 + k3) * c_het) : max(k2 * c_hom, het_c);
  if (e_star < second + kGap) return false;
  const int a1 = t1, a2 = het ? t2 : t1;  // alleles of c*
  // alternative 1: the child is x_d x_m after all
  const int nc_xd = pick4(n[2][0], n[2][1], n[2][2], n[2][3], xd), nc_xm = pick4(n[2][0], n[2][1], n[2][2], n[2][3], xm);
  const int e_alt = xd == xm ? nc_xd * c_hom : (nc_xd + nc_xm) * c_het;
  int ratio = e_star - e_alt;
  if (a1 == xm || a2 == xm) {  // alternative 2: dad is x_d u, u = the child's other allele
    const int u = a1 == xm ? a2 : a1;
    if (u != xd) ratio = min(ratio, nd_x * c_hom - (nd_x + pick4(n[0][0], n[0][1], n[0][2], n[0][3], u)) * c_het);
  }
  if (a1 == xd || a2 == xd) {  // alternative 3: mom is x_m u
    const int u = a1 == xd ? a2 : a1;
    if (u != xm) ratio = min(ratio, nm_x * c_hom - (nm_x + pick4(n[1][0], n[1][1], n[1][2], n[1][3], u)) * c_het);
  }
  return ratio < k.l3_lambda;
}

// level 3 on a site's packed count words
template <bool kU8>
__device__ __forceinline__ bool screen_level3_words(const uint32_t* w, const ScreenConsts& k) {
  int n[3][4];
#pragma unroll
  for (int p = 0; p < 3; p++) {
    if (kU8) {
#pragma unroll
      for (int a = 0; a < 4; a++) n[p][a] = (int)((w[p] >> (8 * a)) & 0xFFu);
    } else {
      n[p][0] = (int)(w[2 * p] & 0xFFFFu); n[p][1] = (int)(w[2 * p] >> 16);
      n[p][2] = (int)(w[2 * p + 1] & 0xFFFFu); n[p][3] = (int)(w[2 * p + 1] >> 16);
    }
  }
  return screen_level3(n, k);
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

// signed-byte dot products with unsigned counts (PTX dp4a / dp2a, SASS IDP): counts may use all 8 / 16 bits
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_us(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// sum_b n_b * wt_b for one person (wt = four signed bytes A,C,G,T)
template <bool kU8>
__device__ __forceinline__ int screen_dot(const uint32_t* w, int p, uint32_t wt) {
  if (kU8) return dp4a_us(w[p], wt, 0);
  return dp2a_lo_us(w[2 * p], wt, dp2a_hi_us(w[2 * p + 1], wt, 0));
}

// level 1 (see the header): w = the site's packed count words.  The candidate X comes from two
// sign tests on dad's pileup (G+T > A+C, C+T > A+G), which name the dominant allele whenever there
// is one; the test itself is  n_X * a >= (N - n_X) * h  per person with signed-byte weights chosen on
// the host so that it implies n_X * log2(r_hom) >= N * log2(r_het) (tables.cpp) — one dot-product
// instruction per count word, no overflow at any depth.  Any X is sound: a wrong guess only parks the site.
template <bool kU8>
__device__ __forceinline__ bool screen_level1(const uint32_t* w, const ScreenConsts& k) {
  const bool hi2 = screen_dot<kU8>(w, 0, 0x0101FFFFu) > 0;  // -A -C +G +T
  const bool odd = screen_dot<kU8>(w, 0, 0x01FF01FFu) > 0;  // -A +C -G +T
  const uint32_t wlo = odd ? k.w[1] : k.w[0], whi = odd ? k.w[3] : k.w[2];
  const uint32_t wt = hi2 ? whi : wlo;
  const int s = screen_dot<kU8>(w, 0, wt) | screen_dot<kU8>(w, 1, wt) | screen_dot<kU8>(w, 2, wt);
  return s >= 0;
}

// Each warp walks a CONTIGUOUS range of 128-site tiles (four sites per lane, 3072 / 1536 bytes per
// tile); TMA bulk copies keep kScreenBufs tiles in flight in warp-private buffers.  Level 2 is
// software pipelined: the counts of 32 parked sites are requested (through L2, where the tile that
// just streamed in still sits) one batch ahead of their evaluation.  Parked and undecided sites
// move through warp-private FIFO rings, so every warp emits its undecided sites in ascending order
// into its own segment of `list` (the segment starts at the warp's first site index, so it cannot
// overflow) and records how many it wrote in wcount[warp]: the concatenation of the segments in
// warp order is the sorted list of undecided sites — no atomics, no sort (trio_list.cuh consumes it).
constexpr int kScreenSites = 128;
constexpr int kScreenPerLane = kScreenSites / 32;
constexpr int kScreenRing = 64;       // ring entries (site indices): at most 31 + 32 waiting
constexpr int kScreenMaxWarps = 4096; // upper bound of gridDim.x * kScreenWarps (sizes wcount and its scan: 16 KB of ListShared)

// tiles [lo, hi) of warp g when n_tiles tiles are dealt to W warps in contiguous, balanced ranges
__host__ __device__ __forceinline__ uint32_t screen_range_lo(uint32_t g, uint32_t base, uint32_t rem) {
  return g * base + (g < rem ? g : rem);
}

// kScreenBufs: tile buffers per warp, refilled as soon as their tile sits in registers
template <bool kU8, int kScreenBufs>
__global__ void __launch_bounds__(kScreenThreads, 4)
trio_screen_kernel(const unsigned char* __restrict__ counts, int64_t n, const ScreenConsts k,
                   uint32_t* __restrict__ list, uint32_t* __restrict__ wcount) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kBps = kU8 ? 12 : 24;
  constexpr int kTile = kScreenSites * kBps;
  constexpr int kWords = kBps / 4;  // 32-bit words per site
  constexpr uint32_t kFull = 0xffffffffu;
  unsigned char* tiles = smem_raw;  // [kScreenWarps][kScreenBufs][kTile]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kScreenWarps * kScreenBufs * kTile);
  uint32_t* rings = reinterpret_cast<uint32_t*>(bars + kScreenWarps * kScreenBufs);  // [kScreenWarps][parked | undecided]
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) ktrace_stamp(k.ktrace, k.ktrace_seq, 0, false);
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned char* my_tiles = tiles + (size_t)warp * kScreenBufs * kTile;
  unsigned long long* my_bars = bars + warp * kScreenBufs;
  uint32_t* my_queue = rings + warp * (2 * kScreenRing);
  uint32_t* my_fails = my_queue + kScreenRing;
  // a chunk holds fewer than 2^31 sites (api.cu: kDeviceChunk), so tile and site indices fit 32 bits
  const uint32_t n_sites = (uint32_t)n;
  const uint32_t n_tiles = (n_sites + kScreenSites - 1) / kScreenSites;
  const uint32_t n_full = (reinterpret_cast<uintptr_t>(counts) & 15) == 0 ? n_sites / kScreenSites : 0u;  // TMA-eligible tiles
  const uint32_t warps_total = gridDim.x * kScreenWarps, g = blockIdx.x * kScreenWarps + warp;
  const uint32_t base = n_tiles / warps_total, rem = n_tiles % warps_total;
  const uint32_t t_lo = screen_range_lo(g, base, rem), t_hi = screen_range_lo(g + 1, base, rem);
  uint32_t* my_list = list + (size_t)t_lo * kScreenSites;
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kScreenBufs; b++) mbar_init(&my_bars[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](uint32_t t, int buf) {
    if (lane == 0 && t < t_hi && t < n_full) {
      mbar_expect_tx(&my_bars[buf], kTile);
      tma_bulk_g2s(my_tiles + buf * kTile, counts + (size_t)t * kTile, kTile, &my_bars[buf]);
    }
  };
  // warp-uniform ring state: parked sites, undecided sites, the batch in flight, sites already written out
  uint32_t qhead = 0, qcount = 0, fhead = 0, fcount = 0, ptake = 0, written = 0;
  uint32_t pw[kWords], psite = 0;  // the batch in flight: one parked site per lane
#pragma unroll
  for (int q = 0; q < kWords; q++) pw[q] = 0;
  // the oldest `take` sites level 2 could not decide: level 3 (32 of them at a time, so the ~1 % of sites that need
  // it cost one warp-wide evaluation per 32), the survivors -> this warp's segment of the list, still ascending
  auto flush = [&](uint32_t take) {
    uint32_t site = 0;
    bool keep = lane < take;
    if (keep) {
      site = my_fails[(fhead + lane) & (kScreenRing - 1)];
      if (k.l3_lambda > 0) {
        uint32_t lw[kWords];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(counts + (size_t)site * kBps);
#pragma unroll
        for (int q = 0; q < kWords; q++) lw[q] = __ldg(src + q);
        keep = !screen_level3_words<kU8>(lw, k);
      }
    }
    const uint32_t km = __ballot_sync(kFull, keep);
    if (keep) my_list[written + __popc(km & lt_mask)] = site;
    written += __popc(km);
    fhead += take;
    fcount -= take;
    __syncwarp();
  };
  auto evaluate = [&]() {  // level 2 on the batch in flight
    const bool fail = lane < ptake && !screen_site<kU8>(pw, k);
    const uint32_t fw = __ballot_sync(kFull, fail);
    ptake = 0;
    if (fw) {
      if (fail) my_fails[(fhead + fcount + __popc(fw & lt_mask)) & (kScreenRing - 1)] = psite;
      fcount += __popc(fw);
      __syncwarp();
      if (fcount >= 32) flush(32);
    }
  };
  auto fetch = [&](uint32_t take) {  // request the counts of the oldest `take` parked sites
    __syncwarp();
    if (lane < take) {
      psite = my_queue[(qhead + lane) & (kScreenRing - 1)];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(counts + (size_t)psite * kBps);
#pragma unroll
      for (int q = 0; q < kWords; q++) pw[q] = __ldg(src + q);
    }
    ptake = take;
    qhead += take;
    qcount -= take;
    __syncwarp();
  };
#pragma unroll
  for (int b = 0; b < kScreenBufs; b++) issue(t_lo + b, b);
  uint32_t phases = 0;
  int buf = 0;
  for (uint32_t t = t_lo; t < t_hi; t++, buf = buf == kScreenBufs - 1 ? 0 : buf + 1) {
    uint32_t w[kScreenPerLane][kWords];
    const uint32_t site0 = t * kScreenSites + lane;
    if (t < n_full) {
      mbar_wait(&my_bars[buf], (phases >> buf) & 1u);
      phases ^= 1u << buf;
      if (kU8) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(my_tiles + buf * kTile);
#pragma unroll
        for (int h = 0; h < kScreenPerLane; h++)
#pragma unroll
          for (int q = 0; q < kWords; q++) w[h][q] = src[(h * 32 + lane) * kWords + q];
      } else {  // 24-byte sites: three conflict-free 64-bit loads
        const uint2* src = reinterpret_cast<const uint2*>(my_tiles + buf * kTile);
#pragma unroll
        for (int h = 0; h < kScreenPerLane; h++)
#pragma unroll
          for (int q = 0; q < 3; q++) {
            const uint2 v = src[(h * 32 + lane) * 3 + q];
            w[h][2 * q] = v.x;
            w[h][2 * q + 1] = v.y;
          }
      }
    } else {
#pragma unroll
      for (int h = 0; h < kScreenPerLane; h++) {
        const uint32_t site = site0 + h * 32;
#pragma unroll
        for (int q = 0; q < kWords; q++)
          w[h][q] = site < n_sites ? __ldg(reinterpret_cast<const uint32_t*>(counts + (size_t)site * kBps) + q) : 0u;
      }
    }
    // The tile now lives in registers: refill the SAME buffer at once, so kScreenBufs tiles per warp
    // stay in flight.  The refill is made data-dependent on the last word read (shared-memory loads
    // of a warp complete in order), so the bulk copy cannot overtake the reads it overwrites.
    uint32_t fence_word;
    asm volatile("shr.b32 %0, %1, %2;" : "=r"(fence_word) : "r"(w[kScreenPerLane - 1][kWords - 1]), "r"(k.shift32));  // == 0
    __syncwarp();
    issue(t + kScreenBufs + fence_word, buf);
#pragma unroll
    for (int h = 0; h < kScreenPerLane; h++) {
      const uint32_t site = site0 + h * 32;
      const bool park = site < n_sites && !screen_level1<kU8>(w[h], k);
      const uint32_t pm = __ballot_sync(kFull, park);
      if (pm) {
        if (park) my_queue[(qhead + qcount + __popc(pm & lt_mask)) & (kScreenRing - 1)] = site;
        qcount += __popc(pm);
        if (qcount >= 32) {
          if (ptake) evaluate();
          fetch(32);
        }
      }
    }
  }
  if (ptake) evaluate();
  if (qcount > 0) {
    fetch(qcount);
    evaluate();
  }
  if (fcount > 0) flush(fcount);
  if (lane == 0) wcount[g] = written;
  if (lane == 0) ktrace_stamp(k.ktrace, k.ktrace_seq, 0, true);
}

// launch configuration of the screen: resident CTAs per SM and tile buffers per warp.
// Default: THREE CTAs per SM, and room for one CTA of the tail kernels (trio_list.cuh) beside them: registers
// work out by themselves (3 * 256 * 64 of 65536), shared memory is padded so that a FOURTH screen CTA — of the
// next chunk, launched on another lane while this one is still running — cannot take that room
// (4 * (57 KB + 1 KB) > 228 KB per SM; 3 * 58 KB leave 54 KB: ListShared is 42 KB).
// DNV_SCREEN_CTAS = 2 | 3 | 4 and DNV_SCREEN_BUFS = 2 | 3 are A/B knobs (4 = the round-1 configuration: four CTAs
// per SM and no room for the tail).
struct ScreenConfig { int per_sm, bufs; };
inline ScreenConfig screen_config() {
  static const ScreenConfig c = [] {
    const char* e = getenv("DNV_SCREEN_CTAS");
    const char* b = getenv("DNV_SCREEN_BUFS");
    ScreenConfig r;
    r.per_sm = e ? atoi(e) : 2;
    if (r.per_sm < 1 || r.per_sm > 4) r.per_sm = 2;
    r.bufs = b ? atoi(b) : 2;
    if (r.bufs < 2 || r.bufs > 4) r.bufs = 2;
    return r;
  }();
  return c;
}

// grid of the screen (and so the warp -> tile-range map the list kernel has to reproduce)
template <bool kU8, int kBufs>
static int screen_grid(dnv_ctx* ctx, int64_t n, size_t* smem_out) {
  const int sms = ctx->sms;
  constexpr int kTile = kScreenSites * (kU8 ? 12 : 24);
  size_t smem = (size_t)kScreenWarps * kBufs * kTile + kScreenWarps * kBufs * 8 + (size_t)kScreenWarps * 2 * kScreenRing * 4;
  const int want_per_sm = screen_config().per_sm;
  if (want_per_sm == 3) smem = std::max<size_t>(smem, 57 * 1024);
  if (want_per_sm == 2) smem = std::max<size_t>(smem, 76 * 1024);
  if (want_per_sm == 1) smem = std::max<size_t>(smem, 114 * 1024);
  auto kern = trio_screen_kernel<kU8, kBufs>;
  int& per_sm = ctx->screen_per_sm[kU8 ? 1 : 0];
  if (per_sm == 0) {  // first launch on this context
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    // the SM keeps its largest shared-memory configuration while the screen is resident: a smaller carve-out (just
    // enough for the screen's own CTAs) would have to be re-made, on a drained SM, before a tail CTA could join
    if (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return -1;
    int q = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kScreenThreads, smem) != cudaSuccess) return -1;
    per_sm = std::max(std::min(q, want_per_sm), 1);
  }
  const int64_t want = ((n + kScreenSites - 1) / kScreenSites + kScreenWarps - 1) / kScreenWarps;
  int64_t grid = std::min<int64_t>(std::max<int64_t>(want, 1), (int64_t)sms * std::max(per_sm, 1));
  grid = std::min<int64_t>(grid, kScreenMaxWarps / kScreenWarps);
  *smem_out = smem;
  return (int)grid;
}

template <bool kU8, int kBufs>
static int launch_screen_bufs(dnv_ctx* ctx, TrioLane& ln, cudaStream_t st, const void* d_counts, int64_t n, const ScreenConsts& sk,
                              int* grid_out) {
  size_t smem = 0;
  const int grid = screen_grid<kU8, kBufs>(ctx, n, &smem);
  if (grid < 1) { set_error("trio_screen_kernel: launch configuration failed"); return DNV_E_CUDA; }
  trio_screen_kernel<kU8, kBufs><<<grid, kScreenThreads, smem, st>>>((const unsigned char*)d_counts, n, sk, ln.d_list, ln.d_wcount);
  DNV_CUDA(cudaGetLastError());
  *grid_out = grid;
  return DNV_OK;
}

template <bool kU8>
static int launch_screen(dnv_ctx* ctx, TrioLane& ln, cudaStream_t st, const void* d_counts, int64_t n, const ScreenConsts& sk,
                         int* grid_out) {
  if (screen_config().bufs == 4) return launch_screen_bufs<kU8, 4>(ctx, ln, st, d_counts, n, sk, grid_out);
  if (screen_config().bufs == 3) return launch_screen_bufs<kU8, 3>(ctx, ln, st, d_counts, n, sk, grid_out);
  return launch_screen_bufs<kU8, 2>(ctx, ln, st, d_counts, n, sk, grid_out);
}

}  // namespace dnv
