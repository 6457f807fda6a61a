// trio_fast.cuh — fused likelihood + trio-posterior kernel (kernel (b) of the
// north star), hand-written for sm_100a.  One thread owns one site; a CTA walks
// tiles of 128 sites whose 3 KB of packed counts are staged into shared memory
// by TMA bulk copies (cp.async.bulk + mbarrier, double buffered).
//
// BAYES / LRT decisions use the factored LINEAR-domain form.
// With N_p reads of person p and mismatch likelihood q = eps/3,
//   exp(L_p[XX]) = q^N_p * r_hom^(n_X)        r_hom = (1-eps)/q
//   exp(L_p[XY]) = q^N_p * r_het^(n_X+n_Y)    r_het = ((1-2eps/3)/2)/q
// Counts are integers, so every power comes from a shared-memory table and NO
// exp() is evaluated: inside a zygosity class values are taken relative to the
// class maximum (r^-(kmax-k), one LDS.64), and the two class maxima r^kmax are
// read as (mantissa, exponent) pairs so each person is rescaled by an exact
// power of two.  The 10x10x10 sum then collapses to the sparse straight-line
// code of trio_sparse_gen.cuh (335 FP64 instructions instead of 2000 dense FMAs
// + 30 exp).  The unshifted magnitudes log2(Lden), log2(Lmen) are tracked so the
// reference's raw-exp underflow semantics are reproduced: a site
// whose decision could depend on rounding (|den - T*men| <= 1e-9 relative), on
// subnormal / flushed terms (gray zone), or whose depth exceeds the tables is
// NOT decided here — it goes to the hard list and is re-evaluated by the exact
// direct-form kernel (trio_exact.cuh).
//
// MAP decisions (and maxGenoType of the dense record) use the exact log-domain
// sweep: bit-identical lls, strict '>' in dad-major order.
#pragma once
#include <stdlib.h>

#include "dnv_internal.cuh"
#include "trio_exact.cuh"

namespace dnv {

constexpr int kFastThreads = 128;
#ifndef DNV_FAST_MINBLOCKS
#define DNV_FAST_MINBLOCKS 3   // resident CTAs per SM the BAYES / LRT all-sites kernel is compiled for (3: 160 registers)
#endif
constexpr int kTileBytes = kFastThreads * 24;

constexpr int kFastWarps = kFastThreads / 32;
constexpr int kWarpTileBytes = 32 * 24;  // one warp's tile: 32 sites x 24 B

struct FastShared {
  double pw_inv[2][kPowTable];  // r_j^-i
  double pw_m[2][kPowTable];    // mantissa of r_j^k
  int32_t pw_e[2][kPowTable];   // exponent of r_j^k
  double rel[1000];             // exact path: getRelationshipLogLikelihood
  uint32_t isden[32];           // exact path: isDenovoMap
  double rmax[110];             // exact MAP pruning bounds
  double warp_L[kFastWarps][32];  // exact path: per-warp read log-likelihoods
  alignas(16) unsigned char tile[kFastWarps][2][kWarpTileBytes];  // sized for the uint16 layout
  alignas(8) unsigned long long bar[kFastWarps][2];
};

// site-invariant scalars, passed by value (constant bank: no global loads in the hot loop)
struct FastConsts {
  double k_men, kd9, kd8, kd7, kd6, thr, mm2, lrt_threshold, hm, ht, mm;
  int method, lin_ok;
};

inline FastConsts make_fast_consts(const Tables& t) {
  FastConsts k;
  k.k_men = t.k_men; k.kd9 = t.k_den[0]; k.kd8 = 2.0 * t.k_den[1]; k.kd7 = 4.0 * t.k_den[2]; k.kd6 = 4.0 * t.k_den[3];
  k.thr = t.lrt_thr; k.mm2 = t.mm2; k.lrt_threshold = t.lrt_threshold; k.hm = t.hm; k.ht = t.ht; k.mm = t.mm;
  k.method = t.method; k.lin_ok = t.fast_ok;
  return k;
}

// 2^sh * m for m in [1,2), sh <= 0 (flushes below 2^-1000: irrelevant at 1e-9, see zone logic)
__device__ __forceinline__ double scale_pow2(double m, int sh) {
  const int hi = __double2hiint(m) + (sh << 20);
  return (sh < -1000) ? 0.0 : __hiloint2double(hi, __double2loint(m));
}

// power-table accessors: shared memory (hot kernel) or global/L1 (per-call records)
struct SmemPow {
  const FastShared& s;
  __device__ __forceinline__ double inv(int j, int i) const { return s.pw_inv[j][i]; }
  __device__ __forceinline__ double mant(int j, int i) const { return s.pw_m[j][i]; }
  __device__ __forceinline__ int expo(int j, int i) const { return s.pw_e[j][i]; }
};
struct GmemPow {
  const Tables* t;
  __device__ __forceinline__ double inv(int j, int i) const { return __ldg(&t->pw_inv[j][i]); }
  __device__ __forceinline__ double mant(int j, int i) const { return __ldg(&t->pw_m[j][i]); }
  __device__ __forceinline__ int expo(int j, int i) const { return __ldg(&t->pw_e[j][i]); }
};

// Scaled genotype likelihoods ev[10] of one person (largest entry in [1,2));
// returns the binary exponent divided out.  All k < kPowTable (checked by caller).
template <typename Pow>
__device__ __forceinline__ int person_linear(const Pow& s, const int* n, double* ev) {
  const int k_hom = max(max(n[0], n[1]), max(n[2], n[3]));
  const int s01 = n[0] + n[1], s02 = n[0] + n[2], s03 = n[0] + n[3], s12 = n[1] + n[2], s13 = n[1] + n[3],
            s23 = n[2] + n[3];
  const int k_het = max(max(max(s01, s02), max(s03, s12)), max(s13, s23));
  const int e1 = s.expo(0, k_hom), e2 = s.expo(1, k_het);
  const int emax = max(e1, e2);
  const double sc1 = scale_pow2(s.mant(0, k_hom), e1 - emax);
  const double sc2 = scale_pow2(s.mant(1, k_het), e2 - emax);
  ev[0] = sc1 * s.inv(0, k_hom - n[0]);
  ev[1] = sc2 * s.inv(1, k_het - s01);
  ev[2] = sc2 * s.inv(1, k_het - s02);
  ev[3] = sc2 * s.inv(1, k_het - s03);
  ev[4] = sc1 * s.inv(0, k_hom - n[1]);
  ev[5] = sc2 * s.inv(1, k_het - s12);
  ev[6] = sc2 * s.inv(1, k_het - s13);
  ev[7] = sc1 * s.inv(0, k_hom - n[2]);
  ev[8] = sc2 * s.inv(1, k_het - s23);
  ev[9] = sc1 * s.inv(0, k_hom - n[3]);
  return emax;
}

// 0 normal, 1 reference value is exactly 0, 2 gray zone (needs the exact path)
__device__ __forceinline__ int magnitude_zone(double v, int s2_floor) {
  if (v >= 0x1p-900) {
    const int ex = ((__double2hiint(v) >> 20) & 0x7FF) - 1023;
    const int lg = s2_floor + ex;          // log2(reference value) in [lg, lg + 2)
    if (lg >= -995) return 0;              // > 2^-995: subnormal terms are below 2^-70 relative
    if (lg + 2 < -1077) return 1;          // every term < 2^-1075: exp() returns +0
    return 2;
  }
  // the scaled sum itself (nearly) underflowed: true value < 2^-899
  return (s2_floor - 899 + 1 < -1077) ? 1 : 2;
}

struct LinOut {
  double den_s, men_s, s2;
  bool dn, hard;
  int zd, zm;
};

// the factored linear-domain evaluation of one site (BAYES / LRT decision)
template <typename Pow>
__device__ __forceinline__ void linear_site(const Pow& s, const int* c, double k_men, double kd9,
                                            double kd8, double kd7, double kd6, double thr, double mm2,
                                            int method, double lrt_threshold, LinOut& o) {
  const int nd = (c[0] + c[1]) + (c[2] + c[3]), nm = (c[4] + c[5]) + (c[6] + c[7]),
            nc = (c[8] + c[9]) + (c[10] + c[11]);
  if (max(nd, max(nm, nc)) >= kPowTable) {  // deeper than the tables: exact path decides
    o.hard = true; o.dn = false; o.zd = o.zm = 2; o.den_s = o.men_s = o.s2 = 0.0;
    return;
  }
  double a[10], b[10], e[10];
  const int ed = person_linear(s, c + 0, a);
  const int em = person_linear(s, c + 4, b);
  const int ec = person_linear(s, c + 8, e);
  const double s2 = fma((double)(nd + nm + nc), mm2, (double)(ed + em + ec));
  const int s2_floor = (int)floor(s2);
  double men_s, den_s;
  {
#include "trio_sparse_gen.cuh"
  }
  const int zd = magnitude_zone(den_s, s2_floor), zm = magnitude_zone(men_s, s2_floor);
  bool dn = false, hard = false;
  if (zd == 2 || zm == 2) {
    hard = true;
  } else if (zd == 0 && zm == 0) {
    const double rhs = thr * men_s;
    dn = den_s > rhs;
    hard = fabs(den_s - rhs) <= 1e-9 * fabs(rhs);
  } else {
    // reference special values: l_den == 0 and/or l_men == 0
    const bool den0 = zd == 1, men0 = zm == 1;
    if (method == DNV_METHOD_LRT) {
      const double lr = den0 ? (men0 ? NAN : 0.0) : INFINITY;
      dn = lr > lrt_threshold;
    } else {
      dn = !den0 && men0;  // p = 1 > 0.5 ; p = 0 or NaN -> false
    }
  }
  o.den_s = den_s; o.men_s = men_s; o.s2 = s2; o.dn = dn; o.hard = hard; o.zd = zd; o.zm = zm;
}

// p, LR and the two log-likelihoods from the factored form (reference special values where a
// raw sum is exactly zero); only valid when !lo.hard
__device__ __forceinline__ void linear_values(const LinOut& lo, double& p, double& lr, double& log_men,
                                              double& log_den) {
  const double ln2 = 0.693147180559945309417;
  const bool den0 = lo.zd == 1, men0 = lo.zm == 1;
  if (!den0 && !men0) {
    p = lo.den_s / (lo.den_s + lo.men_s);
    lr = lo.den_s / lo.men_s;
  } else {
    p = den0 ? (men0 ? NAN : 0.0) : 1.0;
    lr = den0 ? (men0 ? NAN : 0.0) : INFINITY;
  }
  log_den = den0 ? -INFINITY : fma(lo.s2, ln2, log(lo.den_s));
  log_men = men0 ? -INFINITY : fma(lo.s2, ln2, log(lo.men_s));
}

__device__ __forceinline__ void unpack_counts8(const uint32_t* p, int* c) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint32_t v = p[i];
#pragma unroll
    for (int k = 0; k < 4; k++) c[4 * i + k] = (int)((v >> (8 * k)) & 0xFF);
  }
}

__device__ __forceinline__ void unpack_counts(const uint2* p, int* c) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint2 v = p[i];
    c[4 * i + 0] = (int)(v.x & 0xFFFF);
    c[4 * i + 1] = (int)(v.x >> 16);
    c[4 * i + 2] = (int)(v.y & 0xFFFF);
    c[4 * i + 3] = (int)(v.y >> 16);
  }
}

// kLin: evaluate the linear-domain posterior (BAYES/LRT decision, dense values)
// kMap: evaluate the exact log-domain argmax (MAP decision, dense maxGenoType)
// Each WARP walks its own tiles of 32 sites (768 B of counts per tile, TMA bulk copy into a
// warp-private double buffer, warp-private mbarriers): no block-wide barrier in the loop.
template <bool kLin, bool kMap, bool kDense, int kMinBlocks, bool kU8>
__global__ void __launch_bounds__(kFastThreads, kMinBlocks)
trio_fast_kernel(const Tables* __restrict__ tab, const FastConsts k, const unsigned char* __restrict__ counts, int64_t n,
                 uint32_t* __restrict__ bits, unsigned long long* __restrict__ acc,
                 dnv_site_result* __restrict__ dense) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastShared& s = *reinterpret_cast<FastShared*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kBps = kU8 ? 12 : 24;            // bytes per site
  constexpr int kTile = 32 * kBps;               // one warp tile (a multiple of 16)

  const int64_t n_tiles = (n + 31) / 32;
  const bool tma_ok = (reinterpret_cast<uintptr_t>(counts) & 15) == 0;
  if (lane == 0) {
    mbar_init(&s.bar[warp][0], 1);
    mbar_init(&s.bar[warp][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kLin) {
    for (int i = threadIdx.x; i < 2 * kPowTable; i += blockDim.x) {
      (&s.pw_inv[0][0])[i] = (&tab->pw_inv[0][0])[i];
      (&s.pw_m[0][0])[i] = (&tab->pw_m[0][0])[i];
      (&s.pw_e[0][0])[i] = (&tab->pw_e[0][0])[i];
    }
  }
  stage_exact_tables(tab, s.rel, s.isden);
  if (kMap) stage_map_bounds(tab, s.rmax);
  __syncthreads();

  // a full tile (32 sites, 768 B: a multiple of 16) is bulk-copied; the ragged last one is not
  auto issue = [&](int64_t t, int buf) {
    if (lane == 0 && tma_ok && (t + 1) * 32 <= n) {
      mbar_expect_tx(&s.bar[warp][buf], kTile);
      tma_bulk_g2s(s.tile[warp][buf], counts + t * kTile, kTile, &s.bar[warp][buf]);
    }
  };
  const int64_t warps_total = (int64_t)gridDim.x * kFastWarps;
  uint32_t phase0 = 0, phase1 = 0;
  int buf = 0;
  int64_t t = (int64_t)blockIdx.x * kFastWarps + warp;
  if (t < n_tiles) issue(t, 0);
  for (; t < n_tiles; t += warps_total, buf ^= 1) {
    const int64_t site = t * 32 + lane;
    int c[12];
    if (tma_ok && (t + 1) * 32 <= n) {
      if (buf == 0) { mbar_wait(&s.bar[warp][0], phase0); phase0 ^= 1; }
      else { mbar_wait(&s.bar[warp][1], phase1); phase1 ^= 1; }
      if (kU8) unpack_counts8(reinterpret_cast<const uint32_t*>(s.tile[warp][buf] + lane * 12), c);
      else unpack_counts(reinterpret_cast<const uint2*>(s.tile[warp][buf] + lane * 24), c);
    } else if (site < n) {
      if (kU8) unpack_counts8(reinterpret_cast<const uint32_t*>(counts + site * 12), c);
      else unpack_counts(reinterpret_cast<const uint2*>(counts + site * 24), c);
    } else {
#pragma unroll
      for (int q = 0; q < 12; q++) c[q] = 0;
    }
    __syncwarp();  // the whole warp has read the buffer that the next-but-one tile will overwrite
    if (t + warps_total < n_tiles) issue(t + warps_total, buf ^ 1);

    bool dn = false, hard = false;
    dnv_site_result o;
    if (site < n) {
      if (kLin) {
        LinOut lo;
        if (k.lin_ok) {
          linear_site(SmemPow{s}, c, k.k_men, k.kd9, k.kd8, k.kd7, k.kd6, k.thr, k.mm2, k.method, k.lrt_threshold, lo);
        } else {  // eps so large that r <= 1: the table scheme does not apply
          lo.hard = true; lo.dn = false; lo.zd = lo.zm = 2; lo.den_s = lo.men_s = lo.s2 = 0.0;
        }
        dn = lo.dn;
        hard = lo.hard;
        if (kDense && !hard) linear_values(lo, o.p, o.lr, o.log_men, o.log_den);
      }
      if (kMap) {
        // exact log-domain argmax (pruned; bit-identical to the sweep of M/DenovoBayesNet.java:260-278)
        uint16_t cu[12];
#pragma unroll
        for (int q = 0; q < 12; q++) cu[q] = (uint16_t)c[q];
        double Ld[10], Lm[10], Lc[10];
        exact_read_loglik(cu + 0, k.hm, k.ht, k.mm, Ld);
        exact_read_loglik(cu + 4, k.hm, k.ht, k.mm, Lm);
        exact_read_loglik(cu + 8, k.hm, k.ht, k.mm, Lc);
        double max_ll;
        int best;
        exact_map_pruned(Ld, Lm, Lc, s.rel, s.rmax, &best, &max_ll);
        if (k.method == DNV_METHOD_MAP) dn = (s.isden[best >> 5] >> (best & 31)) & 1u;
        if (kDense) {
          o.max_ll = max_ll;
          o.dad = (uint8_t)(best / 100); o.mom = (uint8_t)((best / 10) % 10); o.child = (uint8_t)(best % 10);
        }
      }
      if (kDense && !hard) {
        o.is_denovo = dn;
        o.flags = 0;
        dense[site] = o;
      }
    }
    // hard sites (threshold-ambiguous, underflow gray zone, deeper than the tables): the whole
    // warp evaluates them one at a time with the exact direct form
    uint32_t hword = __ballot_sync(0xffffffffu, hard);
    if (hword) {
      if (lane == 0) atomicAdd(&acc[2], (unsigned long long)__popc(hword));
      while (hword) {
        const int h = __ffs(hword) - 1;
        hword &= hword - 1;
        uint16_t cu[12];
#pragma unroll
        for (int q = 0; q < 12; q++) cu[q] = (uint16_t)__shfl_sync(0xffffffffu, c[q], h);
        ExactOut r;
        exact_site_warp(cu, s.rel, s.isden, s.warp_L[warp], k.hm, k.ht, k.mm, k.method, k.lrt_threshold, &r);
        if (lane == h) {
          dn = r.is_denovo;
          if (kDense) store_dense(dense, site, r, 1u);
        }
      }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, dn);
    if (lane == 0) bits[t] = word;
  }
}

template <bool kLin, bool kMap, bool kDense, int kMinBlocks, bool kU8>
static int launch_fast_kernel(dnv_ctx* ctx, cudaStream_t st, const void* d_counts, int64_t n, dnv_site_result* d_dense,
                              unsigned long long* acc) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const size_t smem = sizeof(FastShared);
  auto kern = trio_fast_kernel<kLin, kMap, kDense, kMinBlocks, kU8>;
  DNV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  DNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFastThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t want = (n + kFastThreads - 1) / kFastThreads;
  const int64_t persistent = (int64_t)sms * per_sm;  // one resident wave; every warp walks its tiles
  const int grid = (int)(want < persistent ? want : persistent);
  const FastConsts k = make_fast_consts(ctx->h_tab);
  kern<<<grid, kFastThreads, smem, st>>>(ctx->d_tab, k, (const unsigned char*)d_counts, n, ctx->d_bits, acc, d_dense);
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

template <bool kLin, bool kMap, bool kDense, int kMinBlocks>
static int launch_fast_variant(dnv_ctx* ctx, cudaStream_t st, const void* d_counts, int64_t n, dnv_site_result* d_dense,
                               unsigned long long* acc) {
  if (ctx->count_bytes == 12) return launch_fast_kernel<kLin, kMap, kDense, kMinBlocks, true>(ctx, st, d_counts, n, d_dense, acc);
  return launch_fast_kernel<kLin, kMap, kDense, kMinBlocks, false>(ctx, st, d_counts, n, d_dense, acc);
}

inline int launch_trio_fast(dnv_ctx* ctx, cudaStream_t st, const void* d_counts, int64_t n, dnv_site_result* d_dense,
                            unsigned long long* acc) {
  const bool map = ctx->prm.method == DNV_METHOD_MAP;
  if (d_dense) return launch_fast_variant<true, true, true, 2>(ctx, st, d_counts, n, d_dense, acc);
  if (map) return launch_fast_variant<false, true, false, 4>(ctx, st, d_counts, n, nullptr, acc);
  return launch_fast_variant<true, false, false, DNV_FAST_MINBLOCKS>(ctx, st, d_counts, n, nullptr, acc);
}

}  // namespace dnv
