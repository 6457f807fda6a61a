// trio_fast.cuh — fused likelihood + trio-posterior kernel (kernel (b) of the
// north star), hand-written for sm_100a.  One thread owns one site.
//
// BAYES / LRT decisions use the factored LINEAR-domain form:
//   exp(L_p[g]) = 2^(N_p*log2(eps/3)) * r_hom^k  (g = XX, k = n_X)
//               = 2^(N_p*log2(eps/3)) * r_het^k  (g = XY, k = n_X + n_Y)
// with r^k = mantissa * 2^exponent read from two small shared-memory tables
// (counts are integers, so no exp() is evaluated at all), each person is
// rescaled by a power of two (exact), and the 10x10x10 sum collapses to the
// sparse straight-line code of trio_sparse_gen.cuh (335 FP64 instructions
// instead of 2000 dense FMAs + 30 exp).  The unshifted magnitudes
// log2(Lden), log2(Lmen) are tracked so the reference's raw-exp underflow
// semantics are reproduced: a site whose decision could depend
// on rounding (|den - T*men| <= 1e-9 relative) or on subnormal / flushed terms
// (gray zone) is NOT decided here — it is appended to the hard list and
// re-evaluated by the exact direct-form kernel (trio_exact.cuh).
//
// MAP decisions (and the maxGenoType of the dense record) use the exact
// log-domain sweep: bit-identical lls, strict '>' in dad-major order.
#pragma once
#include "dnv_internal.cuh"
#include "trio_exact.cuh"

namespace dnv {

constexpr int kFastThreads = 128;

struct FastShared {
  double pw_m[2][kPowTable];
  int32_t pw_e[2][kPowTable];
};

// r_j^k as (mantissa, exponent)
__device__ __forceinline__ void pow_lookup(const FastShared& s, const Tables* __restrict__ tab, int j,
                                           int k, double& m, int& e) {
  if (k < kPowTable) {
    m = s.pw_m[j][k];
    e = s.pw_e[j][k];
  } else {  // deep pileup: k*log2(r) in hi/lo parts, k*hi is exact
    const double kd = (double)k;
    const double t_hi = kd * tab->c_hi[j];
    const double t_lo = kd * tab->c_lo[j];
    const double fl = floor(t_hi + t_lo);
    m = exp2((t_hi - fl) + t_lo);
    e = (int)fl;
  }
}

// per-person scaled genotype likelihoods ev[10] (max in [1,4)), returns the
// binary exponent that was divided out.
__device__ __forceinline__ int person_linear(const FastShared& s, const Tables* __restrict__ tab,
                                             const uint16_t* n, double* ev) {
  double m[10];
  int e[10];
#pragma unroll
  for (int g = 0; g < 10; g++) {
    const int a0 = kGenoA0[g], a1 = kGenoA1[g];
    const bool hom = a0 == a1;
    const int k = hom ? (int)n[a0] : (int)n[a0] + (int)n[a1];
    pow_lookup(s, tab, hom ? 0 : 1, k, m[g], e[g]);
  }
  int emax = e[0];
#pragma unroll
  for (int g = 1; g < 10; g++) emax = max(emax, e[g]);
#pragma unroll
  for (int g = 0; g < 10; g++) {
    const int sh = e[g] - emax;  // <= 0
    const int hi = __double2hiint(m[g]) + (sh << 20);
    ev[g] = (sh < -1000) ? 0.0 : __hiloint2double(hi, __double2loint(m[g]));
  }
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

template <bool kDense>
__global__ void __launch_bounds__(kFastThreads)
trio_fast_kernel(const Tables* __restrict__ tab, const uint16_t* __restrict__ counts, int64_t n,
                 uint32_t* __restrict__ bits, uint32_t* __restrict__ hard_list,
                 unsigned long long* __restrict__ counters, dnv_site_result* __restrict__ dense) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastShared& s = *reinterpret_cast<FastShared*>(smem_raw);
  double* s_rel = reinterpret_cast<double*>(smem_raw + sizeof(FastShared));  // only when MAP/dense
  uint32_t* s_isden = reinterpret_cast<uint32_t*>(s_rel + 1000);

  const int method = tab->method;
  const bool need_map = kDense || method == DNV_METHOD_MAP;
  const bool need_lin = kDense || method != DNV_METHOD_MAP;
  if (need_lin) {
    for (int i = threadIdx.x; i < 2 * kPowTable; i += blockDim.x) {
      (&s.pw_m[0][0])[i] = (&tab->pw_m[0][0])[i];
      (&s.pw_e[0][0])[i] = (&tab->pw_e[0][0])[i];
    }
  }
  if (need_map) stage_exact_tables(tab, s_rel, s_isden);
  __syncthreads();

  const double k_men = tab->k_men;
  const double kd9 = tab->k_den[0], kd8 = 2.0 * tab->k_den[1], kd7 = 4.0 * tab->k_den[2],
               kd6 = 4.0 * tab->k_den[3];
  const double thr = tab->lrt_thr, mm2 = tab->mm2;
  const double hm = tab->hm, ht = tab->ht, mm = tab->mm;

  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t site = base + threadIdx.x;
    bool dn = false, hard = false;
    if (site < n) {
      uint16_t c[12];
      load_counts(counts, site, c);
      dnv_site_result o;
      if (need_lin) {
        double a[10], b[10], e[10];
        const int ed = person_linear(s, tab, c + 0, a);
        const int em = person_linear(s, tab, c + 4, b);
        const int ec = person_linear(s, tab, c + 8, e);
        int ntot = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) ntot += c[k];
        const double s2 = fma((double)ntot, mm2, (double)(ed + em + ec));
        const int s2_floor = (int)floor(fmax(s2, -2.0e9));
        double men_s, den_s;
        {
#include "trio_sparse_gen.cuh"
        }
        const int zd = magnitude_zone(den_s, s2_floor), zm = magnitude_zone(men_s, s2_floor);
        if (zd == 2 || zm == 2) {
          hard = true;
        } else if (zd == 0 && zm == 0) {
          const double rhs = thr * men_s;
          dn = den_s > rhs;
          hard = fabs(den_s - rhs) <= 1e-9 * fabs(rhs);
          if (kDense) {
            o.p = den_s / (den_s + men_s);
            o.lr = den_s / men_s;
            const double ln2 = 0.693147180559945309417;
            o.log_den = fma(s2, ln2, log(den_s));
            o.log_men = fma(s2, ln2, log(men_s));
          }
        } else {
          // reference special values: l_den == 0 and/or l_men == 0
          const bool den0 = zd == 1, men0 = zm == 1;
          if (method == DNV_METHOD_LRT) {
            const double lr = den0 ? (men0 ? NAN : 0.0) : INFINITY;
            dn = lr > tab->lrt_threshold;
          } else {
            dn = !den0 && men0;  // p = 1 > 0.5 ; p = 0 or NaN -> false
          }
          if (kDense) {
            const double ln2 = 0.693147180559945309417;
            o.p = den0 ? (men0 ? NAN : 0.0) : 1.0;
            o.lr = den0 ? (men0 ? NAN : 0.0) : INFINITY;
            o.log_den = den0 ? -INFINITY : fma(s2, ln2, log(den_s));
            o.log_men = men0 ? -INFINITY : fma(s2, ln2, log(men_s));
          }
        }
      }
      if (need_map) {
        // exact log-domain sweep, M/DenovoBayesNet.java:260-278
        double Ld[10], Lm[10], Lc[10];
        exact_read_loglik(c + 0, hm, ht, mm, Ld);
        exact_read_loglik(c + 4, hm, ht, mm, Lm);
        exact_read_loglik(c + 8, hm, ht, mm, Lc);
        double max_ll = -INFINITY;
        int best = 0;
#pragma unroll 1
        for (int d = 0; d < 10; d++) {
#pragma unroll 1
          for (int m = 0; m < 10; m++) {
            const double dm = __dadd_rn(Ld[d], Lm[m]);
            const int b0 = d * 100 + m * 10;
#pragma unroll
            for (int ch = 0; ch < 10; ch++) {
              const double ll = __dadd_rn(__dadd_rn(dm, Lc[ch]), s_rel[b0 + ch]);
              if (ll > max_ll) { max_ll = ll; best = b0 + ch; }
            }
          }
        }
        const bool map_dn = (s_isden[best >> 5] >> (best & 31)) & 1u;
        if (method == DNV_METHOD_MAP) dn = map_dn;
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
    const uint32_t word = __ballot_sync(0xffffffffu, dn && !hard);
    const uint32_t hword = __ballot_sync(0xffffffffu, hard);
    const int lane = threadIdx.x & 31;
    if (lane == 0 && site < n) bits[site >> 5] = word;
    if (hword) {
      unsigned long long slot = 0;
      if (lane == 0) slot = atomicAdd(&counters[1], (unsigned long long)__popc(hword));
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (hard) hard_list[slot + __popc(hword & ((1u << lane) - 1))] = (uint32_t)site;
    }
  }
}

inline int launch_trio_fast(dnv_ctx* ctx, const uint16_t* d_counts, int64_t n, dnv_site_result* d_dense) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const size_t smem = sizeof(FastShared) + 1000 * sizeof(double) + 32 * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set) {
    DNV_CUDA(cudaFuncSetAttribute(trio_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DNV_CUDA(cudaFuncSetAttribute(trio_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t want = (n + kFastThreads - 1) / kFastThreads;
  const int64_t persistent = (int64_t)sms * 4;
  const int grid = (int)(want < persistent ? want : persistent);
  if (d_dense)
    trio_fast_kernel<true><<<grid, kFastThreads, smem, ctx->stream>>>(ctx->d_tab, d_counts, n, ctx->d_bits,
                                                                     ctx->d_hard_list, ctx->d_counters, d_dense);
  else
    trio_fast_kernel<false><<<grid, kFastThreads, smem, ctx->stream>>>(ctx->d_tab, d_counts, n, ctx->d_bits,
                                                                      ctx->d_hard_list, ctx->d_counters, d_dense);
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

}  // namespace dnv
