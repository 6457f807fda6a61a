// trio_exact.cuh — exact direct-form evaluation of one trio site on the device.
//
// This is the reference's performInference (M/DenovoBayesNet.java:247-289) in
// its own operation order: per-genotype read log-likelihoods in A,C,G,T order
// (:193-211), ll = ((Ld + Lm) + Lc) + R for all 1000 triples dad-major (:260-267),
// raw exp() sums with NO max shift (:269-273), strict '>' argmax (:275-278).
// Every add / multiply is an explicit round-to-nearest intrinsic so ptxas can
// not contract them into FMAs (Java never fuses).  It is the decision path for
// "hard" sites (threshold-ambiguous or in the underflow gray zone), the source
// of every value in a dnv_call record, and the DNV_F_FORCE_DIRECT kernel.
#pragma once
#include "dnv_internal.cuh"

namespace dnv {

struct ExactOut {
  double max_ll, p, lr, log_men, log_den;
  int dad, mom, child;
  bool is_denovo;
};

static __device__ __constant__ int8_t kGenoA0[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
static __device__ __constant__ int8_t kGenoA1[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};

// getReadSummaryLogLikelihood (M/DenovoBayesNet.java:193-211): absent alleles
// are skipped by the reference; adding n*bl with n = 0 adds a signed zero,
// which is the identity here (acc starts at +0.0), so the loop is branch free.
__device__ __forceinline__ void exact_read_loglik(const uint16_t* n, double hm, double ht,
                                                  double mm, double* L) {
#pragma unroll
  for (int g = 0; g < 10; g++) {
    const int a0 = kGenoA0[g], a1 = kGenoA1[g];
    const bool hom = a0 == a1;
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double bl = (a == a0 || a == a1) ? (hom ? hm : ht) : mm;
      acc = __dadd_rn(acc, __dmul_rn((double)(int)n[a], bl));
    }
    L[g] = acc;
  }
}

// s_rel: 1000 doubles, s_isden: 32 words (shared memory).
__device__ __noinline__ void exact_site(const uint16_t* cnt, const double* s_rel,
                                        const uint32_t* s_isden, double hm, double ht, double mm,
                                        int method, double lrt_threshold, ExactOut* out) {
  double Ld[10], Lm[10], Lc[10];
  exact_read_loglik(cnt + 0, hm, ht, mm, Ld);
  exact_read_loglik(cnt + 4, hm, ht, mm, Lm);
  exact_read_loglik(cnt + 8, hm, ht, mm, Lc);

  double max_ll = -INFINITY, l_den = 0.0, l_men = 0.0;
  int best = 0;
#pragma unroll 1
  for (int d = 0; d < 10; d++) {
#pragma unroll 1
    for (int m = 0; m < 10; m++) {
      const double dm = __dadd_rn(Ld[d], Lm[m]);
      const int base = d * 100 + m * 10;
#pragma unroll
      for (int c = 0; c < 10; c++) {
        const int idx = base + c;
        const double ll = __dadd_rn(__dadd_rn(dm, Lc[c]), s_rel[idx]);
        const double e = exp(ll);
        if ((s_isden[idx >> 5] >> (idx & 31)) & 1u) l_den = __dadd_rn(l_den, e);
        else l_men = __dadd_rn(l_men, e);
        if (ll > max_ll) { max_ll = ll; best = idx; }
      }
    }
  }
  out->max_ll = max_ll;
  out->dad = best / 100;
  out->mom = (best / 10) % 10;
  out->child = best % 10;
  out->p = __ddiv_rn(l_den, __dadd_rn(l_den, l_men));  // :282
  out->lr = __ddiv_rn(l_den, l_men);                   // :285
  out->log_men = log(l_men);
  out->log_den = log(l_den);
  // InferenceMethod.isDenovo, M/DenovoUtil.java:261-288 (NaN compares false)
  bool dn;
  if (method == DNV_METHOD_MAP) dn = (s_isden[best >> 5] >> (best & 31)) & 1u;
  else if (method == DNV_METHOD_BAYES) dn = out->p > 0.5;
  else dn = out->lr > lrt_threshold;
  out->is_denovo = dn;
}

// cooperative copy of the exact-path tables into shared memory
__device__ __forceinline__ void stage_exact_tables(const Tables* __restrict__ tab, double* s_rel,
                                                   uint32_t* s_isden) {
  for (int i = threadIdx.x; i < 1000; i += blockDim.x) s_rel[i] = tab->rel[i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) s_isden[i] = tab->isden[i];
}

__device__ __forceinline__ void load_counts(const uint16_t* __restrict__ counts, int64_t site,
                                            uint16_t* c) {
  const uint2* p = reinterpret_cast<const uint2*>(counts + site * 12);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const uint2 v = __ldg(p + i);
    c[4 * i + 0] = (uint16_t)(v.x & 0xFFFF);
    c[4 * i + 1] = (uint16_t)(v.x >> 16);
    c[4 * i + 2] = (uint16_t)(v.y & 0xFFFF);
    c[4 * i + 3] = (uint16_t)(v.y >> 16);
  }
}

__device__ __forceinline__ void store_dense(dnv_site_result* dense, int64_t site, const ExactOut& r,
                                            uint32_t flags) {
  dnv_site_result o;
  o.p = r.p; o.lr = r.lr; o.log_men = r.log_men; o.log_den = r.log_den; o.max_ll = r.max_ll;
  o.dad = (uint8_t)r.dad; o.mom = (uint8_t)r.mom; o.child = (uint8_t)r.child;
  o.is_denovo = r.is_denovo; o.flags = flags;
  dense[site] = o;
}

}  // namespace dnv
