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

// Warp-cooperative exact evaluation of ONE site (all 32 lanes must call it with the same
// counts).  Lanes split the 1000 genotype triples (idx = lane + 32*i); every ll is the
// reference's exact add chain, so maxGenoType / max_ll are bit-exact (ties -> lowest
// dad-major index, i.e. the reference's strict '>' sweep); the two exp() sums are lane
// partials combined in a fixed shuffle order (within ~1e-15 relative of the reference's
// sequential sum — the same order as the last-bit differences between exp() implementations).
// s_L: 32 doubles of warp-private shared memory.  The result is valid on every lane.
__device__ __forceinline__ void exact_site_warp(const uint16_t* cnt, const double* s_rel, const uint32_t* s_isden,
                                                double* s_L, double hm, double ht, double mm, int method,
                                                double lrt_threshold, ExactOut* out) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 30) {  // lane = person*10 + genotype: one exact read log-likelihood each
    const int p = lane / 10, g = lane % 10;
    const int a0 = kGenoA0[g], a1 = kGenoA1[g];
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double bl = (a == a0 || a == a1) ? (a0 == a1 ? hm : ht) : mm;
      acc = __dadd_rn(acc, __dmul_rn((double)(int)cnt[p * 4 + a], bl));
    }
    s_L[lane] = acc;
  }
  __syncwarp();
  double max_ll = -INFINITY, l_den = 0.0, l_men = 0.0;
  int best = 0x7fffffff;
  for (int idx = lane; idx < 1000; idx += 32) {
    const int d = idx / 100, m = (idx / 10) % 10, ch = idx % 10;
    const double ll = __dadd_rn(__dadd_rn(__dadd_rn(s_L[d], s_L[10 + m]), s_L[20 + ch]), s_rel[idx]);
    const double e = exp(ll);
    if ((s_isden[idx >> 5] >> (idx & 31)) & 1u) l_den += e; else l_men += e;
    if (ll > max_ll) { max_ll = ll; best = idx; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oll = __shfl_xor_sync(0xffffffffu, max_ll, o);
    const int ob = __shfl_xor_sync(0xffffffffu, best, o);
    if (oll > max_ll || (oll == max_ll && ob < best)) { max_ll = oll; best = ob; }
    l_den += __shfl_xor_sync(0xffffffffu, l_den, o);
    l_men += __shfl_xor_sync(0xffffffffu, l_men, o);
  }
  out->max_ll = max_ll;
  out->dad = best / 100; out->mom = (best / 10) % 10; out->child = best % 10;
  out->p = l_den / (l_den + l_men);
  out->lr = l_den / l_men;
  out->log_men = log(l_men);
  out->log_den = log(l_den);
  bool dn;
  if (method == DNV_METHOD_MAP) dn = (s_isden[best >> 5] >> (best & 31)) & 1u;
  else if (method == DNV_METHOD_BAYES) dn = out->p > 0.5;
  else dn = out->lr > lrt_threshold;
  out->is_denovo = dn;
}

// Exact MAP with pruning (one thread, no exp): maxGenoType and max_ll of one site, bit-identical
// to the reference's 1000-triple sweep (M/DenovoBayesNet.java:260-278) including its tie-break
// (strict '>' in dad-major order = lowest flat index among the maxima).
//   ll(d,m,c) = fl(fl(fl(Ld[d] + Lm[m]) + Lc[c]) + R[d,m,c]).
// Rounded addition is monotone in each operand, so with LmX = max Lm, LcX = max Lc,
//   ll(d,m,c) <= fl(fl(fl(Ld[d] + Lm[m]) + LcX) + max_c R[d,m,.])  =: U_dm
//   U_dm      <= fl(fl(fl(Ld[d] + LmX) + LcX) + max_{m,c} R[d,.,.])  =: U_d
// A (d) row or (d,m) pair is skipped only when its bound is STRICTLY below the best value seen
// so far, so every triple that attains the maximum is evaluated; ties resolve by explicit index.
// s_rel: 1000 doubles, s_rmax: 100 (per d,m) + 10 (per d) doubles, in shared memory.
__device__ __forceinline__ void exact_map_pruned(const double* Ld, const double* Lm, const double* Lc,
                                                 const double* s_rel, const double* s_rmax, int* best_out,
                                                 double* max_ll_out) {
  int d0 = 0, m0 = 0;
  double LmX = Lm[0], LcX = Lc[0], LdX = Ld[0];
#pragma unroll
  for (int g = 1; g < 10; g++) {
    if (Ld[g] > LdX) { LdX = Ld[g]; d0 = g; }
    if (Lm[g] > LmX) { LmX = Lm[g]; m0 = g; }
    LcX = fmax(LcX, Lc[g]);
  }
  // seed: the ten children of the most likely (dad, mom) pair
  double best_v = -INFINITY;
  int best = 0x7fffffff;
  {
    const double dm = __dadd_rn(LdX, LmX);
    const int b0 = d0 * 100 + m0 * 10;
#pragma unroll
    for (int ch = 0; ch < 10; ch++) {
      const double ll = __dadd_rn(__dadd_rn(dm, Lc[ch]), s_rel[b0 + ch]);
      if (ll > best_v) { best_v = ll; best = b0 + ch; }
    }
  }
  // all ten row bounds at once (independent chains), then only the surviving rows
  uint32_t mask_d = 0;
#pragma unroll
  for (int d = 0; d < 10; d++) {
    const double ud = __dadd_rn(__dadd_rn(__dadd_rn(Ld[d], LmX), LcX), s_rmax[100 + d]);
    mask_d |= (ud >= best_v ? 1u : 0u) << d;
  }
  while (mask_d) {
    const int d = __ffs(mask_d) - 1;
    mask_d &= mask_d - 1;
    double ld = Ld[0];  // select chain: keeps Ld[] in registers (no dynamic indexing)
#pragma unroll
    for (int g = 1; g < 10; g++) ld = (d == g) ? Ld[g] : ld;
    uint32_t mask_m = 0;
#pragma unroll
    for (int m = 0; m < 10; m++) {
      const double udm = __dadd_rn(__dadd_rn(__dadd_rn(ld, Lm[m]), LcX), s_rmax[d * 10 + m]);
      mask_m |= (udm >= best_v ? 1u : 0u) << m;
    }
    while (mask_m) {
      const int m = __ffs(mask_m) - 1;
      mask_m &= mask_m - 1;
      double lm = Lm[0];
#pragma unroll
      for (int g = 1; g < 10; g++) lm = (m == g) ? Lm[g] : lm;
      const double dm = __dadd_rn(ld, lm);
      const int b0 = d * 100 + m * 10;
#pragma unroll
      for (int ch = 0; ch < 10; ch++) {
        const double ll = __dadd_rn(__dadd_rn(dm, Lc[ch]), s_rel[b0 + ch]);
        if (ll > best_v || (ll == best_v && b0 + ch < best)) { best_v = ll; best = b0 + ch; }
      }
    }
  }
  *best_out = best;
  *max_ll_out = best_v;
}

// cooperative copy of the exact-path tables into shared memory
__device__ __forceinline__ void stage_exact_tables(const Tables* __restrict__ tab, double* s_rel,
                                                   uint32_t* s_isden) {
  for (int i = threadIdx.x; i < 1000; i += blockDim.x) s_rel[i] = tab->rel[i];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) s_isden[i] = tab->isden[i];
}
// pruning bounds of exact_map_pruned: 100 per-(d,m) maxima followed by 10 per-d maxima
__device__ __forceinline__ void stage_map_bounds(const Tables* __restrict__ tab, double* s_rmax) {
  for (int i = threadIdx.x; i < 100; i += blockDim.x) s_rmax[i] = tab->rmax_dm[i];
  for (int i = threadIdx.x; i < 10; i += blockDim.x) s_rmax[100 + i] = tab->rmax_d[i];
}

// counts of one site: bps = 24 (uint16 x 12) or 12 (uint8 x 12)
__device__ __forceinline__ void load_counts(const void* __restrict__ counts, int bps, int64_t site, uint16_t* c) {
  if (bps == 24) {
    const uint2* p = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(counts) + site * 24);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const uint2 v = __ldg(p + i);
      c[4 * i + 0] = (uint16_t)(v.x & 0xFFFF);
      c[4 * i + 1] = (uint16_t)(v.x >> 16);
      c[4 * i + 2] = (uint16_t)(v.y & 0xFFFF);
      c[4 * i + 3] = (uint16_t)(v.y >> 16);
    }
  } else {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(counts) + site * 12);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const uint32_t v = __ldg(p + i);
#pragma unroll
      for (int k = 0; k < 4; k++) c[4 * i + k] = (uint16_t)((v >> (8 * k)) & 0xFF);
    }
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
