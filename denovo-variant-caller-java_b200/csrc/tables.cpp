// tables.cpp — site-invariant tables of the trio Bayes net (host side, built
// once per ctx; replaces `new DenovoBayesNet(shared)`,
// M/DenovoBayesNet.java:52-63).  Written from the model definition:
//   parent prior  P(g)        = (hom ? 1 : 2) / 16            (:84-85)
//   child CPT     T(d,m,c)    = denovo ? mu / nDen(d,m)
//                                      : ((1 - mu) / 4) * cnt(d,m,c)   (:106-109)
//   relationship  R(d,m,c)    = (log P(d) + log P(m)) + log T(d,m,c)  (:322-330)
//   base log-lik  hm, ht, mm                                          (:159-165)
// Compiled with -ffp-contract=off so R/hm/ht/mm carry the same bits a JVM or
// the CPU oracle computes from the same formulas.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "dnv_internal.cuh"

namespace dnv {

namespace {
// genotype g <-> unordered allele pair (a0 <= a1), ordinal order AA,AC,AG,AT,CC,CG,CT,GG,GT,TT
const int kA0[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
const int kA1[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};

int geno_of(int x, int y) {
  if (x > y) { int t = x; x = y; y = t; }
  for (int g = 0; g < 10; g++)
    if (kA0[g] == x && kA1[g] == y) return g;
  return -1;
}
bool has(int g, int a) { return kA0[g] == a || kA1[g] == a; }

// r^k = m * 2^e with m in [1,2): computed in long double from k*log2(r)
void pow_entry(long double log2r, int k, double* m, int32_t* e) {
  long double t = log2r * (long double)k;
  long double fl = floorl(t);
  long double mant = exp2l(t - fl);
  if (mant >= 2.0L) { mant *= 0.5L; fl += 1.0L; }
  *m = (double)mant;
  if (*m >= 2.0) { *m = 1.0; fl += 1.0L; }
  *e = (int32_t)fl;
}
}  // namespace

void build_tables(const dnv_params& p, Tables* t) {
  std::memset(t, 0, sizeof(*t));
  const double eps = p.seq_err_rate, mu = p.denovo_mut_rate;
  double prior[10];
  for (int g = 0; g < 10; g++) prior[g] = (kA0[g] == kA1[g] ? 1.0 : 2.0) / 16;

  t->hm = std::log(1 - eps);
  t->ht = std::log(1 - 2 * eps / 3) - std::log(2);
  t->mm = std::log(eps) - std::log(3);

  for (int d = 0; d < 10; d++)
    for (int m = 0; m < 10; m++) {
      int cnt[10] = {0};
      const int da[2] = {kA0[d], kA1[d]}, ma[2] = {kA0[m], kA1[m]};
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) cnt[geno_of(da[i], ma[j])]++;
      int keys = 0;
      for (int g = 0; g < 10; g++) keys += cnt[g] > 0;
      const int n_den = 10 - keys;
      t->nden[d * 10 + m] = (uint8_t)n_den;
      for (int c = 0; c < 10; c++) {
        const int c1 = kA0[c], c2 = kA1[c];
        const bool mendel = (has(m, c1) && has(d, c2)) || (has(m, c2) && has(d, c1));
        const int idx = d * 100 + m * 10 + c;
        const double T = mendel ? (1.0 - mu) / 4 * cnt[c] : mu / n_den;
        if (!mendel) t->isden[idx >> 5] |= 1u << (idx & 31);
        double r = 0;
        r += std::log(prior[d]);
        r += std::log(prior[m]);
        r += std::log(T);
        t->rel[idx] = r;
      }
    }

  for (int d = 0; d < 10; d++) {
    t->rmax_d[d] = -INFINITY;
    for (int m = 0; m < 10; m++) {
      double mx = -INFINITY;
      for (int c = 0; c < 10; c++) mx = std::max(mx, t->rel[d * 100 + m * 10 + c]);
      t->rmax_dm[d * 10 + m] = mx;
      t->rmax_d[d] = std::max(t->rmax_d[d], mx);
    }
  }
  // linear-domain path: likelihood ratios against the mismatch likelihood
  const long double l2e = 1.442695040888963407359924681001892137L;
  const long double le = (long double)eps;
  const long double log2r[2] = {
      (logl(1.0L - le) - (logl(le) - logl(3.0L))) * l2e,                       // hom match / mismatch
      ((logl(1.0L - 2.0L * le / 3.0L) - logl(2.0L)) - (logl(le) - logl(3.0L))) * l2e};  // het match / mismatch
  for (int j = 0; j < 2; j++)
    for (int k = 0; k < kPowTable; k++) {
      pow_entry(log2r[j], k, &t->pw_m[j][k], &t->pw_e[j][k]);
      t->pw_inv[j][k] = (double)exp2l(-log2r[j] * (long double)k);
    }
  t->fast_ok = (log2r[0] > 0.0L && log2r[1] > 0.0L) ? 1 : 0;
  // integer screen (trio_screen.cuh): Lden / Lmen <= 5334 mu / (1 - mu) for a Mendelian-consistent
  // most-likely triple; valid when that bound is below the decision threshold and k*log2(r) fits 32 bits.
  // The factor 4 covers the reference's raw-exp sums in the subnormal range, where each of the up to 750
  // de novo terms may round UP by half an ulp (a factor 1.5 would do; see the header of trio_screen.cuh).
  const double thr_eff = p.method == DNV_METHOD_LRT ? p.lrt_threshold : 1.0;
  const bool fits = log2r[0] < 60.0L && log2r[1] < 60.0L;
  t->c_hom_fx = fits ? (uint32_t)llroundl(log2r[0] * 65536.0L) : 0;
  t->c_het_fx = fits ? (uint32_t)llroundl(log2r[1] * 65536.0L) : 0;
  // screen level 1: n_X * a >= (N - n_X) * h  must IMPLY  n_X * log2(r_hom) >= N * log2(r_het), i.e.
  // a / h <= (log2 r_hom - log2 r_het) / log2 r_het; h = 127, a rounded down (signed bytes for dp2a / dp4a)
  {
    int a = 0;
    if (log2r[1] > 0.0L && log2r[0] > log2r[1]) {
      long double q = 127.0L * (log2r[0] - log2r[1]) / log2r[1] * (1.0L - 1e-12L);
      a = q > 127.0L ? 127 : (int)floorl(q);
    }
    for (int x = 0; x < 4; x++) {
      uint32_t w = 0;
      for (int b = 0; b < 4; b++) w |= (uint32_t)(uint8_t)(int8_t)(b == x ? a : -127) << (8 * b);
      t->scr_w[x] = w;
    }
  }
  t->early_ok = (t->fast_ok && fits && log2r[0] > log2r[1] && p.method != DNV_METHOD_MAP && thr_eff > 0.0 && std::isfinite(thr_eff) &&
                 4.0 * 5334.0 * mu / (1.0 - mu) < 0.999 * thr_eff) ? 1 : 0;
  // screen level 3 (trio_screen.cuh: both parents clearly homozygous, child genotype not explained by them):
  //   Lden <= 1.0143 * (mu / 2304) * a* b* e*     and     Lmen >= ((1 - mu) / 256) * (one alternative's likelihoods),
  // so the site is not de novo when log2(likelihood ratio best / alternative) < log2(9 T (1 - mu) / mu) - 0.0205;
  // another 0.05 bits cover the 16.16 fixed point (<= 3 products of depth <= 127 counts, 2^-17 each) and FP64 rounding.
  // Only sites whose smallest relevant term stays a NORMAL double in the reference's raw-exp sums are decided.
  // MAP: the argmax triple is de novo only if some de novo triple beats every Mendelian one; max <= sum, so the same
  // test with T = 1 proves the argmax Mendelian (the log-domain sums of the MAP path round by ~1e-13: inside the margin).
  t->l3_lambda_fx = 0;
  t->l3_max_depth = 0;
  const bool map_ok = t->fast_ok && fits && log2r[0] > log2r[1] && p.method == DNV_METHOD_MAP && mu < 0.01;
  if (t->early_ok || map_ok) {
    const long double lam = log2l(9.0L * (long double)thr_eff * (1.0L - (long double)mu) / (long double)mu) - 0.0705L;
    const long double room = 690.0L - fabsl(logl((long double)mu / 2304.0L));
    const long double per_read = fabsl(logl(le) - logl(3.0L));
    long long nmax = room > 0 ? (long long)floorl(room / per_read) : 0;
    nmax = std::min<long long>(nmax, ((1ll << 31) - (1ll << 21)) / std::max<long long>(1, std::max(t->c_hom_fx, t->c_het_fx)));  // int32 products
    if (lam > 1.0L && lam < 30000.0L && nmax >= 3) {
      t->l3_lambda_fx = (int32_t)floorl(lam * 65536.0L);
      t->l3_max_depth = (int32_t)std::min<long long>(nmax, 4000);
    }
  }
  // MAP through the same screen: the chosen Mendelian triple t* (each genotype within a factor 2 of its
  // person's best) has ll(t*) >= sum Lmax - 3 ln 2 + log((1/16)(1/16)(1-mu)/4); every de novo triple has
  // ll <= sum Lmax + log((1/8)(1/8) mu/6).  The gap is log(0.046875 (1-mu)/mu): with mu < 0.01 it exceeds
  // 1.5 nats, rounding of the four-term sums is ~1e-13, so the argmax is Mendelian: not de novo, exactly.
  t->map_screen_ok = (t->fast_ok && fits && log2r[0] > log2r[1] && p.method == DNV_METHOD_MAP && mu < 0.01) ? 1 : 0;
  t->mm2 = (double)((logl(le) - logl(3.0L)) * l2e);
  t->k_men = (1.0 - mu) / 4 / 64;
  const int nd[4] = {9, 8, 7, 6};
  for (int i = 0; i < 4; i++) t->k_den[i] = mu / 256 / nd[i];
  t->method = p.method;
  t->lrt_threshold = p.lrt_threshold;
  t->lrt_thr = p.method == DNV_METHOD_LRT ? p.lrt_threshold : 1.0;
}

}  // namespace dnv
