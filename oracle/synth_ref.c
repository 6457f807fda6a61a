/*
 * synth_ref.c — CPU restatement of the synthetic trio pileup generator
 * (test infrastructure).  The generator replaces the reference's Google
 * Genomics API client (M/ReadCaller.java:155-186 — network, out of scope); it
 * has no counterpart in the reference, so this file is the *specification*
 * that the product's device generator (csrc/synth.cu) must match bit for bit.
 * SURVEY.md Appendix E.
 *
 * RNG: counter based, stateless, shard invariant:
 *   key(stream)   = splitmix64(seed ^ (stream * PHI))
 *   u(stream, i)  = splitmix64(key(stream) ^ i)
 * Streams: 0 site class, 1 ref base, 2 alt base, 3 parental/transmission bits,
 *          4..6 depth of DAD/MOM/CHILD, 8..10 per-read draws of DAD/MOM/CHILD.
 * Site model: class 0 (90 %) trio hom-ref; class 1 (9 %) one parent ref/alt
 * het, child inherits a random allele of each parent; class 2 (1 %) parents
 * hom-ref, child ref/alt het (true de novo).
 * Depth ~ Poisson(D) by inverse CDF over a caller-supplied table of 2^64-scaled
 * thresholds (so CPU and GPU share the same integers).  Each read picks one of
 * the person's two alleles (bit 0) and with probability err_thr/2^32 is
 * replaced by one of the other three bases.
 */
#include <stdint.h>
#include <string.h>

#define PHI 0x9E3779B97F4A7C15ULL

static uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + PHI;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static uint64_t skey(uint64_t seed, uint64_t stream) {
  return splitmix64(seed ^ (stream * PHI));
}

/* counts: [n][3][4] uint16 (DAD, MOM, CHILD × A,C,G,T) for global site indices
 * first .. first+n-1.  thr: n_thr ascending thresholds, depth = #{k: thr[k] <= u}. */
static void synth_ref_impl(uint64_t seed, int64_t first, const int64_t *index, int64_t n,
                           const uint64_t *thr, int32_t n_thr, uint32_t err_thr,
                           uint16_t *counts) {
  uint64_t k_class = skey(seed, 0), k_ref = skey(seed, 1), k_alt = skey(seed, 2),
           k_par = skey(seed, 3);
  uint64_t k_depth[3] = {skey(seed, 4), skey(seed, 5), skey(seed, 6)};
  uint64_t k_read[3] = {skey(seed, 8), skey(seed, 9), skey(seed, 10)};
  for (int64_t s = 0; s < n; s++) {
    uint64_t i = index ? (uint64_t)index[s] : (uint64_t)(first + s);
    uint32_t c100 = (uint32_t)(((splitmix64(k_class ^ i) >> 32) * 100ULL) >> 32);
    int cls = c100 < 90 ? 0 : (c100 < 99 ? 1 : 2);
    int ref = (int)(splitmix64(k_ref ^ i) & 3);
    int alt = (ref + 1 + (int)(splitmix64(k_alt ^ i) % 3)) & 3;
    uint64_t pb = splitmix64(k_par ^ i);
    int g[3][2] = {{ref, ref}, {ref, ref}, {ref, ref}}; /* DAD, MOM, CHILD */
    if (cls == 1) {
      int het_parent = (int)(pb & 1);     /* 0 DAD, 1 MOM */
      g[het_parent][1] = alt;
      int transmit_alt = (int)((pb >> 1) & 1);
      if (transmit_alt) g[2][1] = alt;
    } else if (cls == 2) {
      g[2][1] = alt;
    }
    uint16_t *out = counts + s * 12;
    memset(out, 0, 12 * sizeof(uint16_t));
    for (int p = 0; p < 3; p++) {
      uint64_t ud = splitmix64(k_depth[p] ^ i);
      int depth = 0;
      while (depth < n_thr && thr[depth] <= ud) depth++;
      for (int j = 0; j < depth; j++) {
        uint64_t u = splitmix64(k_read[p] ^ ((i << 16) | (uint64_t)j));
        int b = g[p][u & 1];
        if ((uint32_t)(u >> 32) < err_thr)
          b = (b + 1 + (int)(((u >> 8) & 0xFFFF) % 3)) & 3;
        out[p * 4 + b]++;
      }
    }
  }
}

void synth_ref_counts(uint64_t seed, int64_t first, int64_t n, const uint64_t *thr,
                      int32_t n_thr, uint32_t err_thr, uint16_t *counts) {
  synth_ref_impl(seed, first, 0, n, thr, n_thr, err_thr, counts);
}
/* arbitrary global site indices (READ-mode candidate sites) */
void synth_ref_counts_at(uint64_t seed, const int64_t *index, int64_t n, const uint64_t *thr,
                         int32_t n_thr, uint32_t err_thr, uint16_t *counts) {
  synth_ref_impl(seed, 0, index, n, thr, n_thr, err_thr, counts);
}
