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

/* ------------------------------------------------------------------------------------------------
 * Synthetic READS at candidate sites: the specification of the host's
 * read source (host/denovo_host.hpp: SyntheticReadSource), which stands in for the three
 * reads().search() calls per candidate (M/ReadCaller.java:155-186).  For every site i of `index`
 * (global position on the concatenated axis) and person p, the d_p(i) reads of the counts generator
 * above become d_p(i) aligned reads of length L that cover the site:
 *   streams 12..14  offset of the site inside read j:  off = u(12 + p, (i << 16) | j) mod L,
 *                   so the read starts at i - off and its base at the site is read j's base above;
 *   stream 16       every other base is the "reference genome" at its own position q:
 *                   "ACGT"[u(16, q) & 3]  (reads of neighbouring candidates therefore agree with each
 *                   other, and they DO contribute to each other's pileups when the candidates are
 *                   closer than L — as real reads would);
 *   stream 17       gaps: base t of read (i, p, j) is '-' when (u(17 + 32 p, ((i << 16 | j) * 1024 + t)) >> 32)
 *                   < gap_thr (skipped by ReadSummary, M/ReadSummary.java:51-53); the site's own base is
 *                   never a gap.
 * Output in the layout dnv_pileup wants: reads sorted by (person, start, site, j).
 * Two calls: bases == NULL returns the number of reads (every read has L bases). */
typedef struct { int64_t start; int64_t site; int32_t j; uint8_t person; uint8_t base; } sread_t;
static int sread_cmp(const void *a, const void *b) {
  const sread_t *x = (const sread_t *)a, *y = (const sread_t *)b;
  if (x->person != y->person) return x->person < y->person ? -1 : 1;
  if (x->start != y->start) return x->start < y->start ? -1 : 1;
  if (x->site != y->site) return x->site < y->site ? -1 : 1;
  return x->j < y->j ? -1 : (x->j > y->j);
}
#include <stdlib.h>
int64_t synth_ref_reads_at(uint64_t seed, const int64_t *index, int64_t n, const uint64_t *thr,
                           int32_t n_thr, uint32_t err_thr, int32_t L, uint32_t gap_thr,
                           int64_t *read_start, int64_t *base_off, char *bases, uint8_t *person) {
  uint64_t k_class = skey(seed, 0), k_ref = skey(seed, 1), k_alt = skey(seed, 2), k_par = skey(seed, 3);
  uint64_t k_fill = skey(seed, 16);
  int64_t n_reads = 0;
  for (int pass = 0; pass < 2; pass++) {
    sread_t *rd = pass ? (sread_t *)malloc((size_t)(n_reads > 0 ? n_reads : 1) * sizeof(sread_t)) : 0;
    int64_t k = 0;
    for (int64_t s = 0; s < n; s++) {
      uint64_t i = (uint64_t)index[s];
      uint32_t c100 = (uint32_t)(((splitmix64(k_class ^ i) >> 32) * 100ULL) >> 32);
      int cls = c100 < 90 ? 0 : (c100 < 99 ? 1 : 2);
      int ref = (int)(splitmix64(k_ref ^ i) & 3);
      int alt = (ref + 1 + (int)(splitmix64(k_alt ^ i) % 3)) & 3;
      uint64_t pb = splitmix64(k_par ^ i);
      int g[3][2] = {{ref, ref}, {ref, ref}, {ref, ref}};
      if (cls == 1) { g[(int)(pb & 1)][1] = alt; if ((pb >> 1) & 1) g[2][1] = alt; }
      else if (cls == 2) g[2][1] = alt;
      for (int p = 0; p < 3; p++) {
        uint64_t ud = splitmix64(skey(seed, 4 + (uint64_t)p) ^ i);
        int depth = 0;
        while (depth < n_thr && thr[depth] <= ud) depth++;
        if (!pass) { k += depth; continue; }
        for (int j = 0; j < depth; j++) {
          uint64_t rid = (i << 16) | (uint64_t)j;
          uint64_t u = splitmix64(skey(seed, 8 + (uint64_t)p) ^ rid);
          int b = g[p][u & 1];
          if ((uint32_t)(u >> 32) < err_thr) b = (b + 1 + (int)(((u >> 8) & 0xFFFF) % 3)) & 3;
          int64_t off = (int64_t)(splitmix64(skey(seed, 12 + (uint64_t)p) ^ rid) % (uint64_t)L);
          rd[k].start = (int64_t)i - off; rd[k].site = (int64_t)i; rd[k].j = j; rd[k].person = (uint8_t)p; rd[k].base = (uint8_t)b;
          k++;
        }
      }
    }
    if (!pass) { n_reads = k; if (!bases) return n_reads; continue; }
    qsort(rd, (size_t)n_reads, sizeof(sread_t), sread_cmp);
    for (int64_t r = 0; r < n_reads; r++) {
      read_start[r] = rd[r].start; person[r] = rd[r].person; base_off[r] = r * (int64_t)L;
      uint64_t rid = ((uint64_t)rd[r].site << 16) | (uint64_t)rd[r].j;
      uint64_t k_gap = skey(seed, 17 + 32 * (uint64_t)rd[r].person);
      for (int t = 0; t < L; t++) {
        int64_t q = rd[r].start + t;
        char c;
        if (q == rd[r].site) c = "ACGT"[rd[r].base];
        else if ((uint32_t)(splitmix64(k_gap ^ (rid * 1024 + (uint64_t)t)) >> 32) < gap_thr) c = '-';
        else c = "ACGT"[splitmix64(k_fill ^ (uint64_t)q) & 3];
        bases[r * (int64_t)L + t] = c;
      }
    }
    base_off[n_reads] = n_reads * (int64_t)L;
    free(rd);
  }
  return n_reads;
}
