/*
 * denovo_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See denovo_oracle.h for scope, parity pin and who may load this.
 *
 * Build: -O2 -ffp-contract=off (Java never fuses a*b+c; SSE2 doubles have no
 * excess precision), see oracle/Makefile.
 *
 * "M/" = /root/reference/src/main/java/com/google/cloud/genomics/denovo/
 */
#include "denovo_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Genotype enum order, M/DenovoUtil.java:195-205 */
static const char GENO[10][3] = {"AA", "AC", "AG", "AT", "CC",
                                 "CG", "CT", "GG", "GT", "TT"};
/* Allele enum order, M/DenovoUtil.java:141-142 */
static const char ALLELE[4] = {'A', 'C', 'G', 'T'};

static int geno_is_hom(int g) { return GENO[g][0] == GENO[g][1]; }
static int geno_contains(int g, char a) {
  return GENO[g][0] == a || GENO[g][1] == a;
}
static int allele_index(char c) {
  switch (c) {
  case 'A': return 0;
  case 'C': return 1;
  case 'G': return 2;
  case 'T': return 3;
  default: return -1; /* Allele.valueOf throws IllegalArgumentException */
  }
}
/* Genotype.valueOfPairAlleles (sort, then valueOf), M/DenovoUtil.java:229-233 */
static int geno_of_alleles(int a, int b) {
  if (a > b) { int t = a; a = b; b = t; }
  for (int g = 0; g < 10; g++)
    if (GENO[g][0] == ALLELE[a] && GENO[g][1] == ALLELE[b]) return g;
  return -1;
}

/* M/DenovoUtil.java:41-50 */
static int is_denovo(int d, int m, int c) {
  char c1 = GENO[c][0], c2 = GENO[c][1];
  int in_parents = geno_contains(m, c1) & geno_contains(d, c2);
  int in_parents_mirror = geno_contains(m, c2) & geno_contains(d, c1);
  return !(in_parents || in_parents_mirror);
}

void orc_is_denovo_table(uint8_t out[1000]) {
  for (int d = 0; d < 10; d++)
    for (int m = 0; m < 10; m++)
      for (int c = 0; c < 10; c++)
        out[d * 100 + m * 10 + c] = (uint8_t)is_denovo(d, m, c);
}

/* M/DenovoBayesNet.java:82-86: (hom ? 1.0 : 2.0) / (4*4) */
void orc_parent_prior(double out[10]) {
  for (int g = 0; g < 10; g++) out[g] = (geno_is_hom(g) ? 1.0 : 2.0) / (4 * 4);
}

/* mendelianGenotypes, M/DenovoBayesNet.java:127-145 */
static int mendelian_genotypes(int d, int m, int cnt[10]) {
  int keys = 0;
  memset(cnt, 0, 10 * sizeof(int));
  for (int ii = 0; ii < 2; ii++)
    for (int jj = 0; jj < 2; jj++) {
      int g = geno_of_alleles(allele_index(GENO[d][ii]), allele_index(GENO[m][jj]));
      if (cnt[g] == 0) keys++;
      cnt[g]++;
    }
  return keys;
}

/* M/DenovoBayesNet.java:87-117 */
void orc_child_cpt(const orc_params *p, double out[1000]) {
  for (int d = 0; d < 10; d++)
    for (int m = 0; m < 10; m++) {
      int cnt[10];
      int keys = mendelian_genotypes(d, m, cnt);
      int num_denovo = 10 - keys; /* :98 */
      int mendelian_count = 0;    /* :99-102 */
      for (int g = 0; g < 10; g++) mendelian_count += cnt[g];
      for (int c = 0; c < 10; c++) {
        double v = is_denovo(d, m, c)
                       ? p->denovo_mut_rate / num_denovo
                       : (1.0 - p->denovo_mut_rate) / mendelian_count * cnt[c];
        out[d * 100 + m * 10 + c] = v; /* :106-109 */
      }
    }
}

/* getBaseLogLikelihood, M/DenovoBayesNet.java:154-166 (Math.log every call) */
static double base_loglik(double e, int g, int allele) {
  char b = ALLELE[allele];
  return geno_is_hom(g)
             ? (geno_contains(g, b) ? log(1 - e) : log(e) - log(3))
             : (geno_contains(g, b) ? log(1 - 2 * e / 3) - log(2)
                                    : log(e) - log(3));
}

void orc_base_loglik(const orc_params *p, double out[40]) {
  for (int g = 0; g < 10; g++)
    for (int a = 0; a < 4; a++) out[g * 4 + a] = base_loglik(p->seq_err_rate, g, a);
}

/* getRelationshipLogLikelihood, M/DenovoBayesNet.java:322-330:
 *   ll = 0; ll += log P(d); ll += log P(m); ll += log T(d,m,c) */
void orc_relationship_loglik(const orc_params *p, double out[1000]) {
  double P[10], T[1000];
  orc_parent_prior(P);
  orc_child_cpt(p, T);
  for (int d = 0; d < 10; d++)
    for (int m = 0; m < 10; m++)
      for (int c = 0; c < 10; c++) {
        double ll = 0;
        ll += log(P[d]);
        ll += log(P[m]);
        ll += log(T[d * 100 + m * 10 + c]);
        out[d * 100 + m * 10 + c] = ll;
      }
}

typedef struct {
  orc_params prm;
  double P[10];
  double T[1000];
  double BL[40];
  double R[1000];
  uint8_t isden[1000];
} tables_t;

static void build_tables(const orc_params *p, tables_t *t) {
  t->prm = *p;
  orc_parent_prior(t->P);
  orc_child_cpt(p, t->T);
  orc_base_loglik(p, t->BL);
  orc_relationship_loglik(p, t->R);
  orc_is_denovo_table(t->isden);
}

/* getReadSummaryLogLikelihood, M/DenovoBayesNet.java:193-211.  The count map
 * is a TreeMap<Allele,Integer> (M/ReadSummary.java:27) holding only alleles
 * seen at least once, iterated in A,C,G,T order. */
static void read_loglik(const tables_t *t, const uint16_t c[4], double out[10],
                        int faithful) {
  for (int g = 0; g < 10; g++) {
    double acc = 0.0;
    for (int a = 0; a < 4; a++) {
      if (c[a] == 0) continue; /* absent key */
      double bl = faithful ? base_loglik(t->prm.seq_err_rate, g, a) : t->BL[g * 4 + a];
      acc += (int)c[a] * bl;
    }
    out[g] = acc;
  }
}

void orc_read_loglik(const orc_params *p, const uint16_t counts[4], double out[10]) {
  tables_t t;
  build_tables(p, &t);
  read_loglik(&t, counts, out, 1);
}

/* performInference, M/DenovoBayesNet.java:247-289 */
static void infer_site(const tables_t *t, const uint16_t *cnt, orc_result *r,
                       int faithful) {
  double Ld[10], Lm[10], Lc[10];
  read_loglik(t, cnt + 0, Ld, faithful);
  read_loglik(t, cnt + 4, Lm, faithful);
  read_loglik(t, cnt + 8, Lc, faithful);

  double max_ll = -INFINITY;
  double l_den = 0.0, l_men = 0.0;
  int bd = -1, bm = -1, bc = -1; /* maxGenoType = null until first strict > */

  for (int d = 0; d < 10; d++)
    for (int m = 0; m < 10; m++)
      for (int c = 0; c < 10; c++) {
        int idx = d * 100 + m * 10 + c;
        /* getTrioGenotypeLogLikelihood :300-312 */
        double trio = 0.0;
        trio += Ld[d];
        trio += Lm[m];
        trio += Lc[c];
        /* getRelationshipLogLikelihood :322-330 */
        double rel;
        if (faithful) {
          rel = 0;
          rel += log(t->P[d]);
          rel += log(t->P[m]);
          rel += log(t->T[idx]);
        } else {
          rel = t->R[idx];
        }
        double ll = 0; /* :264-267 */
        ll += trio;
        ll += rel;
        if (t->isden[idx]) l_den += exp(ll); /* :269-273 */
        else l_men += exp(ll);
        if (ll > max_ll) { /* :275-278 */
          max_ll = ll;
          bd = d; bm = m; bc = c;
        }
      }
  r->dad = bd; r->mom = bm; r->child = bc;
  r->max_ll = max_ll;
  r->p = l_den / (l_den + l_men); /* :282 */
  r->lr = l_den / l_men;          /* :285 */
  r->log_men = log(l_men);        /* :288 */
  r->log_den = log(l_den);
  r->l_den = l_den;
  r->l_men = l_men;
  /* InferenceMethod, M/DenovoUtil.java:261-288 */
  r->denovo_map = (bd >= 0) ? t->isden[bd * 100 + bm * 10 + bc] : 0;
  r->denovo_bayes = r->p > 0.5;
  r->denovo_lrt = r->lr > t->prm.lrt_threshold;
}

typedef struct {
  const tables_t *t;
  const uint16_t *counts;
  orc_result *out;
  int64_t lo, hi;
  int faithful;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (int64_t i = j->lo; i < j->hi; i++)
    infer_site(j->t, j->counts + i * 12, j->out + i, j->faithful);
  return NULL;
}

int orc_infer(const orc_params *p, const uint16_t *counts, int64_t n,
              orc_result *out, int faithful_logs, int threads) {
  if (!p || (!counts && n > 0) || (!out && n > 0) || n < 0) return -1;
  tables_t t;
  build_tables(p, &t);
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads == 1 || n < threads) {
    job_t j = {&t, counts, out, 0, n, faithful_logs};
    worker(&j);
    return 0;
  }
  pthread_t th[256];
  job_t jobs[256];
  for (int k = 0; k < threads; k++) {
    jobs[k].t = &t;
    jobs[k].counts = counts;
    jobs[k].out = out;
    jobs[k].lo = n * k / threads;
    jobs[k].hi = n * (k + 1) / threads;
    jobs[k].faithful = faithful_logs;
    pthread_create(&th[k], NULL, worker, &jobs[k]);
  }
  for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
  return 0;
}

/* ReadSummary(List<Read>, Long), M/ReadSummary.java:45-59 */
int64_t orc_pileup(const int64_t *read_start, const int64_t *base_off,
                   const char *bases, const uint8_t *person, int64_t n_reads,
                   const int64_t *site_pos, int64_t n_sites,
                   uint16_t *counts_out) {
  memset(counts_out, 0, (size_t)n_sites * 12 * sizeof(uint16_t));
  for (int64_t s = 0; s < n_sites; s++) {
    int64_t pos = site_pos[s];
    for (int64_t r = 0; r < n_reads; r++) {
      int64_t len = base_off[r + 1] - base_off[r];
      int64_t off = pos - read_start[r]; /* :48 */
      if (off < 0 || off >= len) continue; /* not returned by reads.search */
      char b = bases[base_off[r] + off];   /* :49 substring(offset, offset+1) */
      if (b == '-') continue;              /* :51-53 */
      int a = allele_index(b);             /* :55 Allele.valueOf */
      if (a < 0) return -(r + 1);
      if (person[r] > 2) return -(r + 1);
      counts_out[s * 12 + person[r] * 4 + a]++;
    }
  }
  return 0;
}

/* The same rule for reads SORTED by (person, start) — what the product's dnv_pileup takes — with the window of
 * reads that can reach a position found by binary search instead of a scan over all reads: for big test
 * inputs.  tests/test_oracle_paths.py checks it against orc_pileup above. */
int64_t orc_pileup_sorted(const int64_t *read_start, const int64_t *base_off,
                          const char *bases, const uint8_t *person, int64_t n_reads,
                          const int64_t *site_pos, int64_t n_sites,
                          uint16_t *counts_out) {
  memset(counts_out, 0, (size_t)n_sites * 12 * sizeof(uint16_t));
  int64_t first[4] = {0, n_reads, n_reads, n_reads}, lmax = 0;
  for (int64_t r = 0; r < n_reads; r++) {
    int64_t len = base_off[r + 1] - base_off[r];
    if (len > lmax) lmax = len;
    if (person[r] > 2) return -(r + 1);
    if (r > 0 && (person[r - 1] > person[r] || (person[r - 1] == person[r] && read_start[r - 1] > read_start[r]))) return -(r + 1);
  }
  for (int p = 1; p <= 2; p++) {
    int64_t r = 0;
    while (r < n_reads && person[r] < p) r++;
    first[p] = r;
  }
  first[3] = n_reads;
  for (int64_t s = 0; s < n_sites; s++) {
    int64_t pos = site_pos[s];
    for (int p = 0; p < 3; p++) {
      int64_t lo = first[p], hi = first[p + 1];
      while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (read_start[mid] <= pos - lmax) lo = mid + 1; else hi = mid; }
      for (int64_t r = lo; r < first[p + 1] && read_start[r] <= pos; r++) {
        int64_t len = base_off[r + 1] - base_off[r];
        int64_t off = pos - read_start[r];
        if (off >= len) continue;
        char b = bases[base_off[r] + off];
        if (b == '-') continue;
        int a = allele_index(b);
        if (a < 0) return -(r + 1);
        counts_out[s * 12 + p * 4 + a]++;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* VARIANT mode                                                        */
/* ------------------------------------------------------------------ */
enum { CHILD = 0, MOM = 1, DAD = 2 };

/* isInsertion, M/VariantsBuffer.java:297-306 (ngt==2 case captured) */
static int vc_is_insertion(const orc_vcall *v) {
  if (v->ngt >= 1 && v->gt0 > 0 && v->alt0_len != 1) return 1;
  if (v->ngt >= 2 && v->gt1 > 0 && v->alt1_len != 1) return 1;
  return 0;
}
/* isDeletion, :288-291 */
static int vc_is_deletion(const orc_vcall *v) { return v->ref_len != 1; }
/* isSnp, :264-282 */
static int vc_is_snp(const orc_vcall *v) {
  if (v->end != v->start + 1) return 0;
  if (v->ref_len != 1) return 0;
  if (vc_is_insertion(v)) return 0;
  return 1;
}
/* callContainsDot, :255-258 */
static int vc_contains_dot(const orc_vcall *v) {
  return (v->ngt >= 1 && v->gt0 == -1) || (v->ngt >= 2 && v->gt1 == -1);
}
/* checkAndAdd, :172-185 */
static int vc_admit(const orc_vcall *v) {
  if ((v->person == CHILD && !vc_is_snp(v)) || vc_contains_dot(v) ||
      !(v->ngt == 2) || !v->filter_pass || vc_is_insertion(v) ||
      vc_is_deletion(v))
    return 0;
  return 1;
}

void orc_mendel_admit(const orc_vcall *calls, int64_t n, uint8_t *admitted) {
  for (int64_t i = 0; i < n; i++) admitted[i] = (uint8_t)vc_admit(&calls[i]);
}

/* getGenotypeFromSNP, :228-239 */
static int vc_genotype(const orc_vcall *v) {
  int a0 = allele_index(v->gt0 > 0 ? v->alt0_base : v->ref_base);
  if (a0 < 0) return -1;
  int a1 = allele_index(v->gt1 > 0 ? v->alt1_base : v->ref_base);
  if (a1 < 0) return -1;
  return geno_of_alleles(a0, a1);
}

typedef struct {
  int64_t *idx; /* record indices, FIFO */
  int64_t head, tail;
} fifo_t;

int64_t orc_mendel_stream(const orc_vcall *calls, int64_t n, orc_pcall *out,
                          int64_t cap) {
  fifo_t q[3];
  int64_t mrs[3] = {0, 0, 0}; /* mostRecentStartPosition, :44-51 */
  for (int k = 0; k < 3; k++) {
    q[k].idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    q[k].head = q[k].tail = 0;
  }
  int64_t n_out = 0, err = 0;

#define Q_EMPTY(k) (q[k].head == q[k].tail)
#define Q_FRONT(k) (&calls[q[k].idx[q[k].head]])

  int64_t i = 0;
  int flushing = 0;
  while (i < n || !flushing) {
    if (i < n) {
      /* offer all calls of one Variant, M/VariantCaller.java:169-172 */
      int32_t vid = calls[i].variant_id;
      while (i < n && calls[i].variant_id == vid) {
        const orc_vcall *v = &calls[i];
        if (v->person <= 2 && vc_admit(v)) { /* push, :56-59 */
          q[v->person].idx[q[v->person].tail++] = i;
          mrs[v->person] = v->start;
        }
        i++;
      }
    } else {
      flushing = 1; /* M/VariantCaller.java:193-209 */
    }
    for (;;) {
      if (Q_EMPTY(CHILD)) break;
      int64_t cstart = Q_FRONT(CHILD)->start;
      if (!flushing) { /* canProcess, :74-78 */
        if (!(mrs[MOM] >= cstart && mrs[DAD] >= cstart)) break;
      }
      /* retrieveNextCall, :196-222 */
      /* evictParents, :147-157 */
      for (int par = MOM; par <= DAD; par++)
        while (!Q_EMPTY(par) && Q_FRONT(par)->end < cstart) q[par].head++;
      const orc_vcall *cs = Q_FRONT(CHILD);
      int64_t cidx = q[CHILD].idx[q[CHILD].head];
      /* getNextSNP asserts isSnp (:329-336) — admission guarantees it */
      int64_t pos = cs->start;
      int gch = vc_genotype(cs);
      if (gch < 0) { err = -(cidx + 1); goto done; }
      int gpar[3] = {gch, -1, -1};
      int missing = 0;
      for (int par = MOM; par <= DAD && !missing; par++) {
        /* getMatchingPair, :314-323: first queued entry with start<=pos<end */
        const orc_vcall *hit = NULL;
        int64_t hidx = -1;
        for (int64_t k = q[par].head; k < q[par].tail; k++) {
          const orc_vcall *pv = &calls[q[par].idx[k]];
          if (pv->start <= pos && pos < pv->end) { hit = pv; hidx = q[par].idx[k]; break; }
        }
        if (!hit) { missing = 1; break; } /* :211-213 return null */
        if (vc_is_snp(hit)) {             /* :216-218 */
          gpar[par] = vc_genotype(hit);
          if (gpar[par] < 0) { err = -(hidx + 1); goto done; }
        } else { /* Genotype.valueOf(ref+ref) with the CHILD's reference base */
          int a = allele_index(cs->ref_base);
          if (a < 0) { err = -(cidx + 1); goto done; }
          gpar[par] = geno_of_alleles(a, a);
        }
      }
      if (!missing) {
        if (n_out < cap) {
          orc_pcall *o = &out[n_out];
          memset(o, 0, sizeof(*o));
          o->position = pos;
          o->child_index = cidx;
          o->child = (uint8_t)gpar[CHILD];
          o->mom = (uint8_t)gpar[MOM];
          o->dad = (uint8_t)gpar[DAD];
          o->is_denovo = (uint8_t)is_denovo(gpar[DAD], gpar[MOM], gpar[CHILD]); /* :352-355 */
        }
        n_out++;
      }
      q[CHILD].head++; /* vbuffer.pop(CHILD) */
    }
    if (flushing) break;
  }
done:
  for (int k = 0; k < 3; k++) free(q[k].idx);
  return err ? err : n_out;
}

_Static_assert(sizeof(orc_vcall) == 32, "orc_vcall must be 32 bytes");
_Static_assert(sizeof(orc_pcall) == 24, "orc_pcall must be 24 bytes");
_Static_assert(sizeof(orc_result) == 80, "orc_result must be 80 bytes");
