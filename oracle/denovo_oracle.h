/*
 * denovo_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's READ-mode trio Bayes-net inference,
 * pileup summariser and VARIANT-mode Mendelian filter
 * (alenzhao/denovo-variant-caller-java).  Every function cites the reference
 * file:line it follows ("M/" = src/main/java/com/google/cloud/genomics/denovo/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libdenovo_b200.so) never
 * links, loads or falls back to it.
 *
 * Parity pin: the restatement is checked against every known-answer case the
 * reference's JUnit suite holds for this path (the JSON files under tests/golden/, extracted
 * from the reference's test sources by tests/golden/make_golden.py).  The
 * reference's Java cannot run here (no JVM), so numeric values of p / LR /
 * logL are pinned only by the formula tests (DenovoBayesNetTest) and by the
 * MAP / isDenovo expectations, exactly as far as the reference's own tests go.
 * VARIANT-mode admission + interval join are NOT pinned by any reference test
 * ("parity unpinned" for that sub-path beyond the 1000-entry truth table).
 */
#ifndef DENOVO_ORACLE_H
#define DENOVO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DenovoShared getters used by the path (M/DenovoShared.java:76-204). */
typedef struct {
  double seq_err_rate;    /* --seq_err_rate    default 1e-2 (M/CommandLine.java:68-70) */
  double denovo_mut_rate; /* --denovo_mut_rate default 1e-8 (M/CommandLine.java:72-74) */
  double lrt_threshold;   /* --lrt_threshold   default 1.0  (M/CommandLine.java:76-78) */
} orc_params;

/* BayesInferenceResult (M/BayesInferenceResult.java:23-51) + the three
 * InferenceMethod decisions (M/DenovoUtil.java:261-288). */
typedef struct {
  int32_t dad, mom, child; /* MAP triple, Genotype ordinals (M/DenovoUtil.java:195-205) */
  int32_t denovo_map, denovo_bayes, denovo_lrt;
  double max_ll;  /* maxLikelihood */
  double p;       /* bayesDenovoProb */
  double lr;      /* likelihoodRatio (raw ratio, not its log) */
  double log_men; /* mendelianLogLikelihood */
  double log_den; /* denovoLogLikelihood */
  double l_den;   /* raw sums, kept so tests can stratify underflow cases */
  double l_men;
} orc_result;

/* One (variant, call) pair of the VARIANT-mode stream, flattened.  Mirrors
 * exactly the fields VariantsBuffer reads (M/VariantsBuffer.java:172-306). */
typedef struct {
  int64_t start;      /* Variant.getStart() */
  int64_t end;        /* Variant.getEnd()   */
  int32_t variant_id; /* consecutive calls with equal id share one Variant */
  int8_t gt0, gt1;    /* call.getGenotype()[0..1]; -1 = '.' */
  uint8_t ngt;        /* call.getGenotype().size() */
  uint8_t person;     /* TrioMember ordinal: 0 CHILD, 1 MOM, 2 DAD (M/DenovoUtil.java:135-137) */
  uint8_t filter_pass;/* 1 iff info has FILTER == ["PASS"] (M/VariantsBuffer.java:245-249) */
  uint8_t ref_len;    /* referenceBases.length(), saturated at 255 */
  char ref_base;      /* referenceBases.charAt(0) */
  uint8_t alt0_len;   /* length of ALT selected by gt0 (0 if gt0<=0) */
  char alt0_base;
  uint8_t alt1_len;   /* length of ALT selected by gt1 (0 if gt1<=0) */
  char alt1_base;
  uint8_t pad_[1];
} orc_vcall; /* 32 bytes */

/* VariantsBuffer.PositionCall (M/VariantsBuffer.java:340-375) */
typedef struct {
  int64_t position;
  int64_t child_index; /* index of the child record in the input stream */
  uint8_t child, mom, dad; /* Genotype ordinals */
  uint8_t is_denovo;
  uint8_t pad_[4];
} orc_pcall; /* 24 bytes */

/* ---- tables (site-invariant) ---- */
/* isDenovoMap, M/DenovoUtil.java:37-55.  index = dad*100+mom*10+child. */
void orc_is_denovo_table(uint8_t out[1000]);
/* Parent prior, M/DenovoBayesNet.java:82-86. */
void orc_parent_prior(double out[10]);
/* Child CPT, M/DenovoBayesNet.java:87-117 (+ mendelianGenotypes :127-145). */
void orc_child_cpt(const orc_params *p, double out[1000]);
/* getBaseLogLikelihood, M/DenovoBayesNet.java:154-166.  out[g*4+allele]. */
void orc_base_loglik(const orc_params *p, double out[40]);
/* getRelationshipLogLikelihood, M/DenovoBayesNet.java:322-330. out[1000]. */
void orc_relationship_loglik(const orc_params *p, double out[1000]);
/* getReadSummaryLogLikelihood, M/DenovoBayesNet.java:193-211.  counts[4] A,C,G,T. */
void orc_read_loglik(const orc_params *p, const uint16_t counts[4], double out[10]);

/* ---- per-site inference ---- */
/* performInference (M/DenovoBayesNet.java:247-289) + InferenceMethod
 * (M/DenovoUtil.java:261-288) over n sites.
 * counts: [n][3][4] uint16, person order DAD, MOM, CHILD; allele order A,C,G,T.
 * faithful_logs != 0: recompute the three Math.log per triple exactly as the
 *   reference does (M/DenovoBayesNet.java:220-233) — same bits, reference cost.
 * threads: worker threads (one contiguous block of sites each; the reference
 *   runs one task per site on a pool, M/ReadCaller.java:66-70,91-95). */
int orc_infer(const orc_params *p, const uint16_t *counts, int64_t n,
              orc_result *out, int faithful_logs, int threads);

/* ---- pileup (ReadSummary, M/ReadSummary.java:45-59) ---- */
/* Reads: start[r], bases = concatenated alignedSequence, base_off[r..r+1],
 * person[r] in packed order 0 DAD, 1 MOM, 2 CHILD.  For each site position the
 * reads of each person that overlap it (start <= pos < start+len — the
 * reads.search interval [pos,pos+1), M/ReadCaller.java:160-161) contribute
 * bases[pos-start]; '-' skipped; any other non-ACGT char is the reference's
 * IllegalArgumentException -> returns -(read index+1).
 * counts_out: [n_sites][3][4] uint16.  Returns 0 on success. */
int64_t orc_pileup(const int64_t *read_start, const int64_t *base_off,
                   const char *bases, const uint8_t *person, int64_t n_reads,
                   const int64_t *site_pos, int64_t n_sites,
                   uint16_t *counts_out);

/* Same rule, reads sorted by (person, start): binary-searched windows (big test inputs). */
int64_t orc_pileup_sorted(const int64_t *read_start, const int64_t *base_off,
                          const char *bases, const uint8_t *person, int64_t n_reads,
                          const int64_t *site_pos, int64_t n_sites,
                          uint16_t *counts_out);

/* ---- VARIANT mode (M/VariantsBuffer.java + M/VariantCaller.java:120-210) ---- */
/* Streaming simulation of one shard: admission, trigger, eviction, interval
 * match, genotype decode.  Emits one orc_pcall per child SNP that found both
 * parents (denovo or not; is_denovo says which).  Returns number emitted, or
 * -(index+1) of the record where the reference would throw
 * (Allele.valueOf on a non-ACGT base). */
int64_t orc_mendel_stream(const orc_vcall *calls, int64_t n, orc_pcall *out,
                          int64_t cap);
/* checkAndAdd admission only (M/VariantsBuffer.java:172-185): admitted[i] 0/1. */
void orc_mendel_admit(const orc_vcall *calls, int64_t n, uint8_t *admitted);

#ifdef __cplusplus
}
#endif
#endif
