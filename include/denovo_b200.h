/*
 * denovo_b200.h — C ABI of the B200-native trio de novo hot path.
 *
 * This is the drop-in boundary: a batched replacement for the three per-site
 * seams of alenzhao/denovo-variant-caller-java
 * ("M/" = src/main/java/com/google/cloud/genomics/denovo/):
 *
 *   BayesCallResult BayesInfer.infer(Map<TrioMember,ReadSummary>, InferenceMethod)
 *                                               M/BayesInfer.java:50-69
 *   BayesInferenceResult DenovoBayesNet.performInference(Map<TrioMember,ReadSummary>)
 *                                               M/DenovoBayesNet.java:247-289
 *   new ReadSummary(List<Read>, Long)           M/ReadSummary.java:45-59
 *   VariantsBuffer.checkAndAdd / retrieveNextCall / PositionCall.isDenovo
 *                                               M/VariantsBuffer.java:172-222,340-355
 *
 * A per-site FFI call would be ~1e5x too slow, so a caller plugin
 * (GpuReadCaller / GpuVariantCaller extends DenovoCaller, M/DenovoCaller.java:24-32)
 * packs a whole shard and makes O(1) calls.  INTEGRATION.md shows the Java
 * (Panama FFM) binding for every entry point below.
 *
 * Conventions
 *  - plain C, no exceptions cross the ABI: every function returns DNV_OK (0) or
 *    a negative DNV_E_* code; dnv_last_error() gives the thread-local message.
 *  - the caller owns every buffer it passes; the library never retains caller
 *    pointers after return and never frees them.  Device scratch lives in dnv_ctx.
 *  - a dnv_ctx is bound to one CUDA device and is NOT re-entrant (one host call at a
 *    time per ctx; internally it is multi-stream); different ctxs may be driven from
 *    different host threads.
 *  - results are returned in INPUT ORDER (deterministic), unlike the reference's
 *    thread-race-ordered output file (M/ReadCaller.java:122-129).
 *  - there is no CPU fallback: without a CUDA device dnv_create fails with
 *    DNV_E_CUDA.
 *
 * Encodings (M/DenovoUtil.java:135-143,195-205)
 *  - allele  A0 C1 G2 T3;  genotype AA0 AC1 AG2 AT3 CC4 CG5 CT6 GG7 GT8 TT9
 *  - packed pileup counts: uint16 [site][person][allele], person order
 *    DAD, MOM, CHILD (the order of maxGenoType, M/DenovoBayesNet.java:277);
 *    24 bytes per site.
 *  - TrioMember ordinals (used in dnv_vcall.person): CHILD0 MOM1 DAD2.
 */
#ifndef DENOVO_B200_H
#define DENOVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNV_ABI_VERSION 2

enum {
  DNV_OK = 0,
  DNV_E_INVALID = -1,  /* null pointer, n<0, eps/mu outside (0,1), bad enum   */
  DNV_E_CUDA = -2,     /* CUDA runtime failure (message holds the driver string) */
  DNV_E_CAPACITY = -3, /* output capacity too small; *n_out holds the required size */
  DNV_E_BADBASE = -4,  /* a base outside {A,C,G,T,-}: the reference throws
                          IllegalArgumentException (M/ReadSummary.java:55) */
  DNV_E_UNSORTED = -5, /* VARIANT stream not sorted by start, or end <= start */
  DNV_E_NOMEM = -6
};

/* InferenceMethod ordinals, M/DenovoUtil.java:261-288 */
enum { DNV_METHOD_MAP = 0, DNV_METHOD_BAYES = 1, DNV_METHOD_LRT = 2 };

/* dnv_params.flags */
enum {
  DNV_F_FORCE_DIRECT = 1u << 0, /* evaluate EVERY site with the exact direct-form
                                   kernel (1000 exp, reference operation order);
                                   validation only, ~30x slower */
  DNV_F_TIME_KERNELS = 1u << 1, /* bracket the kernels with CUDA events on the launching
                                   stream; read dnv_stats.fused_kernel_ms / screen_kernel_ms */
  DNV_F_NO_SCREEN = 1u << 2     /* BAYES/LRT: skip the exact integer screen and evaluate
                                   every site with the fused FP64 kernel (A/B measurements) */
};

/* DenovoShared getters that reach the engine (M/DenovoShared.java:76-204). */
typedef struct {
  double seq_err_rate;    /* getSequenceErrorRate()  --seq_err_rate    (1e-2) */
  double denovo_mut_rate; /* getDenovoMutationRate() --denovo_mut_rate (1e-8) */
  double lrt_threshold;   /* getLrtThreshold()       --lrt_threshold   (1.0)  */
  int32_t method;         /* getInferMethod()        --inference_method (BAYES) */
  int32_t device;         /* CUDA device ordinal */
  uint32_t flags;         /* DNV_F_* */
  uint32_t reserved;
} dnv_params;

/* One called site = BayesCallResult + BayesInferenceResult
 * (M/BayesInfer.java:74-106, M/BayesInferenceResult.java:23-51).
 * dad/mom/child and max_ll always come from the exact log-domain sweep (bit-identical
 * to the reference's argmax, M/DenovoBayesNet.java:260-278).  p, lr, log_men, log_den
 * come from the factored linear form (within 1e-9 relative of the reference's direct
 * form); for HARD sites (flags bit0: threshold-ambiguous, underflow gray zone, deeper
 * than the power tables) and under DNV_F_FORCE_DIRECT they come from the exact
 * direct-form evaluation in the reference's own operation order. */
typedef struct {
  int64_t site;        /* index into the input batch */
  uint16_t counts[12]; /* DAD,MOM,CHILD x A,C,G,T (for the readCounts= field) */
  uint8_t dad, mom, child; /* maxGenoType */
  uint8_t is_denovo;
  uint32_t flags;      /* bit0: decision came from the exact path (hard site) */
  double max_ll, p, lr, log_men, log_den;
} dnv_call; /* 80 bytes */

/* Dense per-site result (optional output). */
typedef struct {
  double p, lr, log_men, log_den, max_ll;
  uint8_t dad, mom, child; /* maxGenoType */
  uint8_t is_denovo;
  uint32_t flags;          /* bit0: exact path */
} dnv_site_result; /* 48 bytes */

/* Counters of the last inference call on this ctx. */
typedef struct {
  int64_t n_sites;
  int64_t n_calls;
  int64_t n_hard;      /* sites re-evaluated by the exact direct kernel
                          (threshold-ambiguous or in the underflow gray zone) */
  int64_t n_launches;  /* kernels launched by the last call */
  double fused_kernel_ms; /* CUDA-event duration of the decision kernels (integer screen +
                             fused likelihood/posterior) of the last call, summed over
                             chunks (only when DNV_F_TIME_KERNELS is set, else 0) */
  double screen_kernel_ms; /* the integer screen's share of fused_kernel_ms */
  int64_t n_screened_out;  /* sites the integer screen proved non-de-novo (no FP64 work) */
} dnv_stats;

/* One (Variant, VariantCall) pair of the VARIANT-mode stream, flattened to the
 * fields VariantsBuffer reads (M/VariantsBuffer.java:172-306). */
typedef struct {
  int64_t start;       /* Variant.getStart() */
  int64_t end;         /* Variant.getEnd() */
  int32_t variant_id;  /* consecutive calls with equal id share one Variant */
  int8_t gt0, gt1;     /* call.getGenotype()[0..1]; -1 = '.' */
  uint8_t ngt;         /* call.getGenotype().size() */
  uint8_t person;      /* TrioMember ordinal 0 CHILD, 1 MOM, 2 DAD */
  uint8_t filter_pass; /* 1 iff info.FILTER == ["PASS"] (:245-249) */
  uint8_t ref_len;     /* referenceBases.length(), saturated at 255 */
  char ref_base;       /* referenceBases.charAt(0) */
  uint8_t alt0_len;    /* length of the ALT selected by gt0 (0 if gt0 <= 0) */
  char alt0_base;
  uint8_t alt1_len;
  char alt1_base;
  uint8_t pad_[1];
} dnv_vcall; /* 32 bytes */

/* VariantsBuffer.PositionCall (M/VariantsBuffer.java:340-375) */
typedef struct {
  int64_t position;
  int64_t child_index; /* index of the child record in the input stream */
  uint8_t child, mom, dad;
  uint8_t is_denovo;
  uint8_t pad_[4];
} dnv_pcall; /* 24 bytes */

typedef struct dnv_ctx dnv_ctx;

/* ---- lifecycle ---------------------------------------------------------- */
/* Replaces `new BayesInfer(shared)` -> `new DenovoBayesNet(shared)`
 * (M/ReadCaller.java:49-52, M/DenovoBayesNet.java:52-63): builds the CPTs once
 * and uploads the site-invariant tables. */
int dnv_create(const dnv_params *params, dnv_ctx **out);
int dnv_destroy(dnv_ctx *ctx);
const char *dnv_last_error(void);
int dnv_abi_version(void);
/* Site-invariant tables as the device holds them (for parity tests):
 * rel[1000] = getRelationshipLogLikelihood (M/DenovoBayesNet.java:322-330),
 * base_ll[40] = getBaseLogLikelihood[g*4+a] (:154-166),
 * is_denovo[1000] = isDenovoMap (M/DenovoUtil.java:37-55). Any may be NULL. */
int dnv_get_tables(dnv_ctx *ctx, double *rel, double *base_ll, uint8_t *is_denovo);
int dnv_get_stats(dnv_ctx *ctx, dnv_stats *out);
/* Run this ctx's kernels on the caller's CUDA stream (a cudaStream_t) so the
 * caller can bracket calls with its own events; NULL restores the ctx's stream. */
int dnv_set_stream(dnv_ctx *ctx, void *cuda_stream);

/* ---- READ mode: pileup counts -> calls ---------------------------------- */
/* Batched BayesInfer.infer (M/BayesInfer.java:50-69) with HOST buffers:
 * counts[n][3][4] -> calls (input order, only is_denovo sites, like the
 * `if (result.isDenovo())` writer, M/ReadCaller.java:122-129) and optionally a
 * dense per-site record.  Host<->device copies are pipelined inside.
 * *n_calls is always set; DNV_E_CAPACITY if it exceeds cap. dense may be NULL. */
int dnv_infer_counts(dnv_ctx *ctx, const uint16_t *counts, int64_t n,
                     dnv_site_result *dense, dnv_call *calls, int64_t cap,
                     int64_t *n_calls);
/* Same with DEVICE buffers already resident in HBM (d_dense may be NULL); d_calls must be
 * 16-byte aligned and d_counts 8-byte aligned (any cudaMalloc pointer is; a 16-byte aligned
 * d_counts also enables the bulk-copy path).
 * n_calls is a host pointer.  Runs on the ctx's own stream and synchronises it. */
int dnv_infer_counts_device(dnv_ctx *ctx, const uint16_t *d_counts, int64_t n,
                            dnv_site_result *d_dense, dnv_call *d_calls,
                            int64_t cap, int64_t *n_calls);

/* Asynchronous form for streams of batches (a genome walked chunk by chunk, a candidate
 * file read block by block): enqueues the whole pipeline and returns without waiting.  The batch
 * starts when the work already enqueued on the ctx's stream has finished (so inputs produced there are
 * ready) and runs on one of the ctx's internal lanes (streams): the bandwidth-bound screen of one batch
 * overlaps the short latency-bound kernels of the previous ones.  Inputs and outputs of a batch must stay
 * untouched until dnv_flush / dnv_sync.  The number of calls of this batch is STORED BY THE LAST KERNEL into
 * *n_calls_out, which must be device-writable: pinned host memory (cudaHostAlloc /
 * cudaHostRegister, mapped under UVA), this GPU's memory, or a peer GPU's memory mapped
 * here (the multi-GPU gather puts it next to the records in rank 0's HBM); call
 * dnv_sync (or synchronise the stream) before reading it or d_calls.  Records beyond cap
 * are not written (compare *n_calls_out with cap).  Scratch inside the ctx is reused in
 * stream order, so batches may be enqueued back to back; each batch needs its own d_calls
 * and n_calls_out if the results are to outlive the next batch.  dnv_get_stats after
 * dnv_sync reports the sums over the batches enqueued since the previous dnv_sync. */
int dnv_infer_counts_device_async(dnv_ctx *ctx, const uint16_t *d_counts, int64_t n,
                                  dnv_call *d_calls, int64_t cap, int64_t *n_calls_out);
int dnv_infer_counts8_device_async(dnv_ctx *ctx, const uint8_t *d_counts8, int64_t n,
                                   dnv_call *d_calls, int64_t cap, int64_t *n_calls_out);
/* Same, with site_base added to dnv_call.site: the batch is one interval shard [site_base, site_base + n)
 * of a larger axis and its records carry GLOBAL site indices (what the gather of section "multi-GPU" and a
 * chunked genome walk want). */
int dnv_infer_counts_device_async_at(dnv_ctx *ctx, const uint16_t *d_counts, int64_t n, int64_t site_base,
                                     dnv_call *d_calls, int64_t cap, int64_t *n_calls_out);
int dnv_infer_counts8_device_async_at(dnv_ctx *ctx, const uint8_t *d_counts8, int64_t n, int64_t site_base,
                                      dnv_call *d_calls, int64_t cap, int64_t *n_calls_out);
/* Device-side join: the ctx's stream (see dnv_set_stream) waits for every batch enqueued so far, the host does
 * not.  Work enqueued on that stream afterwards (the caller's own kernels, copies, timing events) sees the
 * results.  dnv_sync = dnv_flush + host wait + statistics. */
int dnv_flush(dnv_ctx *ctx);
int dnv_sync(dnv_ctx *ctx);

/* Compact input variants: counts as uint8 [n][3][4] (12 bytes per site) for pileups whose
 * per-allele depth is below 256 — halves the host->device traffic that bounds the host path.
 * Same outputs (dnv_call.counts stays uint16). */
int dnv_infer_counts8(dnv_ctx *ctx, const uint8_t *counts8, int64_t n, dnv_site_result *dense,
                      dnv_call *calls, int64_t cap, int64_t *n_calls);
int dnv_infer_counts8_device(dnv_ctx *ctx, const uint8_t *d_counts8, int64_t n,
                             dnv_site_result *d_dense, dnv_call *d_calls, int64_t cap,
                             int64_t *n_calls);

/* ---- READ mode: reads -> pileup counts ---------------------------------- */
/* Batched `new ReadSummary(reads, position)` x3 (M/ReadSummary.java:45-59,
 * M/ReadCaller.java:138-146).  Reads are sorted by (person, start); person in
 * packed order 0 DAD, 1 MOM, 2 CHILD.  bases = concatenated alignedSequence
 * strings ('A','C','G','T','-'); base_off has n_reads+1 entries.  site_pos is
 * ascending.  counts_out[n_sites][3][4] uint16. */
int dnv_pileup(dnv_ctx *ctx, const int64_t *read_start, const int64_t *base_off,
               const char *bases, const uint8_t *person, int64_t n_reads,
               const int64_t *site_pos, int64_t n_sites, uint16_t *counts_out);
int dnv_pileup_device(dnv_ctx *ctx, const int64_t *d_read_start,
                      const int64_t *d_base_off, const char *d_bases,
                      const uint8_t *d_person, int64_t n_reads,
                      const int64_t *d_site_pos, int64_t n_sites,
                      uint16_t *d_counts_out);

/* ---- READ mode: reads -> calls, fused ------------------------------------- */
/* ReadCaller.runBayesDenovoInference for a whole batch of candidate sites (M/ReadCaller.java:111-130):
 * `new ReadSummary(reads, position)` x3 (:138-146, M/ReadSummary.java:45-59) and BayesInfer.infer (:119-120) in
 * one pass.  Same read layout as dnv_pileup; same records as dnv_infer_counts over the counts dnv_pileup would
 * produce (site = index into site_pos).  For consecutive positions the pileup of a tile is accumulated in shared
 * memory and screened there; per-site counts only ever reach HBM for the ~1 % of sites the screen cannot decide. */
int dnv_infer_reads(dnv_ctx *ctx, const int64_t *read_start, const int64_t *base_off, const char *bases,
                    const uint8_t *person, int64_t n_reads, const int64_t *site_pos, int64_t n_sites,
                    dnv_call *calls, int64_t cap, int64_t *n_calls);
int dnv_infer_reads_device(dnv_ctx *ctx, const int64_t *d_read_start, const int64_t *d_base_off,
                           const char *d_bases, const uint8_t *d_person, int64_t n_reads,
                           const int64_t *d_site_pos, int64_t n_sites, dnv_call *d_calls, int64_t cap,
                           int64_t *n_calls);

/* ---- VARIANT mode: Mendelian-inconsistency filter ----------------------- */
/* One shard of VariantCaller.callSimpleDenovo (M/VariantCaller.java:120-210):
 * admission (checkAndAdd), trio interval join (retrieveNextCall) and
 * PositionCall.isDenovo over a start-sorted stream of n (variant, call) pairs.
 * Emits, in stream order, one dnv_pcall per child SNP for which both parents
 * have a covering entry (is_denovo tells which ones the reference would print).
 * *n_out is always set; DNV_E_CAPACITY if it exceeds cap. */
int dnv_mendel_filter(dnv_ctx *ctx, const dnv_vcall *calls, int64_t n,
                      dnv_pcall *out, int64_t cap, int64_t *n_out);
int dnv_mendel_filter_device(dnv_ctx *ctx, const dnv_vcall *d_calls, int64_t n,
                             dnv_pcall *d_out, int64_t cap, int64_t *n_out);

/* ---- synthetic trio source (replaces the Google Genomics client) --------- */
/* Fills d_counts[n][3][4] for global site indices first..first+n-1 with the
 * seeded, shard-invariant generator of SURVEY.md Appendix E (specified by
 * oracle/synth_ref.c).  thr: host array of n_thr 2^64-scaled Poisson CDF
 * thresholds; err_thr = seq_err * 2^32. */
int dnv_synth_counts_device(dnv_ctx *ctx, uint64_t seed, int64_t first, int64_t n,
                            const uint64_t *thr, int32_t n_thr, uint32_t err_thr,
                            uint16_t *d_counts);

/* Same generator for an arbitrary list of global site indices (candidate sites of
 * READ mode: index = contig offset + position).  d_index: device int64[n]. */
int dnv_synth_counts_at_device(dnv_ctx *ctx, uint64_t seed, const int64_t *d_index, int64_t n,
                               const uint64_t *thr, int32_t n_thr, uint32_t err_thr,
                               uint16_t *d_counts);
/* Host-buffer convenience wrapper of the above (index and counts on the host). */
int dnv_synth_counts_at(dnv_ctx *ctx, uint64_t seed, const int64_t *index, int64_t n,
                        const uint64_t *thr, int32_t n_thr, uint32_t err_thr, uint16_t *counts);

/* ---- multi-GPU: interval shards and the gather of call records ---------------- */
/* Sites shard by genomic interval: rank r infers the r-th contiguous range of the axis, so rank order =
 * genomic order and the ranks never exchange data while they compute.  The one exchange is the final gather
 * of the call records to one rank — the reference's shared PrintWriter behind `synchronized writeCalls`
 * (M/DenovoCaller.java:56-59, M/ReadCaller.java:122-129) fed by independent per-site tasks
 * (M/ReadCaller.java:91-95) and per-thread position shards (M/VariantCaller.java:86-103).  It is fused into
 * the producing kernels: every rank's kernels store their records and counts straight into the destination
 * rank's HBM over NVLink; one flag per rank and round tells the destination that everything has landed.
 *
 * One process per GPU (torchrun, MPI):            One process, N GPUs:
 *   every rank   dnv_gather_create                  dnv_group_create / dnv_group_infer_counts (below), or
 *   dst          dnv_gather_export -> 64 bytes      dnv_gather_create per device + dnv_gather_attach
 *   launcher     broadcasts the 64 bytes
 *   others       dnv_gather_import
 * then, per round:  dnv_infer_counts[8]_device_gather (any number of batches, generation = batch % gens)
 *                   dnv_gather_arrive on every rank;  dst: dnv_gather_wait, [dnv_sync,] dnv_gather_fetch.
 * A generation's slots may be rewritten once the destination has fetched them (the caller orders that:
 * program order in one process, its launcher's barrier between processes). */
typedef struct dnv_gather dnv_gather;
#define DNV_GATHER_HANDLE_BYTES 64
/* This rank's end of a gather of `world` ranks to rank `dst`: up to cap records per rank and generation,
 * gens generations of slots.  The end of rank dst allocates the destination buffer in its HBM. */
int dnv_gather_create(dnv_ctx *ctx, int32_t world, int32_t rank, int32_t dst, int64_t cap, int32_t gens,
                      dnv_gather **out);
int dnv_gather_destroy(dnv_gather *g);
/* dst: the 64-byte inter-process handle of the destination buffer; others: map it. */
int dnv_gather_export(dnv_gather *g, void *handle64);
int dnv_gather_import(dnv_gather *g, const void *handle64);
/* Same process: map the destination buffer through peer access (the two ends may share a device). */
int dnv_gather_attach(dnv_gather *g, dnv_gather *dst_end);
/* Where rank `rank` writes generation `gen`, as THIS end's device sees it (on dst: where to read it). */
int dnv_gather_slot(dnv_gather *g, int32_t rank, int32_t gen, dnv_call **records, int64_t **count);
/* dnv_infer_counts[8]_device_async_at whose records and count land in this rank's slot of generation gen:
 * the shard [site_base, site_base + n) of the axis, records carrying GLOBAL site indices. */
int dnv_infer_counts_device_gather(dnv_ctx *ctx, const uint16_t *d_counts, int64_t n, int64_t site_base,
                                   dnv_gather *g, int32_t gen);
int dnv_infer_counts8_device_gather(dnv_ctx *ctx, const uint8_t *d_counts8, int64_t n, int64_t site_base,
                                    dnv_gather *g, int32_t gen);
/* Every rank, after the batches of a round: dnv_flush + a release flag in the destination (on the device,
 * no host wait). */
int dnv_gather_arrive(dnv_gather *g);
/* dst, after its own dnv_gather_arrive: the ctx's stream waits (on the device) until every rank has arrived
 * at this round.  Work enqueued on that stream afterwards — timing events, copies — sees all records. */
int dnv_gather_wait(dnv_gather *g);
/* dst: waits for the ctx's stream, then copies generation gen to the host in rank order = genomic order.
 * per_rank (world entries) may be NULL.  DNV_E_CAPACITY if cap (or a rank's slot) was too small, with the
 * required total in *n_calls; DNV_E_CUDA if a rank never arrived (10 s). */
int dnv_gather_fetch(dnv_gather *g, int32_t gen, dnv_call *calls, int64_t cap, int64_t *n_calls,
                     int64_t *per_rank);

/* N devices driven by one process: what a single-process host (the CLI, the Java binding) uses instead of
 * N processes.  The same device may be listed more than once (tests on a one-GPU box). */
typedef struct dnv_group dnv_group;
int dnv_group_create(const dnv_params *params /* .device ignored */, const int32_t *devices,
                     int32_t n_devices, dnv_group **out);
int dnv_group_destroy(dnv_group *grp);
/* Batched BayesInfer.infer over HOST counts, interval-sharded over the group's devices (one host thread per
 * device for the copies and launches), records gathered through device 0 and returned in input order with
 * site = index into counts.  Same contract as dnv_infer_counts / dnv_infer_counts8 (calls only). */
int dnv_group_infer_counts(dnv_group *grp, const uint16_t *counts, int64_t n, dnv_call *calls,
                           int64_t cap, int64_t *n_calls);
int dnv_group_infer_counts8(dnv_group *grp, const uint8_t *counts8, int64_t n, dnv_call *calls,
                            int64_t cap, int64_t *n_calls);
int dnv_group_size(dnv_group *grp);
int dnv_group_ctx(dnv_group *grp, int32_t rank, dnv_ctx **out); /* borrowed: owned by the group */
int dnv_group_get_stats(dnv_group *grp, dnv_stats *out);         /* sums over the devices, last call */

/* ---- host memory ------------------------------------------------------------ */
/* Page-locked host buffers (cudaHostAlloc, portable across devices): what the host side of a caller plugin
 * should hand to dnv_infer_counts / dnv_group_infer_counts — pageable memory goes through a staging copy at a
 * fraction of the PCIe rate.  (Java: MemorySegment.reinterpret over the returned address.) */
int dnv_alloc_pinned(size_t bytes, void **out);
int dnv_free_pinned(void *p);

/* ---- calibration ----------------------------------------------------------- */
/* Dependent-free DFMA chains on every SM: measured FP64 peak in TFLOP/s
 * (MEASURED_PEAKS.json has no FP64 figure). */
int dnv_measure_fp64_peak(dnv_ctx *ctx, double *tflops, double *ms);
/* `reps` back-to-back launches of the exact integer screen ALONE (the dominant kernel of the READ path) over one
 * chunk of device-resident counts (count_bytes 24: uint16 layout, 12: uint8; n <= 2^27 sites), on the context's
 * stream between one pair of CUDA events: the kernel's average duration with the launch latency of an idle stream
 * amortised.  bench.py's roofline denominator; not part of the caller path. */
int dnv_measure_screen(dnv_ctx *ctx, const void *d_counts, int count_bytes, int64_t n, int reps, double *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* DENOVO_B200_H */
