// dnv_internal.cuh — shared declarations of the B200 trio de novo library.
// Product code: never includes or links anything under oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "denovo_b200.h"

namespace dnv {

constexpr int kPowTable = 1024;  // entries of each r^k table held in shared memory

// Site-invariant tables, built once per ctx on the host (tables.cpp) and kept
// in device global memory; kernels stage what they need into shared memory.
struct Tables {
  // exact log-domain tables (bit-identical to the reference's arithmetic)
  double rel[1000];       // getRelationshipLogLikelihood, M/DenovoBayesNet.java:322-330
  double hm, ht, mm;      // getBaseLogLikelihood's three values, :159-165
  uint32_t isden[32];     // isDenovoMap as a 1000-bit mask, M/DenovoUtil.java:37-55
  uint8_t nden[100];      // 10 - |mendelianGenotypes(d,m)|, M/DenovoBayesNet.java:98
  double rmax_dm[100];    // max over child of rel[d][m][c]  (exact-MAP pruning bounds)
  double rmax_d[10];      // max over mom, child of rel[d][m][c]
  // linear-domain fast path (factored form)
  double pw_inv[2][kPowTable]; // r_j^-i, j=0 hom ratio, j=1 het ratio (may underflow to 0: exact)
  double pw_m[2][kPowTable];   // mantissa in [1,2) of r_j^k
  int32_t pw_e[2][kPowTable];  // binary exponent of r_j^k
  int32_t fast_ok;             // both ratios > 1 (always, unless eps is absurdly large)
  int32_t early_ok;            // the integer screen (trio_screen.cuh) is valid for these parameters (BAYES / LRT)
  int32_t map_screen_ok;       // ... and for the MAP decision
  uint32_t c_hom_fx, c_het_fx; // log2(r_j) in 16.16 fixed point
  uint32_t scr_w[4];           // screen level 1: signed-byte weights (A,C,G,T) when allele X is the candidate
  int32_t l3_lambda_fx;        // screen level 3 (trio_screen.cuh): log2 threshold in 16.16 fixed point, 0 = disabled
  int32_t l3_max_depth;        // ... largest total depth of a site it may decide (everything stays in FP64's normal range)
  double mm2;                  // log2(eps/3)
  double k_men;                // ((1-mu)/4)/64
  double k_den[4];             // (mu/256)/nDen for nDen = 9,8,7,6
  double lrt_thr;              // 1.0 for BAYES, lrt_threshold for LRT
  double lrt_threshold;        // raw --lrt_threshold (exact path)
  int32_t method;
};

void build_tables(const dnv_params& p, Tables* t);

void set_error(const std::string& msg);

}  // namespace dnv

namespace dnv {
// How a chunk's kernels learn where its call records start and publish where they end.  Chunks of one batch
// (and batches among each other) may run on different lanes (streams); the kernels that extend the chain — the
// list kernel or the compaction kernel — are serialised by an event, everything else overlaps.
struct CallChain {
  const unsigned long long* before;  // calls of this batch before this chunk: the previous chunk's pair[1]; nullptr = 0
  unsigned long long* pair;          // this chunk's private words: [0] <- calls before it, [1] <- calls up to and including it
  unsigned long long* acc;           // since the host last collected: [0] calls (added by a batch's last chunk),
                                     // [1] sites the screen left undecided, [2] hard sites
  unsigned long long* work;          // [0] tile / block ticket, [1] CTAs gone (reset by the launch that used them)
  int64_t* count_out;                // where the batch's running call total is published for the caller, or null
  int last;                          // last chunk of its batch
  unsigned long long* host_acc;      // non-null (single-chunk synchronous call): pinned host words that receive acc[0..3] when the
  unsigned long long* fin_gone;      // finish kernel's last CTA leaves (fin_gone counts them); acc restarts at zero
  unsigned long long* ktrace;        // DNV_KTRACE: the lane's time-line slots (6 words per chunk), or null
  const unsigned int* ktrace_seq;    // ... and which slot the lane is at (bumped after every chunk, so graph replays advance)
};

constexpr int kMaxLanes = 4;

// A batch whose arguments repeat (a ring of buffers walked chunk by chunk, a benchmark loop) is captured once into
// a CUDA graph and replayed with ONE launch call instead of a dozen stream operations — at ~50 us of GPU work per
// batch the host would otherwise be the bottleneck.
constexpr int kGraphSlots = 8;
struct GraphKey {
  const void* counts; const void* calls; const void* count_out; const void* dense;
  int64_t n, site_base, cap;
  int bps;
  bool operator==(const GraphKey& o) const {
    return counts == o.counts && calls == o.calls && count_out == o.count_out && dense == o.dense && n == o.n &&
           site_base == o.site_base && cap == o.cap && bps == o.bps;
  }
};
struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t exec;    // null until the key has been seen twice
  int seen;
  unsigned long long stamp;  // last use (LRU)
};

// One lane = one stream + the scratch that lives between a chunk's kernels.  A context feeds its chunks to the
// lanes in turn, so the bandwidth-bound screen of one chunk runs beside the latency-bound list / finish kernels
// of the previous ones (what several contexts on several streams did in round 1, now inside one dnv_ctx).
struct TrioLane {
  cudaStream_t own;        // the lane's screen stream, low priority (unused when the context runs serially)
  cudaStream_t tail;       // the lane's tail stream (list / finish kernels), HIGH priority: its short kernels are
                           // dispatched into the room beside the resident screens ahead of any queued screen CTAs
  cudaEvent_t ev_screen;   // the lane's latest screen has finished
  cudaEvent_t ev_done;     // everything enqueued on the lane so far
  bool used;               // has work the front stream has not waited for yet
  cudaEvent_t ev_flush;    // recorded on the screen stream after the lane's latest submission (what flush waits for)
  int parity;              // which of the lane's two chain pairs the next chunk uses
  uint32_t* d_list;        // sites the integer screen could not decide (chunk-relative), per-warp sorted segments
  int64_t list_cap;
  uint32_t* d_wcount;      // length of every screen warp's segment of d_list
  dnv_call* d_list_stage;  // list kernel: records between its two phases (local HBM)
  int64_t list_stage_cap;
  unsigned long long* d_tile_status;  // decoupled look-back status per list tile / compaction block (all zero between launches)
  int64_t tile_cap;                   // sites the status array is sized for
  size_t tile_status_bytes;
  GraphEntry graphs[kGraphSlots];     // replayable asynchronous batches of this lane
};
}  // namespace dnv

struct dnv_ctx {
  dnv_params prm;
  int device;
  cudaStream_t stream;       // FRONT stream: orders the context against its caller (own_stream unless dnv_set_stream)
  cudaStream_t own_stream;
  cudaEvent_t ev_k0[64], ev_k1[64], ev_s1[64];  // kernel brackets (DNV_F_TIME_KERNELS): start, end, end of the screen
  int n_timed;
  cudaStream_t copy_stream;  // H2D / D2H of the host API
  cudaEvent_t ev_h2d[2], ev_done[2];
  dnv::Tables h_tab;
  dnv::Tables* d_tab;
  // lanes (n_lanes == 1: serial, everything on the front stream — timed contexts, DNV_LANES=1)
  int n_lanes;
  int lane_next;
  dnv::TrioLane lane[dnv::kMaxLanes];
  cudaEvent_t ev_sub;        // front stream at the time of the current submission (inputs are ready)
  cudaEvent_t ev_chain;      // the list / compaction kernel of the previous chunk of the batch being enqueued
  const unsigned long long* chain_prev;  // pair[1] of the previous chunk of the batch being enqueued
  // scratch of the unscreened path (always lane 0)
  int64_t bits_cap;
  uint32_t* d_bits;        // decision bit per site (unscreened path)
  unsigned long long* d_cstatus;      // look-back status of the compaction kernel (tagged words, see next_epoch)
  size_t cstatus_bytes;
  unsigned long long epoch;           // launch tag of the compaction kernel's status words
  bool screened;           // the last call used the screen
  int64_t sites_screened;  // sites that went through the screen since the statistics were last collected
  bool dense_mode;         // recent batches left > 5 % of their sites undecided: the all-sites fused kernel is faster
  int dense_probe;         // chunks since dense_mode last sent one through the screen to re-measure
  bool capturing;          // a CUDA graph of the current submission is being captured (no allocation allowed)
  bool use_graphs;         // DNV_GRAPHS=0 turns graph replay off
  bool use_graphs_async;   // DNV_GRAPHS_ASYNC=1: also replay asynchronous batches
  bool tail_prio;          // DNV_TAIL_PRIO=1: list / finish kernels on a separate high-priority stream per lane
  bool sub_pending;        // ev_sub was recorded for the current submission (the front stream had work pending)
  unsigned long long graph_clock;
  dnv::GraphEntry sync_graphs[dnv::kGraphSlots];  // replayable synchronous calls (dnv_infer_counts*_device)
  int count_bytes;                    // bytes per site of the current call's counts: 24 (uint16) or 12 (uint8)
  int64_t* d_call_idx;     // global site index of every call, input order (unscreened path)
  int64_t call_idx_cap;
  // [0..15] chain pairs (lane x parity x 2), [16..19] acc, [20] finish CTAs gone (host-collecting launches), [23] always zero,
  // [24 + 2 * lane ..] tickets of the lane
  unsigned long long* d_counters;
  unsigned long long* h_acc;       // pinned [16]: [0..3] acc as last collected, [8..15] status words of the last pileup
  // staging of the host API
  uint16_t* d_stage_counts[2];
  dnv_site_result* d_stage_dense[2];
  int64_t stage_cap;
  dnv_call* d_stage_calls;
  int64_t stage_calls_cap;
  // generic scratch for pileup / mendel
  void* d_scratch;
  size_t scratch_bytes;
  dnv_stats stats;
  // launch configuration, queried once per context (attribute and occupancy calls are not free on the host)
  int sms;                  // multiprocessors of the device
  int screen_per_sm[2];     // resident CTAs per SM of trio_screen_kernel<u16 / u8>, 0 = not queried yet
  int list_per_sm[2][2];    // ... of trio_list_kernel<u16 / u8><BAYES-LRT / MAP>
  int scatter_per_sm;       // ... of pileup_screen_kernel
  int scatter_plain_per_sm; // ... of pileup_scatter_kernel
  uint16_t* d_reads_counts; // fused reads -> calls: per-site counts (only the undecided sites' rows are written)
  int64_t reads_counts_cap;
  // DNV_KTRACE=1: device-side time line, 6 words per chunk (screen, list, finish: first CTA start / last CTA end)
  unsigned long long* d_ktrace;   // [lane][kKtraceSlots][6]
  unsigned int* d_ktrace_seq;     // [lane]
  // asynchronous batches (dnv_infer_counts*_device_async)
  int n_async;                 // batches enqueued since the last dnv_sync
  int64_t async_sites;
};

#define DNV_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      dnv::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));     \
      return DNV_E_CUDA;                                                      \
    }                                                                         \
  } while (0)

// statistics (and the dense-mode decision) of everything enqueued since the previous collect; internal, C linkage
extern "C" int dnvi_finish_stats(dnv_ctx* ctx, int64_t n_sites);

namespace dnv {
constexpr int kCtrAcc = 16, kCtrFinGone = 20, kCtrZero = 23, kCtrWork = 24, kCtrWords = 32;

// reads of a chunk whose sites are given as positions (fused reads -> calls path; device pointers)
struct ReadsArgs {
  const int64_t* read_start; const int64_t* base_off; const char* bases; int64_t n_reads;
  const int64_t* site_pos; long long* meta;  // meta: person offsets, max read length, error word (pileup_prep_kernel)
};

// one chunk of sites resident in HBM through the READ pipeline (trio.cu)
struct ChunkArgs {
  const void* d_counts;
  int64_t site_base;       // added to chunk-relative site indices in the records
  int64_t n;
  dnv_site_result* d_dense;
  dnv_call* d_calls;
  int64_t cap;
  bool first, last;        // first / last chunk of its batch
  bool no_sub_wait;        // the caller has already ordered the lane's stream after the context's ev_sub
  bool tail_alone;         // no other chunk's screen will run beside this chunk's list / finish kernels
  cudaEvent_t wait_ev;     // the chunk's input is ready (null: the context's ev_sub)
  int64_t* count_out;
  const struct ReadsArgs* reads; // non-null: d_counts is scratch, the counts of the chunk's sites come from these reads through the
                           // fused pileup-screen kernel (pileup_scatter.cuh) instead of the screen kernel
  dnv_site_result* dense_host;  // host API: where the chunk's dense records go (copied on the chunk's stream), or null
  cudaEvent_t done_ev;     // recorded when the chunk no longer needs its input (staging ring of the host API), or null
  bool collect_to_host = false;  // the chunk is a whole synchronous call: its finish kernel delivers the statistics (CallChain::host_acc)
};
constexpr int kKtraceSlots = 1024;  // per lane
// which: 0 screen, 1 list, 2 finish; first CTA start -> word 2*which (min), last CTA end -> word 2*which+1 (max)
__device__ __forceinline__ void ktrace_stamp(unsigned long long* base, const unsigned int* seq, int which, bool end) {
  if (base) {
    unsigned long long* slot = base + 6 * (*seq & (kKtraceSlots - 1)) + 2 * which;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (end) atomicMax(slot + 1, t); else atomicMin(slot, t);
  }
}
constexpr int kWcountWords = 8192;  // >= kScreenMaxWarps (trio_screen.cuh)
// Enqueues the chunk on the next lane.
int launch_trio_chunk(dnv_ctx* ctx, const ChunkArgs& a);
// reads -> counts with the read-major scatter kernel (pileup_scatter.cuh, compiled in trio.cu)
int launch_pileup_scatter(dnv_ctx* ctx, cudaStream_t st, const int64_t* d_read_start, const int64_t* d_base_off, const char* d_bases,
                          long long* meta, const int64_t* d_site_pos, int64_t n_sites, int64_t n_reads, uint16_t* d_counts);
int ensure_scratch(dnv_ctx* ctx, size_t bytes);
// the pipeline as the entry points see it (api.cu)
int begin_submit(dnv_ctx* ctx);
int flush_lanes(dnv_ctx* ctx);
int collect(dnv_ctx* ctx);
void reset_call_state(dnv_ctx* ctx);
int ensure_stage_calls(dnv_ctx* ctx, int64_t calls_cap);
int enqueue_device_batch(dnv_ctx* ctx, const void* d_counts, int bps, int64_t n, int64_t site_base, dnv_site_result* d_dense,
                         dnv_call* d_calls, int64_t cap, int64_t* count_out, bool split, bool no_sub_wait, bool collect_to_host = false);
bool will_screen(const dnv_ctx* ctx, bool dense);
bool pick_screen(dnv_ctx* ctx, bool dense);
int enqueue_host_batch(dnv_ctx* ctx, const void* counts, int bps, int64_t n, int64_t site_base, dnv_site_result* dense,
                       dnv_call* d_calls, int64_t cap, int64_t* count_out);
inline bool parallel(const dnv_ctx* ctx) { return ctx->n_lanes > 1; }
inline cudaStream_t lane_stream(const dnv_ctx* ctx, int l) { return parallel(ctx) ? ctx->lane[l].own : ctx->stream; }
inline cudaStream_t lane_tail(const dnv_ctx* ctx, int l) { return parallel(ctx) ? (ctx->lane[l].tail ? ctx->lane[l].tail : ctx->lane[l].own) : ctx->stream; }
// The compaction kernel's look-back status words carry a 30-bit launch tag ([63:34]).  Returns the tag of its next
// launch on this context (never 0: zeroed status memory must not look like a published word); when the counter wraps,
// the status array is re-zeroed in stream order first, so a stale word of the same tag cannot be mistaken for a
// fresh one (2^30 launches apart).  (The list kernel needs no tag: it clears what it used.)
inline int next_epoch(dnv_ctx* ctx, cudaStream_t st, unsigned long long* tag_out) {
  constexpr unsigned long long kMask = (1ull << 30) - 1;
  if (((++ctx->epoch) & kMask) == 0) {
    if (ctx->d_cstatus && ctx->cstatus_bytes) DNV_CUDA(cudaMemsetAsync(ctx->d_cstatus, 0, ctx->cstatus_bytes, st));
    ++ctx->epoch;
  }
  *tag_out = ctx->epoch & kMask;
  return DNV_OK;
}

#ifdef __CUDACC__
// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same with a suspend-time hint (ns): the warp sleeps in hardware until the phase completes or the hint expires,
// instead of re-issuing the test every few cycles and taking issue slots from the warps that have work
__device__ __forceinline__ void mbar_wait_sleepy(unsigned long long* bar, uint32_t parity, uint32_t hint_ns) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#endif
}  // namespace dnv
