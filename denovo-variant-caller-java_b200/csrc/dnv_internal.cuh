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

struct dnv_ctx {
  dnv_params prm;
  int device;
  cudaStream_t stream;       // compute (own_stream unless dnv_set_stream)
  cudaStream_t own_stream;
  cudaEvent_t ev_k0[64], ev_k1[64], ev_s1[64];  // kernel brackets (DNV_F_TIME_KERNELS): start, end, end of the screen
  int n_timed;
  cudaStream_t copy_stream;  // H2D / D2H of the host API
  cudaEvent_t ev_h2d[2], ev_done[2];
  dnv::Tables h_tab;
  dnv::Tables* d_tab;
  // per-chunk scratch (sized for chunk_cap sites)
  int64_t chunk_cap;
  uint32_t* d_bits;        // decision bit per site
  uint32_t* d_list;        // sites the integer screen could not decide (chunk-relative), per-warp sorted segments
  uint32_t* d_wcount;      // length of every screen warp's segment of d_list
  dnv_call* d_list_stage;  // list kernel: records between its two phases (local HBM)
  int64_t list_stage_cap;
  bool screened;           // the last call used the screen
  unsigned long long* d_tile_status;  // decoupled look-back status per compaction block
  unsigned long long epoch;           // launch counter tagging d_tile_status entries (low 30 bits, see next_epoch)
  size_t tile_status_bytes;           // bytes of d_tile_status that the look-backs use
  int calls_slot;                     // counters[] slot holding the running call total (0 or 3)
  int count_bytes;                    // bytes per site of the current call's counts: 24 (uint16) or 12 (uint8)
  int64_t* d_call_idx;     // global site index of every call, input order
  int64_t call_idx_cap;
  unsigned long long* d_counters;  // [0]/[3] running call total (ping-pong per chunk) [4] hard sites [6] block ticket
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
  // asynchronous batches (dnv_infer_counts*_device_async): per-batch counters land in a pinned ring
  unsigned long long* h_ring;  // [64][8]
  int n_async;                 // batches enqueued since the last dnv_sync
  int ring_slot[64];           // which counter slot holds each batch's call total
  int64_t async_sites;
  int64_t* count_out;          // where the details kernel stores the batch's call total (async batches), else null
};

#define DNV_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      dnv::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));     \
      return DNV_E_CUDA;                                                      \
    }                                                                         \
  } while (0)

namespace dnv {
// kernels / launchers (defined in the .cu files)
int launch_trio_chunk(dnv_ctx* ctx, const void* d_counts, int64_t site_base, int64_t n,
                      dnv_site_result* d_dense, dnv_call* d_calls, int64_t cap);
int ensure_chunk_scratch(dnv_ctx* ctx, int64_t n);
int ensure_scratch(dnv_ctx* ctx, size_t bytes);
// The look-back status words carry a 30-bit launch tag ([63:34]).  Returns the tag of the next launch on this
// context (never 0: zeroed status memory must not look like a published word); when the counter wraps, the
// status array is re-zeroed in stream order first, so a stale word of the same tag cannot be mistaken for a
// fresh one (2^30 launches apart).
inline int next_epoch(dnv_ctx* ctx, unsigned long long* tag_out) {
  constexpr unsigned long long kMask = (1ull << 30) - 1;
  if (((++ctx->epoch) & kMask) == 0) {
    if (ctx->d_tile_status && ctx->tile_status_bytes)
      DNV_CUDA(cudaMemsetAsync(ctx->d_tile_status, 0, ctx->tile_status_bytes, ctx->stream));
    ++ctx->epoch;
  }
  *tag_out = ctx->epoch & kMask;
  return DNV_OK;
}
}  // namespace dnv
