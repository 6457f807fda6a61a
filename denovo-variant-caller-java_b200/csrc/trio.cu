// trio.cu — READ-mode pipeline on one chunk of sites resident in HBM.
//
// Calls only (the BASELINE workload; BAYES / LRT, and MAP through the same screen) — three kernels:
//   S   trio_screen_kernel (trio_screen.cuh): exact integer screen of EVERY site at HBM speed; the
//         ~1 % it cannot prove non-de-novo leave as per-warp sorted segments
//   L   trio_list_kernel (trio_list.cuh): fused FP64 likelihood + posterior on those sites,
//         order-preserving compaction (decoupled look-back), records staged in input order
//   F   trio_calls_finish_kernel: one thread per call completes the record (exact MAP sweep)
//
// Dense output, DNV_F_NO_SCREEN, DNV_F_FORCE_DIRECT, parameters outside the screen's validity — three kernels:
//   K1  trio_fast_kernel (trio_fast.cuh) or trio_direct_all_kernel: counts -> decision bit per site
//         (+ dense record); hard sites are re-evaluated inside the kernel by the whole warp with the
//         exact direct form
//   K2  compact_calls_kernel: single-pass order-preserving compaction of the decision bits
//   K3  call_details_kernel: one thread per called site -> dnv_call
//
// Replaces the per-site task of ReadCaller.runBayesDenovoInference
// (M/ReadCaller.java:111-130) -> BayesInfer.infer (M/BayesInfer.java:50-69).
// Chunks are fed to the context's lanes in turn (dnv_internal.cuh: TrioLane), so the screen of one chunk runs
// beside the list / finish kernels of the previous ones.  Nothing here synchronises with the host.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "trio_exact.cuh"
#include "trio_fast.cuh"
#include "trio_list.cuh"
#include "trio_screen.cuh"
#include "pileup_scatter.cuh"

namespace dnv {

constexpr int kExactThreads = 128;

// ---------------------------------------------------------------- K1 (exact, all sites)
__global__ void __launch_bounds__(kExactThreads)
trio_direct_all_kernel(const Tables* __restrict__ tab, const void* __restrict__ counts, int bps,
                       int64_t n, uint32_t* __restrict__ bits, dnv_site_result* __restrict__ dense,
                       unsigned long long* __restrict__ acc) {
  __shared__ double s_rel[1000];
  __shared__ uint32_t s_isden[32];
  stage_exact_tables(tab, s_rel, s_isden);
  __syncthreads();
  const double hm = tab->hm, ht = tab->ht, mm = tab->mm, thr = tab->lrt_threshold;
  const int method = tab->method;
  // whole warps iterate together so the ballot below covers 32 consecutive sites
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t site = base + threadIdx.x;
    bool dn = false;
    if (site < n) {
      uint16_t c[12];
      load_counts(counts, bps, site, c);
      ExactOut r;
      exact_site(c, s_rel, s_isden, hm, ht, mm, method, thr, &r);
      dn = r.is_denovo;
      if (dense) store_dense(dense, site, r, 1u);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, dn);
    if ((threadIdx.x & 31) == 0 && site < n) {
      bits[site >> 5] = word;
      atomicAdd(&acc[2], (unsigned long long)min((int64_t)32, n - site));  // every site took the exact path
    }
  }
}

// ---------------------------------------------------------------- K2 compaction
// Single-pass, order-preserving compaction of the decision bits (decoupled look-back over
// blocks of 8192 sites): the global site index of every called site, in INPUT order.
// status word: [63:34] epoch of this launch, [33:32] 1 = block aggregate, 2 = inclusive prefix, [31:0] value.
// ch (CallChain): *before = calls before this chunk, pair[0..1] <- calls before / after it, work[0] block ticket.
constexpr int kCompactThreads = 256;
static_assert(kWcountWords >= kScreenMaxWarps, "wcount buffers hold one word per screen warp");

__global__ void __launch_bounds__(kCompactThreads)
compact_calls_kernel(int64_t site_base, const uint32_t* __restrict__ bits, int64_t n_words,
                     unsigned long long* status, unsigned long long epoch, const CallChain ch,
                     int64_t* __restrict__ call_idx, int64_t call_idx_cap) {
  __shared__ int s_warp_sum[kCompactThreads / 32];
  __shared__ unsigned int s_bid;
  __shared__ unsigned long long s_excl;
  if (threadIdx.x == 0) s_bid = (unsigned int)atomicAdd(&ch.work[0], 1ull);  // blocks ordered by start time
  __syncthreads();
  const unsigned int b = s_bid;
  if (b == gridDim.x - 1 && threadIdx.x == 0) ch.work[0] = 0;  // every ticket is taken: reset for the next launch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long calls_before = ch.before ? *ch.before : 0ull;
  const int64_t w = (int64_t)b * kCompactThreads + threadIdx.x;
  uint32_t word = (w < n_words) ? bits[w] : 0u;
  const int cnt = __popc(word);
  int x = cnt;  // inclusive scan inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp_sum[warp] = x;
  __syncthreads();
  int warp_off = 0, total = 0;
#pragma unroll
  for (int i = 0; i < kCompactThreads / 32; i++) {
    if (i < warp) warp_off += s_warp_sum[i];
    total += s_warp_sum[i];
  }
  if (warp == 0) {  // warp-parallel decoupled look-back: 32 predecessors per probe
    const unsigned long long tag = epoch << 34;
    unsigned long long excl = 0;
    if (b > 0) {
      if (lane == 0) atomicExch(&status[b], tag | (1ull << 32) | (unsigned long long)total);
      int hi = (int)b - 1;  // nearest predecessor not yet accounted for
      for (;;) {
        const int p = hi - lane;
        unsigned long long v = tag | (2ull << 32);  // lanes before block 0 act as an empty inclusive prefix
        if (p >= 0) {
          do { v = *(volatile unsigned long long*)&status[p]; } while ((v >> 34) != epoch);
        }
        const uint32_t is_prefix = __ballot_sync(0xffffffffu, ((v >> 32) & 3ull) == 2ull);
        const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;  // nearest inclusive prefix in this window
        unsigned long long val = lane <= stop ? (v & 0xffffffffull) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (is_prefix) break;
        hi -= 32;
      }
    }
    if (lane == 0) {
      atomicExch(&status[b], tag | (2ull << 32) | (excl + (unsigned long long)total));
      s_excl = excl;
      if (b == gridDim.x - 1) {
        const unsigned long long all = calls_before + excl + (unsigned long long)total;
        ch.pair[0] = calls_before;
        ch.pair[1] = all;
        if (ch.last) atomicAdd(&ch.acc[0], all);
      }
    }
  }
  if (total == 0) return;  // block-uniform
  __syncthreads();
  int64_t pos = (int64_t)calls_before + (int64_t)s_excl + warp_off + x - cnt;
  while (word) {
    const int bit = __ffs(word) - 1;
    word &= word - 1;
    if (pos < call_idx_cap) call_idx[pos] = site_base + w * 32 + bit;
    pos++;
  }
}

// ---------------------------------------------------------------- K3 call records
// One THREAD per called site (calls [pair[0], pair[1]) of this chunk).
// Values: the factored form (power tables read through L1) unless the site is hard, then the
// exact direct form in the reference's own order.  maxGenoType / max_ll: always the exact
// (pruned) log-domain argmax.
constexpr int kDetailThreads = 64;

__global__ void __launch_bounds__(kDetailThreads)
call_details_kernel(const Tables* __restrict__ tab, const void* __restrict__ counts, int bps, int64_t site_base,
                    const int64_t* __restrict__ call_idx, const CallChain ch, dnv_call* __restrict__ calls, int64_t cap,
                    const FastConsts k) {
  int64_t* count_out = ch.count_out;
  const unsigned long long lo = ch.pair[0];
  unsigned long long hi = ch.pair[1];
  // asynchronous batches: the running call total goes where the caller wants it (pinned host memory,
  // or a peer GPU's HBM next to the records) — a plain store, ordered by the stream like the records
  if (count_out && blockIdx.x == 0 && threadIdx.x == 0) *count_out = (int64_t)hi;
  if (hi > (unsigned long long)cap) hi = (unsigned long long)cap;
  if (lo + (unsigned long long)blockIdx.x * kDetailThreads >= hi) return;
  __shared__ double s_rel[1000];
  __shared__ uint32_t s_isden[32];
  __shared__ double s_rmax[110];
  stage_exact_tables(tab, s_rel, s_isden);
  stage_map_bounds(tab, s_rmax);
  __syncthreads();
  const GmemPow pw{tab};
  for (unsigned long long p = lo + (unsigned long long)blockIdx.x * kDetailThreads + threadIdx.x; p < hi;
       p += (unsigned long long)gridDim.x * kDetailThreads) {
    const int64_t gsite = call_idx[p];
    uint16_t c[12];
    load_counts(counts, bps, gsite - site_base, c);
    int ci[12];
#pragma unroll
    for (int q = 0; q < 12; q++) ci[q] = c[q];
    LinOut lo_;
    if (k.lin_ok && k.method != DNV_METHOD_MAP) {
      linear_site(pw, ci, k.k_men, k.kd9, k.kd8, k.kd7, k.kd6, k.thr, k.mm2, k.method, k.lrt_threshold, lo_);
    } else if (k.lin_ok) {  // MAP: the decision never needed the sums; the values still come from them
      linear_site(pw, ci, k.k_men, k.kd9, k.kd8, k.kd7, k.kd6, 1.0, k.mm2, DNV_METHOD_BAYES, 1.0, lo_);
      lo_.hard = lo_.zd == 2 || lo_.zm == 2;
    } else {
      lo_.hard = true;
    }
    dnv_call o;
    if (lo_.hard) {
      ExactOut r;
      exact_site(c, s_rel, s_isden, k.hm, k.ht, k.mm, k.method, k.lrt_threshold, &r);
      o.dad = (uint8_t)r.dad; o.mom = (uint8_t)r.mom; o.child = (uint8_t)r.child;
      o.max_ll = r.max_ll; o.p = r.p; o.lr = r.lr; o.log_men = r.log_men; o.log_den = r.log_den;
      o.flags = 1;
    } else {
      double Ld[10], Lm[10], Lc[10];
      exact_read_loglik(c + 0, k.hm, k.ht, k.mm, Ld);
      exact_read_loglik(c + 4, k.hm, k.ht, k.mm, Lm);
      exact_read_loglik(c + 8, k.hm, k.ht, k.mm, Lc);
      int best;
      double max_ll;
      exact_map_pruned(Ld, Lm, Lc, s_rel, s_rmax, &best, &max_ll);
      o.dad = (uint8_t)(best / 100); o.mom = (uint8_t)((best / 10) % 10); o.child = (uint8_t)(best % 10);
      o.max_ll = max_ll;
      linear_values(lo_, o.p, o.lr, o.log_men, o.log_den);
      o.flags = 0;
    }
    o.site = gsite;
#pragma unroll
    for (int q = 0; q < 12; q++) o.counts[q] = c[q];
    o.is_denovo = 1;  // the decision that put the site on the list stands
    calls[p] = o;
  }
}

__global__ void ktrace_bump_kernel(unsigned int* seq) { *seq += 1; }

// ---------------------------------------------------------------- launcher
// look-back status of the lane (list tiles / compaction blocks), zero between launches; `st` = the stream the next
// kernel that uses it runs on.  Decision bits of the unscreened path live in the context (lane 0 only).
static int ensure_lane_status(dnv_ctx* ctx, TrioLane& ln, int64_t n, cudaStream_t st) {
  if (n <= ln.tile_cap) return DNV_OK;
  if (ctx->capturing) { set_error("internal: scratch growth during graph capture"); return DNV_E_INVALID; }
  cudaFree(ln.d_tile_status);  // (cudaFree synchronises the device)
  ln.d_tile_status = nullptr;
  ln.tile_cap = 0;
  const int64_t n_status = (n + kListThreads - 1) / kListThreads + 1;  // one entry per list tile (kListThreads undecided sites)
  const int64_t n_status_alloc = n_status;
  DNV_CUDA(cudaMalloc(&ln.d_tile_status, (size_t)n_status_alloc * 8));
  DNV_CUDA(cudaMemsetAsync(ln.d_tile_status, 0, (size_t)n_status * 8, st));
  ln.tile_status_bytes = (size_t)n_status * 8;
  ln.tile_cap = n;
  return DNV_OK;
}

// unscreened path: decision bits and the compaction kernel's (tagged) look-back status
static int ensure_bits(dnv_ctx* ctx, int64_t n, cudaStream_t st) {
  if (n <= ctx->bits_cap) return DNV_OK;
  cudaFree(ctx->d_bits); cudaFree(ctx->d_cstatus);
  ctx->d_bits = nullptr; ctx->d_cstatus = nullptr;
  ctx->bits_cap = 0;
  const int64_t n_words = (n + 31) / 32;
  const int64_t n_status = (n_words + kCompactThreads - 1) / kCompactThreads + 1;
  DNV_CUDA(cudaMalloc(&ctx->d_bits, (size_t)(n_words + 1) * 4));
  DNV_CUDA(cudaMalloc(&ctx->d_cstatus, (size_t)n_status * 8));
  DNV_CUDA(cudaMemsetAsync(ctx->d_cstatus, 0, (size_t)n_status * 8, st));
  ctx->cstatus_bytes = (size_t)n_status * 8;
  ctx->bits_cap = n;
  return DNV_OK;
}

static int ensure_lane_list(dnv_ctx* ctx, TrioLane& ln, int64_t n) {
  if (n <= ln.list_cap) return DNV_OK;
  if (ctx->capturing) { set_error("internal: scratch growth during graph capture"); return DNV_E_INVALID; }
  cudaFree(ln.d_list);
  ln.d_list = nullptr;
  ln.list_cap = 0;
  DNV_CUDA(cudaMalloc(&ln.d_list, ((size_t)n + kSC_Tile) * 4));  // a segment starts at its owner's first site: + one tile of slack
  ln.list_cap = n;
  return DNV_OK;
}

static int ensure_call_idx(dnv_ctx* ctx, int64_t cap) {
  if (cap <= ctx->call_idx_cap) return DNV_OK;
  cudaFree(ctx->d_call_idx);
  ctx->d_call_idx = nullptr;
  ctx->call_idx_cap = 0;
  DNV_CUDA(cudaMalloc(&ctx->d_call_idx, (size_t)(cap > 0 ? cap : 1) * 8));
  ctx->call_idx_cap = cap;
  return DNV_OK;
}

int launch_pileup_scatter(dnv_ctx* ctx, cudaStream_t st, const int64_t* d_read_start, const int64_t* d_base_off, const char* d_bases,
                          long long* meta, const int64_t* d_site_pos, int64_t n_sites, int64_t n_reads, uint16_t* d_counts) {
  const size_t smem = sizeof(ScatterShared);
  if (!ctx->scatter_plain_per_sm) {
    DNV_CUDA(cudaFuncSetAttribute(pileup_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int q = 1;
    DNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, pileup_scatter_kernel, kSC_Threads, smem));
    ctx->scatter_plain_per_sm = std::max(q, 1);
  }
  const int grid = scatter_grid(ctx->sms, (n_sites + kSC_Tile - 1) / kSC_Tile, ctx->scatter_plain_per_sm, 1 << 20);
  pileup_scatter_kernel<<<grid, kSC_Threads, smem, st>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites, n_reads,
                                                        d_counts);
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

static ScreenConsts make_screen_consts(const dnv_ctx* ctx) {
  ScreenConsts sk;
  sk.c_hom = ctx->h_tab.c_hom_fx;
  sk.c_het = ctx->h_tab.c_het_fx;
  for (int x = 0; x < 4; x++) sk.w[x] = ctx->h_tab.scr_w[x];
  sk.shift32 = 32;
  static const bool no_l3 = std::getenv("DNV_NO_LEVEL3") != nullptr;  // A/B knob
  sk.l3_lambda = no_l3 ? 0 : ctx->h_tab.l3_lambda_fx;
  sk.l3_max_depth = ctx->h_tab.l3_max_depth;
  sk.ktrace = nullptr;
  sk.ktrace_seq = nullptr;
  return sk;
}

// One chunk (n < 2^31) on the next lane.  The caller has recorded ctx->ev_sub on the front stream (parallel
// contexts) and reads the totals through collect() at the end.
int launch_trio_chunk(dnv_ctx* ctx, const ChunkArgs& a) {
  const int64_t n = a.n;
  if (n <= 0) return DNV_OK;
  const bool direct = (ctx->prm.flags & DNV_F_FORCE_DIRECT) != 0;
  const bool timed = (ctx->prm.flags & DNV_F_TIME_KERNELS) != 0 && ctx->n_timed < 64;
  const bool map = ctx->prm.method == DNV_METHOD_MAP;
  const bool screen = ctx->capturing ? will_screen(ctx, a.d_dense != nullptr) : pick_screen(ctx, a.d_dense != nullptr);
  // the unscreened path keeps its decision bits in shared scratch: always lane 0 (its chunks serialise)
  const int li = screen ? ctx->lane_next : 0;
  if (screen) ctx->lane_next = (ctx->lane_next + 1) % ctx->n_lanes;
  TrioLane& ln = ctx->lane[li];
  if (a.collect_to_host && !screen) { set_error("internal: host-collecting chunk on the unscreened path"); return DNV_E_INVALID; }
  // st: where the chunk starts (the screen, or the whole unscreened path); tl: where its tail runs
  cudaStream_t st = lane_stream(ctx, li), tl = screen ? lane_tail(ctx, li) : st;
  // (the lane's own stream has already waited for the lane's previous tail — see the end of this function — so
  // the lane's scratch is free again)
  if (a.wait_ev) DNV_CUDA(cudaStreamWaitEvent(st, a.wait_ev, 0));
  else if (parallel(ctx) && !a.no_sub_wait && ctx->sub_pending) DNV_CUDA(cudaStreamWaitEvent(st, ctx->ev_sub, 0));
  int rc = screen ? ensure_lane_status(ctx, ln, n, tl) : ensure_bits(ctx, n, st);
  if (rc) return rc;
  const int sms = ctx->sms;

  CallChain ch;
  ch.before = a.first ? nullptr : ctx->chain_prev;
  ch.pair = ctx->d_counters + (li * 2 + ln.parity) * 2;
  ln.parity ^= 1;
  ch.acc = ctx->d_counters + kCtrAcc;
  ch.work = ctx->d_counters + kCtrWork + 2 * li;
  ch.count_out = a.count_out;
  ch.last = a.last ? 1 : 0;
  ch.host_acc = a.collect_to_host ? ctx->h_acc : nullptr;
  ch.fin_gone = ctx->d_counters + kCtrFinGone;
  ch.ktrace = ctx->d_ktrace ? ctx->d_ktrace + (size_t)li * kKtraceSlots * 6 : nullptr;
  ch.ktrace_seq = ctx->d_ktrace ? ctx->d_ktrace_seq + li : nullptr;
  ctx->chain_prev = ch.pair + 1;

  if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_k0[ctx->n_timed], st));
  if (screen) {
    // calls only: exact integer screen at HBM speed, then the list kernel for the sites it could not decide
    // (fused FP64 posterior, ordered compaction) and the per-call finish kernel
    rc = ensure_lane_list(ctx, ln, n);
    if (rc) return rc;
    ScreenConsts sk = make_screen_consts(ctx);
    sk.ktrace = ch.ktrace;
    sk.ktrace_seq = ch.ktrace_seq;
    const bool u8 = ctx->count_bytes == 12;
    int sgrid = 0, n_segments = 0, seg_sites = kScreenSites;
    if (a.reads) {
      // fused: pileup of the chunk's reads in shared memory, screened there; counts written for undecided sites only
      auto kern = pileup_screen_kernel;
      const size_t smem = sizeof(ScatterShared);
      if (!ctx->scatter_per_sm) {
        DNV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int q = 1;
        DNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kSC_Threads, smem));
        ctx->scatter_per_sm = std::max(q, 1);
      }
      sgrid = scatter_grid(sms, (n + kSC_Tile - 1) / kSC_Tile, ctx->scatter_per_sm, kScreenMaxWarps);
      kern<<<sgrid, kSC_Threads, smem, st>>>(a.reads->read_start, a.reads->base_off, a.reads->bases, a.reads->meta,
                                            a.reads->site_pos, n, a.reads->n_reads, (uint16_t*)const_cast<void*>(a.d_counts), sk,
                                            ln.d_list, ln.d_wcount);
      DNV_CUDA(cudaGetLastError());
      n_segments = sgrid;
      seg_sites = kSC_Tile;
    } else {
      rc = u8 ? launch_screen<true>(ctx, ln, st, a.d_counts, n, sk, &sgrid) : launch_screen<false>(ctx, ln, st, a.d_counts, n, sk, &sgrid);
      if (rc) return rc;
      n_segments = sgrid * kScreenWarps;
    }
    if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_s1[ctx->n_timed], st));
    if (tl != st) {
      DNV_CUDA(cudaEventRecord(ln.ev_screen, st));
      DNV_CUDA(cudaStreamWaitEvent(tl, ln.ev_screen, 0));
    }
    const bool beside = parallel(ctx) && !a.tail_alone;
    if (map)
      rc = u8 ? launch_trio_list<true, true>(ctx, ln, tl, a.d_counts, n, a.site_base, n_segments, seg_sites, ch, a.d_calls, a.cap, beside)
              : launch_trio_list<false, true>(ctx, ln, tl, a.d_counts, n, a.site_base, n_segments, seg_sites, ch, a.d_calls, a.cap, beside);
    else
      rc = u8 ? launch_trio_list<true, false>(ctx, ln, tl, a.d_counts, n, a.site_base, n_segments, seg_sites, ch, a.d_calls, a.cap, beside)
              : launch_trio_list<false, false>(ctx, ln, tl, a.d_counts, n, a.site_base, n_segments, seg_sites, ch, a.d_calls, a.cap, beside);
    if (rc) return rc;
    if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_k1[ctx->n_timed++], tl));
    ctx->screened = true;
    if (!ctx->capturing) ctx->sites_screened += n;
  } else {
    rc = ensure_call_idx(ctx, a.cap);
    if (rc) return rc;
    const int64_t n_words = (n + 31) / 32;
    const int n_blocks = (int)((n_words + kCompactThreads - 1) / kCompactThreads);
    if (direct) {
      const int64_t want = (n + kExactThreads - 1) / kExactThreads;
      const int grid = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
      trio_direct_all_kernel<<<grid, kExactThreads, 0, st>>>(ctx->d_tab, a.d_counts, ctx->count_bytes, n, ctx->d_bits,
                                                            a.d_dense, ch.acc);
    } else {
      rc = launch_trio_fast(ctx, st, a.d_counts, n, a.d_dense, ch.acc);
      if (rc) return rc;
    }
    if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_k1[ctx->n_timed++], st));
    unsigned long long tag = 0;
    rc = next_epoch(ctx, st, &tag);
    if (rc) return rc;
    compact_calls_kernel<<<n_blocks, kCompactThreads, 0, st>>>(a.site_base, ctx->d_bits, n_words, ctx->d_cstatus, tag, ch,
                                                              ctx->d_call_idx, ctx->call_idx_cap);
    call_details_kernel<<<sms * 4, kDetailThreads, 0, st>>>(ctx->d_tab, a.d_counts, ctx->count_bytes, a.site_base,
                                                           ctx->d_call_idx, ch, a.d_calls, a.cap,
                                                           make_fast_consts(ctx->h_tab));
  }
  ctx->stats.n_launches += 3;
  DNV_CUDA(cudaGetLastError());
  if (ctx->d_ktrace) ktrace_bump_kernel<<<1, 1, 0, tl>>>(ctx->d_ktrace_seq + li);
  if (a.dense_host && a.d_dense)
    DNV_CUDA(cudaMemcpyAsync(a.dense_host, a.d_dense, (size_t)n * sizeof(dnv_site_result), cudaMemcpyDeviceToHost, tl));
  if (a.done_ev) DNV_CUDA(cudaEventRecord(a.done_ev, tl));
  if (parallel(ctx)) {
    // the lane's screen stream rejoins its tail stream: the next chunk on this lane starts after this one has
    // finished with the lane's scratch, and one event on the screen stream (ev_flush) covers the whole lane
    if (tl != st) {
      DNV_CUDA(cudaEventRecord(ln.ev_done, tl));
      DNV_CUDA(cudaStreamWaitEvent(st, ln.ev_done, 0));
    }
    ln.used = true;
  }
  return DNV_OK;
}

}  // namespace dnv

// Calibration (bench.py's roofline): `reps` back-to-back launches of the exact integer screen ALONE over one chunk of
// counts resident in HBM, on the context's stream, bracketed by one pair of CUDA events — the kernel's average
// duration with the launch latency of an idle stream amortised, as kernels are usually timed.  Uses lane 0's scratch;
// the context must be idle (pending asynchronous batches are collected first).
extern "C" int dnv_measure_screen(dnv_ctx* ctx, const void* d_counts, int count_bytes, int64_t n, int reps, double* ms_per_launch) {
  using namespace dnv;
  if (!ctx || !d_counts || !ms_per_launch || n <= 0 || n > (1ll << 27) || reps < 1 || (count_bytes != 24 && count_bytes != 12)) {
    set_error("dnv_measure_screen: invalid argument");
    return DNV_E_INVALID;
  }
  if (!will_screen(ctx, false)) { set_error("dnv_measure_screen: the screen does not apply to this context's parameters"); return DNV_E_INVALID; }
  if (reinterpret_cast<uintptr_t>(d_counts) & 15) { set_error("dnv_measure_screen: counts must be 16-byte aligned"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(ctx->device));
  int rc = dnv_sync(ctx);
  if (rc) return rc;
  TrioLane& ln = ctx->lane[0];
  rc = ensure_lane_list(ctx, ln, n);
  if (rc) return rc;
  const ScreenConsts sk = make_screen_consts(ctx);
  cudaEvent_t e0, e1;
  DNV_CUDA(cudaEventCreate(&e0));
  DNV_CUDA(cudaEventCreate(&e1));
  int sgrid = 0;
  for (int pass = 0; pass < 2 && rc == DNV_OK; pass++) {  // pass 0 warms up (function attributes, clocks)
    if (pass == 1) cudaEventRecord(e0, ctx->stream);
    for (int r = 0; r < (pass ? reps : 3) && rc == DNV_OK; r++)
      rc = count_bytes == 12 ? launch_screen<true>(ctx, ln, ctx->stream, d_counts, n, sk, &sgrid)
                             : launch_screen<false>(ctx, ln, ctx->stream, d_counts, n, sk, &sgrid);
    if (pass == 1) cudaEventRecord(e1, ctx->stream);
  }
  float ms = 0;
  if (rc == DNV_OK && (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess)) {
    set_error(std::string("dnv_measure_screen: ") + cudaGetErrorString(cudaGetLastError()));
    rc = DNV_E_CUDA;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc == DNV_OK) *ms_per_launch = (double)ms / reps;
  return rc;
}

// DNV_KTRACE=1: the device-side time line of the chunks run so far, lane by lane: out[lane][kKtraceSlots][6]
// (screen start / end, list start / end, finish start / end in %globaltimer nanoseconds; a lane's slots wrap) and
// seq[lane] = chunks the lane has run.  Returns the number of lanes.
extern "C" int dnv_debug_ktrace(dnv_ctx* ctx, unsigned long long* out, unsigned int* seq) {
  if (!ctx || !ctx->d_ktrace) return 0;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (cudaMemcpy(out, ctx->d_ktrace, (size_t)ctx->n_lanes * dnv::kKtraceSlots * 48, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  if (cudaMemcpy(seq, ctx->d_ktrace_seq, (size_t)ctx->n_lanes * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return ctx->n_lanes;
}

