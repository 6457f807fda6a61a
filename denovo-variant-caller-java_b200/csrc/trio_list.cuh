// trio_list.cuh — the tail of the screened READ pipeline: the list kernel and the finish kernel.
//
// Input: the sites the integer screen (trio_screen.cuh) could not decide, as per-warp sorted
// segments of `list` plus their lengths wcount[] — concatenated in warp order they are the
// undecided sites in ascending order (about 1 % of all sites).  For every such site:
//   1. the fused likelihood + posterior in the factored linear form (trio_fast.cuh:
//      linear_site; power tables read through L1 — with ~2 sites per thread, staging them in
//      shared memory would cost more than it saves); hard sites (threshold-ambiguous,
//      underflow gray zone, deeper than the tables) are re-evaluated by the whole warp with
//      the exact direct form (trio_exact.cuh: exact_site_warp);
//   2. an order-preserving compaction of the called sites across the grid (tiles of 128 list
//      items handed out by ticket, decoupled look-back over the tile totals), so that the
//      thread that decided a site stages its dnv_call record at its final, input-ordered index
//      (local HBM): site, counts, p / LR / log-likelihoods from the linear form (everything,
//      from the exact evaluation, for hard sites);
//   3. trio_calls_finish_kernel, ONE THREAD PER CALL, completes the record with the exact pruned
//      log-domain sweep (MAP triple, max_ll) and writes it to the caller's buffer.
//      (MAP method, kMap: the roles swap — the sweep is the decision of the list kernel and the
//      finish kernel adds p / LR / log-likelihoods from the linear form.)
//      Called sites are ~1 in 5 of the list, so doing this densely instead of inside the
//      divergent decision loop takes the sweep off every warp's critical path; as a second,
//      small kernel it also frees most SMs early for whatever runs next (BatchPipeline).
// No decision bitmap, no separate compaction kernel, no sort, no grid-wide barrier.
//
// Replaces, for the ~1 % of sites that need it, BayesInfer.infer -> performInference ->
// isDenovo (M/BayesInfer.java:50-69, M/DenovoBayesNet.java:247-289, M/DenovoUtil.java:261-288)
// and the `if (result.isDenovo()) writeCalls(...)` of M/ReadCaller.java:122-129.
#pragma once
#include <stdlib.h>
#include "dnv_internal.cuh"
#include "trio_exact.cuh"
#include "trio_fast.cuh"
#include "trio_screen.cuh"

namespace dnv {

// The tail kernels are built to run BESIDE the screen of the next chunk.  An SM's register file is four
// sub-partitions of 16384 registers; two resident screen CTAs (trio_screen.cuh) put four 64-register warps on
// each, which leaves 8192 registers per sub-partition: TWO tail warps if — and only if — a tail thread needs at
// most 128 registers.  Hence __launch_bounds__(128, 4) on both kernels (ptxas keeps them at 128 registers; the
// few spills go to L1) and CTAs of four warps, one per sub-partition: two tail CTAs per SM fit next to the
// screen, and the shared-memory request of the screen is padded so that a third screen CTA cannot take their
// room.  Measured (benchmarks/ktrace.py): the list kernel of chunk k then starts ~2 us after its screen ends and
// runs while the screen of chunk k+1 streams.
constexpr int kListThreads = 128;
constexpr int kListWarps = kListThreads / 32;
constexpr int kFinishThreads = 128;
constexpr int kFinishWarps = kFinishThreads / 32;
constexpr int kTailCtasPerSm = 4;  // launch bound of both tail kernels: 128 registers per thread

// A/B switches (benchmarks/gpu/ab.sh builds the variants): staged power tables, raw sums finished by the finish
// kernel, programmatic dependent launch of the finish kernel
#ifndef DNV_LIST_SMEM_POW
#define DNV_LIST_SMEM_POW 1
#endif
#ifndef DNV_RAW_SUMS
#define DNV_RAW_SUMS 1
#endif
#ifndef DNV_FINISH_PDL
#define DNV_FINISH_PDL 1
#endif
constexpr int kListPow = DNV_LIST_SMEM_POW ? 256 : 0;  // power-table entries staged in shared memory (depths below 256; deeper reads L1 / L2)

// BAYES / LRT: the head of the power tables (the dependent table look-ups of person_linear then cost a shared-memory
// round trip instead of an L2 one: beside the screen the SM's L1 is almost all carved out as shared memory); the
// exact tables stay in global memory (hard sites are rare).  MAP: the exact tables, no power tables.
template <bool kMap>
struct ListShared {
  double rel[kMap ? 1000 : 2];    // MAP: getRelationshipLogLikelihood (the decision sweep reads it for every site)
  double rmax[kMap ? 110 : 2];    // MAP: exact pruning bounds
  double pw_inv[kMap ? 1 : 2][(kMap || !kListPow) ? 2 : kListPow];
  double pw_m[kMap ? 1 : 2][(kMap || !kListPow) ? 2 : kListPow];
  int32_t pw_e[kMap ? 1 : 2][(kMap || !kListPow) ? 4 : kListPow];
  double warp_L[kListWarps][32];  // exact path: per-warp read log-likelihoods
  uint32_t isden[32];             // exact path: isDenovoMap
  uint32_t woff[kScreenMaxWarps + 1];  // exclusive scan of wcount[]
  uint32_t part[kListWarps];
  int warp_sum[kListWarps];
  unsigned int ticket[2];  // next tile, handed from thread 0 to the CTA (double buffered by trip parity)
  unsigned long long excl;
};

// power-table accessor of the list kernel: shared memory below kListPow, global memory (L1 / L2) above
template <bool kMap>
struct ListPow {
  const ListShared<kMap>& s;
  const Tables* t;
  __device__ __forceinline__ double inv(int j, int i) const { return i < kListPow ? s.pw_inv[j][i] : __ldg(&t->pw_inv[j][i]); }
  __device__ __forceinline__ double mant(int j, int i) const { return i < kListPow ? s.pw_m[j][i] : __ldg(&t->pw_m[j][i]); }
  __device__ __forceinline__ int expo(int j, int i) const { return i < kListPow ? s.pw_e[j][i] : __ldg(&t->pw_e[j][i]); }
};

// ch (CallChain, dnv_internal.cuh): *before = calls of this batch before this chunk, *after <- calls up to and
//           including it (pair[0..1], private to the chunk), acc[1] += undecided sites, acc[2] += hard sites, work[0] tile ticket, work[1] CTAs gone
//           (the last CTA to leave resets both for the next launch).
// status word: [33:32] 1 = tile aggregate, 2 = inclusive prefix (0 = not published yet), [31:0] value.  The array is
// all zero between launches: the last CTA to leave clears the entries the launch used, so a launch needs no tag and
// can be replayed from a CUDA graph with identical arguments.
template <bool kU8, bool kMap>
__global__ void __launch_bounds__(kListThreads, kTailCtasPerSm)
trio_list_kernel(const Tables* __restrict__ tab, const FastConsts k, const unsigned char* __restrict__ counts,
                 int64_t n, int64_t site_base, const uint32_t* __restrict__ list, const uint32_t* __restrict__ wcount,
                 uint32_t n_warps, uint32_t range_base, uint32_t range_rem, uint32_t seg_sites, unsigned long long* status,
                 const CallChain ch, int64_t cap, dnv_call* __restrict__ stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ListShared<kMap>& s = *reinterpret_cast<ListShared<kMap>*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kBps = kU8 ? 12 : 24;
  if (threadIdx.x == 0) ktrace_stamp(ch.ktrace, ch.ktrace_seq, 1, false);
  // the first tile is claimed before the prologue: by the time any CTA asks for its second tile, every
  // running CTA holds a first one, so tile order stays close to processing order (short look-back waits)
  if (threadIdx.x == 0) s.ticket[0] = (unsigned int)atomicAdd(&ch.work[0], 1ull);

  // Prologue: every global load is issued before the first one is consumed (one round trip, not a
  // dependent chain): wcount[] (coalesced, kScreenMaxWarps / kListThreads words per thread) and this
  // thread's slice of the exact tables; then the block-wide scan of wcount[] in shared memory.
  constexpr int kWLoads = (kScreenMaxWarps + kListThreads - 1) / kListThreads;
  uint32_t wv[kWLoads];
#pragma unroll
  for (int i = 0; i < kWLoads; i++) {
    const uint32_t j = i * kListThreads + threadIdx.x;
    wv[i] = j < n_warps ? __ldg(wcount + j) : 0u;
  }
  // MAP: rel[1000] as 500 double2 + the pruning bounds; BAYES / LRT: the first kListPow entries of the six power
  // tables as 16-byte words (2 x kListPow doubles = kListPow double2 per table pair, ints as int4)
  constexpr int kRelLoads = kMap ? (500 + kListThreads - 1) / kListThreads : 1;
  constexpr int kRmaxLoads = kMap ? (110 + kListThreads - 1) / kListThreads : 1;
  constexpr bool kStagePow = !kMap && kListPow > 0;
  constexpr int kPowD2 = kStagePow ? kListPow / 2 : 1;  // double2 per table row
  constexpr int kPowE4 = kStagePow ? kListPow / 4 : 1;  // int4 per exponent row
  constexpr int kPowLoads = kStagePow ? (4 * kPowD2 + kListThreads - 1) / kListThreads : 1;  // inv[0], inv[1], m[0], m[1]
  constexpr int kPowELoads = kStagePow ? (2 * kPowE4 + kListThreads - 1) / kListThreads : 1;
  double2 relv[kRelLoads];
  double rmaxv[kRmaxLoads];
  double2 powv[kPowLoads];
  int4 powe[kPowELoads];
  if (kMap) {
#pragma unroll
    for (int i = 0; i < kRelLoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      relv[i] = j < 500 ? __ldg(reinterpret_cast<const double2*>(tab->rel) + j) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i = 0; i < kRmaxLoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      rmaxv[i] = j < 110 ? (j < 100 ? tab->rmax_dm[j] : tab->rmax_d[j - 100]) : 0.0;
    }
  } else if (kStagePow) {
#pragma unroll
    for (int i = 0; i < kPowLoads; i++) {
      const int j = threadIdx.x + i * kListThreads;  // row = j / kPowD2: inv[0], inv[1], m[0], m[1]
      const int row = j / kPowD2, col = j % kPowD2;
      const double* src = row < 2 ? tab->pw_inv[row] : tab->pw_m[row & 1];
      powv[i] = row < 4 ? __ldg(reinterpret_cast<const double2*>(src) + col) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int i = 0; i < kPowELoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      const int row = j / kPowE4, col = j % kPowE4;
      powe[i] = row < 2 ? __ldg(reinterpret_cast<const int4*>(tab->pw_e[row]) + col) : make_int4(0, 0, 0, 0);
    }
  }
  const uint32_t isdenv = threadIdx.x < 32 ? tab->isden[threadIdx.x] : 0u;
#pragma unroll
  for (int i = 0; i < kWLoads; i++) {
    const uint32_t j = i * kListThreads + threadIdx.x;
    if (j < n_warps) s.woff[j] = wv[i];
  }
  if (kMap) {
#pragma unroll
    for (int i = 0; i < kRelLoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      if (j < 500) reinterpret_cast<double2*>(s.rel)[j] = relv[i];
    }
#pragma unroll
    for (int i = 0; i < kRmaxLoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      if (j < 110) s.rmax[j] = rmaxv[i];
    }
  } else if (kStagePow) {
#pragma unroll
    for (int i = 0; i < kPowLoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      const int row = j / kPowD2, col = j % kPowD2;
      if (row < 4) reinterpret_cast<double2*>(row < 2 ? s.pw_inv[row] : s.pw_m[row & 1])[col] = powv[i];
    }
#pragma unroll
    for (int i = 0; i < kPowELoads; i++) {
      const int j = threadIdx.x + i * kListThreads;
      const int row = j / kPowE4, col = j % kPowE4;
      if (row < 2) reinterpret_cast<int4*>(s.pw_e[row])[col] = powe[i];
    }
  }
  if (threadIdx.x < 32) s.isden[threadIdx.x] = isdenv;
  __syncthreads();
  // one contiguous run of `per` entries per thread (odd or not, scalar accesses at stride `per` words
  // spread over the banks well enough), block-wide exclusive scan of the run sums by shuffles
  const uint32_t per = (n_warps + kListThreads - 1) / kListThreads;
  const uint32_t lo = min(threadIdx.x * per, n_warps), hi = min(lo + per, n_warps);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += s.woff[i];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s.part[warp] = incl;
  __syncthreads();
  uint32_t before = incl - sum, total_items = 0;
#pragma unroll
  for (int i = 0; i < kListWarps; i++) {
    const uint32_t v = s.part[i];
    if (i < warp) before += v;
    total_items += v;
  }
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t v = s.woff[i];
    s.woff[i] = before;
    before += v;
  }
  if (threadIdx.x == 0) s.woff[n_warps] = total_items;
  __syncthreads();

  const uint32_t n_tiles = (total_items + kListThreads - 1) / kListThreads;
  const unsigned long long calls_before = ch.before ? *ch.before : 0ull;
  constexpr unsigned long long tag = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(&ch.acc[1], (unsigned long long)total_items);  // summed until the host collects the statistics
    if (n_tiles == 0) {  // nothing undecided: the call total carries over
      ch.pair[0] = calls_before;
      ch.pair[1] = calls_before;
      if (ch.count_out) *ch.count_out = (int64_t)calls_before;
      if (ch.last) atomicAdd(&ch.acc[0], calls_before);
    }
  }
  const ListPow<kMap> pw{s, tab};

  // Tiles are handed out by ticket, in increasing order, to CTAs that are running: the look-back below
  // therefore only ever waits for tiles held by running CTAs (no co-residency assumption).  A CTA claims
  // its next tile at the start of each trip and requests that tile's inputs (list entry -> site ->
  // packed counts, two dependent global loads) while the current tile is still being finished.
  auto fetch_site = [&](uint32_t tile) -> uint32_t {
    const uint32_t item = tile * kListThreads + threadIdx.x;
    if (tile >= n_tiles || item >= total_items) return 0u;  // dead lanes read site 0 (harmless) and are masked by `live`
    uint32_t a = 0, b = n_warps;  // largest g with woff[g] <= item
    while (b - a > 1) {
      const uint32_t mid = (a + b) >> 1;
      if (s.woff[mid] <= item) a = mid; else b = mid;
    }
    return __ldg(list + (size_t)screen_range_lo(a, range_base, range_rem) * seg_sites + (item - s.woff[a]));
  };
  auto fetch_counts = [&](uint32_t site, uint32_t* words) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(counts + (size_t)site * kBps);
#pragma unroll
    for (int q = 0; q < kBps / 4; q++) words[q] = __ldg(src + q);
  };
  uint32_t tile = s.ticket[0];  // written before the prologue's barriers
  uint32_t site_cur = fetch_site(tile), site_nxt = 0;
  uint32_t words_cur[kBps / 4];
  fetch_counts(site_cur, words_cur);
  int trip = 0;
  while (tile < n_tiles) {
    trip ^= 1;
    if (threadIdx.x == 0) s.ticket[trip] = (unsigned int)atomicAdd(&ch.work[0], 1ull);  // read after this trip's first barrier
    const uint32_t item = tile * kListThreads + threadIdx.x;
    const bool live = item < total_items;
    const uint32_t site = site_cur;
    int c[12];
    if (kU8) unpack_counts8(words_cur, c);
    else unpack_counts(reinterpret_cast<const uint2*>(words_cur), c);

    LinOut lo_;
    lo_.hard = false; lo_.dn = false; lo_.zd = lo_.zm = 2; lo_.den_s = lo_.men_s = lo_.s2 = 0.0;
    bool dn = false;
    ExactOut ex;
    ex.is_denovo = false;
    int map_best = 0;
    double map_ll = 0.0;
    if (kMap) {
      // MAP: the decision is the exact pruned log-domain sweep itself (no hard sites); p / LR / logs of the
      // called sites are filled in by phase 2
      if (live) {
        uint16_t cu[12];
#pragma unroll
        for (int q = 0; q < 12; q++) cu[q] = (uint16_t)c[q];
        double Ld[10], Lm[10], Lc[10];
        exact_read_loglik(cu + 0, k.hm, k.ht, k.mm, Ld);
        exact_read_loglik(cu + 4, k.hm, k.ht, k.mm, Lm);
        exact_read_loglik(cu + 8, k.hm, k.ht, k.mm, Lc);
        exact_map_pruned(Ld, Lm, Lc, s.rel, s.rmax, &map_best, &map_ll);
        dn = (s.isden[map_best >> 5] >> (map_best & 31)) & 1u;
      }
    } else {
      if (live) {
        if (k.lin_ok) linear_site(pw, c, k.k_men, k.kd9, k.kd8, k.kd7, k.kd6, k.thr, k.mm2, k.method, k.lrt_threshold, lo_);
        else lo_.hard = true;  // eps so large that r <= 1: the table scheme does not apply
      }
      dn = lo_.dn;
      // hard sites: the whole warp evaluates them one at a time with the exact direct form
      uint32_t hword = __ballot_sync(0xffffffffu, lo_.hard);
      if (hword) {
        if (lane == 0) atomicAdd(&ch.acc[2], (unsigned long long)__popc(hword));
        while (hword) {
          const int h = __ffs(hword) - 1;
          hword &= hword - 1;
          uint16_t cu[12];
#pragma unroll
          for (int q = 0; q < 12; q++) cu[q] = (uint16_t)__shfl_sync(0xffffffffu, c[q], h);
          ExactOut r;
          exact_site_warp(cu, kMap ? s.rel : tab->rel, s.isden, s.warp_L[warp], k.hm, k.ht, k.mm, k.method, k.lrt_threshold, &r);
          if (lane == h) { ex = r; dn = r.is_denovo; }
        }
      }
    }
    // order-preserving rank of the called sites inside the tile
    const uint32_t dw = __ballot_sync(0xffffffffu, dn);
    __syncthreads();  // s.warp_sum / s.excl of the previous trip are no longer read; s.ticket is written
    if (lane == 0) s.warp_sum[warp] = __popc(dw);
    const uint32_t tile_nxt = s.ticket[trip];
    site_nxt = fetch_site(tile_nxt);  // the list entry of my item in the next tile: in flight during the look-back
    __syncthreads();
    int warp_off = 0, total = 0;
#pragma unroll
    for (int i = 0; i < kListWarps; i++) {
      if (i < warp) warp_off += s.warp_sum[i];
      total += s.warp_sum[i];
    }
    if (warp == 0) {  // warp-parallel decoupled look-back: 32 predecessor tiles per probe
      unsigned long long excl = 0;
      if (tile > 0) {
        if (lane == 0) atomicExch(&status[tile], tag | (1ull << 32) | (unsigned long long)total);
        int hi_t = (int)tile - 1;
        for (;;) {
          const int p = hi_t - lane;
          unsigned long long v = tag | (2ull << 32);  // lanes before tile 0 act as an empty inclusive prefix
          if (p >= 0) {
            do { v = *(volatile unsigned long long*)&status[p]; } while (((v >> 32) & 3ull) == 0ull);
          }
          const uint32_t is_prefix = __ballot_sync(0xffffffffu, ((v >> 32) & 3ull) == 2ull);
          const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;  // nearest inclusive prefix in this window
          unsigned long long val = lane <= stop ? (v & 0xffffffffull) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
          excl += val;
          if (is_prefix) break;
          hi_t -= 32;
        }
      }
      if (lane == 0) {
        atomicExch(&status[tile], tag | (2ull << 32) | (excl + (unsigned long long)total));
        s.excl = excl;
        if (tile == n_tiles - 1) {
          const unsigned long long all = calls_before + excl + (unsigned long long)total;
          ch.pair[0] = calls_before;
          ch.pair[1] = all;
          if (ch.count_out) *ch.count_out = (int64_t)all;
          if (ch.last) atomicAdd(&ch.acc[0], all);
        }
      }
    }
    if (total != 0) __syncthreads();  // block-uniform: s.excl is published
    site_cur = site_nxt;
    fetch_counts(site_cur, words_cur);  // next tile's counts: in flight while this tile's records are staged
    if (total != 0) {
    if (dn) {
      const unsigned long long pos = calls_before + s.excl + (unsigned long long)(warp_off + __popc(dw & ((1u << lane) - 1u)));
      if (pos < (unsigned long long)cap) {
        dnv_call o;
        if (kMap) {  // phase 2 fills in p, LR and the log-likelihoods
          o.dad = (uint8_t)(map_best / 100); o.mom = (uint8_t)((map_best / 10) % 10); o.child = (uint8_t)(map_best % 10);
          o.max_ll = map_ll;
          o.p = o.lr = o.log_men = o.log_den = 0.0;
          o.flags = 0;
        } else if (lo_.hard) {
          o.dad = (uint8_t)ex.dad; o.mom = (uint8_t)ex.mom; o.child = (uint8_t)ex.child;
          o.max_ll = ex.max_ll; o.p = ex.p; o.lr = ex.lr; o.log_men = ex.log_men; o.log_den = ex.log_den;
          o.flags = 1;
        } else {  // the finish kernel fills in the MAP triple and max_ll, and turns the raw sums into p / LR / logs
          o.dad = o.mom = o.child = 0;
          o.max_ll = 0.0;
#if DNV_RAW_SUMS
          o.p = lo_.den_s; o.lr = lo_.men_s; o.log_men = lo_.s2; o.log_den = 0.0;
          o.flags = 2u | ((uint32_t)lo_.zd << 8) | ((uint32_t)lo_.zm << 12);
#else
          linear_values(lo_, o.p, o.lr, o.log_men, o.log_den);
          o.flags = 0;
#endif
        }
        o.site = site_base + (int64_t)site;
#pragma unroll
        for (int q = 0; q < 12; q++) o.counts[q] = (uint16_t)c[q];
        o.is_denovo = 1;
        stage[pos - calls_before] = o;  // local scratch; trio_calls_finish_kernel completes it and writes the caller's buffer
      }
    }
    }
    tile = tile_nxt;
  }
  __syncthreads();
  // the last CTA to leave clears the status entries this launch used and resets the ticket for the next launch
  // (by then every ticket has been taken and every look-back has finished)
  __shared__ int s_is_last;
  if (threadIdx.x == 0) {
    ktrace_stamp(ch.ktrace, ch.ktrace_seq, 1, true);
    __threadfence();
    s_is_last = atomicAdd(&ch.work[1], 1ull) == (unsigned long long)gridDim.x - 1ull;
  }
  __syncthreads();
  if (s_is_last) {
    for (uint32_t i = threadIdx.x; i < n_tiles; i += kListThreads) status[i] = 0ull;
    if (threadIdx.x == 0) {
      ch.work[1] = 0;
      ch.work[0] = 0;
    }
  }
}

// One thread per call: completes the records staged by trio_list_kernel and writes them, whole, to the
// caller's buffer.  A warp owns 32 consecutive calls: its records (2560 contiguous bytes) go through shared
// memory so that both the read of the staged records and the write of the finished ones are coalesced
// 512-byte accesses — the destination may be a peer GPU's HBM, where scattered 16-byte stores are expensive.
struct FinishShared {
  double rel[1000];
  double rmax[110];
  uint32_t isden[32];
  alignas(16) uint4 rec[kFinishWarps][32 * 5];  // one warp's 32 records
};

template <bool kMap>
__global__ void __launch_bounds__(kFinishThreads, kTailCtasPerSm)
trio_calls_finish_kernel(const Tables* __restrict__ tab, const FastConsts k, const CallChain ch,
                         const dnv_call* __restrict__ stage, dnv_call* __restrict__ calls, int64_t cap) {
  // Launched with programmatic stream serialisation: the CTAs of this kernel are scheduled while the list kernel's
  // last CTAs are still running, stage the site-invariant tables, and only then wait for the list kernel's results
  // (griddepcontrol.wait) — the launch gap and the prologue disappear from the tail's critical path.
  __shared__ FinishShared s;
  stage_exact_tables(tab, s.rel, s.isden);
  stage_map_bounds(tab, s.rmax);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) ktrace_stamp(ch.ktrace, ch.ktrace_seq, 2, false);
  const unsigned long long calls_before = ch.pair[0];  // written by the list kernel of this chunk
  unsigned long long all = ch.pair[1];
  if (all > (unsigned long long)cap) all = (unsigned long long)cap;
  __syncthreads();
  const GmemPow pw{tab};
  (void)pw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (calls_before + (unsigned long long)blockIdx.x * kFinishThreads < all) {  // block-uniform: else no call for this CTA
    // warp-uniform loop: the warp's first call, 32 consecutive calls per trip
    for (unsigned long long w0 = calls_before + (unsigned long long)blockIdx.x * kFinishThreads + (unsigned long long)warp * 32;
         w0 < all; w0 += (unsigned long long)gridDim.x * kFinishThreads) {
      const int n_rec = (int)min((unsigned long long)32, all - w0);
      const unsigned long long pos = w0 + lane;
      // coalesced read of the warp's staged records (local HBM) into shared memory
      const uint4* src = reinterpret_cast<const uint4*>(stage + (w0 - calls_before));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 5; i++) {
        const int q = i * 32 + lane;
        if (q < n_rec * 5) s.rec[warp][q] = __ldcg(src + q);
      }
      __syncwarp();
      union { dnv_call r; uint4 v[5]; } u;
#pragma unroll
      for (int i = 0; i < 5; i++) u.v[i] = s.rec[warp][min(lane, n_rec - 1) * 5 + i];
      if (lane < n_rec) {
      if (kMap) {
        // values of a MAP call: the factored linear form unless the site is hard (gray zone, deeper than
        // the tables), then the exact direct form in the reference's own order — as call_details_kernel does
        int ci[12];
#pragma unroll
        for (int q = 0; q < 12; q++) ci[q] = u.r.counts[q];
        LinOut lv;
        if (k.lin_ok) {
          linear_site(pw, ci, k.k_men, k.kd9, k.kd8, k.kd7, k.kd6, 1.0, k.mm2, DNV_METHOD_BAYES, 1.0, lv);
          lv.hard = lv.zd == 2 || lv.zm == 2;
        } else {
          lv.hard = true;
        }
        if (lv.hard) {
          ExactOut r;
          exact_site(u.r.counts, s.rel, s.isden, k.hm, k.ht, k.mm, k.method, k.lrt_threshold, &r);
          u.r.p = r.p; u.r.lr = r.lr; u.r.log_men = r.log_men; u.r.log_den = r.log_den;
          u.r.flags = 1;
          atomicAdd(&ch.acc[2], 1ull);
        } else {
          linear_values(lv, u.r.p, u.r.lr, u.r.log_men, u.r.log_den);
        }
      } else if (!(u.r.flags & 1u)) {  // hard sites are complete already
        if (u.r.flags & 2u) {  // raw sums of the list kernel -> p, LR, log-likelihoods (log and divisions run here, one dense thread per call)
          LinOut lv;
          lv.den_s = u.r.p; lv.men_s = u.r.lr; lv.s2 = u.r.log_men;
          lv.zd = (int)((u.r.flags >> 8) & 3u); lv.zm = (int)((u.r.flags >> 12) & 3u);
          lv.dn = true; lv.hard = false;
          linear_values(lv, u.r.p, u.r.lr, u.r.log_men, u.r.log_den);
          u.r.flags = 0;
        }
        double Ld[10], Lm[10], Lc[10];
        exact_read_loglik(u.r.counts + 0, k.hm, k.ht, k.mm, Ld);
        exact_read_loglik(u.r.counts + 4, k.hm, k.ht, k.mm, Lm);
        exact_read_loglik(u.r.counts + 8, k.hm, k.ht, k.mm, Lc);
        int best;
        double max_ll;
        exact_map_pruned(Ld, Lm, Lc, s.rel, s.rmax, &best, &max_ll);
        u.r.dad = (uint8_t)(best / 100); u.r.mom = (uint8_t)((best / 10) % 10); u.r.child = (uint8_t)(best % 10);
        u.r.max_ll = max_ll;
      }
      }
      (void)pos;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 5; i++) s.rec[warp][lane * 5 + i] = u.v[i];
      __syncwarp();
      // coalesced write of the finished records (the caller's buffer, possibly in a peer GPU's HBM)
      uint4* dst = reinterpret_cast<uint4*>(calls + w0);
#pragma unroll
      for (int i = 0; i < 5; i++) {
        const int q = i * 32 + lane;
        if (q < n_rec * 5) dst[q] = s.rec[warp][q];
      }
    }
    if (threadIdx.x == 0) ktrace_stamp(ch.ktrace, ch.ktrace_seq, 2, true);
  }
  // Synchronous single-chunk call: the last CTA to leave hands the accumulators to the host (pinned memory) and
  // restarts them — the call then ends with this kernel instead of a copy node and a memset node behind it
  // (measured: 110 -> 93 us per 10 M-site call).  Every CTA arrives, also those without a call.
  if (ch.host_acc) {  // kernel-uniform
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(ch.fin_gone, 1ull) == (unsigned long long)gridDim.x - 1ull) {
        *ch.fin_gone = 0ull;
#pragma unroll
        for (int i = 0; i < 4; i++) ch.host_acc[i] = atomicExch(&ch.acc[i], 0ull);
      }
    }
  }
}

// tail CTAs (four 128-register warps, one per sub-partition) that fit on an SM beside the resident screen CTAs:
// (16384 - screen warps per sub-partition * 32 * 64) / (32 * 128)
inline int tail_ctas_beside_screen() {
  static const int forced = [] { const char* e = getenv("DNV_TAIL_GRID"); return e ? atoi(e) : 0; }();  // A/B knob
  if (forced >= 1 && forced <= 8) return forced;
  const int screen_warps = screen_config().per_sm * kScreenWarps / 4;
  return std::max(1, (16384 - screen_warps * 2048) / 4096);
}

// Launches the list kernel (after the previous chain-extending kernel: ctx->ev_chain) and the finish kernel on
// stream `st`, reading lane `ln`'s list and staging its records in the lane's scratch.
template <bool kU8, bool kMap>
// The undecided sites arrive as n_segments sorted segments of ln.d_list (lengths ln.d_wcount), segment g starting at
// the first site of the g-th balanced range of seg_sites-site tiles: the screen's warps (seg_sites = 128) or the fused
// pileup-screen kernel's CTAs (1024).
static int launch_trio_list(dnv_ctx* ctx, TrioLane& ln, cudaStream_t st, const void* d_counts, int64_t n, int64_t site_base,
                            int n_segments, int seg_sites, const CallChain& ch, dnv_call* d_calls, int64_t cap, bool beside_screen) {
  const int sms = ctx->sms;
  const size_t smem = sizeof(ListShared<kMap>);
  auto kern = trio_list_kernel<kU8, kMap>;
  int& per_sm = ctx->list_per_sm[kU8 ? 1 : 0][kMap ? 1 : 0];
  if (per_sm == 0) {  // first launch on this context
    DNV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // same shared-memory configuration as the screen's, so that both can be resident on one SM (trio_screen.cuh)
    DNV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    DNV_CUDA(cudaFuncSetAttribute(trio_calls_finish_kernel<kMap>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int q = 1;
    DNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kListThreads, smem));
    per_sm = std::max(q, 1);
  }
  // beside_screen: another chunk's screen is (about to be) resident with three CTAs per SM, which leaves room for
  // exactly one tail CTA per SM — a larger grid would only queue CTAs behind that screen and delay the kernel's
  // completion.  Alone (last chunk of a synchronous call, serial contexts) the kernel takes the whole machine.
  const int64_t want = (n + kListThreads - 1) / kListThreads;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * (beside_screen ? tail_ctas_beside_screen() : per_sm)));
  const uint32_t n_tiles = (uint32_t)((n + seg_sites - 1) / seg_sites);
  const uint32_t n_warps = (uint32_t)n_segments;
  if (reinterpret_cast<uintptr_t>(d_calls) & 15) {
    set_error("d_calls must be 16-byte aligned (records are written as 16-byte words)");
    return DNV_E_INVALID;
  }
  // local scratch for the records between the two phases: at most min(cap, n) calls in this chunk
  const int64_t need = std::max<int64_t>(1, std::min(cap, n));
  if (need > ln.list_stage_cap) {
    if (ctx->capturing) { set_error("internal: scratch growth during graph capture"); return DNV_E_INVALID; }
    cudaFree(ln.d_list_stage);  // (synchronises the device: nothing in flight still uses the old buffer)
    ln.d_list_stage = nullptr;
    ln.list_stage_cap = 0;
    DNV_CUDA(cudaMalloc(&ln.d_list_stage, (size_t)need * sizeof(dnv_call)));
    ln.list_stage_cap = need;
  }
  const FastConsts fk = make_fast_consts(ctx->h_tab);
  static const int dbg_skip = [] { const char* e = getenv("DNV_DEBUG_SKIP"); return e ? atoi(e) : 0; }();  // 1: no tail, 2: no finish
  if (dbg_skip == 1) return DNV_OK;
  // chunks of one batch: this chunk's records start where the previous chunk's end (tickets and look-back status
  // are private to the lane, so lists of different batches need no ordering)
  if (parallel(ctx) && ch.before) DNV_CUDA(cudaStreamWaitEvent(st, ctx->ev_chain, 0));
  kern<<<grid, kListThreads, smem, st>>>(ctx->d_tab, fk, (const unsigned char*)d_counts, n, site_base, ln.d_list,
                                        ln.d_wcount, n_warps, n_tiles / n_warps, n_tiles % n_warps, (uint32_t)seg_sites,
                                        ln.d_tile_status, ch, cap, ln.d_list_stage);
  if (parallel(ctx) && !ch.last) DNV_CUDA(cudaEventRecord(ctx->ev_chain, st));
  // one thread per call; CTAs beyond the number of calls leave at once
  const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>((need + kFinishThreads - 1) / kFinishThreads,
                                                                (int64_t)sms * (beside_screen ? tail_ctas_beside_screen() : 4)));
  if (dbg_skip != 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)fgrid);
    cfg.blockDim = dim3(kFinishThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = DNV_FINISH_PDL;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const dnv_call* stage_in = ln.d_list_stage;
    DNV_CUDA(cudaLaunchKernelEx(&cfg, trio_calls_finish_kernel<kMap>, (const Tables*)ctx->d_tab, fk, ch, stage_in, d_calls, cap));
  }
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

}  // namespace dnv
