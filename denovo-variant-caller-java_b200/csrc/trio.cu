// trio.cu — READ-mode pipeline on one chunk of sites resident in HBM:
//
//   K1  trio_fast_kernel / trio_direct_all_kernel : counts -> decision bit per site
//                                                   (+ list of hard sites, + dense record)
//   K2  trio_direct_list_kernel : exact direct form on the hard list, patches bits
//   K3  popc_blocks / scan_blocks / scatter_calls : order-preserving compaction of
//                                                   called sites (input order)
//   K4  call_details_kernel     : exact direct form per called site -> dnv_call
//
// Replaces the per-site task of ReadCaller.runBayesDenovoInference
// (M/ReadCaller.java:111-130) -> BayesInfer.infer (M/BayesInfer.java:50-69).
// Nothing here synchronises with the host; the caller reads counters at the end.
#include <stdio.h>

#include "scan_util.cuh"
#include "trio_exact.cuh"
#include "trio_fast.cuh"

namespace dnv {

constexpr int kExactThreads = 128;

// ---------------------------------------------------------------- K1 (exact, all sites)
__global__ void __launch_bounds__(kExactThreads)
trio_direct_all_kernel(const Tables* __restrict__ tab, const uint16_t* __restrict__ counts,
                       int64_t n, uint32_t* __restrict__ bits, dnv_site_result* __restrict__ dense) {
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
      load_counts(counts, site, c);
      ExactOut r;
      exact_site(c, s_rel, s_isden, hm, ht, mm, method, thr, &r);
      dn = r.is_denovo;
      if (dense) store_dense(dense, site, r, 1u);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, dn);
    if ((threadIdx.x & 31) == 0 && site < n) bits[site >> 5] = word;
  }
}

// ---------------------------------------------------------------- K2 (exact, hard list)
__global__ void __launch_bounds__(kExactThreads)
trio_direct_list_kernel(const Tables* __restrict__ tab, const uint16_t* __restrict__ counts,
                        const uint32_t* __restrict__ hard_list,
                        const unsigned long long* __restrict__ counters, uint32_t* __restrict__ bits,
                        dnv_site_result* __restrict__ dense) {
  const unsigned long long n_hard = counters[1];
  if ((unsigned long long)blockIdx.x * blockDim.x >= n_hard) return;
  __shared__ double s_rel[1000];
  __shared__ uint32_t s_isden[32];
  stage_exact_tables(tab, s_rel, s_isden);
  __syncthreads();
  const double hm = tab->hm, ht = tab->ht, mm = tab->mm, thr = tab->lrt_threshold;
  const int method = tab->method;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_hard;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t site = hard_list[i];
    uint16_t c[12];
    load_counts(counts, site, c);
    ExactOut r;
    exact_site(c, s_rel, s_isden, hm, ht, mm, method, thr, &r);
    const uint32_t mask = 1u << (site & 31);
    if (r.is_denovo) atomicOr(&bits[site >> 5], mask);
    else atomicAnd(&bits[site >> 5], ~mask);
    if (dense) store_dense(dense, site, r, 1u);
  }
}

// ---------------------------------------------------------------- K3 compaction
// single block: exclusive scan of block_sums in place; advances the global call
// counter and publishes the [lo, hi) range of this chunk's calls.
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(uint32_t* __restrict__ block_sums, int n_blocks, unsigned long long* __restrict__ counters) {
  const uint32_t total = block_exclusive_scan_inplace(block_sums, n_blocks);
  if (threadIdx.x == 0) {
    const unsigned long long lo = counters[0];
    counters[2] = lo;
    counters[3] = lo + total;
    counters[0] = lo + total;
    counters[4] += counters[1];
  }
}

__global__ void __launch_bounds__(kScanThreads)
scatter_calls_kernel(const uint32_t* __restrict__ bits, int64_t n_words,
                     const uint32_t* __restrict__ block_offsets,
                     const unsigned long long* __restrict__ counters, int64_t site_base,
                     int64_t* __restrict__ call_idx, int64_t call_idx_cap) {
  const int64_t w = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
  uint32_t word = (w < n_words) ? bits[w] : 0u;
  int64_t pos = (int64_t)counters[2] + word_rank(word, block_offsets);
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    if (pos < call_idx_cap) call_idx[pos] = site_base + w * 32 + b;
    pos++;
  }
}

// ---------------------------------------------------------------- K4 call records
// One WARP per called site: lanes split the 1000 genotype triples (idx = lane + 32*i),
// every ll is the reference's exact add chain, so maxGenoType / max_ll are
// bit-exact (ties -> lowest dad-major index, as the reference's strict '>');
// the two exp() sums are lane partials combined in a fixed shuffle order
// (values agree with the sequential reference sum to ~1e-15 relative).
__global__ void __launch_bounds__(kExactThreads)
call_details_kernel(const Tables* __restrict__ tab, const uint16_t* __restrict__ counts,
                    int64_t site_base, const int64_t* __restrict__ call_idx,
                    const unsigned long long* __restrict__ counters, dnv_call* __restrict__ calls, int64_t cap) {
  const unsigned long long lo = counters[2];
  unsigned long long hi = counters[3];
  if (hi > (unsigned long long)cap) hi = (unsigned long long)cap;
  constexpr int kWarps = kExactThreads / 32;
  if (lo + (unsigned long long)blockIdx.x * kWarps >= hi) return;
  __shared__ double s_rel[1000];
  __shared__ uint32_t s_isden[32];
  __shared__ double s_L[kWarps][32];
  stage_exact_tables(tab, s_rel, s_isden);
  __syncthreads();
  const double hm = tab->hm, ht = tab->ht, mm = tab->mm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned long long i = lo + (unsigned long long)blockIdx.x * kWarps + warp; i < hi;
       i += (unsigned long long)gridDim.x * kWarps) {
    const int64_t gsite = call_idx[i];
    uint16_t c[12];
    load_counts(counts, gsite - site_base, c);
    if (lane < 30) {  // lane = person*10 + genotype: one exact read log-likelihood each
      const int p = lane / 10, g = lane % 10;
      const int a0 = kGenoA0[g], a1 = kGenoA1[g];
      double acc = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const double bl = (a == a0 || a == a1) ? (a0 == a1 ? hm : ht) : mm;
        acc = __dadd_rn(acc, __dmul_rn((double)(int)c[p * 4 + a], bl));
      }
      s_L[warp][lane] = acc;
    }
    __syncwarp();
    double max_ll = -INFINITY, l_den = 0.0, l_men = 0.0;
    int best = 0x7fffffff;
    for (int idx = lane; idx < 1000; idx += 32) {
      const int d = idx / 100, m = (idx / 10) % 10, ch = idx % 10;
      const double ll = __dadd_rn(__dadd_rn(__dadd_rn(s_L[warp][d], s_L[warp][10 + m]), s_L[warp][20 + ch]), s_rel[idx]);
      const double e = exp(ll);
      if ((s_isden[idx >> 5] >> (idx & 31)) & 1u) l_den += e; else l_men += e;
      if (ll > max_ll) { max_ll = ll; best = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oll = __shfl_xor_sync(0xffffffffu, max_ll, o);
      const int ob = __shfl_xor_sync(0xffffffffu, best, o);
      if (oll > max_ll || (oll == max_ll && ob < best)) { max_ll = oll; best = ob; }
      l_den += __shfl_xor_sync(0xffffffffu, l_den, o);
      l_men += __shfl_xor_sync(0xffffffffu, l_men, o);
    }
    __syncwarp();
    if (lane == 0) {
      dnv_call o;
      o.site = gsite;
#pragma unroll
      for (int k = 0; k < 12; k++) o.counts[k] = c[k];
      o.dad = (uint8_t)(best / 100); o.mom = (uint8_t)((best / 10) % 10); o.child = (uint8_t)(best % 10);
      o.is_denovo = 1;  // the decision that put the site on the list stands
      o.flags = 0;
      o.max_ll = max_ll;
      o.p = l_den / (l_den + l_men);
      o.lr = l_den / l_men;
      o.log_men = log(l_men);
      o.log_den = log(l_den);
      calls[i] = o;
    }
  }
}

// ---------------------------------------------------------------- launcher
int ensure_chunk_scratch(dnv_ctx* ctx, int64_t n) {
  if (n <= ctx->chunk_cap) return DNV_OK;
  cudaFree(ctx->d_bits); cudaFree(ctx->d_hard_list); cudaFree(ctx->d_block_sums);
  ctx->d_bits = nullptr; ctx->d_hard_list = nullptr; ctx->d_block_sums = nullptr;
  ctx->chunk_cap = 0;
  const int64_t n_words = (n + 31) / 32;
  const int64_t n_blocks = (n_words + kScanThreads - 1) / kScanThreads;
  DNV_CUDA(cudaMalloc(&ctx->d_bits, (size_t)(n_words + 1) * 4));
  DNV_CUDA(cudaMalloc(&ctx->d_hard_list, (size_t)n * 4));
  DNV_CUDA(cudaMalloc(&ctx->d_block_sums, (size_t)(n_blocks + 1) * 4));
  ctx->chunk_cap = n;
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

// One chunk (n < 2^31).  counters[0] (running call total) must have been
// zeroed by the caller before the first chunk of a batch.
int launch_trio_chunk(dnv_ctx* ctx, const uint16_t* d_counts, int64_t site_base, int64_t n,
                      dnv_site_result* d_dense, dnv_call* d_calls, int64_t cap) {
  if (n <= 0) return DNV_OK;
  int rc = ensure_chunk_scratch(ctx, n);
  if (rc) return rc;
  rc = ensure_call_idx(ctx, cap);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const int64_t n_words = (n + 31) / 32;
  const int n_blocks = (int)((n_words + kScanThreads - 1) / kScanThreads);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);

  DNV_CUDA(cudaMemsetAsync(ctx->d_counters + 1, 0, sizeof(unsigned long long), st));
  const bool direct = (ctx->prm.flags & DNV_F_FORCE_DIRECT) != 0;
  const bool timed = (ctx->prm.flags & DNV_F_TIME_KERNELS) != 0 && ctx->n_timed < 64;
  if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_k0[ctx->n_timed], st));
  if (direct) {
    const int64_t want = (n + kExactThreads - 1) / kExactThreads;
    const int grid = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
    trio_direct_all_kernel<<<grid, kExactThreads, 0, st>>>(ctx->d_tab, d_counts, n, ctx->d_bits, d_dense);
    if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_k1[ctx->n_timed++], st));
    ctx->stats.n_launches += 1;
  } else {
    rc = launch_trio_fast(ctx, d_counts, n, d_dense);
    if (rc) return rc;
    if (timed) DNV_CUDA(cudaEventRecord(ctx->ev_k1[ctx->n_timed++], st));
    trio_direct_list_kernel<<<sms * 4, kExactThreads, 0, st>>>(ctx->d_tab, d_counts, ctx->d_hard_list,
                                                              ctx->d_counters, ctx->d_bits, d_dense);
    ctx->stats.n_launches += 2;
  }
  popc_blocks_kernel<<<n_blocks, kScanThreads, 0, st>>>(ctx->d_bits, n_words, ctx->d_block_sums);
  scan_blocks_kernel<<<1, 1024, 0, st>>>(ctx->d_block_sums, n_blocks, ctx->d_counters);
  scatter_calls_kernel<<<n_blocks, kScanThreads, 0, st>>>(ctx->d_bits, n_words, ctx->d_block_sums,
                                                         ctx->d_counters, site_base, ctx->d_call_idx,
                                                         ctx->call_idx_cap);
  call_details_kernel<<<sms * 8, kExactThreads, 0, st>>>(ctx->d_tab, d_counts, site_base, ctx->d_call_idx,
                                                        ctx->d_counters, d_calls, cap);
  ctx->stats.n_launches += 4;
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

}  // namespace dnv
