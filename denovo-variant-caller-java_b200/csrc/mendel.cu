// mendel.cu — kernel (c): VARIANT-mode Mendelian-inconsistency filter over one
// shard of the start-sorted (variant, call) stream.
//
// Reference semantics (streaming, M/VariantsBuffer.java + M/VariantCaller.java:120-210):
//   admission     checkAndAdd                         :172-185 (+ isSnp :264-282, FILTER :245-249)
//   trigger/evict canProcess / evictParents            :74-78, :147-157
//   join          getMatchingPair = FIRST queued entry (admission order) of each
//                 parent with start <= pos < end       :314-323
//   genotype      getGenotypeFromSNP, or ref+ref of the CHILD's reference base
//                 when the parent entry is not a SNP   :217-219, :228-239
//   call          PositionCall.isDenovo                :352-355
// On a stream sorted by start with end > start the streaming result equals the
// batch definition: for every admitted child SNP match
// each parent against all of that parent's admitted entries, lowest stream
// index first.  Batch form on the GPU, per parent with entries in stream order:
//   ub  = #entries with start <= pos                (binary search on start)
//   i*  = first i with prefix_max(end)[i] > pos     (binary search on the running max)
//   hit = i* < ub   (the first index where the running max exceeds pos is itself
//                    an entry with end > pos, and every earlier entry has end <= pos)
// The input-domain conditions (sorted, end > start for admitted entries) are
// CHECKED on the device and violations fail loudly with DNV_E_UNSORTED.
#include <limits.h>
#include <algorithm>

#include "dnv_internal.cuh"
#include "scan_util.cuh"

namespace dnv {

enum { kChild = 0, kMom = 1, kDad = 2 };
constexpr uint8_t kCodeBlock = 0xF0;  // admitted parent entry that is not a SNP
constexpr uint8_t kCodeBad = 0xFF;    // base outside ACGT (the reference would throw)

__device__ __forceinline__ int allele_of(char c) {
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}
__device__ __forceinline__ int geno_of(int x, int y) {  // Genotype.valueOfPairAlleles, M/DenovoUtil.java:229-233
  if (x > y) { const int t = x; x = y; y = t; }
  return x * 4 - (x * (x - 1)) / 2 + (y - x);
}

struct MendelArrays {
  // per-person admitted entries in stream order
  int64_t* start[3];
  int64_t* end[3];      // parents: running max of end after the scan
  uint8_t* code[3];
  uint8_t* ref;         // child only: allele index of the reference base (or 0xFF)
  int64_t* src;         // child only: index in the input stream
  uint32_t* bits[3];    // admission masks
  uint32_t* block_sums[3];
  unsigned long long* totals;  // [0..2] per-person admitted, [3] found, [4] error index+1, [5] error kind
  // join outputs
  uint32_t* found_bits;
  uint32_t* found_block_sums;
  uint32_t* packed;     // per child: child | mom<<8 | dad<<16
};

// ---- pass A: admission (checkAndAdd) + domain checks
__global__ void __launch_bounds__(kScanThreads)
mendel_admit_kernel(const dnv_vcall* __restrict__ calls, int64_t n, MendelArrays A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool adm = false;
  int person = 3;
  if (i < n) {
    const dnv_vcall v = calls[i];
    person = v.person;
    const bool ins = (v.ngt >= 1 && v.gt0 > 0 && v.alt0_len != 1) || (v.ngt >= 2 && v.gt1 > 0 && v.alt1_len != 1);
    const bool del = v.ref_len != 1;
    const bool snp = (v.end == v.start + 1) && !del && !ins;
    const bool dot = (v.ngt >= 1 && v.gt0 == -1) || (v.ngt >= 2 && v.gt1 == -1);
    adm = person <= 2 && !((person == kChild && !snp) || dot || v.ngt != 2 || !v.filter_pass || ins || del);
    bool bad = false;
    if (i > 0 && calls[i - 1].start > v.start) bad = true;  // stream must be sorted by start
    if (adm && v.end <= v.start) bad = true;                // empty interval: outside the batch-equivalence domain
    if (bad) { atomicMax(&A.totals[4], (unsigned long long)(i + 1)); A.totals[5] = (unsigned long long)(-DNV_E_UNSORTED); }
  }
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const uint32_t w = __ballot_sync(0xffffffffu, adm && person == p);
    if ((threadIdx.x & 31) == 0 && i < n) A.bits[p][i >> 5] = w;
  }
}

// ---- pass C: scatter admitted entries into per-person SoA queues (stream order)
__global__ void __launch_bounds__(kScanThreads)
mendel_scatter_kernel(const dnv_vcall* __restrict__ calls, int64_t n_words, MendelArrays A) {
  const int64_t w = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
#pragma unroll 1
  for (int p = 0; p < 3; p++) {
    uint32_t word = (w < n_words) ? A.bits[p][w] : 0u;
    int64_t pos = word_rank(word, A.block_sums[p]);
    while (word) {
      const int b = __ffs(word) - 1;
      word &= word - 1;
      const int64_t i = w * 32 + b;
      const dnv_vcall v = calls[i];
      const bool snp = v.end == v.start + 1;  // admitted => ref_len == 1 and no insertion
      uint8_t code = kCodeBlock;
      if (snp) {  // getGenotypeFromSNP, :228-239
        const int a0 = allele_of(v.gt0 > 0 ? v.alt0_base : v.ref_base);
        const int a1 = allele_of(v.gt1 > 0 ? v.alt1_base : v.ref_base);
        code = (a0 < 0 || a1 < 0) ? kCodeBad : (uint8_t)geno_of(a0, a1);
      }
      A.start[p][pos] = v.start;
      A.end[p][pos] = v.end;
      A.code[p][pos] = code;
      if (p == kChild) {
        const int r = allele_of(v.ref_base);
        A.ref[pos] = r < 0 ? 0xFF : (uint8_t)r;
        A.src[pos] = i;
      }
      pos++;
    }
  }
}

// ---- pass D: inclusive running max of end[] (three-phase scan)
constexpr int kMaxScanThreads = 256, kMaxScanItems = 8, kMaxScanTile = kMaxScanThreads * kMaxScanItems;

__device__ __forceinline__ long long block_inclusive_max(long long v, long long* s_warp) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v = y > v ? y : v;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
  __syncthreads();
  long long carry = LLONG_MIN;
  for (int i = 0; i < (int)(threadIdx.x >> 5); i++) carry = s_warp[i] > carry ? s_warp[i] : carry;
  __syncthreads();
  return carry > v ? carry : v;
}

__global__ void __launch_bounds__(kMaxScanThreads)
runmax_tile_kernel(int64_t* __restrict__ end, const unsigned long long* __restrict__ n_ptr,
                   long long* __restrict__ tile_max) {
  const int64_t n = (int64_t)*n_ptr;
  const int64_t base = (int64_t)blockIdx.x * kMaxScanTile + (int64_t)threadIdx.x * kMaxScanItems;
  if ((int64_t)blockIdx.x * kMaxScanTile >= n) return;
  __shared__ long long s_warp[kMaxScanThreads / 32];
  long long v[kMaxScanItems];
  long long run = LLONG_MIN;
#pragma unroll
  for (int k = 0; k < kMaxScanItems; k++) {
    const long long x = (base + k < n) ? end[base + k] : LLONG_MIN;
    run = x > run ? x : run;
    v[k] = run;
  }
  const long long incl = block_inclusive_max(run, s_warp);
  // exclusive prefix of this thread = max over previous threads' totals
  long long prev = __shfl_up_sync(0xffffffffu, incl, 1);
  __shared__ long long s_last[kMaxScanThreads / 32];
  if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = incl;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) prev = threadIdx.x ? s_last[(threadIdx.x >> 5) - 1] : LLONG_MIN;
#pragma unroll
  for (int k = 0; k < kMaxScanItems; k++)
    if (base + k < n) end[base + k] = v[k] > prev ? v[k] : prev;
  if (threadIdx.x == kMaxScanThreads - 1) tile_max[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(1024)
runmax_carry_kernel(long long* __restrict__ tile_max, const unsigned long long* __restrict__ n_ptr) {
  // serial over chunks of 1024 tiles; converts tile maxima to exclusive carries
  const int64_t n = (int64_t)*n_ptr;
  const int n_tiles = (int)((n + kMaxScanTile - 1) / kMaxScanTile);
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = LLONG_MIN;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + threadIdx.x;
    const long long v = i < n_tiles ? tile_max[i] : LLONG_MIN;
    const long long incl = block_inclusive_max(v, s_warp);
    long long prev = __shfl_up_sync(0xffffffffu, incl, 1);
    __shared__ long long s_last[32];
    if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = incl;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) prev = threadIdx.x ? s_last[(threadIdx.x >> 5) - 1] : LLONG_MIN;
    const long long carry = s_carry;
    if (i < n_tiles) tile_max[i] = prev > carry ? prev : carry;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl > carry ? incl : carry;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kMaxScanThreads)
runmax_apply_kernel(int64_t* __restrict__ end, const unsigned long long* __restrict__ n_ptr,
                    const long long* __restrict__ tile_carry) {
  const int64_t n = (int64_t)*n_ptr;
  if ((int64_t)blockIdx.x * kMaxScanTile >= n) return;
  const long long carry = tile_carry[blockIdx.x];
  if (carry == LLONG_MIN) return;
  const int64_t base = (int64_t)blockIdx.x * kMaxScanTile;
  for (int k = threadIdx.x; k < kMaxScanTile; k += kMaxScanThreads)
    if (base + k < n) { const long long x = end[base + k]; if (carry > x) end[base + k] = carry; }
}

// ---- pass E: join every admitted child SNP against both parents
__device__ __forceinline__ int parent_code(const int64_t* __restrict__ start, const int64_t* __restrict__ runmax,
                                           const uint8_t* __restrict__ code, int64_t n, int64_t pos) {
  int64_t lo = 0, hi = n;  // ub = #entries with start <= pos
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(start + mid) <= pos) lo = mid + 1; else hi = mid; }
  const int64_t ub = lo;
  lo = 0; hi = ub;         // first i < ub with runmax[i] > pos
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(runmax + mid) > pos) hi = mid; else lo = mid + 1; }
  return lo < ub ? (int)code[lo] : -1;
}

__global__ void __launch_bounds__(kScanThreads)
mendel_join_kernel(MendelArrays A) {
  const int64_t n_child = (int64_t)A.totals[kChild];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool found = false;
  if (i < n_child) {
    const int64_t pos = A.start[kChild][i];
    const int cg = A.code[kChild][i];
    const int rb = A.ref[i];
    int g[3] = {cg, -1, -1};
    bool bad = cg == kCodeBad;
    found = true;
    for (int par = kMom; par <= kDad && found; par++) {  // MOM then DAD, :203-220
      const int c = parent_code(A.start[par], A.end[par], A.code[par], (int64_t)A.totals[par], pos);
      if (c < 0) { found = false; break; }                // no covering entry -> no PositionCall
      if (c == kCodeBlock) { if (rb == 0xFF) bad = true; g[par] = geno_of(rb & 3, rb & 3); }
      else { if (c == kCodeBad) bad = true; g[par] = c; }
    }
    // the reference evaluates the child genotype first, so a bad child base throws even without parents
    if (bad) { atomicMax(&A.totals[4], (unsigned long long)(A.src[i] + 1)); A.totals[5] = (unsigned long long)(-DNV_E_BADBASE); found = false; }
    if (found) A.packed[i] = (uint32_t)g[kChild] | ((uint32_t)g[kMom] << 8) | ((uint32_t)g[kDad] << 16);
  }
  const uint32_t w = __ballot_sync(0xffffffffu, found);
  if ((threadIdx.x & 31) == 0 && i < n_child) A.found_bits[i >> 5] = w;
}

// ---- pass F: PositionCalls in stream order
__global__ void __launch_bounds__(kScanThreads)
mendel_emit_kernel(const Tables* __restrict__ tab, MendelArrays A, dnv_pcall* __restrict__ out, int64_t cap) {
  const int64_t n_child = (int64_t)A.totals[kChild];
  const int64_t n_words = (n_child + 31) / 32;
  const int64_t w = (int64_t)blockIdx.x * kScanThreads + threadIdx.x;
  if ((int64_t)blockIdx.x * kScanThreads >= n_words) return;
  uint32_t word = (w < n_words) ? A.found_bits[w] : 0u;
  int64_t pos = word_rank(word, A.found_block_sums);
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    const int64_t i = w * 32 + b;
    if (pos < cap) {
      const uint32_t pk = A.packed[i];
      const int c = pk & 0xFF, m = (pk >> 8) & 0xFF, d = (pk >> 16) & 0xFF;
      const int idx = d * 100 + m * 10 + c;
      dnv_pcall o;
      o.position = A.start[kChild][i];
      o.child_index = A.src[i];
      o.child = (uint8_t)c; o.mom = (uint8_t)m; o.dad = (uint8_t)d;
      o.is_denovo = (tab->isden[idx >> 5] >> (idx & 31)) & 1u;  // PositionCall.isDenovo, :352-355
      o.pad_[0] = o.pad_[1] = o.pad_[2] = o.pad_[3] = 0;
      out[pos] = o;
    }
    pos++;
  }
}

// found_block_sums of child blocks past the live range must be defined for the scan
__global__ void mendel_found_popc_kernel(MendelArrays A, int n_blocks_cap) {
  const int64_t n_child = (int64_t)A.totals[kChild];
  const int64_t n_words = (n_child + 31) / 32;
  for (int b = blockIdx.x; b < n_blocks_cap; b += gridDim.x) {
    const int64_t w = (int64_t)b * kScanThreads + threadIdx.x;
    int v = (w < n_words) ? __popc(A.found_bits[w]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ int s[kScanThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < kScanThreads / 32; k++) t += s[k];
      A.found_block_sums[b] = (uint32_t)t;
    }
    __syncthreads();
  }
}

}  // namespace dnv

using namespace dnv;

extern "C" {

int dnv_mendel_filter_device(dnv_ctx* ctx, const dnv_vcall* d_calls, int64_t n, dnv_pcall* d_out, int64_t cap,
                             int64_t* n_out) {
  if (!ctx || !n_out || n < 0 || cap < 0 || (n > 0 && !d_calls) || (cap > 0 && !d_out) || n >= (1ll << 36)) {
    set_error("dnv_mendel_filter_device: invalid argument");
    return DNV_E_INVALID;
  }
  *n_out = 0;
  if (n == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t n_words = (n + 31) / 32;
  const int n_blocks = (int)((n_words + kScanThreads - 1) / kScanThreads);
  const int64_t n_tiles = (n + kMaxScanTile - 1) / kMaxScanTile;

  // carve the scratch arena (every person's queue is sized for n entries)
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_start[3], o_end[3], o_code[3], o_bits[3], o_bsum[3];
  for (int p = 0; p < 3; p++) {
    o_start[p] = take((size_t)n * 8); o_end[p] = take((size_t)n * 8); o_code[p] = take((size_t)n);
    o_bits[p] = take((size_t)(n_words + 1) * 4); o_bsum[p] = take((size_t)(n_blocks + 1) * 4);
  }
  const size_t o_ref = take((size_t)n), o_src = take((size_t)n * 8), o_tot = take(64);
  const size_t o_fbits = take((size_t)(n_words + 1) * 4), o_fbsum = take((size_t)(n_blocks + 1) * 4);
  const size_t o_packed = take((size_t)n * 4), o_tmax = take((size_t)(n_tiles + 1) * 8);
  int rc = ensure_scratch(ctx, off);
  if (rc) return rc;
  char* base = (char*)ctx->d_scratch;
  MendelArrays A;
  for (int p = 0; p < 3; p++) {
    A.start[p] = (int64_t*)(base + o_start[p]); A.end[p] = (int64_t*)(base + o_end[p]);
    A.code[p] = (uint8_t*)(base + o_code[p]); A.bits[p] = (uint32_t*)(base + o_bits[p]);
    A.block_sums[p] = (uint32_t*)(base + o_bsum[p]);
  }
  A.ref = (uint8_t*)(base + o_ref); A.src = (int64_t*)(base + o_src);
  A.totals = (unsigned long long*)(base + o_tot);
  A.found_bits = (uint32_t*)(base + o_fbits); A.found_block_sums = (uint32_t*)(base + o_fbsum);
  A.packed = (uint32_t*)(base + o_packed);
  long long* tile_max = (long long*)(base + o_tmax);

  DNV_CUDA(cudaMemsetAsync(A.totals, 0, 64, st));
  const int grid_n = (int)((n + kScanThreads - 1) / kScanThreads);
  mendel_admit_kernel<<<grid_n, kScanThreads, 0, st>>>(d_calls, n, A);
  for (int p = 0; p < 3; p++) {
    popc_blocks_kernel<<<n_blocks, kScanThreads, 0, st>>>(A.bits[p], n_words, A.block_sums[p]);
    scan_blocks_generic_kernel<<<1, 1024, 0, st>>>(A.block_sums[p], n_blocks, A.totals + p);
  }
  mendel_scatter_kernel<<<n_blocks, kScanThreads, 0, st>>>(d_calls, n_words, A);
  for (int par = kMom; par <= kDad; par++) {
    runmax_tile_kernel<<<(int)n_tiles, kMaxScanThreads, 0, st>>>(A.end[par], A.totals + par, tile_max);
    runmax_carry_kernel<<<1, 1024, 0, st>>>(tile_max, A.totals + par);
    runmax_apply_kernel<<<(int)n_tiles, kMaxScanThreads, 0, st>>>(A.end[par], A.totals + par, tile_max);
  }
  mendel_join_kernel<<<grid_n, kScanThreads, 0, st>>>(A);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  mendel_found_popc_kernel<<<std::min(n_blocks, sms * 8), kScanThreads, 0, st>>>(A, n_blocks);
  scan_blocks_generic_kernel<<<1, 1024, 0, st>>>(A.found_block_sums, n_blocks, A.totals + 3);
  mendel_emit_kernel<<<n_blocks, kScanThreads, 0, st>>>(ctx->d_tab, A, d_out, cap);
  DNV_CUDA(cudaGetLastError());
  unsigned long long h[8];
  DNV_CUDA(cudaMemcpyAsync(h, A.totals, sizeof(h), cudaMemcpyDeviceToHost, st));
  DNV_CUDA(cudaStreamSynchronize(st));
  ctx->stats = dnv_stats();
  ctx->stats.n_sites = (int64_t)h[0];
  ctx->stats.n_calls = (int64_t)h[3];
  ctx->stats.n_launches = 18;
  if (h[4] != 0) {
    const int kind = -(int)h[5];
    set_error(kind == DNV_E_BADBASE
                  ? "mendel filter: base outside ACGT at record " + std::to_string(h[4] - 1) +
                        " (the reference throws IllegalArgumentException, M/VariantsBuffer.java:235-236)"
                  : "mendel filter: stream not sorted by start, or admitted record with end <= start, at record " +
                        std::to_string(h[4] - 1));
    return kind == DNV_E_BADBASE ? DNV_E_BADBASE : DNV_E_UNSORTED;
  }
  *n_out = (int64_t)h[3];
  if ((int64_t)h[3] > cap) {
    set_error("mendel filter: output capacity too small: need " + std::to_string(h[3]));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
}

int dnv_mendel_filter(dnv_ctx* ctx, const dnv_vcall* calls, int64_t n, dnv_pcall* out, int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out || n < 0 || cap < 0 || (n > 0 && !calls) || (cap > 0 && !out)) {
    set_error("dnv_mendel_filter: invalid argument");
    return DNV_E_INVALID;
  }
  *n_out = 0;
  if (n == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  dnv_vcall* d_calls = nullptr;
  dnv_pcall* d_out = nullptr;
  const int64_t dev_cap = std::min(cap, n);
  if (cudaMalloc(&d_calls, (size_t)n * sizeof(dnv_vcall)) != cudaSuccess ||
      cudaMalloc(&d_out, (size_t)std::max<int64_t>(dev_cap, 1) * sizeof(dnv_pcall)) != cudaSuccess) {
    cudaFree(d_calls); cudaFree(d_out);
    set_error("dnv_mendel_filter: device allocation failed");
    return DNV_E_NOMEM;
  }
  int rc = DNV_OK;
  if (cudaMemcpyAsync(d_calls, calls, (size_t)n * sizeof(dnv_vcall), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    set_error("dnv_mendel_filter: H2D copy failed"); rc = DNV_E_CUDA;
  }
  if (rc == DNV_OK) rc = dnv_mendel_filter_device(ctx, d_calls, n, d_out, dev_cap, n_out);
  const int64_t got = std::min<int64_t>(*n_out, dev_cap);
  if ((rc == DNV_OK || rc == DNV_E_CAPACITY) && got > 0) {
    if (cudaMemcpyAsync(out, d_out, (size_t)got * sizeof(dnv_pcall), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error("dnv_mendel_filter: D2H copy failed"); rc = DNV_E_CUDA;
    }
  }
  cudaFree(d_calls); cudaFree(d_out);
  return rc;
}

}  // extern "C"
