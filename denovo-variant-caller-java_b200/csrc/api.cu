// api.cu — the extern "C" boundary (include/denovo_b200.h): context lifecycle
// and the READ-mode entry points.  Plain pointers and sizes only.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "dnv_internal.cuh"

namespace dnv {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int ensure_scratch(dnv_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return DNV_OK;
  cudaFree(ctx->d_scratch);
  ctx->d_scratch = nullptr;
  ctx->scratch_bytes = 0;
  DNV_CUDA(cudaMalloc(&ctx->d_scratch, bytes));
  ctx->scratch_bytes = bytes;
  return DNV_OK;
}

// sites per pipeline chunk of the device path (scratch is sized for one chunk)
constexpr int64_t kDeviceChunkDefault = 1ll << 27;
// DNV_DEVICE_CHUNK (sites, a multiple of 128 so that chunk boundaries keep the 16-byte alignment the
// bulk copies want) shrinks the chunk so the tests can drive the multi-chunk path with small inputs
static int64_t device_chunk() {
  static const int64_t v = [] {
    const char* e = getenv("DNV_DEVICE_CHUNK");
    const long long x = e ? atoll(e) : 0;
    return (x >= 128 && x <= kDeviceChunkDefault) ? (int64_t)(x / 128 * 128) : kDeviceChunkDefault;
  }();
  return v;
}
// sites per H2D/compute/D2H stage of the host path
constexpr int64_t kHostChunk = 1ll << 21;

static int ensure_stage(dnv_ctx* ctx, int64_t n, bool dense, int64_t calls_cap) {  // stage buffers hold the uint16 layout
  if (n > ctx->stage_cap) {
    for (int i = 0; i < 2; i++) {
      cudaFree(ctx->d_stage_counts[i]);
      cudaFree(ctx->d_stage_dense[i]);
      ctx->d_stage_counts[i] = nullptr;
      ctx->d_stage_dense[i] = nullptr;
    }
    ctx->stage_cap = 0;
    for (int i = 0; i < 2; i++) DNV_CUDA(cudaMalloc(&ctx->d_stage_counts[i], (size_t)n * 24));
    ctx->stage_cap = n;
  }
  if (dense)
    for (int i = 0; i < 2; i++)
      if (!ctx->d_stage_dense[i])
        DNV_CUDA(cudaMalloc(&ctx->d_stage_dense[i], (size_t)ctx->stage_cap * sizeof(dnv_site_result)));
  if (calls_cap > ctx->stage_calls_cap) {
    cudaFree(ctx->d_stage_calls);
    ctx->d_stage_calls = nullptr;
    ctx->stage_calls_cap = 0;
    DNV_CUDA(cudaMalloc(&ctx->d_stage_calls, (size_t)std::max<int64_t>(calls_cap, 1) * sizeof(dnv_call)));
    ctx->stage_calls_cap = calls_cap;
  }
  return DNV_OK;
}

// everything dnv_create allocates; on failure the caller hands the half-built context to dnv_destroy
static int init_ctx(dnv_ctx* ctx) {
  const dnv_params* p = &ctx->prm;
  build_tables(*p, &ctx->h_tab);
  ctx->sms = 148;
  cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, p->device);
  DNV_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  if (p->flags & DNV_F_TIME_KERNELS)
    for (int i = 0; i < 64; i++) {
      DNV_CUDA(cudaEventCreate(&ctx->ev_k0[i]));
      DNV_CUDA(cudaEventCreate(&ctx->ev_k1[i]));
      DNV_CUDA(cudaEventCreate(&ctx->ev_s1[i]));
    }
  DNV_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    DNV_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
    DNV_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
  }
  DNV_CUDA(cudaMalloc(&ctx->d_tab, sizeof(Tables)));
  DNV_CUDA(cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
  DNV_CUDA(cudaMalloc(&ctx->d_wcount, 8192 * sizeof(uint32_t)));
  DNV_CUDA(cudaHostAlloc((void**)&ctx->h_ring, 64 * 8 * sizeof(unsigned long long), cudaHostAllocDefault));
  // the kernels run on non-blocking streams, which do not order against the legacy default stream: the
  // uploads go through the context's own stream and are complete before dnv_create returns
  DNV_CUDA(cudaMemcpyAsync(ctx->d_tab, &ctx->h_tab, sizeof(Tables), cudaMemcpyHostToDevice, ctx->own_stream));
  DNV_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->own_stream));
  DNV_CUDA(cudaStreamSynchronize(ctx->own_stream));
  return DNV_OK;
}

}  // namespace dnv

using namespace dnv;

extern "C" {

const char* dnv_last_error(void) { return g_last_error.c_str(); }
int dnv_abi_version(void) { return DNV_ABI_VERSION; }

int dnv_create(const dnv_params* p, dnv_ctx** out) {
  if (!p || !out) { set_error("dnv_create: null argument"); return DNV_E_INVALID; }
  *out = nullptr;
  if (!(p->seq_err_rate > 0.0 && p->seq_err_rate < 1.0) ||
      !(p->denovo_mut_rate > 0.0 && p->denovo_mut_rate < 1.0) || isnan(p->lrt_threshold) ||
      p->method < DNV_METHOD_MAP || p->method > DNV_METHOD_LRT) {
    set_error("dnv_create: seq_err_rate and denovo_mut_rate must lie in (0,1), method in {0,1,2}");
    return DNV_E_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("dnv_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    return DNV_E_CUDA;
  }
  if (p->device < 0 || p->device >= ndev) { set_error("dnv_create: bad device ordinal"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(p->device));
  dnv_ctx* ctx = new (std::nothrow) dnv_ctx();
  if (!ctx) return DNV_E_NOMEM;
  memset(ctx, 0, sizeof(*ctx));
  ctx->prm = *p;
  ctx->device = p->device;
  const int rc = init_ctx(ctx);
  if (rc != DNV_OK) {  // nothing leaks on a failed create: dnv_destroy copes with a half-built context
    const std::string msg = g_last_error;
    dnv_destroy(ctx);
    set_error(msg);
    return rc;
  }
  *out = ctx;
  return DNV_OK;
}

int dnv_destroy(dnv_ctx* ctx) {
  if (!ctx) return DNV_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  cudaFree(ctx->d_tab); cudaFree(ctx->d_bits); cudaFree(ctx->d_tile_status); cudaFree(ctx->d_list);
  cudaFree(ctx->d_call_idx); cudaFree(ctx->d_counters); cudaFree(ctx->d_wcount); cudaFree(ctx->d_list_stage); cudaFree(ctx->d_stage_calls); cudaFree(ctx->d_scratch);
  for (int i = 0; i < 2; i++) {
    cudaFree(ctx->d_stage_counts[i]); cudaFree(ctx->d_stage_dense[i]);
    if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
  }
  for (int i = 0; i < 64; i++) {
    if (ctx->ev_k0[i]) cudaEventDestroy(ctx->ev_k0[i]);
    if (ctx->ev_k1[i]) cudaEventDestroy(ctx->ev_k1[i]);
    if (ctx->ev_s1[i]) cudaEventDestroy(ctx->ev_s1[i]);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
  cudaGetLastError();  // a half-built context may have produced sticky-free errors above: leave none behind
  delete ctx;
  return DNV_OK;
}

int dnv_get_tables(dnv_ctx* ctx, double* rel, double* base_ll, uint8_t* is_denovo) {
  if (!ctx) { set_error("dnv_get_tables: null ctx"); return DNV_E_INVALID; }
  Tables t;  // read back what the device actually holds
  DNV_CUDA(cudaSetDevice(ctx->device));
  DNV_CUDA(cudaMemcpy(&t, ctx->d_tab, sizeof(Tables), cudaMemcpyDeviceToHost));
  if (rel) memcpy(rel, t.rel, sizeof(t.rel));
  if (base_ll) {
    static const int a0[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3}, a1[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};
    for (int g = 0; g < 10; g++)
      for (int a = 0; a < 4; a++)
        base_ll[g * 4 + a] = (a == a0[g] || a == a1[g]) ? (a0[g] == a1[g] ? t.hm : t.ht) : t.mm;
  }
  if (is_denovo)
    for (int i = 0; i < 1000; i++) is_denovo[i] = (t.isden[i >> 5] >> (i & 31)) & 1u;
  return DNV_OK;
}

int dnv_set_stream(dnv_ctx* ctx, void* cuda_stream) {
  if (!ctx) { set_error("dnv_set_stream: null ctx"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(ctx->device));
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return DNV_OK;
}

int dnv_get_stats(dnv_ctx* ctx, dnv_stats* out) {
  if (!ctx || !out) { set_error("dnv_get_stats: null argument"); return DNV_E_INVALID; }
  *out = ctx->stats;
  return DNV_OK;
}

static void collect_kernel_times(dnv_ctx* ctx) {
  ctx->stats.fused_kernel_ms = 0;
  ctx->stats.screen_kernel_ms = 0;
  for (int i = 0; i < ctx->n_timed; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev_k0[i], ctx->ev_k1[i]) == cudaSuccess) ctx->stats.fused_kernel_ms += ms;
    if (ctx->screened && cudaEventElapsedTime(&ms, ctx->ev_k0[i], ctx->ev_s1[i]) == cudaSuccess)
      ctx->stats.screen_kernel_ms += ms;
  }
}

static int finish_batch(dnv_ctx* ctx, int64_t n, int64_t cap, int64_t* n_calls) {
  unsigned long long h[8];
  DNV_CUDA(cudaMemcpyAsync(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  collect_kernel_times(ctx);
  ctx->stats.n_screened_out = ctx->screened ? n - (int64_t)h[1] : 0;
  ctx->stats.n_sites = n;
  const unsigned long long total = h[ctx->calls_slot];
  ctx->stats.n_calls = (int64_t)total;
  ctx->stats.n_hard = (int64_t)h[4];
  *n_calls = (int64_t)total;
  if ((int64_t)total > cap) {
    set_error("call capacity too small: need " + std::to_string(total) + ", have " + std::to_string(cap));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
}

// batches enqueued by the async entry points must be collected before a synchronous call reuses the ctx
static int drain_async(dnv_ctx* ctx) {
  if (ctx->n_async == 0) return DNV_OK;
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  collect_kernel_times(ctx);
  int64_t calls = 0, hard = 0, undecided = 0;
  for (int i = 0; i < ctx->n_async; i++) {
    const unsigned long long* h = ctx->h_ring + i * 8;
    calls += (int64_t)h[ctx->ring_slot[i]];
    hard += (int64_t)h[4];
    undecided += (int64_t)h[1];
  }
  ctx->stats.n_sites = ctx->async_sites;
  ctx->stats.n_calls = calls;
  ctx->stats.n_hard = hard;
  ctx->stats.n_screened_out = ctx->screened ? ctx->async_sites - undecided : 0;
  ctx->n_async = 0;
  ctx->async_sites = 0;
  return DNV_OK;
}

static int infer_device_async_impl(dnv_ctx* ctx, const void* d_counts, int bps, int64_t n, dnv_call* d_calls,
                                   int64_t cap, int64_t* h_n_calls) {
  if (!ctx || !h_n_calls || n < 0 || cap < 0 || (n > 0 && !d_counts) || (cap > 0 && !d_calls)) {
    set_error("dnv_infer_counts_device_async: invalid argument");
    return DNV_E_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(d_counts) & (bps == 24 ? 7 : 3)) {  // the kernels read sites as 8- / 4-byte words
    set_error("d_counts must be 8-byte aligned (uint16 layout) or 4-byte aligned (uint8 layout)");
    return DNV_E_INVALID;
  }
  if (ctx->n_async >= 64) {
    set_error("dnv_infer_counts_device_async: 64 batches in flight, call dnv_sync first");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(ctx->device));
  if (ctx->n_async == 0) {
    ctx->stats = dnv_stats();
    ctx->n_timed = 0;
    ctx->screened = false;
    ctx->async_sites = 0;
  }
  ctx->calls_slot = 0;
  ctx->count_bytes = bps;
  DNV_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  ctx->count_out = h_n_calls;
  const int64_t kDeviceChunk = device_chunk();
  for (int64_t off = 0; off < n; off += kDeviceChunk) {
    const int64_t m = std::min(kDeviceChunk, n - off);
    int rc = launch_trio_chunk(ctx, (const unsigned char*)d_counts + off * bps, off, m, nullptr, d_calls, cap);
    if (rc) { ctx->count_out = nullptr; return rc; }
  }
  ctx->count_out = nullptr;
  if (n == 0)  // no kernel ran: the (zeroed) total still has to reach the caller's slot
    DNV_CUDA(cudaMemcpyAsync(h_n_calls, ctx->d_counters, sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
  unsigned long long* ring = ctx->h_ring + ctx->n_async * 8;
  DNV_CUDA(cudaMemcpyAsync(ring, ctx->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->ring_slot[ctx->n_async] = ctx->calls_slot;
  ctx->n_async++;
  ctx->async_sites += n;
  return DNV_OK;
}

static int infer_device_impl(dnv_ctx* ctx, const void* d_counts, int bps, int64_t n, dnv_site_result* d_dense,
                             dnv_call* d_calls, int64_t cap, int64_t* n_calls) {
  if (!ctx || !n_calls || n < 0 || cap < 0 || (n > 0 && !d_counts) || (cap > 0 && !d_calls)) {
    set_error("dnv_infer_counts_device: invalid argument");
    return DNV_E_INVALID;
  }
  if (reinterpret_cast<uintptr_t>(d_counts) & (bps == 24 ? 7 : 3)) {  // the kernels read sites as 8- / 4-byte words
    set_error("d_counts must be 8-byte aligned (uint16 layout) or 4-byte aligned (uint8 layout)");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(ctx->device));
  { int rc = drain_async(ctx); if (rc) return rc; }
  ctx->stats = dnv_stats();
  ctx->n_timed = 0;
  ctx->calls_slot = 0;
  ctx->screened = false;
  ctx->count_bytes = bps;
  DNV_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  const int64_t kDeviceChunk = device_chunk();
  for (int64_t off = 0; off < n; off += kDeviceChunk) {
    const int64_t m = std::min(kDeviceChunk, n - off);
    int rc = launch_trio_chunk(ctx, (const unsigned char*)d_counts + off * bps, off, m,
                               d_dense ? d_dense + off : nullptr, d_calls, cap);
    if (rc) return rc;
  }
  return finish_batch(ctx, n, cap, n_calls);
}

static int infer_host_impl(dnv_ctx* ctx, const void* counts, int bps, int64_t n, dnv_site_result* dense,
                           dnv_call* calls, int64_t cap, int64_t* n_calls) {
  if (!ctx || !n_calls || n < 0 || cap < 0 || (n > 0 && !counts) || (cap > 0 && !calls)) {
    set_error("dnv_infer_counts: invalid argument");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(ctx->device));
  { int rc = drain_async(ctx); if (rc) return rc; }
  ctx->stats = dnv_stats();
  ctx->n_timed = 0;
  ctx->calls_slot = 0;
  ctx->screened = false;
  ctx->count_bytes = bps;
  const int64_t chunk = std::min<int64_t>(kHostChunk, std::max<int64_t>(n, 1));
  const int64_t dev_cap = std::min(cap, n);
  int rc = ensure_stage(ctx, chunk, dense != nullptr, dev_cap);
  if (rc) return rc;
  DNV_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  // two-stage ring: H2D of chunk k+1 (copy stream) overlaps compute of chunk k
  int k = 0;
  for (int64_t off = 0; off < n; off += chunk, k ^= 1) {
    const int64_t m = std::min(chunk, n - off);
    // buffer k is free once the compute (and dense D2H) that last used it finished
    DNV_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k], 0));
    DNV_CUDA(cudaMemcpyAsync(ctx->d_stage_counts[k], (const unsigned char*)counts + off * bps, (size_t)m * bps,
                             cudaMemcpyHostToDevice, ctx->copy_stream));
    DNV_CUDA(cudaEventRecord(ctx->ev_h2d[k], ctx->copy_stream));
    DNV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[k], 0));
    rc = launch_trio_chunk(ctx, ctx->d_stage_counts[k], off, m, dense ? ctx->d_stage_dense[k] : nullptr,
                           ctx->d_stage_calls, dev_cap);
    if (rc) return rc;
    if (dense)
      DNV_CUDA(cudaMemcpyAsync(dense + off, ctx->d_stage_dense[k], (size_t)m * sizeof(dnv_site_result),
                               cudaMemcpyDeviceToHost, ctx->stream));
    DNV_CUDA(cudaEventRecord(ctx->ev_done[k], ctx->stream));
  }
  rc = finish_batch(ctx, n, cap, n_calls);
  const int64_t got = std::min<int64_t>(*n_calls, dev_cap);
  if (got > 0) {
    DNV_CUDA(cudaMemcpyAsync(calls, ctx->d_stage_calls, (size_t)got * sizeof(dnv_call), cudaMemcpyDeviceToHost,
                             ctx->stream));
    DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return rc;
}

int dnv_infer_counts_device(dnv_ctx* ctx, const uint16_t* d_counts, int64_t n, dnv_site_result* d_dense,
                            dnv_call* d_calls, int64_t cap, int64_t* n_calls) {
  return infer_device_impl(ctx, d_counts, 24, n, d_dense, d_calls, cap, n_calls);
}
int dnv_infer_counts8_device(dnv_ctx* ctx, const uint8_t* d_counts8, int64_t n, dnv_site_result* d_dense,
                             dnv_call* d_calls, int64_t cap, int64_t* n_calls) {
  return infer_device_impl(ctx, d_counts8, 12, n, d_dense, d_calls, cap, n_calls);
}
int dnv_infer_counts_device_async(dnv_ctx* ctx, const uint16_t* d_counts, int64_t n, dnv_call* d_calls, int64_t cap,
                                  int64_t* h_n_calls) {
  return infer_device_async_impl(ctx, d_counts, 24, n, d_calls, cap, h_n_calls);
}
int dnv_infer_counts8_device_async(dnv_ctx* ctx, const uint8_t* d_counts8, int64_t n, dnv_call* d_calls, int64_t cap,
                                   int64_t* h_n_calls) {
  return infer_device_async_impl(ctx, d_counts8, 12, n, d_calls, cap, h_n_calls);
}
int dnv_sync(dnv_ctx* ctx) {
  if (!ctx) { set_error("dnv_sync: null ctx"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(ctx->device));
  if (ctx->n_async == 0) { DNV_CUDA(cudaStreamSynchronize(ctx->stream)); return DNV_OK; }
  return drain_async(ctx);
}
int dnv_infer_counts(dnv_ctx* ctx, const uint16_t* counts, int64_t n, dnv_site_result* dense, dnv_call* calls,
                     int64_t cap, int64_t* n_calls) {
  return infer_host_impl(ctx, counts, 24, n, dense, calls, cap, n_calls);
}
int dnv_infer_counts8(dnv_ctx* ctx, const uint8_t* counts8, int64_t n, dnv_site_result* dense, dnv_call* calls,
                      int64_t cap, int64_t* n_calls) {
  return infer_host_impl(ctx, counts8, 12, n, dense, calls, cap, n_calls);
}

}  // extern "C"
