// api.cu — the extern "C" boundary (include/denovo_b200.h): context lifecycle
// and the READ-mode entry points.  Plain pointers and sizes only.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

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

int ensure_stage_calls(dnv_ctx* ctx, int64_t calls_cap) {
  if (calls_cap > ctx->stage_calls_cap) {
    cudaFree(ctx->d_stage_calls);
    ctx->d_stage_calls = nullptr;
    ctx->stage_calls_cap = 0;
    DNV_CUDA(cudaMalloc(&ctx->d_stage_calls, (size_t)std::max<int64_t>(calls_cap, 1) * sizeof(dnv_call)));
    ctx->stage_calls_cap = calls_cap;
  }
  return DNV_OK;
}

static int ensure_stage(dnv_ctx* ctx, int64_t n, bool dense) {  // stage buffers hold the uint16 layout
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
  return DNV_OK;
}

// lanes of a new context: DNV_LANES (1..4) overrides; timed contexts run serially so that their event
// brackets measure each kernel alone
static int lanes_for(const dnv_params& p) {
  if (p.flags & DNV_F_TIME_KERNELS) return 1;
  const char* e = getenv("DNV_LANES");
  const int x = e ? atoi(e) : 0;
  return (x >= 1 && x <= kMaxLanes) ? x : 3;  // three: screen k+1 starts while the tail of k still holds lane k
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
  ctx->n_lanes = lanes_for(*p);
  { const char* e = getenv("DNV_GRAPHS"); ctx->use_graphs = ctx->n_lanes > 1 && !(e && atoi(e) == 0); }
  // Asynchronous batches are enqueued eagerly unless DNV_GRAPHS_ASYNC=1: replayed graphs of different lanes were
  // measured to serialise against each other (benchmarks/ktrace.py: the list kernel of one lane's graph waits for
  // the screen of the next lane's), which costs more than the launch calls they save.
  { const char* e = getenv("DNV_GRAPHS_ASYNC"); ctx->use_graphs_async = ctx->use_graphs && e && atoi(e) == 1; }
  // list / finish kernels on a high-priority stream per lane: their CTAs are dispatched into the room beside the
  // resident screen ahead of the queued CTAs of the next screen (54.6 vs 58.6 us per 10 M-site batch); DNV_TAIL_PRIO=0
  // keeps a lane on one stream
  { const char* e = getenv("DNV_TAIL_PRIO"); ctx->tail_prio = !(e && atoi(e) == 0); }
  DNV_CUDA(cudaEventCreateWithFlags(&ctx->ev_sub, cudaEventDisableTiming));
  DNV_CUDA(cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
  for (int l = 0; l < ctx->n_lanes; l++) {
    TrioLane& ln = ctx->lane[l];
    if (ctx->n_lanes > 1) {
      int least = 0, greatest = 0;
      DNV_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      DNV_CUDA(cudaStreamCreateWithPriority(&ln.own, cudaStreamNonBlocking, least));
      if (ctx->tail_prio) DNV_CUDA(cudaStreamCreateWithPriority(&ln.tail, cudaStreamNonBlocking, greatest));
      DNV_CUDA(cudaEventCreateWithFlags(&ln.ev_screen, cudaEventDisableTiming));
      DNV_CUDA(cudaEventCreateWithFlags(&ln.ev_done, cudaEventDisableTiming));
      DNV_CUDA(cudaEventCreateWithFlags(&ln.ev_flush, cudaEventDisableTiming));
    }
    DNV_CUDA(cudaMalloc(&ln.d_wcount, kWcountWords * sizeof(uint32_t)));
  }
  DNV_CUDA(cudaMalloc(&ctx->d_tab, sizeof(Tables)));
  DNV_CUDA(cudaMalloc(&ctx->d_counters, kCtrWords * sizeof(unsigned long long)));
  DNV_CUDA(cudaHostAlloc((void**)&ctx->h_acc, 16 * sizeof(unsigned long long), cudaHostAllocDefault));  // [0..3] acc, [8..15] pileup status (pileup.cu)
  // the kernels run on non-blocking streams, which do not order against the legacy default stream: the
  // uploads go through the context's own stream and are complete before dnv_create returns
  DNV_CUDA(cudaMemcpyAsync(ctx->d_tab, &ctx->h_tab, sizeof(Tables), cudaMemcpyHostToDevice, ctx->own_stream));
  DNV_CUDA(cudaMemsetAsync(ctx->d_counters, 0, kCtrWords * sizeof(unsigned long long), ctx->own_stream));
  if (getenv("DNV_KTRACE")) {
    std::vector<unsigned long long> init((size_t)kMaxLanes * kKtraceSlots * 6);
    for (size_t i = 0; i < init.size(); i++) init[i] = (i & 1) ? 0ull : ~0ull;  // min slots start at +inf, max slots at 0
    DNV_CUDA(cudaMalloc(&ctx->d_ktrace, init.size() * 8));
    DNV_CUDA(cudaMalloc(&ctx->d_ktrace_seq, kMaxLanes * 4));
    DNV_CUDA(cudaMemcpyAsync(ctx->d_ktrace, init.data(), init.size() * 8, cudaMemcpyHostToDevice, ctx->own_stream));
    DNV_CUDA(cudaMemsetAsync(ctx->d_ktrace_seq, 0, kMaxLanes * 4, ctx->own_stream));
    DNV_CUDA(cudaStreamSynchronize(ctx->own_stream));  // `init` goes out of scope
  }
  DNV_CUDA(cudaStreamSynchronize(ctx->own_stream));
  return DNV_OK;
}

// ---- the pipeline as the entry points see it ------------------------------------------------------------------
// begin_submit   the front stream's position becomes "inputs are ready" for whatever is enqueued next
// enqueue_*      chunks of one batch onto the lanes; nothing waits on the host
// flush_lanes    the front stream waits (on the device) for everything enqueued so far
// collect        flush + accumulators to the host + host wait; they restart at zero
int begin_submit(dnv_ctx* ctx) {
  DNV_CUDA(cudaSetDevice(ctx->device));
  ctx->sub_pending = false;
  if (parallel(ctx)) {
    // nothing pending on the front stream (the usual case in a loop of asynchronous batches): the inputs are
    // ready and the lanes need not wait for anything — one query instead of a record and a wait per batch
    if (!ctx->capturing && cudaStreamQuery(ctx->stream) == cudaSuccess) return DNV_OK;
    cudaGetLastError();  // cudaErrorNotReady is not an error
    DNV_CUDA(cudaEventRecord(ctx->ev_sub, ctx->stream));
    ctx->sub_pending = true;
  }
  return DNV_OK;
}

int flush_lanes(dnv_ctx* ctx) {
  if (!parallel(ctx)) return DNV_OK;
  for (int l = 0; l < ctx->n_lanes; l++)
    if (ctx->lane[l].used) {  // the lane's screen stream has rejoined its tail stream after every chunk
      DNV_CUDA(cudaEventRecord(ctx->lane[l].ev_flush, ctx->lane[l].own));
      DNV_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->lane[l].ev_flush, 0));
      ctx->lane[l].used = false;
    }
  return DNV_OK;
}

// the calls-only pipeline (screen -> list -> finish) is what a graph may hold: its kernels keep no launch tag
bool will_screen(const dnv_ctx* ctx, bool dense) {
  const bool direct = (ctx->prm.flags & DNV_F_FORCE_DIRECT) != 0;
  const bool map = ctx->prm.method == DNV_METHOD_MAP;
  return !direct && (map ? ctx->h_tab.map_screen_ok : ctx->h_tab.early_ok) && !dense && !(ctx->prm.flags & DNV_F_NO_SCREEN);
}

// Per chunk: the screen, unless recent batches were screen-hostile.  The screen pays when it decides most sites
// (99 % of a genome-wide or candidate-file pileup); on inputs that are mostly de-novo-pattern sites — e.g. the
// Mendelian-inconsistent candidates FULL mode passes from the VARIANT filter to READ mode
// (M/DenovoRunner.java:173-182) — everything it leaves goes through the list and finish kernels, built for short
// lists (measured: 53 us + ~0.7 us per 1000 undecided sites of a 10 M-site batch), and above ~5 % undecided the
// all-sites fused kernel (trio_fast.cuh, ~34 ns per site) wins.  Both paths return identical
// records; while in dense mode every 32nd chunk is screened again to re-measure the mix.
bool pick_screen(dnv_ctx* ctx, bool dense) {
  if (!will_screen(ctx, dense)) return false;
  if (!ctx->dense_mode) return true;
  if (++ctx->dense_probe >= 32) { ctx->dense_probe = 0; return true; }
  return false;
}

// entry of `key` in a graph cache: the existing one, else the least recently used slot recycled for it
static GraphEntry* graph_slot(dnv_ctx* ctx, GraphEntry* tab, const GraphKey& key) {
  GraphEntry* lru = &tab[0];
  for (int i = 0; i < kGraphSlots; i++) {
    if (tab[i].seen && tab[i].key == key) { tab[i].stamp = ++ctx->graph_clock; return &tab[i]; }
    if (tab[i].stamp < lru->stamp) lru = &tab[i];
  }
  if (lru->exec) cudaGraphExecDestroy(lru->exec);
  lru->exec = nullptr;
  lru->key = key;
  lru->seen = 0;
  lru->stamp = ++ctx->graph_clock;
  return lru;
}

static int end_capture(dnv_ctx* ctx, cudaStream_t origin, int enqueue_rc, cudaGraphExec_t* exec) {
  ctx->capturing = false;
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(origin, &graph);
  if (enqueue_rc) { if (graph) cudaGraphDestroy(graph); return enqueue_rc; }
  if (e != cudaSuccess || !graph) { set_error(std::string("graph capture failed: ") + cudaGetErrorString(e)); cudaGetLastError(); return DNV_E_CUDA; }
  const cudaError_t e2 = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess) { set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2)); return DNV_E_CUDA; }
  return DNV_OK;
}

// device side of collect(): everything joins the front stream, the accumulators go to the host and restart
static int enqueue_collect(dnv_ctx* ctx) {
  int rc = flush_lanes(ctx);
  if (rc) return rc;
  DNV_CUDA(cudaMemcpyAsync(ctx->h_acc, ctx->d_counters + kCtrAcc, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           ctx->stream));
  DNV_CUDA(cudaMemsetAsync(ctx->d_counters + kCtrAcc, 0, 4 * sizeof(unsigned long long), ctx->stream));
  return DNV_OK;
}

int collect(dnv_ctx* ctx) {
  int rc = enqueue_collect(ctx);
  if (rc) return rc;
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  return DNV_OK;
}

// chunk size of one device-resident batch.  Round 2 first split a synchronous call over the lanes (the tail of one
// chunk hides behind the screen of the next).  Measured against the whole call in one chunk
// (benchmarks/time_sync_call.py; us per call whole / in three chunks): 10 M sites 92 / 130, 40 M 236 / 263,
// 67 M 370 / 389, 100 M 523 / 543 — the chunks' list kernels are chained (record ranges) at ~20 us of latency per
// link, and only a call that is ONE chunk ends with its finish kernel (CallChain::host_acc).  So a call is split only
// where it has to be (the scratch is sized for device_chunk() sites); DNV_SYNC_MIN_CHUNK=<sites> asks for one chunk
// per that many sites, up to one per lane (A/B knob, and the tests' way onto the lane-split path).
static int64_t chunk_for(const dnv_ctx* ctx, int64_t n, bool split) {
  const int64_t full = device_chunk();
  if (!split || !parallel(ctx) || getenv("DNV_DEVICE_CHUNK")) return full;
  static const int64_t kSplitSites = [] { const char* e = getenv("DNV_SYNC_MIN_CHUNK"); const long long v = e ? atoll(e) : 0; return v > 0 ? (int64_t)v : (int64_t)0; }();
  if (kSplitSites == 0) return full;
  const int64_t k = std::max<int64_t>(1, std::min<int64_t>(ctx->n_lanes, n / kSplitSites));
  int64_t c = (n + k - 1) / k;
  c = (c + 127) / 128 * 128;  // chunk boundaries keep the 16-byte alignment the bulk copies want
  return std::min(std::max<int64_t>(c, 128), full);
}

int enqueue_device_batch(dnv_ctx* ctx, const void* d_counts, int bps, int64_t n, int64_t site_base, dnv_site_result* d_dense,
                         dnv_call* d_calls, int64_t cap, int64_t* count_out, bool split, bool no_sub_wait, bool collect_to_host) {
  ctx->count_bytes = bps;
  const int64_t chunk = chunk_for(ctx, n, split);
  for (int64_t off = 0; off < n; off += chunk) {
    ChunkArgs a;
    a.d_counts = (const unsigned char*)d_counts + off * bps;
    a.site_base = site_base + off;
    a.n = std::min(chunk, n - off);
    a.d_dense = d_dense ? d_dense + off : nullptr;
    a.d_calls = d_calls;
    a.cap = cap;
    a.first = off == 0;
    a.last = off + chunk >= n;
    a.tail_alone = split && a.last;  // synchronous call: nothing follows the last chunk
    a.no_sub_wait = no_sub_wait;
    a.wait_ev = nullptr;
    a.count_out = count_out;
    a.reads = nullptr;
    a.dense_host = nullptr;
    a.done_ev = nullptr;
    a.collect_to_host = collect_to_host && a.first && a.last;
    const int rc = launch_trio_chunk(ctx, a);
    if (rc) return rc;
  }
  if (n == 0 && count_out)  // no kernel ran: the (zero) total still has to reach the caller's slot
    DNV_CUDA(cudaMemcpyAsync(count_out, ctx->d_counters + kCtrZero, sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
  return DNV_OK;
}

// HOST counts -> device records: two-stage ring, the H2D copy of stage k+1 (copy stream) overlaps the
// compute of stage k; every stage is one chunk on the next lane.  d_calls: device buffer of the records.
int enqueue_host_batch(dnv_ctx* ctx, const void* counts, int bps, int64_t n, int64_t site_base, dnv_site_result* dense,
                       dnv_call* d_calls, int64_t cap, int64_t* count_out) {
  ctx->count_bytes = bps;
  const int64_t chunk = std::min<int64_t>(kHostChunk, std::max<int64_t>(n, 1));
  int rc = ensure_stage(ctx, chunk, dense != nullptr);
  if (rc) return rc;
  if (parallel(ctx) && ctx->sub_pending) DNV_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sub, 0));
  int k = 0;
  for (int64_t off = 0; off < n; off += chunk, k ^= 1) {
    const int64_t m = std::min(chunk, n - off);
    // buffer k is free once the chunk (and dense D2H) that last used it has finished
    DNV_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k], 0));
    DNV_CUDA(cudaMemcpyAsync(ctx->d_stage_counts[k], (const unsigned char*)counts + off * bps, (size_t)m * bps,
                             cudaMemcpyHostToDevice, ctx->copy_stream));
    DNV_CUDA(cudaEventRecord(ctx->ev_h2d[k], ctx->copy_stream));
    ChunkArgs a;
    a.d_counts = ctx->d_stage_counts[k];
    a.site_base = site_base + off;
    a.n = m;
    a.d_dense = dense ? ctx->d_stage_dense[k] : nullptr;
    a.d_calls = d_calls;
    a.cap = cap;
    a.first = off == 0;
    a.last = off + chunk >= n;
    a.tail_alone = false;
    a.no_sub_wait = false;
    a.wait_ev = ctx->ev_h2d[k];
    a.count_out = count_out;
    a.reads = nullptr;
    a.dense_host = dense ? dense + off : nullptr;
    a.done_ev = ctx->ev_done[k];
    rc = launch_trio_chunk(ctx, a);
    if (rc) return rc;
  }
  if (n == 0 && count_out)
    DNV_CUDA(cudaMemcpyAsync(count_out, ctx->d_counters + kCtrZero, sizeof(int64_t), cudaMemcpyDefault, ctx->stream));
  return DNV_OK;
}

void reset_call_state(dnv_ctx* ctx) {
  ctx->stats = dnv_stats();
  ctx->n_timed = 0;
  ctx->screened = false;
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
  // One hardware work queue per stream of a context (front, copy, and a screen + a tail stream per lane): with
  // the default of 8 connections, streams that share a queue serialise against each other (measured: the list
  // kernel of one lane waits for the screen of the next).  Only effective if CUDA has not been initialised yet in
  // this process; a host that initialises CUDA first sets the variable itself (bench.py does).
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
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
  cudaDeviceSynchronize();
  for (int g = 0; g < kGraphSlots; g++) if (ctx->sync_graphs[g].exec) cudaGraphExecDestroy(ctx->sync_graphs[g].exec);
  cudaFree(ctx->d_tab); cudaFree(ctx->d_bits); cudaFree(ctx->d_cstatus);
  cudaFree(ctx->d_ktrace); cudaFree(ctx->d_ktrace_seq); cudaFree(ctx->d_reads_counts);
  cudaFree(ctx->d_call_idx); cudaFree(ctx->d_counters); cudaFree(ctx->d_stage_calls); cudaFree(ctx->d_scratch);
  for (int l = 0; l < kMaxLanes; l++) {
    TrioLane& ln = ctx->lane[l];
    cudaFree(ln.d_list); cudaFree(ln.d_wcount); cudaFree(ln.d_list_stage);
    if (ln.ev_done) cudaEventDestroy(ln.ev_done);
    if (ln.ev_flush) cudaEventDestroy(ln.ev_flush);
    cudaFree(ln.d_tile_status);
    for (int g = 0; g < kGraphSlots; g++) if (ln.graphs[g].exec) cudaGraphExecDestroy(ln.graphs[g].exec);
    if (ln.ev_screen) cudaEventDestroy(ln.ev_screen);
    if (ln.own) cudaStreamDestroy(ln.own);
    if (ln.tail) cudaStreamDestroy(ln.tail);
  }
  if (ctx->ev_sub) cudaEventDestroy(ctx->ev_sub);
  if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
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
  if (ctx->h_acc) cudaFreeHost(ctx->h_acc);
  cudaGetLastError();  // a half-built context may have left an error above: leave none behind
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

// statistics of everything enqueued since the previous collect (h_acc holds the accumulators)
int dnvi_finish_stats(dnv_ctx* ctx, int64_t n_sites) {
  collect_kernel_times(ctx);
  ctx->stats.n_sites = n_sites;
  ctx->stats.n_calls = (int64_t)ctx->h_acc[0];
  ctx->stats.n_hard = (int64_t)ctx->h_acc[2];
  ctx->stats.n_screened_out = ctx->sites_screened - (int64_t)ctx->h_acc[1];
  if (ctx->sites_screened >= 4096 && !getenv("DNV_NO_DENSE_MODE"))
    ctx->dense_mode = (double)ctx->h_acc[1] > 0.05 * (double)ctx->sites_screened;
  ctx->sites_screened = 0;
  return DNV_OK;
}
// host wait + statistics
static int finish_all(dnv_ctx* ctx, int64_t n_sites) {
  const int rc = collect(ctx);
  return rc ? rc : dnvi_finish_stats(ctx, n_sites);
}

// batches enqueued by the async entry points must be collected before a synchronous call reuses the ctx
static int drain_async(dnv_ctx* ctx) {
  if (ctx->n_async == 0) return DNV_OK;
  const int rc = finish_all(ctx, ctx->async_sites);
  ctx->n_async = 0;
  ctx->async_sites = 0;
  return rc;
}

int dnv_set_stream(dnv_ctx* ctx, void* cuda_stream) {
  if (!ctx) { set_error("dnv_set_stream: null ctx"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(ctx->device));
  { const int rc = flush_lanes(ctx); if (rc) return rc; }
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return DNV_OK;
}

int dnv_get_stats(dnv_ctx* ctx, dnv_stats* out) {
  if (!ctx || !out) { set_error("dnv_get_stats: null argument"); return DNV_E_INVALID; }
  *out = ctx->stats;
  return DNV_OK;
}

static bool bad_alignment(const void* d_counts, int bps) {
  if (reinterpret_cast<uintptr_t>(d_counts) & (bps == 24 ? 7 : 3)) {  // the kernels read sites as 8- / 4-byte words
    set_error("d_counts must be 8-byte aligned (uint16 layout) or 4-byte aligned (uint8 layout)");
    return true;
  }
  return false;
}

static int infer_device_async_impl(dnv_ctx* ctx, const void* d_counts, int bps, int64_t n, int64_t site_base, dnv_call* d_calls,
                                   int64_t cap, int64_t* n_calls_out) {
  if (!ctx || !n_calls_out || n < 0 || cap < 0 || (n > 0 && !d_counts) || (cap > 0 && !d_calls)) {
    set_error("dnv_infer_counts_device_async: invalid argument");
    return DNV_E_INVALID;
  }
  if (bad_alignment(d_counts, bps)) return DNV_E_INVALID;
  int rc = begin_submit(ctx);
  if (rc) return rc;
  if (ctx->n_async == 0) {
    reset_call_state(ctx);
    ctx->async_sites = 0;
  }
  // A batch of one chunk on the calls-only path is replayable: the second time its arguments are seen it is
  // captured on its lane's stream, from then on it costs the host one wait + one graph launch.
  const bool graphable = ctx->use_graphs_async && !ctx->dense_mode && n > 0 && n <= device_chunk() && will_screen(ctx, false);
  if (graphable) {
    TrioLane& ln = ctx->lane[ctx->lane_next];
    const GraphKey key{d_counts, d_calls, n_calls_out, nullptr, n, site_base, cap, bps};
    GraphEntry* ge = graph_slot(ctx, ln.graphs, key);
    ge->seen++;
    if (!ge->exec && ge->seen >= 2) {
      const int lane_before = ctx->lane_next;
      ln.parity = 0;
      DNV_CUDA(cudaStreamBeginCapture(ln.own, cudaStreamCaptureModeRelaxed));
      ctx->capturing = true;
      rc = enqueue_device_batch(ctx, d_counts, bps, n, site_base, nullptr, d_calls, cap, n_calls_out, /*split=*/false, /*no_sub_wait=*/true);
      rc = end_capture(ctx, ln.own, rc, &ge->exec);
      ctx->lane_next = lane_before;
      ln.parity = 0;
      if (rc) return rc;
    }
    if (ge->exec) {
      if (ctx->sub_pending) DNV_CUDA(cudaStreamWaitEvent(ln.own, ctx->ev_sub, 0));
      DNV_CUDA(cudaGraphLaunch(ge->exec, ln.own));
      ln.used = true;
      ctx->lane_next = (ctx->lane_next + 1) % ctx->n_lanes;
      ctx->screened = true;
      ctx->sites_screened += n;
      ctx->stats.n_launches += 3;
      ctx->n_async++;
      ctx->async_sites += n;
      return DNV_OK;
    }
    ln.parity = 0;  // the eager pass uses the words the graph will use
  }
  rc = enqueue_device_batch(ctx, d_counts, bps, n, site_base, nullptr, d_calls, cap, n_calls_out, /*split=*/false, false);
  if (rc) return rc;
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
  if (bad_alignment(d_counts, bps)) return DNV_E_INVALID;
  DNV_CUDA(cudaSetDevice(ctx->device));
  int rc = drain_async(ctx);
  if (rc) return rc;
  reset_call_state(ctx);
  // A call that is ONE chunk on the calls-only pipeline ends with its finish kernel: the kernel's last CTA hands the
  // accumulators to the host (CallChain::host_acc), no copy / memset behind it (DNV_NO_HOST_COLLECT=1: A/B knob)
  static const bool no_host_collect = getenv("DNV_NO_HOST_COLLECT") != nullptr || getenv("DNV_DEBUG_SKIP") != nullptr;
  const bool one = !no_host_collect && parallel(ctx) && !ctx->dense_mode && n > 0 && n <= chunk_for(ctx, n, true) &&
                   will_screen(ctx, d_dense != nullptr);
  // the whole call — chunks over the lanes, join, statistics to the host — is replayable when its arguments repeat
  bool launched = false;
  if (ctx->use_graphs && !ctx->dense_mode && n > 0 && will_screen(ctx, d_dense != nullptr)) {
    const GraphKey key{d_counts, d_calls, nullptr, d_dense, n, 0, cap, bps};
    GraphEntry* ge = graph_slot(ctx, ctx->sync_graphs, key);
    ge->seen++;
    ctx->lane_next = 0;
    for (int l = 0; l < ctx->n_lanes; l++) ctx->lane[l].parity = 0;
    if (!ge->exec && ge->seen >= 2) {
      DNV_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
      ctx->capturing = true;
      rc = begin_submit(ctx);
      if (!rc) rc = enqueue_device_batch(ctx, d_counts, bps, n, 0, d_dense, d_calls, cap, nullptr, /*split=*/true, false, one);
      if (!rc) rc = one ? flush_lanes(ctx) : enqueue_collect(ctx);
      rc = end_capture(ctx, ctx->stream, rc, &ge->exec);
      if (rc) return rc;
      ctx->lane_next = 0;
      for (int l = 0; l < ctx->n_lanes; l++) ctx->lane[l].parity = 0;
    }
    if (ge->exec) {
      DNV_CUDA(cudaGraphLaunch(ge->exec, ctx->stream));
      DNV_CUDA(cudaStreamSynchronize(ctx->stream));
      const int64_t chunk = chunk_for(ctx, n, true);
      ctx->stats.n_launches += 3 * ((n + chunk - 1) / chunk);
      ctx->screened = true;
      ctx->sites_screened += n;
      launched = true;
    }
  }
  if (!launched) {
    rc = begin_submit(ctx);
    if (rc) return rc;
    rc = enqueue_device_batch(ctx, d_counts, bps, n, 0, d_dense, d_calls, cap, nullptr, /*split=*/true, false, one);
    if (rc) return rc;
    if (one) {
      rc = flush_lanes(ctx);
      if (rc) return rc;
      DNV_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
      rc = collect(ctx);
      if (rc) return rc;
    }
  }
  rc = dnvi_finish_stats(ctx, n);
  if (rc) return rc;
  *n_calls = ctx->stats.n_calls;
  if (*n_calls > cap) {
    set_error("call capacity too small: need " + std::to_string(*n_calls) + ", have " + std::to_string(cap));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
}

static int infer_host_impl(dnv_ctx* ctx, const void* counts, int bps, int64_t n, dnv_site_result* dense,
                           dnv_call* calls, int64_t cap, int64_t* n_calls) {
  if (!ctx || !n_calls || n < 0 || cap < 0 || (n > 0 && !counts) || (cap > 0 && !calls)) {
    set_error("dnv_infer_counts: invalid argument");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(ctx->device));
  int rc = drain_async(ctx);
  if (rc) return rc;
  reset_call_state(ctx);
  const int64_t dev_cap = std::min(cap, n);
  rc = ensure_stage_calls(ctx, dev_cap);
  if (rc) return rc;
  rc = begin_submit(ctx);
  if (rc) return rc;
  rc = enqueue_host_batch(ctx, counts, bps, n, 0, dense, ctx->d_stage_calls, dev_cap, nullptr);
  if (rc) return rc;
  rc = finish_all(ctx, n);
  if (rc) return rc;
  *n_calls = ctx->stats.n_calls;
  const int64_t got = std::min<int64_t>(*n_calls, dev_cap);
  if (got > 0) {
    DNV_CUDA(cudaMemcpyAsync(calls, ctx->d_stage_calls, (size_t)got * sizeof(dnv_call), cudaMemcpyDeviceToHost,
                             ctx->stream));
    DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (*n_calls > cap) {
    set_error("call capacity too small: need " + std::to_string(*n_calls) + ", have " + std::to_string(cap));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
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
                                  int64_t* n_calls_out) {
  return infer_device_async_impl(ctx, d_counts, 24, n, 0, d_calls, cap, n_calls_out);
}
int dnv_infer_counts8_device_async(dnv_ctx* ctx, const uint8_t* d_counts8, int64_t n, dnv_call* d_calls, int64_t cap,
                                   int64_t* n_calls_out) {
  return infer_device_async_impl(ctx, d_counts8, 12, n, 0, d_calls, cap, n_calls_out);
}
int dnv_infer_counts_device_async_at(dnv_ctx* ctx, const uint16_t* d_counts, int64_t n, int64_t site_base, dnv_call* d_calls,
                                     int64_t cap, int64_t* n_calls_out) {
  return infer_device_async_impl(ctx, d_counts, 24, n, site_base, d_calls, cap, n_calls_out);
}
int dnv_infer_counts8_device_async_at(dnv_ctx* ctx, const uint8_t* d_counts8, int64_t n, int64_t site_base, dnv_call* d_calls,
                                      int64_t cap, int64_t* n_calls_out) {
  return infer_device_async_impl(ctx, d_counts8, 12, n, site_base, d_calls, cap, n_calls_out);
}
int dnv_flush(dnv_ctx* ctx) {
  if (!ctx) { set_error("dnv_flush: null ctx"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(ctx->device));
  return flush_lanes(ctx);
}
int dnv_sync(dnv_ctx* ctx) {
  if (!ctx) { set_error("dnv_sync: null ctx"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(ctx->device));
  if (ctx->n_async == 0) {
    { const int rc = flush_lanes(ctx); if (rc) return rc; }
    DNV_CUDA(cudaStreamSynchronize(ctx->stream));
    return DNV_OK;
  }
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

int dnv_alloc_pinned(size_t bytes, void** out) {
  if (!out) { set_error("dnv_alloc_pinned: null argument"); return DNV_E_INVALID; }
  *out = nullptr;
  DNV_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return DNV_OK;
}
int dnv_free_pinned(void* p) {
  if (p) DNV_CUDA(cudaFreeHost(p));
  return DNV_OK;
}

}  // extern "C"
