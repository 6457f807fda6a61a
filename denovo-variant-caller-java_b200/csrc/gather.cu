// gather.cu — multi-GPU under the C ABI: interval shards, the gather of call records to one device, and the
// single-process device group.
//
// Sites shard by genomic interval: rank r owns the r-th contiguous range of the axis, so
// rank order = genomic order and there is NO data-path collective.  The only exchange is the final gather of call
// records + their counts to one rank — what the reference does with one shared PrintWriter behind
// `synchronized writeCalls` (M/DenovoCaller.java:56-59, M/ReadCaller.java:122-129), fed by independent per-site
// tasks (M/ReadCaller.java:91-95) and per-thread position shards (M/VariantCaller.java:86-103).  Here it is FUSED
// into the producing kernels: the finish kernel of every rank stores its finished records, and the list kernel
// their count, straight into the destination rank's HBM over NVLink (peer stores); one release flag per rank and
// round tells the destination that everything has landed.  Ranks never wait for each other; only the destination
// waits, once per round, on the device.
//
//   dnv_gather      one rank's end of the gather.  The destination's end owns the buffer
//                   [flags | generation x rank x (64-byte header with the int64 count | cap records)];
//                   the other ends map it: cudaIpc* handles between processes (one process per GPU: torchrun,
//                   MPI — the 64-byte handle travels through whatever the launcher offers), peer access inside
//                   one process.
//   dnv_group       N devices driven by ONE process (the host CLI, the Java binding): a dnv_ctx and a gather end
//                   per device, host batches split by interval, one host thread per device for the enqueue.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "dnv_internal.cuh"

namespace dnv {

constexpr size_t kGatherFlagBytes = 1024;  // [0..63] arrival round of every rank, [64] first rank that timed out + 1
constexpr size_t kGatherHeader = 64;
constexpr int kGatherMaxWorld = 64;
constexpr unsigned long long kGatherTimeoutNs = 10ull * 1000 * 1000 * 1000;

__global__ void gather_arrive_kernel(unsigned long long* flags, int rank, unsigned long long round) {
  __threadfence_system();  // this rank's records and counts (earlier kernels of the stream) before the flag
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags + rank), "l"(round) : "memory");
}

// one warp on the destination: lane r waits until rank r has arrived at `round` (bounded: a dead rank must not
// hang the GPU — the failure is reported by dnv_gather_fetch)
__global__ void gather_wait_kernel(unsigned long long* flags, int world, unsigned long long round) {
  for (int r = threadIdx.x; r < world; r += 32) {
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
      if (v >= round) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > kGatherTimeoutNs) { atomicMax(flags + 64, (unsigned long long)(r + 1)); break; }
      __nanosleep(200);
    }
  }
}

}  // namespace dnv

using namespace dnv;

struct dnv_gather {
  dnv_ctx* ctx;
  int world, rank, dst, gens;
  int64_t cap;
  size_t slot_bytes, total_bytes;
  unsigned char* base;  // the destination's buffer as this rank's device sees it
  bool owns, ipc;
  unsigned long long round;  // arrivals (wait: rounds waited for) issued so far
};

static unsigned char* slot_ptr(const dnv_gather* g, int rank, int gen) {
  return g->base + kGatherFlagBytes + ((size_t)(gen % g->gens) * g->world + rank) * g->slot_bytes;
}

struct dnv_group {
  int n;
  dnv_ctx* ctx[kGatherMaxWorld];
  dnv_gather* end[kGatherMaxWorld];
  int64_t cap;  // records per rank the gather ends are sized for
  dnv_stats stats;
};

extern "C" {

int dnv_gather_create(dnv_ctx* ctx, int32_t world, int32_t rank, int32_t dst, int64_t cap, int32_t gens, dnv_gather** out) {
  if (!ctx || !out || world < 1 || world > kGatherMaxWorld || rank < 0 || rank >= world || dst < 0 || dst >= world ||
      cap < 0 || gens < 1) {
    set_error("dnv_gather_create: invalid argument (1 <= world <= 64, 0 <= rank, dst < world, gens >= 1)");
    return DNV_E_INVALID;
  }
  *out = nullptr;
  DNV_CUDA(cudaSetDevice(ctx->device));
  dnv_gather* g = new (std::nothrow) dnv_gather();
  if (!g) return DNV_E_NOMEM;
  memset(g, 0, sizeof(*g));
  g->ctx = ctx; g->world = world; g->rank = rank; g->dst = dst; g->gens = gens; g->cap = cap;
  g->slot_bytes = (kGatherHeader + (size_t)cap * sizeof(dnv_call) + 255) / 256 * 256;
  g->total_bytes = kGatherFlagBytes + (size_t)gens * world * g->slot_bytes;
  if (rank == dst) {
    if (cudaMalloc(&g->base, g->total_bytes) != cudaSuccess || cudaMemset(g->base, 0, g->total_bytes) != cudaSuccess) {
      set_error(std::string("dnv_gather_create: ") + cudaGetErrorString(cudaGetLastError()));
      cudaFree(g->base);
      delete g;
      return DNV_E_CUDA;
    }
    g->owns = true;
  }
  *out = g;
  return DNV_OK;
}

int dnv_gather_destroy(dnv_gather* g) {
  if (!g) return DNV_OK;
  cudaSetDevice(g->ctx->device);
  cudaDeviceSynchronize();
  if (g->owns) cudaFree(g->base);
  else if (g->ipc && g->base) cudaIpcCloseMemHandle(g->base);
  cudaGetLastError();
  delete g;
  return DNV_OK;
}

int dnv_gather_export(dnv_gather* g, void* handle64) {
  if (!g || !handle64 || !g->owns) { set_error("dnv_gather_export: only the destination rank's end exports"); return DNV_E_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == DNV_GATHER_HANDLE_BYTES, "IPC handle size");
  DNV_CUDA(cudaSetDevice(g->ctx->device));
  cudaIpcMemHandle_t h;
  DNV_CUDA(cudaIpcGetMemHandle(&h, g->base));
  memcpy(handle64, &h, sizeof(h));
  return DNV_OK;
}

int dnv_gather_import(dnv_gather* g, const void* handle64) {
  if (!g || !handle64 || g->owns || g->base) { set_error("dnv_gather_import: a non-destination end that is not mapped yet"); return DNV_E_INVALID; }
  DNV_CUDA(cudaSetDevice(g->ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  DNV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  g->base = (unsigned char*)p;
  g->ipc = true;
  return DNV_OK;
}

int dnv_gather_attach(dnv_gather* g, dnv_gather* dst_end) {
  if (!g || !dst_end || !dst_end->owns || g->owns || g->base || g->world != dst_end->world || g->cap != dst_end->cap ||
      g->gens != dst_end->gens) {
    set_error("dnv_gather_attach: needs an unmapped non-destination end and the destination end of the same gather");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(g->ctx->device));
  if (g->ctx->device != dst_end->ctx->device) {
    int can = 0;
    DNV_CUDA(cudaDeviceCanAccessPeer(&can, g->ctx->device, dst_end->ctx->device));
    if (!can) { set_error("dnv_gather_attach: no peer access between the two devices"); return DNV_E_CUDA; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(dst_end->ctx->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      return DNV_E_CUDA;
    }
    cudaGetLastError();
  }
  g->base = dst_end->base;
  return DNV_OK;
}

int dnv_gather_slot(dnv_gather* g, int32_t rank, int32_t gen, dnv_call** records, int64_t** count) {
  if (!g || !g->base || rank < 0 || rank >= g->world || gen < 0) { set_error("dnv_gather_slot: unmapped end or bad rank / generation"); return DNV_E_INVALID; }
  unsigned char* s = slot_ptr(g, rank, gen);
  if (count) *count = (int64_t*)s;
  if (records) *records = (dnv_call*)(s + kGatherHeader);
  return DNV_OK;
}

static int gather_infer(dnv_ctx* ctx, const void* d_counts, int bps, int64_t n, int64_t site_base, dnv_gather* g, int32_t gen) {
  if (!g || g->ctx != ctx || !g->base || gen < 0) { set_error("dnv_infer_counts_device_gather: unmapped gather end, or not this context's"); return DNV_E_INVALID; }
  unsigned char* s = slot_ptr(g, g->rank, gen);
  dnv_call* recs = (dnv_call*)(s + kGatherHeader);
  return bps == 24 ? dnv_infer_counts_device_async_at(ctx, (const uint16_t*)d_counts, n, site_base, recs, g->cap, (int64_t*)s)
                   : dnv_infer_counts8_device_async_at(ctx, (const uint8_t*)d_counts, n, site_base, recs, g->cap, (int64_t*)s);
}
int dnv_infer_counts_device_gather(dnv_ctx* ctx, const uint16_t* d_counts, int64_t n, int64_t site_base, dnv_gather* g, int32_t gen) {
  return gather_infer(ctx, d_counts, 24, n, site_base, g, gen);
}
int dnv_infer_counts8_device_gather(dnv_ctx* ctx, const uint8_t* d_counts8, int64_t n, int64_t site_base, dnv_gather* g, int32_t gen) {
  return gather_infer(ctx, d_counts8, 12, n, site_base, g, gen);
}

int dnv_gather_arrive(dnv_gather* g) {
  if (!g || !g->base) { set_error("dnv_gather_arrive: unmapped gather end"); return DNV_E_INVALID; }
  dnv_ctx* ctx = g->ctx;
  DNV_CUDA(cudaSetDevice(ctx->device));
  const int rc = flush_lanes(ctx);
  if (rc) return rc;
  gather_arrive_kernel<<<1, 1, 0, ctx->stream>>>((unsigned long long*)g->base, g->rank, ++g->round);
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

int dnv_gather_wait(dnv_gather* g) {
  if (!g || !g->owns) { set_error("dnv_gather_wait: only the destination rank waits"); return DNV_E_INVALID; }
  dnv_ctx* ctx = g->ctx;
  DNV_CUDA(cudaSetDevice(ctx->device));
  gather_wait_kernel<<<1, 32, 0, ctx->stream>>>((unsigned long long*)g->base, g->world, g->round);
  DNV_CUDA(cudaGetLastError());
  return DNV_OK;
}

int dnv_gather_fetch(dnv_gather* g, int32_t gen, dnv_call* calls, int64_t cap, int64_t* n_calls, int64_t* per_rank) {
  if (!g || !g->owns || !n_calls || cap < 0 || (cap > 0 && !calls) || gen < 0) { set_error("dnv_gather_fetch: only on the destination rank"); return DNV_E_INVALID; }
  dnv_ctx* ctx = g->ctx;
  DNV_CUDA(cudaSetDevice(ctx->device));
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  unsigned long long timed_out = 0;
  DNV_CUDA(cudaMemcpy(&timed_out, g->base + 64 * 8, 8, cudaMemcpyDeviceToHost));
  if (timed_out) {
    set_error("gather: rank " + std::to_string(timed_out - 1) + " did not arrive within 10 s");
    cudaMemset(g->base + 64 * 8, 0, 8);
    return DNV_E_CUDA;
  }
  std::vector<int64_t> cnt((size_t)g->world);
  DNV_CUDA(cudaMemcpy2D(cnt.data(), 8, slot_ptr(g, 0, gen), g->slot_bytes, 8, (size_t)g->world, cudaMemcpyDeviceToHost));
  int64_t total = 0;
  bool over = false;
  for (int r = 0; r < g->world; r++) {
    if (per_rank) per_rank[r] = cnt[(size_t)r];
    if (cnt[(size_t)r] > g->cap) over = true;  // that rank produced more records than a slot holds
    const int64_t have = std::min(cnt[(size_t)r], g->cap);
    const int64_t room = std::max<int64_t>(0, std::min(have, cap - total));
    if (room > 0)
      DNV_CUDA(cudaMemcpyAsync(calls + total, slot_ptr(g, r, gen) + kGatherHeader, (size_t)room * sizeof(dnv_call),
                               cudaMemcpyDeviceToHost, ctx->stream));
    total += cnt[(size_t)r];
  }
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_calls = total;
  if (over || total > cap) {
    set_error("gather: call capacity too small: need " + std::to_string(total) + " in all" + (over ? " (a rank's slot overflowed)" : ""));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
}

// ---- dnv_group ---------------------------------------------------------------------------------------------------

int dnv_group_destroy(dnv_group* grp) {
  if (!grp) return DNV_OK;
  for (int r = 0; r < grp->n; r++) dnv_gather_destroy(grp->end[r]);
  for (int r = 0; r < grp->n; r++) dnv_destroy(grp->ctx[r]);
  delete grp;
  return DNV_OK;
}

int dnv_group_create(const dnv_params* params, const int32_t* devices, int32_t n_devices, dnv_group** out) {
  if (!params || !devices || !out || n_devices < 1 || n_devices > kGatherMaxWorld) {
    set_error("dnv_group_create: invalid argument (1 <= n_devices <= 64)");
    return DNV_E_INVALID;
  }
  *out = nullptr;
  dnv_group* grp = new (std::nothrow) dnv_group();
  if (!grp) return DNV_E_NOMEM;
  memset(grp, 0, sizeof(*grp));
  for (int r = 0; r < n_devices; r++) {
    dnv_params p = *params;
    p.device = devices[r];
    const int rc = dnv_create(&p, &grp->ctx[r]);
    if (rc) { const std::string msg = dnv_last_error(); dnv_group_destroy(grp); set_error(msg); return rc; }
    grp->n = r + 1;
  }
  *out = grp;
  return DNV_OK;
}

static int group_ensure_gather(dnv_group* grp, int64_t cap) {
  if (grp->end[0] && cap <= grp->cap) return DNV_OK;
  for (int r = 0; r < grp->n; r++) { dnv_gather_destroy(grp->end[r]); grp->end[r] = nullptr; }
  grp->cap = 0;
  for (int r = 0; r < grp->n; r++) {
    const int rc = dnv_gather_create(grp->ctx[r], grp->n, r, 0, cap, 1, &grp->end[r]);
    if (rc) return rc;
    if (r > 0) { const int rc2 = dnv_gather_attach(grp->end[r], grp->end[0]); if (rc2) return rc2; }
  }
  grp->cap = cap;
  return DNV_OK;
}

static int group_infer(dnv_group* grp, const void* counts, int bps, int64_t n, dnv_call* calls, int64_t cap, int64_t* n_calls) {
  if (!grp || !n_calls || n < 0 || cap < 0 || (n > 0 && !counts) || (cap > 0 && !calls)) {
    set_error("dnv_group_infer_counts: invalid argument");
    return DNV_E_INVALID;
  }
  const int world = grp->n;
  const int64_t base = n / world, rem = n % world;
  auto lo_of = [&](int r) { return (int64_t)r * base + std::min<int64_t>(r, rem); };
  // a rank cannot produce more records than it has sites, nor more than the caller wants in all
  int rc = group_ensure_gather(grp, std::max<int64_t>(1, std::min(cap, base + 1)));
  if (rc) return rc;
  // one host thread per device: the enqueue (and, for pageable input, the staged host-to-device copies) of the
  // shards proceed in parallel
  std::vector<int> rcs((size_t)world, DNV_OK);
  std::vector<std::string> msgs((size_t)world);
  auto work = [&](int r) {
    dnv_ctx* ctx = grp->ctx[r];
    const int64_t lo = lo_of(r), m = lo_of(r + 1) - lo;
    dnv_call* recs = nullptr;
    int64_t* cnt = nullptr;
    int e = dnv_gather_slot(grp->end[r], r, 0, &recs, &cnt);
    if (!e) e = cudaSetDevice(ctx->device) == cudaSuccess ? DNV_OK : DNV_E_CUDA;
    if (!e) { reset_call_state(ctx); e = begin_submit(ctx); }
    if (!e) e = enqueue_host_batch(ctx, (const unsigned char*)counts + lo * bps, bps, m, lo, nullptr, recs, grp->cap, cnt);
    if (!e) e = dnv_gather_arrive(grp->end[r]);
    rcs[(size_t)r] = e;
    if (e) msgs[(size_t)r] = dnv_last_error();
  };
  if (world == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++) th.emplace_back(work, r);
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < world; r++)
    if (rcs[(size_t)r]) { set_error("device " + std::to_string(grp->ctx[r]->device) + ": " + msgs[(size_t)r]); return rcs[(size_t)r]; }
  rc = dnv_gather_wait(grp->end[0]);
  if (rc) return rc;
  const int frc = dnv_gather_fetch(grp->end[0], 0, calls, cap, n_calls, nullptr);
  // statistics: every context is collected (that is also the host wait for ranks > 0)
  grp->stats = dnv_stats();
  for (int r = 0; r < world; r++) {
    dnv_ctx* ctx = grp->ctx[r];
    rc = cudaSetDevice(ctx->device) == cudaSuccess ? collect(ctx) : DNV_E_CUDA;
    if (rc) return rc;
    const int64_t m = lo_of(r + 1) - lo_of(r);
    dnvi_finish_stats(ctx, m);
    grp->stats.n_sites += m;
    grp->stats.n_calls += ctx->stats.n_calls;
    grp->stats.n_hard += ctx->stats.n_hard;
    grp->stats.n_screened_out += ctx->stats.n_screened_out;
    grp->stats.n_launches += ctx->stats.n_launches;
  }
  return frc;
}

int dnv_group_infer_counts(dnv_group* grp, const uint16_t* counts, int64_t n, dnv_call* calls, int64_t cap, int64_t* n_calls) {
  return group_infer(grp, counts, 24, n, calls, cap, n_calls);
}
int dnv_group_infer_counts8(dnv_group* grp, const uint8_t* counts8, int64_t n, dnv_call* calls, int64_t cap, int64_t* n_calls) {
  return group_infer(grp, counts8, 12, n, calls, cap, n_calls);
}
int dnv_group_size(dnv_group* grp) { return grp ? grp->n : 0; }
int dnv_group_ctx(dnv_group* grp, int32_t rank, dnv_ctx** out) {
  if (!grp || !out || rank < 0 || rank >= grp->n) { set_error("dnv_group_ctx: bad rank"); return DNV_E_INVALID; }
  *out = grp->ctx[rank];
  return DNV_OK;
}
int dnv_group_get_stats(dnv_group* grp, dnv_stats* out) {
  if (!grp || !out) { set_error("dnv_group_get_stats: null argument"); return DNV_E_INVALID; }
  *out = grp->stats;
  return DNV_OK;
}

}  // extern "C"
