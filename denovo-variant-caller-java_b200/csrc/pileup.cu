// pileup.cu — kernel (a): reads -> packed per-site A/C/G/T counts.
//
// Batched `new ReadSummary(reads, position)` for DAD, MOM, CHILD
// (M/ReadSummary.java:45-59, M/ReadCaller.java:138-146): for every candidate
// position, every read of the person that overlaps it contributes
// alignedSequence[pos - readStart]; '-' is skipped; anything outside ACGT is the
// reference's IllegalArgumentException (reported as DNV_E_BADBASE).
//
// Layout: reads sorted by (person, start); per person the reads overlapping a position are a
// contiguous window found by binary search, so the kernel is a GATHER: every (site, person)
// count vector is accumulated in registers by one thread and written once — no atomics, no
// inter-thread contention, deterministic.  The kernels live in pileup_pipe.cuh (default: producer /
// consumer pipeline over dense tiles), pileup_stage.cuh and pileup_scatter.cuh; this file holds the
// validation / metadata pass over the read descriptors and the C entry points.
#include <stdlib.h>
#include <algorithm>
#include "dnv_internal.cuh"
#include "pileup_stage.cuh"
#include "pileup_pipe.cuh"

namespace dnv {

// meta[0..3] person offsets, meta[4] max read length, meta[5] error word (pileup_err: index and kind in ONE atomicMax)
__global__ void __launch_bounds__(256)
pileup_prep_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const uint8_t* __restrict__ person, int64_t n_reads, long long* __restrict__ meta) {
  long long lmax = 0;
  // four reads per thread and trip, every load of a trip issued before the first is used (the kernel is a
  // 17 B/read stream: it wants bytes in flight, not instructions)
  constexpr int kU = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < n_reads; r0 += kU * stride) {
    long long b0[kU], b1[kU], s0[kU], s1[kU];
    int pc[kU], pv[kU];
#pragma unroll
    for (int k = 0; k < kU; k++) {
      const int64_t r = r0 + k * stride;
      const bool ok = r < n_reads, has_prev = ok && r > 0;
      b0[k] = ok ? __ldg(base_off + r) : 0;
      b1[k] = ok ? __ldg(base_off + r + 1) : 0;
      s1[k] = ok ? __ldg(read_start + r) : 0;
      s0[k] = has_prev ? __ldg(read_start + r - 1) : 0;
      pc[k] = ok ? (int)__ldg(person + r) : 0;
      pv[k] = has_prev ? (int)__ldg(person + r - 1) : 0;
    }
#pragma unroll
    for (int k = 0; k < kU; k++) {
      const int64_t r = r0 + k * stride;
      if (r >= n_reads) continue;
      const long long len = b1[k] - b0[k];
      lmax = len > lmax ? len : lmax;
      const int p = pc[k];
      if (p > 2 || len < 0) { atomicMax((unsigned long long*)&meta[5], pileup_err(r, false, kPileErrInvalid)); continue; }
      if (r == 0) {
        for (int q = 0; q <= p; q++) meta[q] = 0;
      } else {
        const int pp = pv[k];
        if (pp > p || (pp == p && s0[k] > s1[k])) {
          atomicMax((unsigned long long*)&meta[5], pileup_err(r, false, kPileErrUnsorted));
        }
        for (int q = pp + 1; q <= p; q++) meta[q] = r;
      }
      if (r == n_reads - 1)
        for (int q = p + 1; q <= 3; q++) meta[q] = n_reads;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long y = __shfl_xor_sync(0xffffffffu, lmax, o);
    lmax = y > lmax ? y : lmax;
  }
  if ((threadIdx.x & 31) == 0 && lmax > 0) atomicMax(&meta[4], lmax);
}

}  // namespace dnv

using namespace dnv;

extern "C" {

static int pileup_status(long long* h);

// Everything dnv_pileup_device does on the device, enqueued on the context's stream without a host wait: validation /
// metadata pass, windows pre-pass, pileup kernel, and the status words (meta[0..7]) on their way to the context's
// pinned status buffer.  The caller waits for the stream and then decodes them with pileup_status(pile_host(ctx)).
static long long* pile_host(dnv_ctx* ctx) { return reinterpret_cast<long long*>(ctx->h_acc + 8); }

static int pileup_enqueue(dnv_ctx* ctx, const int64_t* d_read_start, const int64_t* d_base_off, const char* d_bases,
                          const uint8_t* d_person, int64_t n_reads, const int64_t* d_site_pos, int64_t n_sites,
                          uint16_t* d_counts_out) {
  // scratch: meta[8] (see pileup_prep_kernel), then one record per tile of the pipeline kernel
  const int64_t n_tiles = (n_sites + kPP_Consumers * 8 - 1) / (kPP_Consumers * 8);
  int rc = ensure_scratch(ctx, 64 + (size_t)n_tiles * kPP_TileWords * sizeof(long long));
  if (rc) return rc;
  long long* meta = (long long*)ctx->d_scratch;
  DNV_CUDA(cudaMemsetAsync(meta, 0, 64, ctx->stream));
  const int sms = ctx->sms;
  if (n_reads > 0) {
    const int grid = (int)std::min<int64_t>((n_reads + 255) / 256, (int64_t)sms * 8);
    pileup_prep_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_person, n_reads, meta);
  }
  // dense tiles: the producer / consumer pipeline of pileup_pipe.cuh (it bulk-copies slices of read_start / base_off:
  // 16-byte aligned arrays); otherwise, or with DNV_PILEUP_KERNEL=stage, the position-major gather of
  // pileup_stage.cuh; DNV_PILEUP_KERNEL=scatter selects the read-major formulation with shared-memory atomics
  // (pileup_scatter.cuh — its tile routine is what the fused reads -> calls kernel builds on); sparse tiles: one site
  // at a time in all of them
  static const std::string which = [] { const char* e = getenv("DNV_PILEUP_KERNEL"); return std::string(e ? e : "pipe"); }();
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(d_read_start) | reinterpret_cast<uintptr_t>(d_base_off)) & 15) == 0;
  bool status_sent = false;
  if (which == "pipe" && aligned16) {
    // (the pipeline kernel's last warp to leave writes the status words to the pinned buffer itself)
    DNV_CUDA(launch_pileup_pipe<8>(ctx->stream, sms, d_read_start, d_base_off, d_bases, meta, meta + 8, d_site_pos, n_sites,
                                   n_reads, d_counts_out, pile_host(ctx)));
    status_sent = true;
  } else if (which == "pipe" || which == "stage") {
    DNV_CUDA(launch_pileup_stage<8>(ctx->stream, sms, d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites, n_reads,
                                    d_counts_out));
  } else {
    const int rc2 = launch_pileup_scatter(ctx, ctx->stream, d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites, n_reads,
                                          d_counts_out);
    if (rc2) return rc2;
  }
  DNV_CUDA(cudaGetLastError());
  if (!status_sent) DNV_CUDA(cudaMemcpyAsync(pile_host(ctx), meta, 8 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  return DNV_OK;
}

int dnv_pileup_device(dnv_ctx* ctx, const int64_t* d_read_start, const int64_t* d_base_off, const char* d_bases,
                      const uint8_t* d_person, int64_t n_reads, const int64_t* d_site_pos, int64_t n_sites,
                      uint16_t* d_counts_out) {
  if (!ctx || n_reads < 0 || n_sites < 0 || (n_sites > 0 && (!d_site_pos || !d_counts_out)) ||
      (n_reads > 0 && (!d_read_start || !d_base_off || !d_bases || !d_person))) {
    set_error("dnv_pileup_device: invalid argument");
    return DNV_E_INVALID;
  }
  if (n_sites == 0) return DNV_OK;
  if ((reinterpret_cast<uintptr_t>(d_counts_out) & 7) || (reinterpret_cast<uintptr_t>(d_read_start) & 7) ||
      (reinterpret_cast<uintptr_t>(d_base_off) & 7) || (reinterpret_cast<uintptr_t>(d_site_pos) & 7)) {
    set_error("dnv_pileup_device: counts, read_start, base_off and site_pos must be 8-byte aligned");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(ctx->device));
  const int rc = pileup_enqueue(ctx, d_read_start, d_base_off, d_bases, d_person, n_reads, d_site_pos, n_sites, d_counts_out);
  if (rc) return rc;
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  return pileup_status(pile_host(ctx));
}

// decodes the error word the pileup kernels leave in meta[5]
static int pileup_status(long long* h) {
  if (h[5] != 0) {
    const unsigned long long w = (unsigned long long)h[5];
    const int kind = (w & 7) == kPileErrBadBase ? DNV_E_BADBASE : (w & 7) == kPileErrUnsorted ? DNV_E_UNSORTED : DNV_E_INVALID;
    const bool at_site = (w >> 3) & 1;
    h[5] = (long long)(w >> 4);
    set_error(kind == DNV_E_BADBASE
                  ? std::string("pileup: base outside {A,C,G,T,-} ") + (at_site ? "at site " : "in read ") + std::to_string(h[5] - 1) +
                        " (the reference throws IllegalArgumentException, M/ReadSummary.java:55)"
                  : kind == DNV_E_UNSORTED ? "pileup: reads must be sorted by (person, start); first violation at read " +
                                                 std::to_string(h[5] - 1)
                                           : "pileup: invalid read record or depth above 65535");
    return kind;
  }
  return DNV_OK;
}

int dnv_pileup(dnv_ctx* ctx, const int64_t* read_start, const int64_t* base_off, const char* bases,
               const uint8_t* person, int64_t n_reads, const int64_t* site_pos, int64_t n_sites,
               uint16_t* counts_out) {
  if (!ctx || n_reads < 0 || n_sites < 0 || (n_sites > 0 && (!site_pos || !counts_out)) ||
      (n_reads > 0 && (!read_start || !base_off || !bases || !person))) {
    set_error("dnv_pileup: invalid argument");
    return DNV_E_INVALID;
  }
  if (n_sites == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  const int64_t n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  int64_t *d_start = nullptr, *d_off = nullptr, *d_pos = nullptr;
  char* d_bases = nullptr;
  uint8_t* d_person = nullptr;
  uint16_t* d_counts = nullptr;
  int rc = DNV_OK;
#define PL_TRY(x) do { if (rc == DNV_OK && (x) != cudaSuccess) { set_error(std::string("dnv_pileup: ") + cudaGetErrorString(cudaGetLastError())); rc = DNV_E_CUDA; } } while (0)
  PL_TRY(cudaMalloc(&d_start, (size_t)std::max<int64_t>(n_reads, 1) * 8));
  PL_TRY(cudaMalloc(&d_off, (size_t)(n_reads + 1) * 8));
  PL_TRY(cudaMalloc(&d_bases, (size_t)std::max<int64_t>(n_bases, 1)));
  PL_TRY(cudaMalloc(&d_person, (size_t)std::max<int64_t>(n_reads, 1)));
  PL_TRY(cudaMalloc(&d_pos, (size_t)n_sites * 8));
  PL_TRY(cudaMalloc(&d_counts, (size_t)n_sites * 24));
  if (n_reads > 0) {
    PL_TRY(cudaMemcpyAsync(d_start, read_start, (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
    PL_TRY(cudaMemcpyAsync(d_off, base_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    PL_TRY(cudaMemcpyAsync(d_bases, bases, (size_t)n_bases, cudaMemcpyHostToDevice, ctx->stream));
    PL_TRY(cudaMemcpyAsync(d_person, person, (size_t)n_reads, cudaMemcpyHostToDevice, ctx->stream));
  }
  PL_TRY(cudaMemcpyAsync(d_pos, site_pos, (size_t)n_sites * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (rc == DNV_OK) rc = dnv_pileup_device(ctx, d_start, d_off, d_bases, d_person, n_reads, d_pos, n_sites, d_counts);
  PL_TRY(cudaMemcpyAsync(counts_out, d_counts, (size_t)n_sites * 24, cudaMemcpyDeviceToHost, ctx->stream));
  PL_TRY(cudaStreamSynchronize(ctx->stream));
#undef PL_TRY
  cudaFree(d_start); cudaFree(d_off); cudaFree(d_bases); cudaFree(d_person); cudaFree(d_pos); cudaFree(d_counts);
  return rc;
}

// ---- reads -> calls, fused -------------------------------------------------------------
// ReadCaller.runBayesDenovoInference for a whole batch of candidate sites (M/ReadCaller.java:111-130):
// getReadSummaryMap (`new ReadSummary(reads, position)` x3, :138-146) and BayesInfer.infer (:119-120) in ONE pass —
// the pileup of a tile of positions is accumulated in shared memory, screened there by the exact integer screen, and
// only the ~1 % undecided sites leave the SM (counts + sorted site list) for the fused FP64 list kernel.
int dnv_infer_reads_device(dnv_ctx* ctx, const int64_t* d_read_start, const int64_t* d_base_off, const char* d_bases,
                           const uint8_t* d_person, int64_t n_reads, const int64_t* d_site_pos, int64_t n_sites,
                           dnv_call* d_calls, int64_t cap, int64_t* n_calls) {
  if (!ctx || !n_calls || n_reads < 0 || n_sites < 0 || cap < 0 || (n_sites > 0 && !d_site_pos) || (cap > 0 && !d_calls) ||
      (n_reads > 0 && (!d_read_start || !d_base_off || !d_bases || !d_person))) {
    set_error("dnv_infer_reads_device: invalid argument");
    return DNV_E_INVALID;
  }
  *n_calls = 0;
  if (n_sites == 0) return DNV_OK;
  if ((reinterpret_cast<uintptr_t>(d_read_start) & 7) || (reinterpret_cast<uintptr_t>(d_base_off) & 7) ||
      (reinterpret_cast<uintptr_t>(d_site_pos) & 7)) {
    set_error("dnv_infer_reads_device: read_start, base_off and site_pos must be 8-byte aligned");
    return DNV_E_INVALID;
  }
  DNV_CUDA(cudaSetDevice(ctx->device));
  { const int rc = dnv_sync(ctx); if (rc) return rc; }  // batches of the asynchronous entry points are collected first
  // the per-site counts: only the rows of undecided sites are ever written (fused path)
  if (n_sites > ctx->reads_counts_cap) {
    cudaFree(ctx->d_reads_counts);
    ctx->d_reads_counts = nullptr;
    ctx->reads_counts_cap = 0;
    DNV_CUDA(cudaMalloc(&ctx->d_reads_counts, (size_t)n_sites * 24));
    ctx->reads_counts_cap = n_sites;
  }
  // Default: the pileup pipeline into the context's counts buffer, then the READ pipeline on it — fusing the two saves
  // only the 48 B/site of writing and re-reading the counts (of ~185), while the pileup formulation that CAN be fused
  // with the screen (read-major scatter, pileup_scatter.cuh) costs 2.4x the instructions of the pipeline kernel.
  // DNV_READS_FUSED=1 selects the single-kernel variant (pileup_screen_kernel) where the screen applies.
  const char* fused_env = getenv("DNV_READS_FUSED");
  const bool want_fused = fused_env && fused_env[0] == '1';
  if (!want_fused || !will_screen(ctx, false) || ctx->dense_mode) {
    // one host wait for both: the pileup's status words reach the host in stream order before the statistics of
    // the inference, which starts on the device as soon as the counts are complete (a pileup that failed leaves
    // counts that are numbers all the same; its error is what the caller gets)
    int rc = pileup_enqueue(ctx, d_read_start, d_base_off, d_bases, d_person, n_reads, d_site_pos, n_sites, ctx->d_reads_counts);
    if (rc) return rc;
    const int rc_infer = dnv_infer_counts_device(ctx, ctx->d_reads_counts, n_sites, nullptr, d_calls, cap, n_calls);
    DNV_CUDA(cudaStreamSynchronize(ctx->stream));  // (already idle unless the inference failed before its own wait)
    rc = pileup_status(pile_host(ctx));
    if (rc) { *n_calls = 0; return rc; }
    return rc_infer;
  }
  int rc = ensure_scratch(ctx, 64);
  if (rc) return rc;
  long long* meta = (long long*)ctx->d_scratch;
  DNV_CUDA(cudaMemsetAsync(meta, 0, 64, ctx->stream));
  if (n_reads > 0) {
    const int grid = (int)std::min<int64_t>((n_reads + 255) / 256, (int64_t)ctx->sms * 8);
    pileup_prep_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_person, n_reads, meta);
  }
  reset_call_state(ctx);
  rc = begin_submit(ctx);  // the lanes start after the validation / metadata pass
  if (rc) return rc;
  ctx->count_bytes = 24;
  ReadsArgs ra{d_read_start, d_base_off, d_bases, n_reads, d_site_pos, meta};
  const int64_t chunk = 1ll << 27;  // a multiple of the fused kernel's tile
  for (int64_t off = 0; off < n_sites; off += chunk) {
    ReadsArgs rc_args = ra;
    rc_args.site_pos = d_site_pos + off;
    ChunkArgs a;
    a.d_counts = ctx->d_reads_counts + off * 12;
    a.site_base = off;
    a.n = std::min(chunk, n_sites - off);
    a.d_dense = nullptr;
    a.d_calls = d_calls;
    a.cap = cap;
    a.first = off == 0;
    a.last = off + chunk >= n_sites;
    a.no_sub_wait = false;
    a.tail_alone = a.last;
    a.wait_ev = nullptr;
    a.count_out = nullptr;
    a.reads = &rc_args;
    a.dense_host = nullptr;
    a.done_ev = nullptr;
    rc = launch_trio_chunk(ctx, a);
    if (rc) return rc;
  }
  rc = collect(ctx);
  if (rc) return rc;
  dnvi_finish_stats(ctx, n_sites);
  long long h[8];
  DNV_CUDA(cudaMemcpyAsync(h, meta, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  rc = pileup_status(h);
  if (rc) return rc;
  *n_calls = ctx->stats.n_calls;
  if (*n_calls > cap) {
    set_error("call capacity too small: need " + std::to_string(*n_calls) + ", have " + std::to_string(cap));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
}

int dnv_infer_reads(dnv_ctx* ctx, const int64_t* read_start, const int64_t* base_off, const char* bases,
                    const uint8_t* person, int64_t n_reads, const int64_t* site_pos, int64_t n_sites, dnv_call* calls,
                    int64_t cap, int64_t* n_calls) {
  if (!ctx || !n_calls || n_reads < 0 || n_sites < 0 || cap < 0 || (n_sites > 0 && !site_pos) || (cap > 0 && !calls) ||
      (n_reads > 0 && (!read_start || !base_off || !bases || !person))) {
    set_error("dnv_infer_reads: invalid argument");
    return DNV_E_INVALID;
  }
  *n_calls = 0;
  if (n_sites == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  const int64_t n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  const int64_t dev_cap = std::min(cap, n_sites);
  int64_t *d_start = nullptr, *d_off = nullptr, *d_pos = nullptr;
  char* d_bases = nullptr;
  uint8_t* d_person = nullptr;
  dnv_call* d_calls = nullptr;
  int rc = DNV_OK;
#define RD_TRY(x) do { if (rc == DNV_OK && (x) != cudaSuccess) { set_error(std::string("dnv_infer_reads: ") + cudaGetErrorString(cudaGetLastError())); rc = DNV_E_CUDA; } } while (0)
  RD_TRY(cudaMalloc(&d_start, (size_t)std::max<int64_t>(n_reads, 1) * 8));
  RD_TRY(cudaMalloc(&d_off, (size_t)(n_reads + 1) * 8));
  RD_TRY(cudaMalloc(&d_bases, (size_t)std::max<int64_t>(n_bases, 1)));
  RD_TRY(cudaMalloc(&d_person, (size_t)std::max<int64_t>(n_reads, 1)));
  RD_TRY(cudaMalloc(&d_pos, (size_t)n_sites * 8));
  RD_TRY(cudaMalloc(&d_calls, (size_t)std::max<int64_t>(dev_cap, 1) * sizeof(dnv_call)));
  if (n_reads > 0) {
    RD_TRY(cudaMemcpyAsync(d_start, read_start, (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
    RD_TRY(cudaMemcpyAsync(d_off, base_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    RD_TRY(cudaMemcpyAsync(d_bases, bases, (size_t)n_bases, cudaMemcpyHostToDevice, ctx->stream));
    RD_TRY(cudaMemcpyAsync(d_person, person, (size_t)n_reads, cudaMemcpyHostToDevice, ctx->stream));
  }
  RD_TRY(cudaMemcpyAsync(d_pos, site_pos, (size_t)n_sites * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (rc == DNV_OK) rc = dnv_infer_reads_device(ctx, d_start, d_off, d_bases, d_person, n_reads, d_pos, n_sites, d_calls, dev_cap, n_calls);
  const int64_t got = std::min<int64_t>(*n_calls, dev_cap);
  if ((rc == DNV_OK || rc == DNV_E_CAPACITY) && got > 0) {
    const int keep = rc;
    rc = DNV_OK;
    RD_TRY(cudaMemcpyAsync(calls, d_calls, (size_t)got * sizeof(dnv_call), cudaMemcpyDeviceToHost, ctx->stream));
    RD_TRY(cudaStreamSynchronize(ctx->stream));
    if (rc == DNV_OK) rc = keep;
  }
  if (rc == DNV_OK && *n_calls > cap) {
    set_error("call capacity too small: need " + std::to_string(*n_calls) + ", have " + std::to_string(cap));
    rc = DNV_E_CAPACITY;
  }
  cudaStreamSynchronize(ctx->stream);
#undef RD_TRY
  cudaFree(d_start); cudaFree(d_off); cudaFree(d_bases); cudaFree(d_person); cudaFree(d_pos); cudaFree(d_calls);
  return rc;
}

}  // extern "C"
