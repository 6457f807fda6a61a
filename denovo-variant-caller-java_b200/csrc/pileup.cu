// pileup.cu — kernel (a): reads -> packed per-site A/C/G/T counts.
//
// Batched `new ReadSummary(reads, position)` for DAD, MOM, CHILD
// (M/ReadSummary.java:45-59, M/ReadCaller.java:138-146): for every candidate
// position, every read of the person that overlaps it contributes
// alignedSequence[pos - readStart]; '-' is skipped; anything outside ACGT is the
// reference's IllegalArgumentException (reported as DNV_E_BADBASE).
//
// Layout: reads sorted by (person, start); per person the overlapping reads of a
// position are a contiguous window found by binary search, so the kernel is a
// GATHER: one thread owns one (site, person) count vector in registers and
// writes it once — no atomics, no inter-thread contention, deterministic.  For
// dense candidate lists neighbouring threads walk the same reads at neighbouring
// byte offsets, so the byte loads of a warp fall in the same 128-byte lines.
#include <stdlib.h>
#include <algorithm>
#include "dnv_internal.cuh"

namespace dnv {

// meta[0..3] person offsets, meta[4] max read length, meta[5] error (1+read index), meta[6] error kind
__global__ void __launch_bounds__(256)
pileup_prep_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const uint8_t* __restrict__ person, int64_t n_reads, long long* __restrict__ meta) {
  long long lmax = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
    const long long len = base_off[r + 1] - base_off[r];
    lmax = len > lmax ? len : lmax;
    const int p = person[r];
    if (p > 2 || len < 0) { atomicMax((unsigned long long*)&meta[5], (unsigned long long)(r + 1)); meta[6] = DNV_E_INVALID; continue; }
    if (r == 0) {
      for (int q = 0; q <= p; q++) meta[q] = 0;
    } else {
      const int pp = person[r - 1];
      if (pp > p || (pp == p && read_start[r - 1] > read_start[r])) {
        atomicMax((unsigned long long*)&meta[5], (unsigned long long)(r + 1));
        meta[6] = DNV_E_UNSORTED;
      }
      for (int q = pp + 1; q <= p; q++) meta[q] = r;
    }
    if (r == n_reads - 1)
      for (int q = p + 1; q <= 3; q++) meta[q] = n_reads;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long y = __shfl_xor_sync(0xffffffffu, lmax, o);
    lmax = y > lmax ? y : lmax;
  }
  if ((threadIdx.x & 31) == 0 && lmax > 0) atomicMax(&meta[4], lmax);
}

__global__ void __launch_bounds__(256)
pileup_gather_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                     const char* __restrict__ bases, long long* meta,
                     const int64_t* __restrict__ site_pos, int64_t n_sites, uint16_t* __restrict__ counts) {
  const long long* meta_in = meta;
  const long long lmax = meta_in[4];
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sites; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = site_pos[s];
    uint32_t out[12];
#pragma unroll
    for (int p = 0; p < 3; p++) {
      const int64_t r0 = meta_in[p], r1 = meta_in[p + 1];
      // first read of this person that can still reach pos: start > pos - lmax
      int64_t lo = r0, hi = r1;
      const int64_t key = pos - lmax;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(read_start + mid) <= key) lo = mid + 1; else hi = mid;
      }
      uint32_t cnt[4] = {0, 0, 0, 0};
      for (int64_t r = lo; r < r1; r++) {
        const int64_t st = __ldg(read_start + r);
        if (st > pos) break;
        const int64_t b0 = __ldg(base_off + r);
        const int64_t off = pos - st;  // M/ReadSummary.java:48
        if (off >= __ldg(base_off + r + 1) - b0) continue;  // read ends before pos
        const char b = __ldg(bases + b0 + off);  // :49
        if (b == '-') continue;                  // :51-53
        const int a = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;  // :55
        if (a < 0) {
          atomicMax((unsigned long long*)&meta[5], (unsigned long long)(r + 1));
          meta[6] = DNV_E_BADBASE;
          continue;
        }
        cnt[0] += a == 0; cnt[1] += a == 1; cnt[2] += a == 2; cnt[3] += a == 3;
      }
#pragma unroll
      for (int a = 0; a < 4; a++) {
        if (cnt[a] > 65535u) { meta[6] = DNV_E_INVALID; atomicMax((unsigned long long*)&meta[5], 1ull); cnt[a] = 65535u; }
        out[p * 4 + a] = cnt[a];
      }
    }
    uint2* dst = reinterpret_cast<uint2*>(counts + s * 12);
#pragma unroll
    for (int q = 0; q < 3; q++)
      dst[q] = make_uint2(out[4 * q] | (out[4 * q + 1] << 16), out[4 * q + 2] | (out[4 * q + 3] << 16));
  }
}

// Warp-cooperative variant for dense candidate lists (the every-position / dense-candidate
// layouts of SURVEY.md E.3): a warp owns 32 consecutive sites whose positions span less than
// ~2 read lengths.  The reads overlapping the warp's window are walked ONCE per warp: lanes
// load 32 read descriptors at a time (coalesced), each descriptor is broadcast by shuffle, and
// every lane fetches its own base of that read — 32 neighbouring bytes, one sector, and the
// loads of successive reads are independent (no pointer chasing), so they pipeline.
__global__ void __launch_bounds__(256)
pileup_warp_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const char* __restrict__ bases, long long* meta, const int64_t* __restrict__ site_pos,
                   int64_t n_sites, uint16_t* __restrict__ counts) {
  const long long lmax = meta[4];
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (n_sites + 31) / 32;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < n_tiles; tile += warps_total) {
    const int64_t s = tile * 32 + lane;
    const bool live = s < n_sites;
    const int64_t pos = site_pos[live ? s : n_sites - 1];
    const int64_t pmin = __shfl_sync(0xffffffffu, pos, 0);
    const int64_t pmax = __shfl_sync(0xffffffffu, pos, 31);
    uint32_t out[12];
    if (pmax - pmin <= 2 * lmax + 64) {
      const int my = (int)(pos - pmin);
      bool bad = false;
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int64_t r0 = meta[p], r1 = meta[p + 1];
        int64_t lo = r0, hi = r1;
        while (lo < hi) {  // first read with start > pmin - lmax (warp-uniform search)
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(read_start + mid) <= pmin - lmax) lo = mid + 1; else hi = mid;
        }
        const int64_t first = lo;
        hi = r1;
        while (lo < hi) {  // first read with start > pmax
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(read_start + mid) <= pmax) lo = mid + 1; else hi = mid;
        }
        const int64_t last = lo;
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        for (int64_t c = first; c < last; c += 32) {
          const int64_t r = c + lane;
          const bool v = r < last;
          const int64_t b0 = v ? __ldg(base_off + r) : 0;
          const int len = v ? (int)(__ldg(base_off + r + 1) - b0) : 0;
          const int rel = v ? (int)(__ldg(read_start + r) - pmin) : 0;
          const int m = (int)min((int64_t)32, last - c);
#pragma unroll 4
          for (int j = 0; j < m; j++) {
            const int o = my - __shfl_sync(0xffffffffu, rel, j);         // M/ReadSummary.java:48
            const int l = __shfl_sync(0xffffffffu, len, j);
            const int64_t bb = __shfl_sync(0xffffffffu, b0, j);
            if ((unsigned)o < (unsigned)l) {
              const char b = __ldg(bases + bb + o);                       // :49
              const bool iA = b == 'A', iC = b == 'C', iG = b == 'G', iT = b == 'T';
              c0 += iA; c1 += iC; c2 += iG; c3 += iT;
              bad |= !(iA | iC | iG | iT | (b == '-'));                   // :51-55
            }
          }
        }
        out[p * 4 + 0] = c0; out[p * 4 + 1] = c1; out[p * 4 + 2] = c2; out[p * 4 + 3] = c3;
      }
      if (bad && live) { atomicMax((unsigned long long*)&meta[5], (unsigned long long)(s + 1)); meta[6] = DNV_E_BADBASE; meta[7] = 1; }
    } else {
      // sparse candidates: every lane walks only its own window
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int64_t r0 = meta[p], r1 = meta[p + 1];
        int64_t lo = r0, hi = r1;
        const int64_t key = pos - lmax;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(read_start + mid) <= key) lo = mid + 1; else hi = mid;
        }
        uint32_t cnt[4] = {0, 0, 0, 0};
        for (int64_t r = lo; r < r1; r++) {
          const int64_t st = __ldg(read_start + r);
          if (st > pos) break;
          const int64_t b0 = __ldg(base_off + r);
          const int64_t off = pos - st;
          if (off >= __ldg(base_off + r + 1) - b0) continue;
          const char b = __ldg(bases + b0 + off);
          if (b == '-') continue;
          const int a = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1;
          if (a < 0) {
            if (live) { atomicMax((unsigned long long*)&meta[5], (unsigned long long)(r + 1)); meta[6] = DNV_E_BADBASE; }
            continue;
          }
          cnt[0] += a == 0; cnt[1] += a == 1; cnt[2] += a == 2; cnt[3] += a == 3;
        }
#pragma unroll
        for (int a = 0; a < 4; a++) out[p * 4 + a] = cnt[a];
      }
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < 12; k++)
        if (out[k] > 65535u) { meta[6] = DNV_E_INVALID; atomicMax((unsigned long long*)&meta[5], 1ull); out[k] = 65535u; }
      uint2* dst = reinterpret_cast<uint2*>(counts + s * 12);
#pragma unroll
      for (int q = 0; q < 3; q++)
        dst[q] = make_uint2(out[4 * q] | (out[4 * q + 1] << 16), out[4 * q + 2] | (out[4 * q + 3] << 16));
    }
  }
}

}  // namespace dnv

using namespace dnv;

extern "C" {

int dnv_pileup_device(dnv_ctx* ctx, const int64_t* d_read_start, const int64_t* d_base_off, const char* d_bases,
                      const uint8_t* d_person, int64_t n_reads, const int64_t* d_site_pos, int64_t n_sites,
                      uint16_t* d_counts_out) {
  if (!ctx || n_reads < 0 || n_sites < 0 || (n_sites > 0 && (!d_site_pos || !d_counts_out)) ||
      (n_reads > 0 && (!d_read_start || !d_base_off || !d_bases || !d_person))) {
    set_error("dnv_pileup_device: invalid argument");
    return DNV_E_INVALID;
  }
  if (n_sites == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_scratch(ctx, 64);
  if (rc) return rc;
  long long* meta = (long long*)ctx->d_scratch;
  DNV_CUDA(cudaMemsetAsync(meta, 0, 64, ctx->stream));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  if (n_reads > 0) {
    const int grid = (int)std::min<int64_t>((n_reads + 255) / 256, (int64_t)sms * 8);
    pileup_prep_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_person, n_reads, meta);
  }
  const int grid = (int)std::min<int64_t>((n_sites + 255) / 256, (int64_t)sms * 16);
  static const bool legacy = getenv("DNV_PILEUP_LEGACY") != nullptr;  // per-site gather (first version), kept for A/B runs
  if (legacy)
    pileup_gather_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites,
                                                        d_counts_out);
  else
    pileup_warp_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites,
                                                      d_counts_out);
  DNV_CUDA(cudaGetLastError());
  long long h[8];
  DNV_CUDA(cudaMemcpyAsync(h, meta, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h[5] != 0) {
    const int kind = (int)h[6] ? (int)h[6] : DNV_E_INVALID;
    set_error(kind == DNV_E_BADBASE
                  ? std::string("pileup: base outside {A,C,G,T,-} ") + (h[7] ? "at site " : "in read ") + std::to_string(h[5] - 1) +
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

}  // extern "C"
