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
// inter-thread contention, deterministic.  The kernel itself lives in pileup_stage.cuh; this
// file holds the validation / metadata pass over the read descriptors and the C entry points.
#include <stdlib.h>
#include <algorithm>
#include "dnv_internal.cuh"
#include "pileup_stage.cuh"

namespace dnv {

// meta[0..3] person offsets, meta[4] max read length, meta[5] error word (pileup_err: index and kind in ONE atomicMax)
__global__ void __launch_bounds__(256)
pileup_prep_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const uint8_t* __restrict__ person, int64_t n_reads, long long* __restrict__ meta) {
  long long lmax = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
    const long long len = base_off[r + 1] - base_off[r];
    lmax = len > lmax ? len : lmax;
    const int p = person[r];
    if (p > 2 || len < 0) { atomicMax((unsigned long long*)&meta[5], pileup_err(r, false, kPileErrInvalid)); continue; }
    if (r == 0) {
      for (int q = 0; q <= p; q++) meta[q] = 0;
    } else {
      const int pp = person[r - 1];
      if (pp > p || (pp == p && read_start[r - 1] > read_start[r])) {
        atomicMax((unsigned long long*)&meta[5], pileup_err(r, false, kPileErrUnsorted));
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
  if ((reinterpret_cast<uintptr_t>(d_counts_out) & 7) || (reinterpret_cast<uintptr_t>(d_read_start) & 7) ||
      (reinterpret_cast<uintptr_t>(d_base_off) & 7) || (reinterpret_cast<uintptr_t>(d_site_pos) & 7)) {
    set_error("dnv_pileup_device: counts, read_start, base_off and site_pos must be 8-byte aligned");
    return DNV_E_INVALID;
  }
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
  // dense tiles: bases staged in shared memory, eight sites per thread (pileup_stage.cuh); sparse tiles:
  // one site at a time inside the same kernel
  DNV_CUDA(launch_pileup_stage<8>(ctx->stream, sms, d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites, n_reads,
                                  d_counts_out));
  DNV_CUDA(cudaGetLastError());
  long long h[8];
  DNV_CUDA(cudaMemcpyAsync(h, meta, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DNV_CUDA(cudaStreamSynchronize(ctx->stream));
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

}  // extern "C"
