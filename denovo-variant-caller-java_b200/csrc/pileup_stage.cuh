// pileup_stage.cuh — pileup kernel for DENSE candidate lists (consecutive positions: the
// every-position layout of BASELINE config 4 and the dense-candidate layout of config 3),
// bases staged through shared memory.
//
// Reads are sorted by (person, start) and their alignedSequence strings are concatenated in
// that order, so the reads that can reach a tile of consecutive positions are ONE contiguous
// byte range of `bases` per person.  A CTA of 128 threads owns a tile of 128*S consecutive
// positions and, per person and per chunk of reads:
//   1. copies that byte range into shared memory with 16-byte loads (several CTAs per SM cover each other's copies) (every base is read from
//      HBM once per tile it touches: (tile + read length) / tile ~ 1.1x the algorithmic bytes),
//      and the read descriptors next to it (start relative to the tile, offset in the stage, length);
//   2. every thread owns S neighbouring sites: it binary-searches its sub-window of reads in
//      shared memory and, per read, takes S+4 staged bytes as aligned words + funnel shifts —
//      S pileup columns per visit, no global loads in the loop;
//   3. each base goes through a 256-entry LUT (one LDS) whose entry is the increment of a
//      packed accumulator: 6-bit lanes A, C, G, T and "not a base" (anything outside
//      {A,C,G,T,-}: the reference's IllegalArgumentException, M/ReadSummary.java:55), flushed
//      into 16-bit totals every 63 reads: 4 instructions per base.
// No atomics, deterministic, counts written once.  Tiles that are not consecutive positions, and
// reads longer than the stage, take the per-site path of pileup.cu.
//
// Replaces `new ReadSummary(reads, position)` x3 per site (M/ReadSummary.java:45-59).
#pragma once
#include "dnv_internal.cuh"

namespace dnv {

constexpr int kPS_Threads = 128;
constexpr int kPS_ReadCap = 384;            // read descriptors per chunk
constexpr int kPS_StageBytes = 20 * 1024;   // staged bases per chunk (small: ~7 CTAs per SM cover each other's staging)
constexpr int kPS_Pad = 32;                 // readable slack on both sides of the staged bytes

// error word of the pileup kernels: (index + 1) << 4 | at_site << 3 | kind — one atomicMax carries the kind
// together with the index it belongs to
constexpr unsigned long long kPileErrInvalid = 1, kPileErrUnsorted = 2, kPileErrBadBase = 3;
__host__ __device__ __forceinline__ unsigned long long pileup_err(long long index, bool at_site, unsigned long long kind) {
  return ((unsigned long long)(index + 1) << 4) | (at_site ? 8ull : 0ull) | kind;
}

template <int S>
struct PileStageShared {
  alignas(16) unsigned char stage[kPS_Pad + kPS_StageBytes + kPS_Pad];
  uint2 tot[S][kPS_Threads];  // 16-bit totals (A | C << 16, G | T << 16) of thread t's site k: [k][t], conflict free
  int rel[kPS_ReadCap];    // read start - P0
  int soff[kPS_ReadCap];   // offset of the read's first base in stage[]
  int len[kPS_ReadCap];
  uint32_t lut[256];
  long long win[3][2];
  int dense;
};

// first index in [lo, hi) of the sorted array whose value is > key; 32-ary search by one warp
__device__ __forceinline__ int64_t ps_warp_upper_bound(const int64_t* __restrict__ a, int64_t lo, int64_t hi,
                                                       int64_t key, int lane) {
  while (hi - lo > 32) {
    const int64_t step = (hi - lo + 32) / 33;
    const int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
    const bool le = idx < hi ? (__ldg(a + idx) <= key) : false;
    const uint32_t m = __ballot_sync(0xffffffffu, le);
    const int c = __popc(m);
    const int64_t new_lo = c ? lo + (int64_t)c * step : lo;
    const int64_t new_hi = c < 32 ? min(hi, lo + (int64_t)(c + 1) * step - 1) : hi;
    lo = new_lo; hi = new_hi;
  }
  const bool le = lo + lane < hi ? (__ldg(a + lo + lane) <= key) : false;
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

// the six window bounds of a tile at once: out[2p] = first read of person p with start > key_lo, out[2p+1] = first
// with start > key_hi (person p's reads are [meta[p], meta[p+1])); 32-ary, the loads of all six searches of a round
// in flight together (4 round trips for 600 k reads per person instead of 24)
__device__ __forceinline__ void ps_warp_upper_bound6(const int64_t* __restrict__ a, const long long* meta, int64_t key_lo,
                                                     int64_t key_hi, int lane, int64_t (&out)[6]) {
  constexpr long long kInf = 0x7fffffffffffffffll;
  int64_t lo[6], hi[6];
#pragma unroll
  for (int q = 0; q < 6; q++) { lo[q] = meta[q >> 1]; hi[q] = meta[(q >> 1) + 1]; }
  for (;;) {
    bool big = false;
#pragma unroll
    for (int q = 0; q < 6; q++) big |= hi[q] - lo[q] > 32;
    if (!big) break;
    int64_t step[6];
    long long v[6];
#pragma unroll
    for (int q = 0; q < 6; q++) {
      step[q] = (hi[q] - lo[q] + 32) / 33;
      const int64_t idx = lo[q] + (int64_t)(lane + 1) * step[q] - 1;
      v[q] = (hi[q] - lo[q] > 32 && idx < hi[q]) ? __ldg(a + idx) : kInf;
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
      if (hi[q] - lo[q] > 32) {  // warp uniform
        const int c = __popc(__ballot_sync(0xffffffffu, v[q] <= ((q & 1) ? key_hi : key_lo)));
        const int64_t new_lo = c ? lo[q] + (int64_t)c * step[q] : lo[q];
        const int64_t new_hi = c < 32 ? min(hi[q], lo[q] + (int64_t)(c + 1) * step[q] - 1) : hi[q];
        lo[q] = new_lo; hi[q] = new_hi;
      }
    }
  }
  long long v[6];
#pragma unroll
  for (int q = 0; q < 6; q++) v[q] = lo[q] + lane < hi[q] ? __ldg(a + lo[q] + lane) : kInf;
#pragma unroll
  for (int q = 0; q < 6; q++) out[q] = lo[q] + __popc(__ballot_sync(0xffffffffu, v[q] <= ((q & 1) ? key_hi : key_lo)));
}

// one person of one dense tile: out01[k] = A | C << 16, out23[k] = G | T << 16 for my S sites
template <int S>
__device__ __forceinline__ void ps_person(PileStageShared<S>& sh, int64_t c_first, int64_t c_last,
                                          const int64_t* __restrict__ read_start,
                                          const int64_t* __restrict__ base_off, const char* __restrict__ bases,
                                          long long n_bases, int64_t P0, int my_off, int n_live, int lmax,
                                          bool& bad, bool& over) {
  constexpr int NW = S / 4;
  uint32_t acc[S];
#pragma unroll
  for (int k = 0; k < S; k++) { acc[k] = 0; sh.tot[k][threadIdx.x] = make_uint2(0u, 0u); }
  uint32_t badacc = 0;
  int since_flush = 0, visits = 0;
  auto flush = [&]() {  // 6-bit lanes -> this thread's 16-bit totals in shared memory (once per 63 visits)
#pragma unroll
    for (int k = 0; k < S; k++) {
      const uint32_t a = acc[k];
      uint2 t = sh.tot[k][threadIdx.x];
      t.x += (a & 0x3Fu) | ((a << 10) & 0x3F0000u);
      t.y += ((a >> 12) & 0x3Fu) | ((a >> 2) & 0x3F0000u);
      sh.tot[k][threadIdx.x] = t;
      badacc |= a >> 24;
      acc[k] = 0;
    }
    since_flush = 0;
  };
  const uintptr_t bases_abs = reinterpret_cast<uintptr_t>(bases);
  const int chunk_reads = min(kPS_ReadCap, (kPS_StageBytes - 16) / max(lmax, 1));  // >= 1: the caller checked lmax
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(sh.stage + kPS_Pad);
  const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(sh.lut);
  for (int64_t c0 = c_first; c0 < c_last; c0 += chunk_reads) {
    const int m = (int)min((int64_t)chunk_reads, c_last - c0);
    __syncthreads();  // the previous chunk (or person) is no longer being read
    const long long g0 = __ldg(base_off + c0), g1 = __ldg(base_off + c0 + m);
    const long long a0 = g0 - (long long)((bases_abs + (uintptr_t)g0) & 15);  // stage[i] <-> bases[a0 + i], 16-byte aligned
    for (int i = threadIdx.x; i < m; i += kPS_Threads) {
      const long long b0 = __ldg(base_off + c0 + i);
      sh.rel[i] = (int)(__ldg(read_start + c0 + i) - P0);
      sh.len[i] = (int)(__ldg(base_off + c0 + i + 1) - b0);
      sh.soff[i] = (int)(b0 - a0);
    }
    const int n16 = (int)((g1 - a0 + 15) >> 4);
    for (int v = threadIdx.x; v < n16; v += kPS_Threads) {
      const long long gi = a0 + 16ll * v;
      uint4 x;
      if (gi >= 0 && gi + 16 <= n_bases) {
        x = __ldg(reinterpret_cast<const uint4*>(bases + gi));
      } else {  // first / last 16 bytes of the buffer: byte loads inside [0, n_bases)
        uint32_t w[4] = {0x2D2D2D2Du, 0x2D2D2D2Du, 0x2D2D2D2Du, 0x2D2D2D2Du};
        for (int b = 0; b < 16; b++) {
          const long long gb = gi + b;
          if (gb >= 0 && gb < n_bases)
            w[b >> 2] = (w[b >> 2] & ~(0xFFu << (8 * (b & 3)))) | ((uint32_t)(unsigned char)__ldg(bases + gb) << (8 * (b & 3)));
        }
        x = make_uint4(w[0], w[1], w[2], w[3]);
      }
      *reinterpret_cast<uint4*>(sh.stage + kPS_Pad + 16 * v) = x;
    }
    __syncthreads();
    int lo = 0, hi = m;  // first staged read with rel > my_off - lmax (earlier ones end before my first site)
    const int key = my_off - lmax;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sh.rel[mid] <= key) lo = mid + 1; else hi = mid;
    }
    for (int j = lo; j < m; j++) {
      const int rel = sh.rel[j];
      if (rel > my_off + S - 1) break;
      const int q = my_off - rel;  // offset of my first site inside the read (M/ReadSummary.java:48)
      const int len = sh.len[j];
      if (q >= len) continue;      // read ends before my first site
      const int addr = sh.soff[j] + q;  // >= -(S-1): inside the front pad
      const uint32_t wa = stage_s + (uint32_t)(addr & ~3);
      const uint32_t shft = ((uint32_t)addr & 3u) * 8u;
      uint32_t x[NW + 1], w[NW];
#pragma unroll
      for (int i = 0; i <= NW; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"(wa + 4 * i));
#pragma unroll
      for (int i = 0; i < NW; i++) w[i] = __funnelshift_r(x[i], x[i + 1], shft);
      // bytes of my window outside [0, len) of this read are not its bases, and window positions at or past
      // n_live (last, partial tile) are not requested sites — the reference only looks at the base AT a
      // candidate position (M/ReadSummary.java:48-55), so an 'N' next to the last site must not fail the
      // batch: turn both kinds into '-' (branch free).  t_lo in [0, S-1], t_hi <= S (<= 0: nothing of mine)
      const int t_lo = max(0, -q), t_hi = min(min(S, len - q), n_live - my_off);
#pragma unroll
      for (int i = 0; i < NW; i += 2) {  // 8 window bytes at a time as one 64-bit mask
        const int a = min(max(t_lo - 4 * i, 0), 8), b = min(max(t_hi - 4 * i, 0), 8);
        const unsigned long long from = a >= 8 ? 0ull : ~0ull << (8 * a);
        const unsigned long long to = b >= 8 ? ~0ull : ~(~0ull << (8 * b));
        const unsigned long long keep = from & to;
        const uint32_t k0 = (uint32_t)keep, k1 = (uint32_t)(keep >> 32);
        w[i] = (w[i] & k0) | (0x2D2D2D2Du & ~k0);
        w[i + 1] = (w[i + 1] & k1) | (0x2D2D2D2Du & ~k1);
      }
#pragma unroll
      for (int i = 0; i < NW; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t byte = __byte_perm(w[i], 0, 0x4440 + k);  // PRMT: byte k, zero extended
          uint32_t e;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(lut_s + byte * 4u));  // '-' adds nothing (:51-53)
          acc[4 * i + k] += e;
        }
      }
      visits++;
      if (++since_flush >= 63) flush();
    }
  }
  flush();
  bad |= badacc != 0;
  over |= visits > 65535;  // a 16-bit total could have wrapped: fail loudly (depth above 65535)
}

template <int S>
__global__ void __launch_bounds__(kPS_Threads)
pileup_stage_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                    const char* __restrict__ bases, long long* meta, const int64_t* __restrict__ site_pos,
                    int64_t n_sites, int64_t n_reads, uint16_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  PileStageShared<S>& sh = *reinterpret_cast<PileStageShared<S>*>(ps_smem);
  constexpr int kSites = kPS_Threads * S;
  const int lmax = (int)meta[4];
  const long long n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (n_sites + kSites - 1) / kSites;
  // LUT: increment of the packed accumulator (6-bit lanes A, C, G, T; bit 24 = not a base)
  for (int i = threadIdx.x; i < 256; i += kPS_Threads)
    sh.lut[i] = i == 'A' ? 1u : i == 'C' ? 1u << 6 : i == 'G' ? 1u << 12 : i == 'T' ? 1u << 18
              : i == '-' ? 0u : 1u << 24;
  const bool stageable = lmax > 0 && lmax <= kPS_StageBytes - 16;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s0 = tile * kSites;
    const int n_live = (int)min((int64_t)kSites, n_sites - s0);
    const int64_t P0 = site_pos[s0];
    __syncthreads();
    if (threadIdx.x == 0) sh.dense = (stageable && site_pos[s0 + n_live - 1] - P0 == n_live - 1) ? 1 : 0;
    // window of reads per person: start in (P0 - lmax, P0 + n_live - 1]; six searches over four warps
    for (int q = warp; q < 6; q += kPS_Threads / 32) {
      const int p = q >> 1, which = q & 1;
      const int64_t key = which ? P0 + n_live - 1 : P0 - lmax;
      const int64_t r = ps_warp_upper_bound(read_start, meta[p], meta[p + 1], key, lane);
      if (lane == 0) sh.win[p][which] = r;
    }
    __syncthreads();
    // strictly increasing positions with last - first == n - 1  <=>  consecutive
    int ok = 1;
    for (int i = threadIdx.x; i < n_live - 1; i += kPS_Threads) ok &= site_pos[s0 + i + 1] > site_pos[s0 + i];
    if (!ok) sh.dense = 0;  // benign race: only ever cleared
    __syncthreads();
    const bool dense = sh.dense != 0;
    const int my_off = threadIdx.x * S;  // offset of my first site from P0
    bool bad = false, over = false;
    if (dense) {
#pragma unroll 1
      for (int p = 0; p < 3; p++) {
        ps_person<S>(sh, sh.win[p][0], sh.win[p][1], read_start, base_off, bases, n_bases, P0, my_off, n_live, lmax, bad, over);
#pragma unroll
        for (int k = 0; k < S; k++) {
          const int64_t s = s0 + my_off + k;
          if (s < n_sites) *reinterpret_cast<uint2*>(counts + s * 12 + p * 4) = sh.tot[k][threadIdx.x];
        }
      }
    } else {
      // sparse tile: every thread walks its own windows, one site at a time
      for (int k = 0; k < S; k++) {
        const int64_t s = s0 + my_off + k;
        if (s >= n_sites) break;
        const int64_t pos = site_pos[s];
#pragma unroll
        for (int p = 0; p < 3; p++) {
          uint32_t c4[4] = {0, 0, 0, 0};
          const int64_t r1 = meta[p + 1];
          int64_t lo = meta[p], hi = r1;
          const int64_t key = pos - lmax;
          while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(read_start + mid) <= key) lo = mid + 1; else hi = mid;
          }
          for (int64_t r = lo; r < r1; r++) {
            const int64_t st = __ldg(read_start + r);
            if (st > pos) break;
            const int64_t b0 = __ldg(base_off + r);
            const int64_t off = pos - st;
            if (off >= __ldg(base_off + r + 1) - b0) continue;
            const char b = __ldg(bases + b0 + off);
            if (b == '-') continue;
            const bool iA = b == 'A', iC = b == 'C', iG = b == 'G', iT = b == 'T';
            c4[0] += iA; c4[1] += iC; c4[2] += iG; c4[3] += iT;
            bad |= !(iA | iC | iG | iT);
          }
#pragma unroll
          for (int a = 0; a < 4; a++) over |= c4[a] > 65535u;
          *reinterpret_cast<uint2*>(counts + s * 12 + p * 4) =
              make_uint2((c4[0] & 0xFFFF) | (c4[1] << 16), (c4[2] & 0xFFFF) | (c4[3] << 16));
        }
      }
    }
    if (bad && my_off < n_live) {
      atomicMax((unsigned long long*)&meta[5], pileup_err(s0 + my_off, true, kPileErrBadBase));
    } else if (over && my_off < n_live) {  // more than 65535 reads over one site
      atomicMax((unsigned long long*)&meta[5], pileup_err(0, false, kPileErrInvalid));
    }
  }
}

template <int S>
static cudaError_t launch_pileup_stage(cudaStream_t st, int sms, const int64_t* d_read_start, const int64_t* d_base_off,
                                       const char* d_bases, long long* meta, const int64_t* d_site_pos, int64_t n_sites,
                                       int64_t n_reads, uint16_t* d_counts) {
  auto kern = pileup_stage_kernel<S>;
  const size_t smem = sizeof(PileStageShared<S>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPS_Threads, smem);
  if (e != cudaSuccess) return e;
  const int64_t n_tiles = (n_sites + kPS_Threads * S - 1) / (kPS_Threads * S);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)sms * std::max(per_sm, 1)));
  kern<<<grid, kPS_Threads, smem, st>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites, n_reads, d_counts);
  return cudaGetLastError();
}

}  // namespace dnv
