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
#include "pileup_stage.cuh"

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

// ---------------------------------------------------------------------------------------------
// Tiled kernel for DENSE candidate lists (consecutive positions: the every-position layout of
// BASELINE config 4 and the dense-candidate layout of config 3).  A CTA of 128 threads owns a
// tile of 512 consecutive positions; every thread owns FOUR neighbouring sites, so one
// unaligned 32-bit fetch from a read serves four pileup columns.  Per person the read
// descriptors of the tile's window are staged once in shared memory (start relative to the
// tile, length, base offset), each thread binary-searches its own sub-window there and walks
// ~depth reads.  The four bases are classified with exact per-byte SWAR compares and counted
// in packed byte accumulators (flushed every 128 reads).  No atomics; counts written once,
// 96 contiguous bytes per thread.  Tiles that are not dense fall back to the warp kernel's path.
constexpr int kPT_Threads = 128, kPT_Sites = 512, kPT_ReadCap = 1024;

struct PileTileShared {
  int2 rl[kPT_ReadCap];            // (start - P0, length usable by the 4-byte fast path | 0, ...) see stage
  int len[kPT_ReadCap];            // true length
  unsigned long long ptr[kPT_ReadCap];  // address of base (start - P0 == 0 column): bases + base_off - rel
  uint2 lut[256];                  // byte -> (one-hot A/C/G/T increment in byte lanes, invalid flag)
  long long win[3][2];
  int dense;
};

// 0x80 in every byte of x that is zero (exact, no cross-byte carries)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

// first index in [lo, hi) of the sorted array whose value is > key; 32-ary search by one warp
// (5 dependent probes for 32M reads instead of 25)
__device__ __forceinline__ int64_t warp_upper_bound(const int64_t* __restrict__ a, int64_t lo, int64_t hi,
                                                    int64_t key, int lane) {
  while (hi - lo > 32) {
    const int64_t step = (hi - lo + 32) / 33;          // 32 pivots split the range into 33 parts
    const int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
    const bool le = idx < hi ? (__ldg(a + idx) <= key) : false;
    const uint32_t m = __ballot_sync(0xffffffffu, le);  // a prefix of lanes is true (sorted input)
    const int c = __popc(m);
    const int64_t new_lo = c ? lo + (int64_t)c * step : lo;
    const int64_t new_hi = c < 32 ? min(hi, lo + (int64_t)(c + 1) * step - 1) : hi;
    lo = new_lo; hi = new_hi;
  }
  const bool le = lo + lane < hi ? (__ldg(a + lo + lane) <= key) : false;
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

template <int P>
__device__ __forceinline__ void pile_person(PileTileShared& sh, int64_t c_first, int64_t c_last,
                                            const int64_t* __restrict__ read_start,
                                            const int64_t* __restrict__ base_off, const char* __restrict__ bases,
                                            long long n_bases, int64_t P0, int my_off, int lmax,
                                            uint32_t (&cnt)[4][2], bool& bad, bool& over) {
  const unsigned long long bases_addr = (unsigned long long)reinterpret_cast<uintptr_t>(bases);
  uint32_t acc[4] = {0, 0, 0, 0};  // per my-site k: byte lanes = A, C, G, T counters
  uint32_t tot[4][4];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int a = 0; a < 4; a++) tot[k][a] = 0;
  uint32_t badw = 0;
  int since_flush = 0;
  auto flush = [&]() {
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
      for (int a = 0; a < 4; a++) tot[k][a] += (acc[k] >> (8 * a)) & 0xFF;
      acc[k] = 0;
    }
    since_flush = 0;
  };
  for (long long c0 = c_first; c0 < c_last; c0 += kPT_ReadCap) {
    const int m = (int)min((long long)kPT_ReadCap, c_last - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += kPT_Threads) {
      const long long b0 = __ldg(base_off + c0 + i);
      const int rel = (int)(__ldg(read_start + c0 + i) - P0);
      const int len = (int)(__ldg(base_off + c0 + i + 1) - b0);
      // the fast path reads two aligned words: keep it 8 bytes away from the end of the buffer
      // ... and off the first word, so the aligned-down address never precedes the buffer
      const int fast_len = (b0 >= 4 && b0 + len + 8 <= n_bases) ? len : 0;
      sh.rl[i] = make_int2(rel, fast_len);
      sh.len[i] = len;
      sh.ptr[i] = bases_addr + (unsigned long long)b0 - (unsigned long long)(long long)rel;
    }
    __syncthreads();
    int lo = 0, hi = m;  // first staged read with rel > my_off - lmax
    const int key = my_off - lmax;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sh.rl[mid].x <= key) lo = mid + 1; else hi = mid;
    }
    // four reads per trip: fetch all four words first (independent loads), then classify
    for (int j = lo; j < m && sh.rl[j].x <= my_off + 3; j += 4) {
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        w[u] = 0;
        const int jj = j + u;
        if (jj < m) {
          const int2 rl = sh.rl[jj];
          const int q = my_off - rl.x;  // offset of my first site inside the read (M/ReadSummary.java:48)
          const unsigned long long addr = sh.ptr[jj] + (unsigned long long)my_off;
          if (q >= 0 && q <= rl.y - 4) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~3ull);
            const uint32_t shft = ((uint32_t)addr & 3u) * 8u;
            w[u] = __funnelshift_r(__ldg(wp), __ldg(wp + 1), shft);
          } else {
            const int len = sh.len[jj];
            if (q > -4 && q < len) {  // read edge: bytes outside [0, len) contribute nothing
              const unsigned char* bp = reinterpret_cast<const unsigned char*>(addr);
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int o = q + k;
                if (o >= 0 && o < len) w[u] |= (uint32_t)__ldg(bp + k) << (8 * k);
              }
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint2 e = sh.lut[(w[u] >> (8 * k)) & 0xFF];  // '-' and "no base" add nothing (:51-53)
          acc[k] += e.x;
          badw |= e.y;                                       // anything else: Allele.valueOf throws (:55)
        }
      }
      since_flush += 4;
      if (since_flush >= 252) flush();
    }
  }
  flush();
  bad |= badw != 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int a = 0; a < 4; a++) over |= tot[k][a] > 65535u;
    cnt[k][0] = (tot[k][0] & 0xFFFF) | (tot[k][1] << 16);
    cnt[k][1] = (tot[k][2] & 0xFFFF) | (tot[k][3] << 16);
  }
}

__global__ void __launch_bounds__(kPT_Threads)
pileup_tile_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const char* __restrict__ bases, long long* meta, const int64_t* __restrict__ site_pos,
                   int64_t n_sites, int64_t n_reads, uint16_t* __restrict__ counts) {
  __shared__ PileTileShared sh;
  const int lmax = (int)meta[4];
  const long long n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (n_sites + kPT_Sites - 1) / kPT_Sites;
  for (int i = threadIdx.x; i < 256; i += kPT_Threads)
    sh.lut[i] = i == 'A' ? make_uint2(1u, 0u) : i == 'C' ? make_uint2(1u << 8, 0u) : i == 'G' ? make_uint2(1u << 16, 0u)
              : i == 'T' ? make_uint2(1u << 24, 0u) : (i == '-' || i == 0) ? make_uint2(0u, 0u) : make_uint2(0u, 1u);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s0 = tile * kPT_Sites;
    const int n_live = (int)min((int64_t)kPT_Sites, n_sites - s0);
    const int64_t P0 = site_pos[s0];
    __syncthreads();
    if (threadIdx.x == 0) sh.dense = (site_pos[s0 + n_live - 1] - P0 == n_live - 1) ? 1 : 0;
    // window of reads per person: start in (P0 - lmax, P0 + n_live - 1]; six searches over four warps
    for (int q = warp; q < 6; q += kPT_Threads / 32) {
      const int p = q >> 1, which = q & 1;
      const int64_t key = which ? P0 + n_live - 1 : P0 - lmax;
      const int64_t r = warp_upper_bound(read_start, meta[p], meta[p + 1], key, lane);
      if (lane == 0) sh.win[p][which] = r;
    }
    __syncthreads();
    // strictly increasing positions with last - first == n - 1  <=>  consecutive
    int ok = 1;
    for (int i = threadIdx.x; i < n_live - 1; i += kPT_Threads) ok &= site_pos[s0 + i + 1] > site_pos[s0 + i];
    if (!ok) sh.dense = 0;  // benign race: only ever cleared
    __syncthreads();
    const bool dense = sh.dense != 0;
    const int my_off = threadIdx.x * 4;  // offset of my first site from P0
    uint32_t cnt[3][4][2];               // [person][my site][packed u16 pairs]
    bool bad = false, over = false;
    if (dense) {
      pile_person<0>(sh, sh.win[0][0], sh.win[0][1], read_start, base_off, bases, n_bases, P0, my_off, lmax, cnt[0], bad, over);
      pile_person<1>(sh, sh.win[1][0], sh.win[1][1], read_start, base_off, bases, n_bases, P0, my_off, lmax, cnt[1], bad, over);
      pile_person<2>(sh, sh.win[2][0], sh.win[2][1], read_start, base_off, bases, n_bases, P0, my_off, lmax, cnt[2], bad, over);
    } else {
      // sparse tile: every thread walks its own windows, one site at a time
#pragma unroll
      for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int p = 0; p < 3; p++) {
          uint32_t c4[4] = {0, 0, 0, 0};
          const int64_t s = s0 + my_off + k;
          if (s < n_sites) {
            const int64_t pos = site_pos[s];
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
          }
#pragma unroll
          for (int a = 0; a < 4; a++) over |= c4[a] > 65535u;
          cnt[p][k][0] = (c4[0] & 0xFFFF) | (c4[1] << 16);
          cnt[p][k][1] = (c4[2] & 0xFFFF) | (c4[3] << 16);
        }
      }
    }
    if (bad && my_off < n_live) {
      atomicMax((unsigned long long*)&meta[5], (unsigned long long)(s0 + my_off + 1));
      meta[6] = DNV_E_BADBASE; meta[7] = 1;
    } else if (over && my_off < n_live) {  // more than 65535 reads of one allele at one site
      atomicMax((unsigned long long*)&meta[5], 1ull);
      meta[6] = DNV_E_INVALID;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int64_t s = s0 + my_off + k;
      if (s < n_sites) {
        uint2* dst = reinterpret_cast<uint2*>(counts + s * 12);
#pragma unroll
        for (int p = 0; p < 3; p++) dst[p] = make_uint2(cnt[p][k][0], cnt[p][k][1]);
      }
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
  static const int variant = getenv("DNV_PILEUP_VARIANT") ? atoi(getenv("DNV_PILEUP_VARIANT")) : 3;  // A/B knob
  if (variant == 0)  // v0: one thread per site, per-lane windows
    pileup_gather_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites,
                                                        d_counts_out);
  else if (variant == 1)  // v1: warp-cooperative window walk
    pileup_warp_kernel<<<grid, 256, 0, ctx->stream>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites,
                                                      d_counts_out);
  else if (variant == 3 || variant == 4) {  // v3 / v4: bases staged in shared memory, 8 / 16 sites per thread
    cudaError_t e = variant == 3 ? launch_pileup_stage<8>(ctx->stream, sms, d_read_start, d_base_off, d_bases, meta, d_site_pos,
                                                          n_sites, n_reads, d_counts_out)
                                 : launch_pileup_stage<16>(ctx->stream, sms, d_read_start, d_base_off, d_bases, meta, d_site_pos,
                                                           n_sites, n_reads, d_counts_out);
    DNV_CUDA(e);
  } else {  // v2: 512-site tiles, 4 sites per thread, descriptors staged in shared memory
    const int64_t n_tiles = (n_sites + kPT_Sites - 1) / kPT_Sites;
    const int tgrid = (int)std::min<int64_t>(n_tiles, (int64_t)sms * 16);
    pileup_tile_kernel<<<tgrid, kPT_Threads, 0, ctx->stream>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos,
                                                              n_sites, n_reads, d_counts_out);
  }
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
