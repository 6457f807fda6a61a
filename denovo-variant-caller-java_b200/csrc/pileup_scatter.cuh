// pileup_scatter.cuh — pileup kernel (a), second formulation, and its fusion with the integer screen.
//
// Batched `new ReadSummary(reads, position)` x3 (M/ReadSummary.java:45-59) for DENSE candidate lists
// (consecutive positions: the every-position layout of BASELINE config 4 and the dense-candidate layout of
// config 3).  pileup_stage.cuh gathers: a thread owns eight positions and visits every read that reaches them
// (~84 instructions per visit of 8 bases, most of them finding and aligning the read).  Here the work is
// READ-MAJOR: a CTA owns a tile of 1024 consecutive positions and, per person, walks the reads that reach the
// tile in rounds of at most 255 reads:
//   1. the round's bases — ONE contiguous byte range, the reads being sorted by (person, start) and their
//      strings concatenated in that order — are staged in shared memory with 16-byte loads, the read
//      descriptors next to them (as in pileup_stage.cuh);
//   2. every thread takes one read and streams its bases as aligned words + a funnel shift; each base goes
//      through a 256-entry table (one 8-byte shared-memory load: increment of a packed counter with 8-bit lanes
//      A, C, G, T — zero for '-', M/ReadSummary.java:51-53 — and a "not a base" flag, the reference's
//      IllegalArgumentException, :55) and ONE shared-memory atomic add (SASS ATOMS.ADD, fire and forget) on the
//      counter word of its position.  The north star's "warp-aggregated atomics" live here, in shared memory:
//      4.5 instructions per base, no search, no per-visit masks;
//   3. 255 reads cannot overflow an 8-bit lane, so after the round the counters are folded into 16-bit totals
//      (the output layout: A | C << 16, G | T << 16 per person), also in shared memory.
// The totals of all three persons of the tile stay in shared memory: the plain kernel writes them out (24 bytes
// per site, coalesced); the FUSED kernel screens every site straight from shared memory
// with the exact integer screen of trio_screen.cuh and writes only what the ~1 % undecided sites need — their
// counts and their ascending site list, one sorted segment per CTA — for the list / finish kernels of
// trio_list.cuh: reads -> calls without the 24 B/site round trip through HBM.
// Tiles that are not consecutive positions take the per-site path (one site per thread and person).
#pragma once
#include "dnv_internal.cuh"
#include "pileup_stage.cuh"
#include "trio_screen.cuh"

namespace dnv {

constexpr int kSC_Threads = 256;
constexpr int kSC_Tile = 1024;             // positions per tile (4 per thread)
constexpr int kSC_ReadCap = 255;           // reads per round: an 8-bit counter lane cannot overflow within a round
constexpr int kSC_StageBytes = 28 * 1024;  // staged bases per round
constexpr int kSC_Pad = 16;                // readable slack on both sides of the staged bytes

struct ScatterShared {
  alignas(16) unsigned char stage[kSC_Pad + kSC_StageBytes + kSC_Pad];
  uint2 tot[3][kSC_Tile];           // 16-bit totals per person: A | C << 16, G | T << 16 (the output layout)
  uint32_t cnt[kSC_Tile + 8];       // the round's counters: 8-bit lanes A, C, G, T (+8: masked tail bytes add 0 there)
  int rel[kSC_ReadCap + 1];         // read start - P0
  int soff[kSC_ReadCap + 1];        // offset of the read's first base in stage[]
  int len[kSC_ReadCap + 1];
  unsigned long long lut[256];      // [31:0] increment of the packed counter, [32] not a base
  long long win[3][2];
  int dense, bad, over;
  uint32_t warp_cnt[kSC_Threads / 32];
  uint32_t written;                 // fused kernel: undecided sites this CTA has emitted so far
};

__device__ __forceinline__ void sc_init_lut(ScatterShared& sh) {
  for (int i = threadIdx.x; i < 256; i += kSC_Threads)
    sh.lut[i] = i == 'A' ? 1ull : i == 'C' ? 1ull << 8 : i == 'G' ? 1ull << 16 : i == 'T' ? 1ull << 24
              : i == '-' ? 0ull : 1ull << 32;
}

// windows of reads per person that can reach the tile, and whether the tile is consecutive positions
__device__ __forceinline__ void sc_tile_setup(ScatterShared& sh, const int64_t* __restrict__ read_start, const long long* meta,
                                              const int64_t* __restrict__ site_pos, int64_t s0, int n_live, int lmax, bool stageable) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t P0 = site_pos[s0];
  __syncthreads();
  if (threadIdx.x == 0) {
    sh.dense = (stageable && site_pos[s0 + n_live - 1] - P0 == n_live - 1) ? 1 : 0;
    sh.bad = 0; sh.over = 0;
  }
  if (warp < 6) {  // six searches, one warp each: start in (P0 - lmax, P0 + n_live - 1]
    const int p = warp >> 1, which = warp & 1;
    const int64_t key = which ? P0 + n_live - 1 : P0 - lmax;
    const int64_t r = ps_warp_upper_bound(read_start, meta[p], meta[p + 1], key, lane);
    if (lane == 0) sh.win[p][which] = r;
  }
  __syncthreads();
  int ok = 1;  // strictly increasing positions with last - first == n - 1  <=>  consecutive
  for (int i = threadIdx.x; i < n_live - 1; i += kSC_Threads) ok &= site_pos[s0 + i + 1] > site_pos[s0 + i];
  if (!ok) sh.dense = 0;  // benign race: only ever cleared
  for (int i = threadIdx.x; i < kSC_Tile; i += kSC_Threads) {
    sh.tot[0][i] = make_uint2(0, 0); sh.tot[1][i] = make_uint2(0, 0); sh.tot[2][i] = make_uint2(0, 0);
  }
  __syncthreads();
}

// dense tile, one person: rounds of reads [c_first, c_last) -> sh.tot[p]
__device__ __forceinline__ void sc_person(ScatterShared& sh, int p, int64_t c_first, int64_t c_last,
                                          const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                                          const char* __restrict__ bases, long long n_bases, int64_t P0, int n_live, int lmax) {
  const uintptr_t bases_abs = reinterpret_cast<uintptr_t>(bases);
  const int round_reads = min(kSC_ReadCap, (kSC_StageBytes - 16) / max(lmax, 1));  // >= 1: the caller checked lmax
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(sh.stage + kSC_Pad);
  const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(sh.lut);
  const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(sh.cnt);
  const bool check_over = c_last - c_first > 60000;  // only then can a 16-bit total wrap
  uint32_t bad = 0, over = 0;
  for (int64_t c0 = c_first; c0 < c_last; c0 += round_reads) {
    const int m = (int)min((int64_t)round_reads, c_last - c0);
    __syncthreads();  // the previous round (or person) has been folded
    for (int i = threadIdx.x; i < kSC_Tile + 8; i += kSC_Threads) sh.cnt[i] = 0;
    const long long g0 = __ldg(base_off + c0), g1 = __ldg(base_off + c0 + m);
    const long long a0 = g0 - (long long)((bases_abs + (uintptr_t)g0) & 15);  // stage[i] <-> bases[a0 + i], 16-byte aligned
    for (int i = threadIdx.x; i < m; i += kSC_Threads) {
      const long long b0 = __ldg(base_off + c0 + i);
      sh.rel[i] = (int)(__ldg(read_start + c0 + i) - P0);
      sh.len[i] = (int)(__ldg(base_off + c0 + i + 1) - b0);
      sh.soff[i] = (int)(b0 - a0);
    }
    const int n16 = (int)((g1 - a0 + 15) >> 4);
    for (int v = threadIdx.x; v < n16; v += kSC_Threads) {
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
      *reinterpret_cast<uint4*>(sh.stage + kSC_Pad + 16 * v) = x;
    }
    __syncthreads();
    // Work items are QUARTERS of reads (4 m items over 256 threads: the round is balanced whatever m is): bases
    // [t_lo, t_hi) of a read land on positions [rel + t_lo, rel + t_hi) of the tile, item q takes the q-th quarter
    for (int item = threadIdx.x; item < 4 * m; item += kSC_Threads) {
      const int r = item >> 2, q = item & 3;
      const int rel = sh.rel[r], len = sh.len[r];
      const int t_lo = max(0, -rel), t_hi = min(len, n_live - rel);
      const int span = t_hi - t_lo;
      if (span <= 0) continue;
      const int b_lo = t_lo + ((span * q) >> 2), b_hi = t_lo + ((span * (q + 1)) >> 2);
      const int nb = b_hi - b_lo;
      if (nb <= 0) continue;
      const int addr = sh.soff[r] + b_lo;
      uint32_t wa = stage_s + (uint32_t)(addr & ~3);
      const uint32_t shft = ((uint32_t)addr & 3u) * 8u;
      uint32_t ca = cnt_s + (uint32_t)(rel + b_lo) * 4u;
      uint32_t x0;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x0) : "r"(wa));
      for (int left = nb; left > 0; left -= 4) {
        uint32_t x1;
        wa += 4;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x1) : "r"(wa));
        uint32_t w = __funnelshift_r(x0, x1, shft);
        x0 = x1;
        if (left < 4) {  // bytes past the item's last base are not its bases: '-' adds nothing
          const uint32_t keep = (1u << (8 * left)) - 1u;
          w = (w & keep) | (0x2D2D2D2Du & ~keep);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t byte = __byte_perm(w, 0, 0x4440 + k);  // PRMT: byte k, zero extended
          uint32_t inc, flag;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(inc), "=r"(flag) : "r"(lut_s + byte * 8u));
          bad |= flag;
          asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ca + 4u * k), "r"(inc) : "memory");
        }
        ca += 16;
      }
    }
    __syncthreads();
    // fold the round's 8-bit lanes into the 16-bit totals
#pragma unroll
    for (int k = 0; k < kSC_Tile / kSC_Threads; k++) {
      const int pos = threadIdx.x + k * kSC_Threads;
      const uint32_t c = sh.cnt[pos];
      if (c) {
        uint2 t = sh.tot[p][pos];
        const uint32_t ax = (c & 0xFFu) | ((c << 8) & 0xFF0000u), ay = ((c >> 16) & 0xFFu) | ((c >> 8) & 0xFF0000u);
        if (check_over)
          over |= ((t.x & 0xFFFFu) + (ax & 0xFFFFu)) >> 16 | ((t.x >> 16) + (ax >> 16)) >> 16 |
                  ((t.y & 0xFFFFu) + (ay & 0xFFFFu)) >> 16 | ((t.y >> 16) + (ay >> 16)) >> 16;
        t.x += ax; t.y += ay;
        sh.tot[p][pos] = t;
      }
    }
  }
  if (bad) sh.bad = 1;    // benign races: only ever set
  if (over) sh.over = 1;
}

// tile that is not consecutive positions: one site at a time per thread and person (as pileup_stage.cuh)
__device__ __forceinline__ void sc_sparse_tile(ScatterShared& sh, const int64_t* __restrict__ read_start,
                                               const int64_t* __restrict__ base_off, const char* __restrict__ bases,
                                               const long long* meta, const int64_t* __restrict__ site_pos, int64_t s0,
                                               int n_live, int lmax, long long* meta_err) {
  for (int k = 0; k < kSC_Tile / kSC_Threads; k++) {
    const int i = threadIdx.x + k * kSC_Threads;
    if (i >= n_live) break;
    const int64_t pos = site_pos[s0 + i];
    bool bad = false, over = false;
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
      sh.tot[p][i] = make_uint2((c4[0] & 0xFFFF) | (c4[1] << 16), (c4[2] & 0xFFFF) | (c4[3] << 16));
    }
    if (bad) atomicMax((unsigned long long*)&meta_err[5], pileup_err(s0 + i, true, kPileErrBadBase));
    else if (over) atomicMax((unsigned long long*)&meta_err[5], pileup_err(0, false, kPileErrInvalid));
  }
}

// the pileup of one tile -> sh.tot (both kernels)
__device__ __forceinline__ void sc_pile_tile(ScatterShared& sh, const int64_t* __restrict__ read_start,
                                             const int64_t* __restrict__ base_off, const char* __restrict__ bases, long long* meta,
                                             const int64_t* __restrict__ site_pos, int64_t s0, int n_live, int lmax, long long n_bases,
                                             bool stageable) {
  sc_tile_setup(sh, read_start, meta, site_pos, s0, n_live, lmax, stageable);
  if (sh.dense) {
    const int64_t P0 = site_pos[s0];
#pragma unroll 1
    for (int p = 0; p < 3; p++) sc_person(sh, p, sh.win[p][0], sh.win[p][1], read_start, base_off, bases, n_bases, P0, n_live, lmax);
    __syncthreads();
    if (threadIdx.x == 0) {
      // the staged path sees a bad base anywhere in the tile's positions, all of which are requested sites
      if (sh.bad) atomicMax((unsigned long long*)&meta[5], pileup_err(s0, true, kPileErrBadBase));
      else if (sh.over) atomicMax((unsigned long long*)&meta[5], pileup_err(0, false, kPileErrInvalid));
    }
  } else {
    sc_sparse_tile(sh, read_start, base_off, bases, meta, site_pos, s0, n_live, lmax, meta);
    __syncthreads();
  }
}

// ---- plain kernel: reads -> counts ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSC_Threads, 3)
pileup_scatter_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                      const char* __restrict__ bases, long long* meta, const int64_t* __restrict__ site_pos,
                      int64_t n_sites, int64_t n_reads, uint16_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char sc_smem[];
  ScatterShared& sh = *reinterpret_cast<ScatterShared*>(sc_smem);
  const int lmax = (int)meta[4];
  const long long n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  const int64_t n_tiles = (n_sites + kSC_Tile - 1) / kSC_Tile;
  sc_init_lut(sh);
  const bool stageable = lmax > 0 && lmax <= kSC_StageBytes - 16;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t s0 = tile * kSC_Tile;
    const int n_live = (int)min((int64_t)kSC_Tile, n_sites - s0);
    sc_pile_tile(sh, read_start, base_off, bases, meta, site_pos, s0, n_live, lmax, n_bases, stageable);
#pragma unroll
    for (int k = 0; k < kSC_Tile / kSC_Threads; k++) {
      const int i = threadIdx.x + k * kSC_Threads;
      if (i < n_live) {
        uint2* dst = reinterpret_cast<uint2*>(counts + (s0 + i) * 12);
        dst[0] = sh.tot[0][i]; dst[1] = sh.tot[1][i]; dst[2] = sh.tot[2][i];
      }
    }
  }
}

// ---- fused kernel: reads -> screened sites -------------------------------------------------------------------
// Each CTA owns a CONTIGUOUS range of tiles (balanced like the screen's warps: screen_range_lo) and emits the
// undecided sites of its range in ascending order into its own segment of `list`, which starts at the CTA's first
// site index (so it cannot overflow); wcount[cta] = how many.  The concatenation of the segments in CTA order is
// the sorted list of undecided sites — what trio_list_kernel consumes (segment granule kSC_Tile sites).
// counts: only the rows of undecided sites are written.
__global__ void __launch_bounds__(kSC_Threads, 3)
pileup_screen_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                     const char* __restrict__ bases, long long* meta, const int64_t* __restrict__ site_pos,
                     int64_t n_sites, int64_t n_reads, uint16_t* __restrict__ counts, const ScreenConsts k,
                     uint32_t* __restrict__ list, uint32_t* __restrict__ wcount) {
  extern __shared__ __align__(16) unsigned char sc_smem[];
  ScatterShared& sh = *reinterpret_cast<ScatterShared*>(sc_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lmax = (int)meta[4];
  const long long n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  const uint32_t n_tiles = (uint32_t)((n_sites + kSC_Tile - 1) / kSC_Tile);
  const uint32_t base = n_tiles / gridDim.x, rem = n_tiles % gridDim.x;
  const uint32_t t_lo = screen_range_lo(blockIdx.x, base, rem), t_hi = screen_range_lo(blockIdx.x + 1, base, rem);
  uint32_t* my_list = list + (size_t)t_lo * kSC_Tile;
  sc_init_lut(sh);
  if (threadIdx.x == 0) sh.written = 0;
  const bool stageable = lmax > 0 && lmax <= kSC_StageBytes - 16;
  for (uint32_t tile = t_lo; tile < t_hi; tile++) {
    const int64_t s0 = (int64_t)tile * kSC_Tile;
    const int n_live = (int)min((int64_t)kSC_Tile, n_sites - s0);
    sc_pile_tile(sh, read_start, base_off, bases, meta, site_pos, s0, n_live, lmax, n_bases, stageable);
    // screen 256 consecutive positions per round, in order, so the emitted sites stay ascending
#pragma unroll 1
    for (int kk = 0; kk < kSC_Tile / kSC_Threads; kk++) {
      const int i = threadIdx.x + kk * kSC_Threads;
      uint32_t w[6];
      bool undecided = false;
      if (i < n_live) {
        const uint2 d = sh.tot[0][i], m = sh.tot[1][i], c = sh.tot[2][i];
        w[0] = d.x; w[1] = d.y; w[2] = m.x; w[3] = m.y; w[4] = c.x; w[5] = c.y;
        undecided = !screen_level1<false>(w, k) && !screen_site<false>(w, k) &&
                    !(k.l3_lambda > 0 && screen_level3_words<false>(w, k));
      }
      const uint32_t uw = __ballot_sync(0xffffffffu, undecided);
      if (lane == 0) sh.warp_cnt[warp] = __popc(uw);
      __syncthreads();
      uint32_t before = sh.written, total = 0;
#pragma unroll
      for (int q = 0; q < kSC_Threads / 32; q++) {
        const uint32_t v = sh.warp_cnt[q];
        if (q < warp) before += v;
        total += v;
      }
      if (undecided) {
        const uint32_t site = (uint32_t)(s0 + i);
        my_list[before + __popc(uw & ((1u << lane) - 1u))] = site;
        uint2* dst = reinterpret_cast<uint2*>(counts + (size_t)site * 12);
        dst[0] = make_uint2(w[0], w[1]); dst[1] = make_uint2(w[2], w[3]); dst[2] = make_uint2(w[4], w[5]);
      }
      __syncthreads();
      if (threadIdx.x == 0) sh.written += total;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) wcount[blockIdx.x] = sh.written;
}

inline int scatter_grid(int sms, int64_t n_tiles, int per_sm, int max_ctas) {
  int64_t g = std::min<int64_t>(std::max<int64_t>(n_tiles, 1), (int64_t)sms * std::max(per_sm, 1));
  return (int)std::min<int64_t>(g, max_ctas);
}

}  // namespace dnv
