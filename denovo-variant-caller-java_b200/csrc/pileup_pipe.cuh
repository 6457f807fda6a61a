// pileup_pipe.cuh — the pileup kernel for DENSE candidate lists (consecutive positions: the every-position layout
// of BASELINE config 4 and the dense-candidate layout of config 3), as a producer / consumer pipeline inside the CTA.
//
// Reads are sorted by (person, start) and their alignedSequence strings are concatenated in that order, so the reads
// that can reach a tile of consecutive positions are ONE contiguous byte range of `bases` per person.  A CTA owns a
// tile of 128*S consecutive positions at a time (tiles are handed out by a ticket) and has
//   * ONE PRODUCER WARP that only moves data: per tile it reads the tile's record (position, density, the three read
//     windows: pileup_windows_kernel, a pre-pass with one warp per tile), and per person it waits until the consumers
//     are done with the stage and starts THREE TMA bulk copies (cp.async.bulk -> one mbarrier): ALL the person's bases
//     for the tile, and the slices of read_start[] and base_off[] that describe them, raw — every base is read from
//     HBM once per tile it touches ((tile + read length) / tile of the algorithmic bytes);
//   * FOUR CONSUMER WARPS that never touch global memory on the dense path except to write the counts: a thread owns
//     S neighbouring sites, binary-searches its sub-window of the staged reads and per read takes S+4 staged bytes as
//     aligned words + funnel shifts; each base goes through a 256-entry shared-memory LUT whose entry is the
//     increment of a packed accumulator (6-bit lanes A, C, G, T and "not a base" — anything outside {A,C,G,T,-}: the
//     reference's IllegalArgumentException, M/ReadSummary.java:55), flushed into the 16-bit counts every 63 reads.
// Staging a whole person-tile at once matters: every lane of a warp then walks ITS reads in the same pass (lane
// efficiency = mean / max trip count, ~0.7); with smaller chunks only the lanes whose windows overlap the chunk work
// while the warp still runs the longest lane's trip count (0.37-0.5 measured, profiles/r02_pileup_pipe_ncu.md).
// The stage is single buffered (36 KB of bases + 6 KB of descriptors: five CTAs per SM cover each other's refills);
// one full / one empty mbarrier hand it back and forth, so the consumers' instruction stream has no global-load
// latency and no CTA-wide barrier in it, and a refill costs one TMA round trip (the two global loads the producer
// needs for the copy — the byte range of the reads — are issued before it waits for the stage).  A person-tile that does not fit the stage (deep pileups) is
// cut into chunks of reads; tiles that are not consecutive positions, and reads longer than the stage, take the
// per-site path (the consumers walk their own windows).  No atomics on the counts, deterministic, counts written once.
//
// Replaces `new ReadSummary(reads, position)` x3 per site (M/ReadSummary.java:45-59).
#pragma once
#include "dnv_internal.cuh"
#include "pileup_stage.cuh"

namespace dnv {

constexpr int kPP_Consumers = 128;
constexpr int kPP_Threads = kPP_Consumers + 32;
#ifndef DNV_PP_STAGE_KB
#define DNV_PP_STAGE_KB 36
#endif
constexpr int kPP_StageBytes = DNV_PP_STAGE_KB * 1024;  // staged bases per step
constexpr int kPP_ReadCap = 384;                        // read descriptors per step
constexpr int kPP_Pad = 64;                             // readable slack on both sides of the staged bytes

enum { kPP_Dense = 0, kPP_Sparse = 1, kPP_Stop = 2 };
struct PipeStep {    // what the producer tells the consumers about one step
  long long s0;      // first site of the tile
  int P0_lo;         // low word of the tile's first position (read starts are used relative to it: 32 bits do)
  int n_live;        // sites of the tile
  int person, m;     // dense: person and number of staged reads
  int kind;
  int first, last;   // first / last chunk of this person: zero / write the totals
  int skip;          // the staged slices of read_start[] / base_off[] begin `skip` (0 / 1) entries before the first read
  int bias;          // offset of a read's first base in stage[] = low word of its base_off + bias
};
// record of a tile (pileup_windows_kernel): [0..5] = per person the first read with start > P0 - lmax and the first
// with start > P0 + n_live - 1; [6] = P0; [7] = 1 if the tile's positions are consecutive
constexpr int kPP_TileWords = 8;

template <int S>
struct PipeShared {
  static_assert(S == 8, "the edge-mask table is for 8-site windows");
  alignas(128) unsigned char stage[kPP_Pad + kPP_StageBytes + kPP_Pad];
  alignas(16) long long rs_raw[kPP_ReadCap + 2];  // read_start[] of the staged reads (16-byte aligned slice)
  alignas(16) long long bo_raw[kPP_ReadCap + 4];  // base_off[] of the staged reads and one more
  uint32_t lut[256];
  unsigned long long keep[8][9];  // keep[t_lo][t_hi]: 0xFF in bytes t_lo <= k < t_hi of an 8-byte window
  PipeStep step[2];
  alignas(8) unsigned long long full, empty;
};

// Every warp of the pipeline kernel passes here on its way out: the last one hands the status words (person
// offsets, longest read, error word) to the host's pinned buffer, so a call ends with the kernel instead of a copy
// behind it.  meta[7] counts the warps gone (zeroed with meta by the caller).
__device__ __forceinline__ void pipe_leave(long long* meta, long long* host_meta, int lane) {
  if (!host_meta) return;
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    const unsigned long long all = (unsigned long long)gridDim.x * (kPP_Threads / 32);
    if (atomicAdd((unsigned long long*)&meta[7], 1ull) == all - 1ull) {
      __threadfence();
#pragma unroll
      for (int q = 0; q < 6; q++) host_meta[q] = *(volatile long long*)&meta[q];
    }
  }
}

template <int S>
__global__ void __launch_bounds__(kPP_Threads, 5)
pileup_pipe_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const char* __restrict__ bases, long long* meta, const long long* __restrict__ tiles,
                   const int64_t* __restrict__ site_pos, int64_t n_sites, int64_t n_reads,
                   uint16_t* __restrict__ counts, long long* host_meta) {
  extern __shared__ __align__(128) unsigned char pp_smem[];
  PipeShared<S>& sh = *reinterpret_cast<PipeShared<S>*>(pp_smem);
  constexpr int kSites = kPP_Consumers * S;
  constexpr uint32_t kFull = 0xffffffffu;
  const int lmax = (int)meta[4];
  const long long n_bases = n_reads > 0 ? base_off[n_reads] : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (n_sites + kSites - 1) / kSites;
  // LUT: increment of the packed accumulator (6-bit lanes A, C, G, T; bit 24 = not a base)
  for (int i = threadIdx.x; i < 256; i += kPP_Threads)
    sh.lut[i] = i == 'A' ? 1u : i == 'C' ? 1u << 6 : i == 'G' ? 1u << 12 : i == 'T' ? 1u << 18
              : i == '-' ? 0u : 1u << 24;
  for (int i = threadIdx.x; i < 72; i += kPP_Threads) {
    const int t_lo = i / 9, t_hi = i % 9;
    const unsigned long long below_hi = t_hi >= 8 ? ~0ull : ~(~0ull << (8 * t_hi));
    sh.keep[t_lo][t_hi] = t_hi > t_lo ? (below_hi & (~0ull << (8 * t_lo))) : 0ull;
  }
  if (threadIdx.x == 0) {
    mbar_init(&sh.full, 1);
    mbar_init(&sh.empty, kPP_Consumers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t it = 0;  // steps so far: descriptor buffer it & 1, barrier phase it & 1

  if (warp == kPP_Consumers / 32) {
    // ------------------------------------------------ producer warp ------------------------------------------------
    // One flat loop over steps.  The global loads a step needs (read descriptors) are issued BEFORE the wait for a
    // free buffer, so their latency overlaps the consumers' work on the buffers already filled.
    const uintptr_t bases_abs = reinterpret_cast<uintptr_t>(bases);
    const int chunk_reads = min(kPP_ReadCap, (kPP_StageBytes - 16) / max(lmax, 1));
    struct TileInfo { int64_t s0, P0, win[6]; int n_live; bool dense; };
    auto fetch = [&](TileInfo& t) -> bool {  // ticket + one round trip for the tile's record
      long long tile = 0;
      if (lane == 0) tile = (long long)atomicAdd((unsigned long long*)&meta[6], 1ull);
      tile = __shfl_sync(kFull, tile, 0);
      if (tile >= n_tiles) return false;
      t.s0 = tile * kSites;
      t.n_live = (int)min((int64_t)kSites, n_sites - t.s0);
      const long long w = lane < kPP_TileWords ? __ldg(tiles + tile * kPP_TileWords + lane) : 0;
#pragma unroll
      for (int q = 0; q < 6; q++) t.win[q] = __shfl_sync(kFull, w, q);
      t.P0 = __shfl_sync(kFull, w, 6);
      t.dense = __shfl_sync(kFull, w, 7) != 0;
      return true;
    };
    TileInfo cur, nxt;
    bool more = fetch(cur);
    while (more) {
      // the steps of this tile: one per person and chunk of reads (at least one per person, even without reads: the
      // consumers zero and write the counts in it), or a single "sparse tile" step.  The NEXT tile's ticket and
      // record are fetched after the first step is on its way, while the consumers are busy with it.
      bool have_nxt = false, more_nxt = false;
      int p = 0;
      int64_t c_first = cur.win[0], c_last = cur.win[1], c0 = c_first;
      for (int k = 0, tile_done = 0; !tile_done; k++) {
        const int b = (int)(it & 1u);
        const int m = cur.dense ? (int)min((int64_t)chunk_reads, c_last - c0) : 0;
        long long g0 = 0, g1 = 0;
        if (m > 0) {  // the byte range of the reads: in flight while we wait for the stage
          g0 = __ldg(base_off + c0);
          g1 = __ldg(base_off + c0 + m);
        }
        if (!have_nxt && k > 0) {
          more_nxt = fetch(nxt);
          have_nxt = true;
        }
        mbar_wait_sleepy(&sh.empty, (it & 1u) ^ 1u, 20000u);  // the consumers are done with the previous step: the stage is free
        if (!cur.dense) {
          if (lane == 0) {
            sh.step[b] = PipeStep{cur.s0, (int)cur.P0, cur.n_live, 0, 0, kPP_Sparse, 1, 1, 0, 0};
            mbar_arrive(&sh.full);
          }
          tile_done = 1;
        } else {
          const int align = (int)((bases_abs + (uintptr_t)g0) & 15);
          const long long a0 = g0 - align;  // stage[i] <-> bases[a0 + i], 16-byte aligned
          const int n16 = m > 0 ? (int)((g1 - a0 + 15) >> 4) : 0;
          // the first / last 16 bytes of the buffer may straddle its ends: byte loads inside [0, n_bases) for those
          int v_lo = 0, v_hi = n16;
          if (n16 > 0 && a0 < 0) v_lo = 1;
          if (n16 > v_lo && a0 + 16ll * n16 > n_bases) v_hi = n16 - 1;
          if (lane < 16) {
            if (v_lo == 1) {
              const long long gb = a0 + lane;
              sh.stage[kPP_Pad + lane] = (gb >= 0 && gb < n_bases) ? (unsigned char)__ldg(bases + gb) : (unsigned char)'-';
            }
            if (v_hi < n16 && v_hi >= v_lo) {
              const long long gb = a0 + 16ll * v_hi + lane;
              sh.stage[kPP_Pad + 16 * v_hi + lane] = (gb >= 0 && gb < n_bases) ? (unsigned char)__ldg(bases + gb) : (unsigned char)'-';
            }
          }
          // descriptor slices: entries [ce, ce + n) of read_start[] and [ce, ce + n + ...) of base_off[], ce even
          // (the arrays are 16-byte aligned: launch_pileup_pipe checks); a pair that would straddle the end of an
          // array is loaded by hand
          const int skip = (int)(c0 & 1);
          const int64_t ce = c0 - skip;
          int n_rs = m > 0 ? (skip + m + 1) & ~1 : 0;      // entries ce .. ce + n_rs - 1 cover reads c0 .. c0 + m - 1
          int n_bo = m > 0 ? (skip + m + 2) & ~1 : 0;      // ... and base_off[c0 + m]
          if (ce + n_rs > n_reads) {
            n_rs -= 2;
            if (lane < 2) sh.rs_raw[n_rs + lane] = ce + n_rs + lane < n_reads ? __ldg(read_start + ce + n_rs + lane) : 0;
          }
          if (ce + n_bo > n_reads + 1) {
            n_bo -= 2;
            if (lane < 2) sh.bo_raw[n_bo + lane] = ce + n_bo + lane < n_reads + 1 ? __ldg(base_off + ce + n_bo + lane) : 0;
          }
          __syncwarp();  // the lanes' edge writes happen before lane 0's release below
          if (lane == 0) {
            sh.step[b] = PipeStep{cur.s0, (int)cur.P0, cur.n_live, p, m, kPP_Dense, c0 == c_first ? 1 : 0,
                                  c0 + m >= c_last ? 1 : 0, skip, align - (int)g0};
            const uint32_t tx_b = v_hi > v_lo ? (uint32_t)(v_hi - v_lo) * 16u : 0u;
            const uint32_t tx = tx_b + (uint32_t)(n_rs + n_bo) * 8u;
            if (tx) {
              mbar_expect_tx(&sh.full, tx);
              if (tx_b) tma_bulk_g2s(sh.stage + kPP_Pad + 16 * v_lo, bases + a0 + 16ll * v_lo, tx_b, &sh.full);
              if (n_rs) tma_bulk_g2s(sh.rs_raw, read_start + ce, (uint32_t)n_rs * 8u, &sh.full);
              if (n_bo) tma_bulk_g2s(sh.bo_raw, base_off + ce, (uint32_t)n_bo * 8u, &sh.full);
            } else {
              mbar_arrive(&sh.full);
            }
          }
          c0 += m;
          if (c0 >= c_last) {
            if (++p == 3) {
              tile_done = 1;
            } else {
              c0 = c_first = p == 1 ? cur.win[2] : cur.win[4];
              c_last = p == 1 ? cur.win[3] : cur.win[5];
            }
          }
        }
        it++;
      }
      if (!have_nxt) more_nxt = fetch(nxt);
      cur = nxt;
      more = more_nxt;
    }
    mbar_wait_sleepy(&sh.empty, (it & 1u) ^ 1u, 20000u);
    if (lane == 0) {
      sh.step[it & 1u].kind = kPP_Stop;
      mbar_arrive(&sh.full);
    }
    pipe_leave(meta, host_meta, lane);
    return;
  }

  // -------------------------------------------------- consumer warps -------------------------------------------------
  constexpr int NW = S / 4;
  const int my_off = threadIdx.x * S;  // offset of my first site from the tile's first position
  const uint32_t lut_s = smem_u32(sh.lut);
  uint32_t acc[S];
  uint32_t badacc = 0;
  int since_flush = 0, visits = 0;
  bool wrote = false;  // this person's counts of my sites are in global memory already (more than 63 reads over them)
  // 6-bit lanes -> 16-bit totals (A | C << 16, G | T << 16), straight into the output: a plain store the first time
  // (the usual case: once per person), my own read-modify-write after that (every 63 visits of a deep pileup)
  auto flush = [&](int64_t s0, int p, int lim) {
#pragma unroll
    for (int k = 0; k < S; k++) {
      const uint32_t a = acc[k];
      if (k < lim) {
        uint2* dst = reinterpret_cast<uint2*>(counts + (s0 + my_off + k) * 12 + p * 4);
        uint2 t = make_uint2(0u, 0u);
        if (wrote) t = *dst;
        t.x += (a & 0x3Fu) | ((a << 10) & 0x3F0000u);
        t.y += ((a >> 12) & 0x3Fu) | ((a >> 2) & 0x3F0000u);
        *dst = t;
      }
      badacc |= a >> 24;
      acc[k] = 0;
    }
    wrote = true;
    visits += since_flush;
    since_flush = 0;
  };
#pragma unroll
  for (int k = 0; k < S; k++) acc[k] = 0;
  for (;; it++) {
    const int b = (int)(it & 1u);
    mbar_wait_sleepy(&sh.full, it & 1u, 20000u);
    const int kind = sh.step[b].kind;
    if (kind == kPP_Stop) break;
    const int64_t s0 = sh.step[b].s0;
    const int n_live = sh.step[b].n_live;
    if (kind == kPP_Sparse) {
      mbar_arrive(&sh.empty);
      // sparse tile: every thread walks its own windows, one site at a time
      bool bad = false, over = false;
      for (int k = 0; k < S; k++) {
        const int64_t s = s0 + my_off + k;
        if (my_off + k >= n_live) break;
        const int64_t pos = site_pos[s];
#pragma unroll 1
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
            const char c = __ldg(bases + b0 + off);
            if (c == '-') continue;
            const bool iA = c == 'A', iC = c == 'C', iG = c == 'G', iT = c == 'T';
            c4[0] += iA; c4[1] += iC; c4[2] += iG; c4[3] += iT;
            bad |= !(iA | iC | iG | iT);
          }
#pragma unroll
          for (int a = 0; a < 4; a++) over |= c4[a] > 65535u;
          *reinterpret_cast<uint2*>(counts + s * 12 + p * 4) =
              make_uint2((c4[0] & 0xFFFF) | (c4[1] << 16), (c4[2] & 0xFFFF) | (c4[3] << 16));
        }
      }
      if (bad && my_off < n_live) {
        atomicMax((unsigned long long*)&meta[5], pileup_err(s0 + my_off, true, kPileErrBadBase));
      } else if (over && my_off < n_live) {  // more than 65535 reads over one site
        atomicMax((unsigned long long*)&meta[5], pileup_err(0, false, kPileErrInvalid));
      }
      continue;
    }
    // dense chunk
    const int m = sh.step[b].m, p = sh.step[b].person;
    const int is_last = sh.step[b].last;
    if (sh.step[b].first) {
#pragma unroll
      for (int k = 0; k < S; k++) acc[k] = 0;
      badacc = 0; since_flush = 0; visits = 0; wrote = false;
    }
    {
      const int skip = sh.step[b].skip;
      const int P0_lo = sh.step[b].P0_lo;
      const int* rs_lo = reinterpret_cast<const int*>(sh.rs_raw + skip);  // low words: entry j at [2 * j]
      const int* bo_lo = reinterpret_cast<const int*>(sh.bo_raw + skip);
      const uint32_t stage_s = smem_u32(sh.stage + kPP_Pad) + (uint32_t)sh.step[b].bias;
      const int lim = n_live - my_off;  // my sites at or past it are not requested (last, partial tile)
      const int off0 = P0_lo + my_off;  // low word of my first site's position (differences of low words are exact)
      int lo = 0, hi = m;  // first staged read that starts after my_off - lmax (earlier ones end before my first site)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (off0 - rs_lo[2 * mid] >= lmax) lo = mid + 1; else hi = mid;
      }
      // The descriptors of the next read are fetched while this one is processed (no load latency in front of the
      // two tests), and base_off of the next read is this read's end.  Blocks of at most 63 iterations: the 6-bit
      // lanes cannot overflow inside one, so the loop itself carries no flush counter.
      const uint32_t keep_s = smem_u32(sh.keep);
      int rs_n = rs_lo[2 * lo], b0_n = bo_lo[2 * lo], b1_n = bo_lo[2 * lo + 2];
      int skipped = 0;
      bool done = false;
      for (int jb = lo; jb < m && !done; jb += 63) {
        const int je = min(jb + 63, m);
        int j = jb;
        for (; j < je; j++) {
          const int q = off0 - rs_n;  // offset of my first site inside the read (M/ReadSummary.java:48)
          const int b0 = b0_n, b1 = b1_n;
          rs_n = rs_lo[2 * j + 2];    // (one entry past the last read at the end: inside the arrays, unused)
          b0_n = b1;
          b1_n = bo_lo[2 * j + 4];
          if (q < -(S - 1)) { done = true; break; }  // this read and all later ones start after my last site
          const int len = b1 - b0;
          if (q >= len) { skipped++; continue; }  // read ends before my first site
          const uint32_t addr = stage_s + (uint32_t)(b0 + q);  // >= -(S-1) from the staged bytes: inside the front pad
          const uint32_t wa = addr & ~3u;
          const uint32_t shft = (addr & 3u) * 8u;
          uint32_t x[NW + 1], w[NW];
#pragma unroll
          for (int i = 0; i <= NW; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"(wa + 4 * i));
#pragma unroll
          for (int i = 0; i < NW; i++) w[i] = __funnelshift_r(x[i], x[i + 1], shft);
          // bytes of my window outside [0, len) of this read are not its bases, and window positions at or past
          // n_live (last, partial tile) are not requested sites — the reference only looks at the base AT a
          // candidate position (M/ReadSummary.java:48-55), so an 'N' next to the last site must not fail the
          // batch: turn both kinds into '-' (branch free, one table look-up).  t_lo in [0, 7], t_hi in [0, 8]
          const int t_lo = max(0, -q), t_hi = max(min(min(S, len - q), lim), 0);
          uint32_t k0, k1;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(k0), "=r"(k1) : "r"(keep_s + (uint32_t)(t_lo * 9 + t_hi) * 8u));
          w[0] = (w[0] & k0) | (0x2D2D2D2Du & ~k0);
          w[1] = (w[1] & k1) | (0x2D2D2D2Du & ~k1);
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
        }
        since_flush = j - jb;
        if (!done && je < m) flush(s0, p, lim);  // more reads to come in this step
      }
      visits -= skipped;
    }
    mbar_arrive(&sh.empty);  // my last read of the stage and of this step's descriptors is behind me
    flush(s0, p, n_live - my_off);  // every step ends with empty lanes (a person has several only in deep pileups)
    if (is_last) {
      if (badacc != 0 && my_off < n_live) {
        atomicMax((unsigned long long*)&meta[5], pileup_err(s0 + my_off, true, kPileErrBadBase));
      } else if (visits > 65535 && my_off < n_live) {  // a 16-bit total could have wrapped: fail loudly
        atomicMax((unsigned long long*)&meta[5], pileup_err(0, false, kPileErrInvalid));
      }
    }
  }
  pipe_leave(meta, host_meta, lane);
}

// pre-pass: one warp per tile writes the tile's record (kPP_TileWords words): the six window bounds and whether the
// tile's positions are consecutive (site_pos[s0 + i] == P0 + i: 32 coalesced loads per lane, all issued before the
// searches' dependent rounds so that they are in flight together).  Round 2 had a separate streaming kernel for the
// density flags (18 us + a memset of the records per 8 M positions).
__global__ void __launch_bounds__(256)
pileup_windows_kernel(const int64_t* __restrict__ read_start, const long long* __restrict__ meta,
                      const int64_t* __restrict__ site_pos, int64_t n_sites, int tile_sites, int stage_limit,
                      long long* __restrict__ tiles) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_tiles = (n_sites + tile_sites - 1) / tile_sites;
  if (tile >= n_tiles) return;
  const int lmax = (int)meta[4];
  const int64_t s0 = tile * tile_sites;
  const int n_live = (int)min((int64_t)tile_sites, n_sites - s0);
  const int64_t P0 = __ldg(site_pos + s0);
  bool gap = false;
  for (int i0 = 0; i0 < n_live; i0 += 32 * 8) {
    long long v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = i0 + 32 * k + lane;
      v[k] = i < n_live ? __ldg(site_pos + s0 + i) : P0 + i;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) gap |= v[k] != P0 + (i0 + 32 * k + lane);
  }
  int64_t win[6];
  ps_warp_upper_bound6(read_start, meta, P0 - lmax, P0 + n_live - 1, lane, win);
  const bool dense = lmax > 0 && lmax <= stage_limit && !__any_sync(0xffffffffu, gap);
  long long w = lane == 6 ? P0 : lane == 7 ? (dense ? 1 : 0) : 0;
#pragma unroll
  for (int q = 0; q < 6; q++) w = lane == q ? win[q] : w;
  if (lane < kPP_TileWords) tiles[tile * kPP_TileWords + lane] = w;
}

// The pipeline kernel wants 16-byte aligned read_start / base_off (its bulk copies of their slices).
template <int S>
static cudaError_t launch_pileup_pipe(cudaStream_t st, int sms, const int64_t* d_read_start, const int64_t* d_base_off,
                                      const char* d_bases, long long* meta, long long* tiles, const int64_t* d_site_pos,
                                      int64_t n_sites, int64_t n_reads, uint16_t* d_counts, long long* host_meta) {
  auto kern = pileup_pipe_kernel<S>;
  const size_t smem = sizeof(PipeShared<S>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPP_Threads, smem);
  if (e != cudaSuccess) return e;
  constexpr int kSites = kPP_Consumers * S;
  const int64_t n_tiles = (n_sites + kSites - 1) / kSites;
  const int64_t wblocks = (n_tiles * 32 + 255) / 256;
  pileup_windows_kernel<<<(unsigned)wblocks, 256, 0, st>>>(d_read_start, meta, d_site_pos, n_sites, kSites,
                                                           std::min(kPP_StageBytes - 16, 0x7FFFFFFF), tiles);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)sms * std::max(per_sm, 1)));
  kern<<<grid, kPP_Threads, smem, st>>>(d_read_start, d_base_off, d_bases, meta, tiles, d_site_pos, n_sites, n_reads, d_counts, host_meta);
  return cudaGetLastError();
}

}  // namespace dnv
