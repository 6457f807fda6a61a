// pileup_pipe.cuh — the pileup kernel for DENSE candidate lists (consecutive positions: the every-position layout
// of BASELINE config 4 and the dense-candidate layout of config 3), as a producer / consumer pipeline inside the CTA.
//
// Reads are sorted by (person, start) and their alignedSequence strings are concatenated in that order, so the reads
// that can reach a tile of consecutive positions are ONE contiguous byte range of `bases` per person.  A CTA owns a
// tile of 128*S consecutive positions at a time (tiles are handed out by a ticket) and has
//   * ONE PRODUCER WARP: per tile it finds the three read windows (six 32-ary searches, their loads in flight
//     together); per person it writes the read descriptors (start relative to the tile, offset in the stage, length)
//     into the descriptor buffer the consumers are NOT reading, waits until they are done with the stage buffer and
//     starts ONE TMA bulk copy (cp.async.bulk -> mbarrier) of ALL the person's bytes for the tile — every base is
//     read from HBM once per tile it touches ((tile + read length) / tile of the algorithmic bytes);
//   * FOUR CONSUMER WARPS that never touch global memory on the dense path except to write the counts: a thread owns
//     S neighbouring sites, binary-searches its sub-window of the staged reads and per read takes S+4 staged bytes as
//     aligned words + funnel shifts; each base goes through a 256-entry shared-memory LUT whose entry is the
//     increment of a packed accumulator (6-bit lanes A, C, G, T and "not a base" — anything outside {A,C,G,T,-}: the
//     reference's IllegalArgumentException, M/ReadSummary.java:55), flushed into 16-bit totals every 63 reads.
// Staging a whole person-tile at once matters: every lane of a warp then walks ITS reads in the same pass (lane
// efficiency = mean / max trip count, ~0.7); with smaller chunks only the lanes whose windows overlap the chunk work
// while the warp still runs the longest lane's trip count (0.37-0.5 measured, profiles/r02_pileup_pipe_ncu.md).
// The stage buffer is single (36 KB: four CTAs per SM cover each other's refills), the descriptors are double
// buffered; one full / one empty mbarrier hand the stage back and forth, so the consumers' instruction stream has no
// global-load latency and no CTA-wide barrier in it.  A person-tile that does not fit the stage (deep pileups) is
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
  long long s0, P0;  // first site of the tile, its position
  int n_live;        // sites of the tile
  int person, m;     // dense: person and number of staged reads
  int kind;
  int first, last;   // first / last chunk of this person: zero / write the totals
};

template <int S>
struct PipeShared {
  alignas(128) unsigned char stage[kPP_Pad + kPP_StageBytes + kPP_Pad];
  int2 desc[2][kPP_ReadCap];    // x = read start - P0; y = offset of the read's first base in stage[] | length << 16
  uint2 tot[S][kPP_Consumers];  // 16-bit totals (A | C << 16, G | T << 16) of thread t's site k: conflict free
  uint32_t lut[256];
  PipeStep step[2];
  alignas(8) unsigned long long full, empty;
};

template <int S>
__global__ void __launch_bounds__(kPP_Threads, 4)
pileup_pipe_kernel(const int64_t* __restrict__ read_start, const int64_t* __restrict__ base_off,
                   const char* __restrict__ bases, long long* meta, const int64_t* __restrict__ site_pos,
                   int64_t n_sites, int64_t n_reads, uint16_t* __restrict__ counts) {
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
  if (threadIdx.x == 0) {
    mbar_init(&sh.full, 1);
    mbar_init(&sh.empty, kPP_Consumers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool stageable = lmax > 0 && lmax <= min(kPP_StageBytes - 16, 0xFFFF);
  uint32_t it = 0;  // steps so far: descriptor buffer it & 1, barrier phase it & 1

  if (warp == kPP_Consumers / 32) {
    // ------------------------------------------------ producer warp ------------------------------------------------
    // One flat loop over steps.  The global loads a step needs (read descriptors) are issued BEFORE the wait for a
    // free buffer, so their latency overlaps the consumers' work on the buffers already filled.
    const uintptr_t bases_abs = reinterpret_cast<uintptr_t>(bases);
    const int chunk_reads = min(kPP_ReadCap, (kPP_StageBytes - 16) / max(lmax, 1));
    int64_t s0 = 0, P0 = 0, win[6] = {0, 0, 0, 0, 0, 0}, c0 = 0, c_first = 0, c_last = 0;
    int n_live = 0, p = 0;
    bool dense = false;
    auto next_tile = [&]() -> bool {
      long long tile = 0;
      if (lane == 0) tile = (long long)atomicAdd((unsigned long long*)&meta[6], 1ull);
      tile = __shfl_sync(kFull, tile, 0);
      if (tile >= n_tiles) return false;
      s0 = tile * kSites;
      n_live = (int)min((int64_t)kSites, n_sites - s0);
      P0 = __ldg(site_pos + s0);
      // windows of reads per person: start in (P0 - lmax, P0 + n_live - 1]; the six searches run together
      ps_warp_upper_bound6(read_start, meta, P0 - lmax, P0 + n_live - 1, lane, win);
      // strictly increasing positions with last - first == n - 1  <=>  consecutive
      dense = stageable && __ldg(site_pos + s0 + n_live - 1) - P0 == n_live - 1;
      if (dense) {
        int ok = 1;
#pragma unroll 8
        for (int i = lane; i < n_live - 1; i += 32) ok &= __ldg(site_pos + s0 + i + 1) > __ldg(site_pos + s0 + i);
        dense = __all_sync(kFull, ok);
      }
      p = 0;
      c0 = c_first = win[0];
      c_last = win[1];
      return true;
    };
    bool more = next_tile();
    while (more) {
      const int b = (int)(it & 1u);
      const int m = dense ? (int)min((int64_t)chunk_reads, c_last - c0) : 0;  // may be 0: a person without reads here
      long long g0 = 0, g1 = 0, a0 = 0;
      if (m > 0) {
        g0 = __ldg(base_off + c0);
        g1 = __ldg(base_off + c0 + m);
        a0 = g0 - (long long)((bases_abs + (uintptr_t)g0) & 15);  // stage[i] <-> bases[a0 + i], 16-byte aligned
#pragma unroll 4
        for (int i = lane; i < m; i += 32) {  // descriptor buffer b: the consumers are at most reading the other one
          const long long b0 = __ldg(base_off + c0 + i);
          const int len = (int)(__ldg(base_off + c0 + i + 1) - b0);
          sh.desc[b][i] = make_int2((int)(__ldg(read_start + c0 + i) - P0), (int)(b0 - a0) | (len << 16));
        }
      }
      mbar_wait(&sh.empty, (it & 1u) ^ 1u);  // the consumers are done with the previous step: the stage is free
      if (!dense) {
        if (lane == 0) {
          sh.step[b] = PipeStep{s0, P0, n_live, 0, 0, kPP_Sparse, 1, 1};
          mbar_arrive(&sh.full);
        }
      } else {
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
        __syncwarp();  // the lanes' descriptor and edge writes happen before lane 0's release below
        if (lane == 0) {
          sh.step[b] = PipeStep{s0, P0, n_live, p, m, kPP_Dense, c0 == c_first ? 1 : 0, c0 + m >= c_last ? 1 : 0};
          const uint32_t tx = v_hi > v_lo ? (uint32_t)(v_hi - v_lo) * 16u : 0u;
          if (tx) {
            mbar_expect_tx(&sh.full, tx);
            tma_bulk_g2s(sh.stage + kPP_Pad + 16 * v_lo, bases + a0 + 16ll * v_lo, tx, &sh.full);
          } else {
            mbar_arrive(&sh.full);
          }
        }
      }
      it++;
      // advance: next chunk, next person (at least one step each, even without reads: the consumers zero and write
      // the totals in it), next tile
      if (!dense) {
        more = next_tile();
      } else {
        c0 += m;
        if (c0 >= c_last) {
          if (++p == 3) {
            more = next_tile();
          } else {
            c0 = c_first = p == 1 ? win[2] : win[4];
            c_last = p == 1 ? win[3] : win[5];
          }
        }
      }
    }
    mbar_wait(&sh.empty, (it & 1u) ^ 1u);
    if (lane == 0) {
      sh.step[it & 1u].kind = kPP_Stop;
      mbar_arrive(&sh.full);
    }
    return;
  }

  // -------------------------------------------------- consumer warps -------------------------------------------------
  constexpr int NW = S / 4;
  const int my_off = threadIdx.x * S;  // offset of my first site from the tile's first position
  const uint32_t lut_s = smem_u32(sh.lut);
  uint32_t acc[S];
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
#pragma unroll
  for (int k = 0; k < S; k++) acc[k] = 0;
  for (;; it++) {
    const int b = (int)(it & 1u);
    mbar_wait(&sh.full, it & 1u);
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
      for (int k = 0; k < S; k++) { acc[k] = 0; sh.tot[k][threadIdx.x] = make_uint2(0u, 0u); }
      badacc = 0; since_flush = 0; visits = 0;
    }
    {
      const int2* desc_b = sh.desc[b];
      const uint32_t stage_s = smem_u32(sh.stage + kPP_Pad);
      const int lim = n_live - my_off;  // my sites at or past it are not requested (last, partial tile)
      int lo = 0, hi = m;  // first staged read with rel > my_off - lmax (earlier ones end before my first site)
      const int key = my_off - lmax;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (desc_b[mid].x <= key) lo = mid + 1; else hi = mid;
      }
      for (int j = lo; j < m; j++) {
        const int2 d = desc_b[j];
        if (d.x > my_off + S - 1) break;
        const int q = my_off - d.x;  // offset of my first site inside the read (M/ReadSummary.java:48)
        const int len = (int)((uint32_t)d.y >> 16);
        if (q >= len) continue;      // read ends before my first site
        const int addr = (d.y & 0xFFFF) + q;  // >= -(S-1): inside the front pad
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
        const int t_lo = max(0, -q), t_hi = min(min(S, len - q), lim);
#pragma unroll
        for (int i = 0; i < NW; i += 2) {  // 8 window bytes at a time as one 64-bit mask
          const int a = min(max(t_lo - 4 * i, 0), 8), e = min(max(t_hi - 4 * i, 0), 8);
          const unsigned long long from = a >= 8 ? 0ull : ~0ull << (8 * a);
          const unsigned long long to = e >= 8 ? ~0ull : ~(~0ull << (8 * e));
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
    mbar_arrive(&sh.empty);  // my last read of the stage and of this step's descriptors is behind me
    if (is_last) {
      flush();
#pragma unroll
      for (int k = 0; k < S; k++) {
        if (my_off + k < n_live)
          *reinterpret_cast<uint2*>(counts + (s0 + my_off + k) * 12 + p * 4) = sh.tot[k][threadIdx.x];
      }
      if (badacc != 0 && my_off < n_live) {
        atomicMax((unsigned long long*)&meta[5], pileup_err(s0 + my_off, true, kPileErrBadBase));
      } else if (visits > 65535 && my_off < n_live) {  // a 16-bit total could have wrapped: fail loudly
        atomicMax((unsigned long long*)&meta[5], pileup_err(0, false, kPileErrInvalid));
      }
    }
  }
}

template <int S>
static cudaError_t launch_pileup_pipe(cudaStream_t st, int sms, const int64_t* d_read_start, const int64_t* d_base_off,
                                      const char* d_bases, long long* meta, const int64_t* d_site_pos, int64_t n_sites,
                                      int64_t n_reads, uint16_t* d_counts) {
  auto kern = pileup_pipe_kernel<S>;
  const size_t smem = sizeof(PipeShared<S>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPP_Threads, smem);
  if (e != cudaSuccess) return e;
  const int64_t n_tiles = (n_sites + kPP_Consumers * S - 1) / (kPP_Consumers * S);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)sms * std::max(per_sm, 1)));
  kern<<<grid, kPP_Threads, smem, st>>>(d_read_start, d_base_off, d_bases, meta, d_site_pos, n_sites, n_reads, d_counts);
  return cudaGetLastError();
}

}  // namespace dnv
