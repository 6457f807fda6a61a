// mendel.cu — kernel (c): VARIANT-mode Mendelian-inconsistency filter over one
// shard of the start-sorted (variant, call) stream.
//
// Reference semantics (streaming, M/VariantsBuffer.java + M/VariantCaller.java:120-210):
//   admission     checkAndAdd                         :172-185 (+ isSnp :264-282, FILTER :245-249)
//   trigger/evict canProcess / evictParents            :74-78, :147-157
//   join          getMatchingPair = FIRST queued entry (admission order) of each
//                 parent with start <= pos < end       :314-323
//   genotype      getGenotypeFromSNP, or ref+ref of the CHILD's reference base
//                 when the parent entry is not a SNP   :217-219, :228-239
//   call          PositionCall.isDenovo                :352-355
// On a stream sorted by start with end > start the streaming result equals the
// batch definition: for every admitted child SNP match
// each parent against all of that parent's admitted entries, lowest stream
// index first.  Batch form on the GPU, per parent with entries in stream order:
//   ub  = #entries with start <= pos                (binary search on start)
//   i*  = first i with prefix_max(end)[i] > pos     (binary search on the running max)
//   hit = i* < ub   (the first index where the running max exceeds pos is itself
//                    an entry with end > pos, and every earlier entry has end <= pos)
// The input-domain conditions (sorted, end > start for admitted entries) are
// CHECKED on the device and violations fail loudly with DNV_E_UNSORTED.
//
// Two kernels, each ONE pass with decoupled look-back (no separate scan passes):
//   A  mendel_admit_compact_kernel : reads each 32-byte record once; admission; ordered
//      compaction into the CHILD/MOM/DAD queues (3 sum chains) and the running max of
//      `end` of each parent queue (2 max chains), written straight from registers
//   B  mendel_join_emit_kernel     : per child SNP the two searches per parent, confined to
//      the block's window; truth-table lookup; ordered emission of PositionCalls
#include <limits.h>
#include <algorithm>

#include "dnv_internal.cuh"

namespace dnv {

enum { kChild = 0, kMom = 1, kDad = 2 };
constexpr uint8_t kCodeBlock = 0xF0;  // admitted parent entry that is not a SNP
constexpr uint8_t kCodeBad = 0xFF;    // base outside ACGT (the reference would throw)

__device__ __forceinline__ int allele_of(char c) {
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}
__device__ __forceinline__ int geno_of(int x, int y) {  // Genotype.valueOfPairAlleles, M/DenovoUtil.java:229-233
  if (x > y) { const int t = x; x = y; y = t; }
  return x * 4 - (x * (x - 1)) / 2 + (y - x);
}

constexpr int kMB = 256;  // records (kernel A) / child SNPs (kernel B) per block
#ifndef DNV_MENDEL_JI
#define DNV_MENDEL_JI 2
#endif
constexpr int kJI = DNV_MENDEL_JI;     // child SNPs per thread in kernel B
constexpr int kJTile = kJI * kMB;      // 512 children per block

struct MendelArrays {
  // per-person admitted entries in stream order (SoA queues)
  int64_t* start[3];
  int64_t* runmax[3];   // parents: running max of `end` over the queue up to and including the entry
  uint8_t* code[3];
  uint8_t* ref;         // child only: allele index of the reference base (or 0xFF)
  int64_t* src;         // child only: index in the input stream
  // decoupled look-back status, zeroed before every call: [63:62] 1 aggregate / 2 inclusive prefix, [61:0] value
  unsigned long long* st_sum[3];   // admitted counts per person
  unsigned long long* st_max[3];   // running max of end (+1; 0 = none) per parent
  unsigned long long* st_found;    // PositionCalls
  unsigned long long* win;         // per block of kernel B: [lo_mom, hi_mom, lo_dad, hi_dad] + 1 each, 0 = not written
  unsigned long long* totals;      // [0..2] admitted per person, [3] PositionCalls, [4] error: (record index + 1) << 2 | kind,
                                   // [5] blocks of kernel B gone, [6],[7] tickets
  unsigned long long* host_totals; // pinned host words: kernel B's last block to leave copies totals[0..4] there
};

// error word: index and kind travel in ONE atomicMax, so the reported kind always belongs to the reported record
constexpr unsigned long long kErrUnsorted = 1, kErrBadBase = 2;
constexpr unsigned long long kStAgg = 1ull << 62, kStPre = 2ull << 62, kStVal = (1ull << 62) - 1;

#ifndef DNV_MENDEL_SPIN_NS
#define DNV_MENDEL_SPIN_NS 0
#endif
// one warp: exclusive combination (sum or max) of all predecessor blocks' values
template <bool kMax>
__device__ __forceinline__ unsigned long long lookback(unsigned long long* status, unsigned int b,
                                                       unsigned long long mine, int lane) {
  unsigned long long excl = 0;
  if (b > 0) {
    if (lane == 0) atomicExch(&status[b], kStAgg | mine);
    int hi = (int)b - 1;
    for (;;) {
      const int p = hi - lane;
      unsigned long long v = kStPre;  // before block 0: empty inclusive prefix
      if (p >= 0) {
        // a lane whose predecessor has not published yet backs off: the kernels are bound by integer issue slots,
        // which a tight polling loop would take from the warps that still have records to decode (DNV_MENDEL_SPIN_NS)
        for (;;) {
          v = *(volatile unsigned long long*)&status[p];
          if ((v >> 62) != 0) break;
          if (DNV_MENDEL_SPIN_NS > 0) __nanosleep(DNV_MENDEL_SPIN_NS);
        }
      }
      const uint32_t is_pre = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
      const int stop = is_pre ? __ffs(is_pre) - 1 : 31;
      unsigned long long val = lane <= stop ? (v & kStVal) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, val, o);
        val = kMax ? (y > val ? y : val) : val + y;
      }
      excl = kMax ? (val > excl ? val : excl) : excl + val;
      if (is_pre) break;
      hi -= 32;
    }
  }
  if (lane == 0) {
    const unsigned long long incl = kMax ? (mine > excl ? mine : excl) : excl + mine;
    atomicExch(&status[b], kStPre | incl);
  }
  return excl;
}

// ---- kernel A: admission (checkAndAdd) + domain checks + ordered compaction into the three
// queues + running max of end for the parents, in ONE pass over the 32-byte records.
// A block owns kMTile consecutive records: ONE TMA bulk copy brings them into shared memory
// (32 KB in flight per block, five blocks per SM: the kernel no longer waits on its own loads
// record by record); thread t then handles records t, t + 256, ... (a warp = 32 consecutive
// records per round, which is what the ordered ranks are built on).  The essentials of a
// thread's records stay in registers while the block's totals go through the look-back chains.
// The running max of `end` is scanned inside a warp on 32-bit values relative to the tile's
// first start; a tile holding an admitted entry that ends 2^32 or more past it (no such genome)
// takes the same code on 64-bit values.
constexpr int kMI = 4;                    // records per thread
constexpr int kMTile = kMI * kMB;         // 1024 records = 32 KB
constexpr int kMW = kMB / 32;             // warps per block

// allele index of a base, -1 outside ACGT: (c >> 1) & 3 is a perfect hash of A, C, G, T (0, 1, 3, 2)
__device__ __forceinline__ int allele_fast(uint32_t c) {
  const uint32_t h = (c >> 1) & 3u;
  const uint32_t canon = (0x47544341u >> (8 * h)) & 0xFFu;  // "ACTG"[h]
  return canon == c ? (int)((0x2310u >> (4 * h)) & 3u) : -1;
}

template <typename T>
__device__ __forceinline__ void mendel_admit_tile(const uint4* s_rec, const dnv_vcall* __restrict__ calls, int64_t n,
                                                  const MendelArrays& A, unsigned int b, int n_here,
                                                  int (*s_cnt)[kMI * kMW + 1], unsigned long long (*s_max)[kMI * kMW + 1],
                                                  unsigned long long* s_excl, unsigned long long* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile0 = (int64_t)b * kMTile;
  const int64_t base = (int64_t)(((unsigned long long)s_rec[0].y << 32) | s_rec[0].x);  // first start of the tile
  int64_t st[kMI];
  uint32_t meta[kMI];  // bit0 admitted, bits 2:1 person, bits 15:8 genotype code, bits 23:16 ref allele, bits 31:24 rank in warp
  T myrun[kMI];        // warp-inclusive running max of (end + 1 - base) over entries of my own parent queue; 0 = none
#pragma unroll
  for (int j = 0; j < kMI; j++) {
    const int r = j * kMB + threadIdx.x;
    bool adm = false;
    int person = 3;
    int64_t start = 0, end = 0;
    uint32_t code = kCodeBlock, ref = 0xFF;
    if (r < n_here) {
      const uint4 lo = s_rec[2 * r], hi = s_rec[2 * r + 1];
      start = (int64_t)(((unsigned long long)lo.y << 32) | lo.x);
      end = (int64_t)(((unsigned long long)lo.w << 32) | lo.z);
      // hi.x variant_id; hi.y = gt0 | gt1 << 8 | ngt << 16 | person << 24;
      // hi.z = filter_pass | ref_len << 8 | ref_base << 16 | alt0_len << 24; hi.w = alt0_base | alt1_len << 8 | alt1_base << 16
      const int gt0 = (int)(int8_t)(hi.y & 0xFFu), gt1 = (int)(int8_t)((hi.y >> 8) & 0xFFu);
      const uint32_t ngt = (hi.y >> 16) & 0xFFu;
      person = (int)(hi.y >> 24);
      const uint32_t filter_pass = hi.z & 0xFFu, ref_len = (hi.z >> 8) & 0xFFu, ref_base = (hi.z >> 16) & 0xFFu;
      const uint32_t alt0_len = hi.z >> 24, alt0_base = hi.w & 0xFFu, alt1_len = (hi.w >> 8) & 0xFFu, alt1_base = (hi.w >> 16) & 0xFFu;
      const bool ins = (ngt >= 1 && gt0 > 0 && alt0_len != 1) || (ngt >= 2 && gt1 > 0 && alt1_len != 1);
      const bool del = ref_len != 1;
      const bool snp = (end == start + 1) && !del && !ins;
      const bool dot = (ngt >= 1 && gt0 == -1) || (ngt >= 2 && gt1 == -1);
      adm = person <= 2 && !((person == kChild && !snp) || dot || ngt != 2 || !filter_pass || ins || del);
      // stream must be sorted by start (the record before the tile's first one comes from global memory)
      int64_t prev = start;
      if (r > 0) prev = (int64_t)(((unsigned long long)s_rec[2 * r - 2].y << 32) | s_rec[2 * r - 2].x);
      else if (tile0 > 0) prev = calls[tile0 - 1].start;
      bool bad = prev > start;
      if (adm && (end <= start || start < 0 || end >= (1ll << 60))) bad = true;  // outside the batch-equivalence domain
      if (bad) atomicMax(&A.totals[4], ((unsigned long long)(tile0 + r + 1) << 2) | kErrUnsorted);
      if (adm) {
        if (end == start + 1) {  // admitted => ref_len == 1 and no insertion: getGenotypeFromSNP, :228-239
          const int a0 = allele_fast(gt0 > 0 ? alt0_base : ref_base);
          const int a1 = allele_fast(gt1 > 0 ? alt1_base : ref_base);
          code = (a0 < 0 || a1 < 0) ? kCodeBad : (uint32_t)geno_of(a0, a1);
        }
        const int rb = allele_fast(ref_base);
        ref = rb < 0 ? 0xFFu : (uint32_t)rb;
      }
    }
    int rank = 0;
#pragma unroll
    for (int p = 0; p < 3; p++) {
      const uint32_t m = __ballot_sync(0xffffffffu, adm && person == p);
      if (person == p) rank = __popc(m & ((1u << lane) - 1));
      if (lane == 0) s_cnt[p][j * kMW + warp] = __popc(m);
    }
    // running max of end (+1, relative to base; 0 = none) per parent queue: shfl_up hands lanes below the offset
    // their own value back, so no predicate is needed
    const T rel = (adm && person != kChild) ? (T)(unsigned long long)(end + 1 - base) : (T)0;
    T mine = 0;
#pragma unroll
    for (int par = 0; par < 2; par++) {
      T x = person == par + 1 ? rel : (T)0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const T y = __shfl_up_sync(0xffffffffu, x, o);
        x = y > x ? y : x;
      }
      if (person == par + 1) mine = x;
      if (lane == 31) s_max[par][j * kMW + warp] = x ? (unsigned long long)base + (unsigned long long)x : 0ull;
    }
    st[j] = start; myrun[j] = mine;
    meta[j] = (adm ? 1u : 0u) | ((uint32_t)(person & 3) << 1) | (code << 8) | (ref << 16) | ((uint32_t)rank << 24);
  }
  __syncthreads();
  // exclusive scans over the kMI*kMW (round, warp) cells: five small scans, one warp each
  unsigned long long block_total = 0;
  static_assert(kMI * kMW == 32, "one warp-wide scan covers the cells");
  if (warp < 3) {
    const int v0 = s_cnt[warp][lane];
    int x = v0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    s_cnt[warp][lane] = x - v0;
    block_total = (unsigned long long)__shfl_sync(0xffffffffu, x, 31);
  } else if (warp < 5) {
    const unsigned long long v0 = s_max[warp - 3][lane];
    unsigned long long x = v0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x = y > x ? y : x; }
    unsigned long long ex = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) ex = 0;
    s_max[warp - 3][lane] = ex;
    block_total = __shfl_sync(0xffffffffu, x, 31);
  }
  // five look-back chains, one warp each
  if (warp < 3) {
    const unsigned long long e = lookback<false>(A.st_sum[warp], b, block_total, lane);
    if (lane == 0) { s_excl[warp] = e; s_tot[warp] = block_total; if (b == gridDim.x - 1) A.totals[warp] = e + block_total; }
  } else if (warp < 5) {
    const unsigned long long e = lookback<true>(A.st_max[warp - 2], b, block_total, lane);
    if (lane == 0) s_excl[warp] = e;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kMI; j++) {
    if (!(meta[j] & 1u)) continue;
    const int p = (meta[j] >> 1) & 3;
    const int64_t pos = (int64_t)s_excl[p] + s_cnt[p][j * kMW + warp] + (int)(meta[j] >> 24);
    A.start[p][pos] = st[j];
    A.code[p][pos] = (uint8_t)(meta[j] >> 8);
    if (p == kChild) {
      A.ref[pos] = (uint8_t)(meta[j] >> 16);
      A.src[pos] = tile0 + (int64_t)j * kMB + threadIdx.x;
      // kernel B's window of each parent queue for the block of children this one opens / closes: every parent entry
      // of an earlier tile starts at or before this child, and (up to ties, which B checks) none of a later tile does
      const int64_t jb = pos / kJTile;
      const int wi = (int)(pos % kJTile);
      if (wi == 0) { A.win[jb * 4 + 0] = s_excl[kMom] + 1; A.win[jb * 4 + 2] = s_excl[kDad] + 1; }
      if (wi == kJTile - 1) { A.win[jb * 4 + 1] = s_excl[kMom] + s_tot[kMom] + 1; A.win[jb * 4 + 3] = s_excl[kDad] + s_tot[kDad] + 1; }
    } else {
      unsigned long long rm = myrun[j] ? (unsigned long long)base + (unsigned long long)myrun[j] : 0ull;
      const unsigned long long e1 = s_max[p - 1][j * kMW + warp], e2 = s_excl[2 + p];
      rm = e1 > rm ? e1 : rm;
      rm = e2 > rm ? e2 : rm;
      A.runmax[p][pos] = (int64_t)rm - 1;
    }
  }
}

__global__ void __launch_bounds__(kMB, 5)
mendel_admit_compact_kernel(const dnv_vcall* __restrict__ calls, int64_t n, MendelArrays A) {
  extern __shared__ __align__(128) unsigned char mendel_smem[];
  uint4* s_rec = reinterpret_cast<uint4*>(mendel_smem);   // the tile's records
  __shared__ unsigned int s_bid;
  __shared__ alignas(8) unsigned long long s_bar;
  __shared__ int s_cnt[3][kMI * kMW + 1];                 // per (round, warp) admitted counts -> exclusive offsets
  __shared__ unsigned long long s_max[2][kMI * kMW + 1];  // per (round, warp) max of end+1 -> exclusive running max
  __shared__ unsigned long long s_excl[5];
  __shared__ unsigned long long s_tot[3];
  if (threadIdx.x == 0) {
    const unsigned int t = (unsigned int)atomicAdd(&A.totals[6], 1ull);
    s_bid = t;
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)min((int64_t)kMTile, n - (int64_t)t * kMTile) * (uint32_t)sizeof(dnv_vcall);
    mbar_expect_tx(&s_bar, bytes);
    tma_bulk_g2s(s_rec, calls + (int64_t)t * kMTile, bytes, &s_bar);
  }
  __syncthreads();
  const unsigned int b = s_bid;
  const int n_here = (int)min((int64_t)kMTile, n - (int64_t)b * kMTile);
  mbar_wait(&s_bar, 0);
  // does any admitted-looking entry end 2^32 or more past the tile's first start?  (block uniform)
  const int64_t base = (int64_t)(((unsigned long long)s_rec[0].y << 32) | s_rec[0].x);
  int wide = 0;
#pragma unroll
  for (int j = 0; j < kMI; j++) {
    const int r = j * kMB + threadIdx.x;
    if (r < n_here) {
      const uint4 lo = s_rec[2 * r];
      const int64_t end = (int64_t)(((unsigned long long)lo.w << 32) | lo.z);
      wide |= (unsigned long long)(end + 1 - base) >= (1ull << 32);  // (negative differences land here too)
    }
  }
  if (__syncthreads_or(wide)) mendel_admit_tile<unsigned long long>(s_rec, calls, n, A, b, n_here, s_cnt, s_max, s_excl, s_tot);
  else mendel_admit_tile<uint32_t>(s_rec, calls, n, A, b, n_here, s_cnt, s_max, s_excl, s_tot);
}

// ---- kernel B: join every admitted child SNP against both parents and emit the PositionCalls
// in stream order (one pass; ordered compaction by look-back).  The binary searches of a block
// are confined to the window spanned by its first and last child position (both bounds are
// monotone in the position), so they run out of L1/L2.
// first index in [lo, hi) with a[idx] > key: 32-ary search by one warp (5 dependent probes per 32M entries)
__device__ __forceinline__ int64_t warp_upper_bound64(const int64_t* __restrict__ a, int64_t lo, int64_t hi,
                                                      int64_t key, int lane) {
  while (hi - lo > 32) {
    const int64_t step = (hi - lo + 32) / 33;
    const int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
    const bool le = idx < hi ? (__ldg(a + idx) <= key) : false;
    const int c = __popc(__ballot_sync(0xffffffffu, le));
    const int64_t new_lo = c ? lo + (int64_t)c * step : lo;
    const int64_t new_hi = c < 32 ? min(hi, lo + (int64_t)(c + 1) * step - 1) : hi;
    lo = new_lo; hi = new_hi;
  }
  const bool le = lo + lane < hi ? (__ldg(a + lo + lane) <= key) : false;
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

__device__ __forceinline__ int64_t upper_bound_le(const int64_t* __restrict__ a, int64_t lo, int64_t hi, int64_t key) {
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid; }
  return lo;  // first index with a[idx] > key
}

#ifndef DNV_MENDEL_JCAP
#define DNV_MENDEL_JCAP 1536
#endif
constexpr int kJCap = DNV_MENDEL_JCAP;  // parent entries staged per parent (start + running max as 32-bit offsets, code: 13.5 KB)
static_assert(kJI * (kMB / 32) <= 32, "one warp scans the per-(round, warp) counts of kernel B");

__device__ __forceinline__ int smem_upper_bound_le32(const uint32_t* a, int lo, int hi, uint32_t key) {
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= key) lo = mid + 1; else hi = mid; }
  return lo;
}
// v - base saturated to [0, 2^32 - 1]: every comparison against a key in (0, 2^32 - 1) keeps its outcome
__device__ __forceinline__ uint32_t sat_rel(int64_t v, int64_t base) {
  const int64_t d = v - base;
  return d < 0 ? 0u : d > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)d;
}

__global__ void __launch_bounds__(kMB)
mendel_join_emit_kernel(const Tables* __restrict__ tab, MendelArrays A, dnv_pcall* __restrict__ out, int64_t cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_start = reinterpret_cast<uint32_t*>(smem_raw);  // [2][kJCap] start - base
  uint32_t* s_rmax = s_start + 2 * kJCap;                       // [2][kJCap] running max of end - base
  uint8_t* s_code = reinterpret_cast<uint8_t*>(s_rmax + 2 * kJCap);  // [2][kJCap]
  __shared__ unsigned int s_bid;
  __shared__ long long s_win[2][4];
  __shared__ int s_cnt[kJI * kMW + 32];
  __shared__ unsigned long long s_excl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_child = (int64_t)A.totals[kChild];
  // persistent blocks: tiles of 512 children are handed out by a ticket (the look-back below needs every earlier
  // tile to be running or done), until the children run out — their number is only known on the device
  for (;;) {
  __syncthreads();  // the previous tile's shared state is no longer in use
  if (threadIdx.x == 0) s_bid = (unsigned int)atomicAdd(&A.totals[7], 1ull);
  __syncthreads();
  const unsigned int b = s_bid;
  const int64_t i0 = (int64_t)b * kJTile;
  if (i0 >= n_child) {  // no child SNP left
    if (n_child == 0 && b == 0 && threadIdx.x == 0) A.totals[3] = 0;
    // the last block to leave hands the totals and the error word to the host: the call ends with this kernel, not
    // with a copy behind it (every block's results — totals[3], atomicMax on totals[4] — precede its arrival)
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&A.totals[5], 1ull) == (unsigned long long)gridDim.x - 1ull) {
        __threadfence();
#pragma unroll
        for (int q = 0; q < 5; q++) A.host_totals[q] = *(volatile unsigned long long*)&A.totals[q];
      }
    }
    break;
  }
  const bool last_live = n_child <= i0 + kJTile;
  const int64_t i_last = min(n_child, i0 + kJTile) - 1;
  const int64_t pos0 = A.start[kChild][i0], pos_last = A.start[kChild][i_last];
  // Window of each parent queue that the block's children can touch.  Fast path: kernel A left the tile-granular
  // bracket of "#entries with start <= pos" for the block's first and last child (every entry of an earlier tile of
  // the stream starts at or before the child; up to ties — checked — none of a later tile does), and the first entry
  // whose running max of `end` exceeds pos0 is rarely more than a few entries before that (checked 64 back).  Every
  // thread evaluates the same few loads (broadcasts).  The exact bounds are then found by 32-ary searches inside the
  // bracket (a parent whose checks fail: inside its whole queue).
  long long br_lo[2], br_g[2], br_hi[2];  // bracket per parent: ub in [lo, hi], first covering entry in [g, hi]
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int par = q + 1;
    const int64_t np = (int64_t)A.totals[par];
    const unsigned long long w_lo = A.win[(int64_t)b * 4 + 2 * q], w_hi = A.win[(int64_t)b * 4 + 2 * q + 1];
    const int64_t lo_ub = (int64_t)w_lo - 1;                  // written: the block's first child exists
    const int64_t hi_ub = w_hi ? (int64_t)w_hi - 1 : np;      // the stream's last, partial block of children
    const int64_t g = max(lo_ub - 64, (int64_t)0);
    const bool ok = w_lo != 0 && (g == 0 || __ldg(A.runmax[par] + g - 1) <= pos0) && (hi_ub >= np || __ldg(A.start[par] + hi_ub) > pos_last);
    br_lo[q] = ok ? lo_ub : 0; br_g[q] = ok ? g : 0; br_hi[q] = ok ? hi_ub : np;
  }
  {  // the four exact bounds per parent inside its bracket: one warp each (two probes for a bracket of a few thousand)
    const int q = warp >> 2, kind = warp & 3;
    const int64_t key = (kind & 1) ? pos_last : pos0;
    const int64_t r = warp_upper_bound64(kind < 2 ? A.start[q + 1] : A.runmax[q + 1], kind < 2 ? br_lo[q] : br_g[q], br_hi[q], key, lane);
    if (lane == 0) s_win[q][kind] = r;
  }
  __syncthreads();
  // stage [lo, hi) of start, running max and code of both parents when the windows fit; the positions as 32-bit
  // offsets from pos0 - 2^30, saturated (sat_rel), which needs the block's children within 2^31 of each other
  long long wlo[2], whi[2];
  bool staged = pos_last - pos0 < (1ll << 31);
  const int64_t base = pos0 - (1ll << 30);
#pragma unroll
  for (int q = 0; q < 2; q++) {
    wlo[q] = min(s_win[q][0], s_win[q][2]);
    whi[q] = max(s_win[q][1], s_win[q][3]);
    if (whi[q] - wlo[q] > kJCap) staged = false;
  }
  if (staged) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int m = (int)(whi[q] - wlo[q]);
      for (int k = threadIdx.x; k < m; k += kMB) {
        s_start[q * kJCap + k] = sat_rel(A.start[q + 1][wlo[q] + k], base);
        s_rmax[q * kJCap + k] = sat_rel(A.runmax[q + 1][wlo[q] + k], base);
        s_code[q * kJCap + k] = A.code[q + 1][wlo[q] + k];
      }
    }
  }
  __syncthreads();
  uint32_t packed[kJI];  // bit31 found | rank<<24 | dad<<16 | mom<<8 | child
#pragma unroll
  for (int j = 0; j < kJI; j++) {
    const int64_t i = i0 + (int64_t)j * kMB + threadIdx.x;
    bool found = false;
    uint32_t pk = 0;
    if (i < n_child) {
      const int64_t pos = A.start[kChild][i];
      const int cg = A.code[kChild][i];
      const int rb = A.ref[i];
      int g[3] = {cg, -1, -1};
      bool bad = cg == kCodeBad;
      found = true;
#pragma unroll
      for (int par = kMom; par <= kDad; par++) {  // MOM then DAD, :203-220
        if (!found) break;
        const int q = par - 1;
        int64_t ub, is;
        int c;
        if (staged) {
          const uint32_t key = (uint32_t)(pos - base);  // in [2^30, 2^30 + 2^31)
          ub = wlo[q] + smem_upper_bound_le32(s_start + q * kJCap, (int)(s_win[q][0] - wlo[q]), (int)(s_win[q][1] - wlo[q]), key);  // #entries with start <= pos
          is = wlo[q] + smem_upper_bound_le32(s_rmax + q * kJCap, (int)(s_win[q][2] - wlo[q]), (int)(s_win[q][3] - wlo[q]), key);   // first with running max of end > pos
          if (is >= ub) { found = false; break; }  // no covering entry -> no PositionCall (:211-213)
          c = s_code[q * kJCap + (int)(is - wlo[q])];
        } else {
          ub = upper_bound_le(A.start[par], s_win[q][0], s_win[q][1], pos);
          is = upper_bound_le(A.runmax[par], s_win[q][2], s_win[q][3], pos);
          if (is >= ub) { found = false; break; }
          c = A.code[par][is];
        }
        if (c == kCodeBlock) { if (rb == 0xFF) bad = true; g[par] = geno_of(rb & 3, rb & 3); }  // ref+ref of the CHILD's base (:217-219)
        else { if (c == kCodeBad) bad = true; g[par] = c; }
      }
      // the reference evaluates the child genotype first, so a bad child base throws even without parents
      if (bad) { atomicMax(&A.totals[4], ((unsigned long long)(A.src[i] + 1) << 2) | kErrBadBase); found = false; }
      if (found) pk = (uint32_t)g[kChild] | ((uint32_t)g[kMom] << 8) | ((uint32_t)g[kDad] << 16);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, found);
    if (lane == 0) s_cnt[j * kMW + warp] = __popc(m);
    packed[j] = pk | ((uint32_t)__popc(m & ((1u << lane) - 1)) << 24) | (found ? 0x80000000u : 0u);
  }
  __syncthreads();
  if (warp == 0) {
    const int v0 = lane < kJI * kMW ? s_cnt[lane] : 0;
    int x = v0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    s_cnt[lane] = x - v0;
    const int total = __shfl_sync(0xffffffffu, x, 31);
    const unsigned long long e = lookback<false>(A.st_found, b, (unsigned long long)total, lane);
    if (lane == 0) { s_excl = e; if (last_live) A.totals[3] = e + (unsigned long long)total; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kJI; j++) {
    if (!(packed[j] & 0x80000000u)) continue;
    const int64_t i = i0 + (int64_t)j * kMB + threadIdx.x;
    const int64_t o_pos = (int64_t)s_excl + s_cnt[j * kMW + warp] + (int)((packed[j] >> 24) & 0x7F);
    if (o_pos < cap) {
      const int c = packed[j] & 0xFF, mo = (packed[j] >> 8) & 0xFF, d = (packed[j] >> 16) & 0xFF;
      const int idx = d * 100 + mo * 10 + c;
      dnv_pcall o;
      o.position = A.start[kChild][i];
      o.child_index = A.src[i];
      o.child = (uint8_t)c; o.mom = (uint8_t)mo; o.dad = (uint8_t)d;
      o.is_denovo = (tab->isden[idx >> 5] >> (idx & 31)) & 1u;  // PositionCall.isDenovo, :352-355
      o.pad_[0] = o.pad_[1] = o.pad_[2] = o.pad_[3] = 0;
      out[o_pos] = o;
    }
  }
  }  // next tile
}

}  // namespace dnv

using namespace dnv;

extern "C" {

int dnv_mendel_filter_device(dnv_ctx* ctx, const dnv_vcall* d_calls, int64_t n, dnv_pcall* d_out, int64_t cap,
                             int64_t* n_out) {
  if (!ctx || !n_out || n < 0 || cap < 0 || (n > 0 && !d_calls) || (cap > 0 && !d_out) || n >= (1ll << 36)) {
    set_error("dnv_mendel_filter_device: invalid argument");
    return DNV_E_INVALID;
  }
  *n_out = 0;
  if (n == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // carve the scratch arena (every person's queue is sized for n entries)
  const int n_blocks = (int)((n + kMTile - 1) / kMTile);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  size_t o_start[3], o_rmax[3], o_code[3];
  for (int p = 0; p < 3; p++) { o_start[p] = take((size_t)n * 8); o_rmax[p] = take((size_t)n * 8); o_code[p] = take((size_t)n); }
  const size_t o_ref = take((size_t)n), o_src = take((size_t)n * 8);
  constexpr int kFiner = kJTile < kMTile ? kJTile : kMTile;
  const size_t status_bytes = (size_t)((n + kFiner - 1) / kFiner + 1) * 8;  // sized for the finer of the two grids
  const size_t win_bytes = (size_t)((n + kJTile - 1) / kJTile + 1) * 4 * 8;  // kernel B's window records
  const size_t o_status = take(7 * status_bytes + 64 + win_bytes);  // 3 sums, 3 maxima (child unused), found, the totals, the windows
  int rc = ensure_scratch(ctx, off);
  if (rc) return rc;
  char* base = (char*)ctx->d_scratch;
  MendelArrays A;
  for (int p = 0; p < 3; p++) {
    A.start[p] = (int64_t*)(base + o_start[p]); A.runmax[p] = (int64_t*)(base + o_rmax[p]);
    A.code[p] = (uint8_t*)(base + o_code[p]);
    A.st_sum[p] = (unsigned long long*)(base + o_status + (size_t)p * status_bytes);
    A.st_max[p] = (unsigned long long*)(base + o_status + (size_t)(3 + p) * status_bytes);
  }
  A.st_found = (unsigned long long*)(base + o_status + 6 * status_bytes);
  A.totals = (unsigned long long*)(base + o_status + 7 * status_bytes);
  A.win = (unsigned long long*)(base + o_status + 7 * status_bytes + 64);
  A.ref = (uint8_t*)(base + o_ref); A.src = (int64_t*)(base + o_src);
  A.host_totals = ctx->h_acc + 8;  // pinned status words of the context's synchronous calls (shared with the pileup)

  DNV_CUDA(cudaMemsetAsync(base + o_status, 0, 7 * status_bytes + 64 + win_bytes, st));
  if (reinterpret_cast<uintptr_t>(d_calls) & 15) {
    set_error("dnv_mendel_filter_device: the record array must be 16-byte aligned");
    return DNV_E_INVALID;
  }
  const size_t a_smem = (size_t)kMTile * sizeof(dnv_vcall);
  DNV_CUDA(cudaFuncSetAttribute(mendel_admit_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a_smem));
  mendel_admit_compact_kernel<<<n_blocks, kMB, a_smem, st>>>(d_calls, n, A);
  const int j_blocks = (int)((n + kJTile - 1) / kJTile);  // children <= records; empty blocks retire at once
  const size_t j_smem = (size_t)2 * kJCap * (2 * sizeof(uint32_t) + 1);
  DNV_CUDA(cudaFuncSetAttribute(mendel_join_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)j_smem));
  int j_per_sm = 1;
  DNV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&j_per_sm, mendel_join_emit_kernel, kMB, j_smem));
  const int j_grid = std::max(1, std::min(j_blocks, ctx->sms * std::max(j_per_sm, 1)));
  mendel_join_emit_kernel<<<j_grid, kMB, j_smem, st>>>(ctx->d_tab, A, d_out, cap);
  DNV_CUDA(cudaGetLastError());
  DNV_CUDA(cudaStreamSynchronize(st));
  unsigned long long h[5];
  for (int q = 0; q < 5; q++) h[q] = A.host_totals[q];  // written by kernel B's last block
  ctx->stats = dnv_stats();
  ctx->stats.n_sites = (int64_t)h[0];
  ctx->stats.n_calls = (int64_t)h[3];
  ctx->stats.n_launches = 2;
  if (h[4] != 0) {
    const int kind = (h[4] & 3) == kErrBadBase ? DNV_E_BADBASE : DNV_E_UNSORTED;
    h[4] >>= 2;
    set_error(kind == DNV_E_BADBASE
                  ? "mendel filter: base outside ACGT at record " + std::to_string(h[4] - 1) +
                        " (the reference throws IllegalArgumentException, M/VariantsBuffer.java:235-236)"
                  : "mendel filter: stream not sorted by start, or admitted record with end <= start, at record " +
                        std::to_string(h[4] - 1));
    return kind == DNV_E_BADBASE ? DNV_E_BADBASE : DNV_E_UNSORTED;
  }
  *n_out = (int64_t)h[3];
  if ((int64_t)h[3] > cap) {
    set_error("mendel filter: output capacity too small: need " + std::to_string(h[3]));
    return DNV_E_CAPACITY;
  }
  return DNV_OK;
}

int dnv_mendel_filter(dnv_ctx* ctx, const dnv_vcall* calls, int64_t n, dnv_pcall* out, int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out || n < 0 || cap < 0 || (n > 0 && !calls) || (cap > 0 && !out)) {
    set_error("dnv_mendel_filter: invalid argument");
    return DNV_E_INVALID;
  }
  *n_out = 0;
  if (n == 0) return DNV_OK;
  DNV_CUDA(cudaSetDevice(ctx->device));
  dnv_vcall* d_calls = nullptr;
  dnv_pcall* d_out = nullptr;
  const int64_t dev_cap = std::min(cap, n);
  if (cudaMalloc(&d_calls, (size_t)n * sizeof(dnv_vcall)) != cudaSuccess ||
      cudaMalloc(&d_out, (size_t)std::max<int64_t>(dev_cap, 1) * sizeof(dnv_pcall)) != cudaSuccess) {
    cudaFree(d_calls); cudaFree(d_out);
    set_error("dnv_mendel_filter: device allocation failed");
    return DNV_E_NOMEM;
  }
  int rc = DNV_OK;
  if (cudaMemcpyAsync(d_calls, calls, (size_t)n * sizeof(dnv_vcall), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
    set_error("dnv_mendel_filter: H2D copy failed"); rc = DNV_E_CUDA;
  }
  if (rc == DNV_OK) rc = dnv_mendel_filter_device(ctx, d_calls, n, d_out, dev_cap, n_out);
  const int64_t got = std::min<int64_t>(*n_out, dev_cap);
  if ((rc == DNV_OK || rc == DNV_E_CAPACITY) && got > 0) {
    if (cudaMemcpyAsync(out, d_out, (size_t)got * sizeof(dnv_pcall), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error("dnv_mendel_filter: D2H copy failed"); rc = DNV_E_CUDA;
    }
  }
  cudaFree(d_calls); cudaFree(d_out);
  return rc;
}

}  // extern "C"
