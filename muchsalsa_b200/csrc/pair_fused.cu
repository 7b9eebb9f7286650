// pair_fused.cu — candidate-pair generation AND keyed merge without ever sorting in HBM (K6 + K7, the product variant).
//
// Replaces MatchMap::calculateEdges / processScaffold (MatchMap.cpp:161-224) + Graph::addEdge (Graph.cpp:212-230,
// 291-311) + MatchMap::addEdgeMatch (MatchMap.cpp:109-134).
//
// The reference walks every unitig group and throws each accepted pair at two global hash maps.  The sort-based
// variant (pair_sort.cu) writes every candidate to HBM in unitig order and moves it through 4-6 radix passes: ~250 B
// of HBM traffic per edge match for 25 B of result.  Here the work is turned around so that every output row is
// produced directly at its final position:
//
//   * The reduction key is (first, second), first < second (read ids; MatchMap.cpp:204-213).  Call the edge matches
//     whose `first` is read r "row r".  Row r is the union, over the matches a of r (its "sources", one per unitig),
//     of the members b of a's unitig group with read(b) > r and unitig overlap > 100 (MatchMap.cpp:189-192).
//   * A pre-pass re-orders every unitig group by read id (rank by counting inside the group).  Then the partners of a
//     source are exactly the SUFFIX of its group behind it, already sorted by `second`, and sum(suffix lengths) is
//     exactly the number of candidate pairs sum k(k-1)/2.
//   * Rows are cut into tiles of <= 4096 candidates.  Light rows are packed, several consecutive rows per tile.  A row
//     heavier than a tile (a read that crosses a high-multiplicity repeat) is split into sub-tiles by VALUE RANGES of
//     `second`: the boundaries are found by bisection on the merged rank over the row's sorted suffixes, so every
//     sub-tile is bounded, no edge straddles two sub-tiles and the sub-tiles of a row are in output order.
//   * pair_tiles_merge_kernel, one CTA per tile at a time (one resident wave of CTAs, tiles handed out by a ticket
//     counter), no dependency between tiles: stage the sources, enumerate the
//     candidates into shared memory, merge-sort them there by (row, second) starting from the slices, which are
//     sorted runs already — stable, and candidates are generated in unitig order, so rows of one edge come out by
//     unitig — find the run heads (= edges) and write the tile's surviving candidates as compact 8-byte records
//     (partner row, source, head flag, heads before) plus two counts.
//   * two single-pass scans turn the per-tile counts into the global offset of every tile's first edge match and edge.
//   * pair_emit_rows_kernel, embarrassingly parallel: each record gathers its two matches (one 32-byte sector each),
//     evaluates the reference's score expression in IEEE double and writes the edge match — and, for a run head, the
//     edge and its CSR offset — exactly once, coalesced, in (first, second, unitig) order.
//   (An earlier version did all of it in one persistent kernel with a decoupled look-back for the offsets; ncu showed
//    27 % of its stall samples at the barrier behind the look-back, because a tile's counts exist only after its sort.
//    The 8-byte records cost +0.66 GB of traffic at yeast scale and remove every inter-CTA wait.)
//
// Algorithmic HBM traffic = every match record once + 25 B per edge match + 16 B per edge; actual traffic adds one
// write and one read of 8 B per edge match.  The k^2 work never leaves the SM.  The same kernels serve multi-GPU runs:
// a rank processes a contiguous range of tiles, so its output is a contiguous slice of the global
// (first, second)-sorted result and no candidate ever crosses NVLink.
#include <cstdlib>
#include <string>

#include "../../include/msgpu.h"
#include "internal.h"
#include "scan.cuh"

namespace msgpu {

constexpr int kThOverlapF = 100; // MatchMap.cpp:41

constexpr int      kFusedBlock = 512;
constexpr int      kFusedWarps = kFusedBlock / 32;
#ifndef MSGPU_IDX_BITS
#define MSGPU_IDX_BITS 12
#endif
constexpr int      kIdxBits    = MSGPU_IDX_BITS;
constexpr uint32_t kTileCap    = 1u << kIdxBits; // candidates per tile
constexpr uint32_t kSrcCap     = 512;            // sources per tile (the product instantiation of the tile kernel)
constexpr int      kSrcBitsWide = 12;            // ... of the instantiation for rows with more sources: up to 4095 — an
constexpr uint32_t kSrcCapWide = 1u << kSrcBitsWide; // edge cannot hold more edge matches than a tile holds candidates
constexpr uint32_t kHalfCap    = kTileCap / 2;
constexpr uint32_t kSrcRatio   = kHalfCap / (kSrcCap / 2); // a source weighs at least this much when tiles are packed
constexpr int      kItems      = kTileCap / kFusedBlock;   // sort items per thread
#ifndef MSGPU_MERGE_CTAS
#define MSGPU_MERGE_CTAS (MSGPU_IDX_BITS <= 12 ? 4 : 2)
#endif
#ifndef MSGPU_TICKETS
#define MSGPU_TICKETS 1 // 0: tiles strided over the grid (the A/B of the ticket counters)
#endif
constexpr int      kMergeCtas  = MSGPU_MERGE_CTAS;          // resident CTAs of the merge kernel per SM (narrow words)

// ---------------------------------------------------------------------------------------------------------------------
// pre-pass 1: order every group by read id, build the pair rows, and accumulate per read its number of sources and
// its number of candidates
// ---------------------------------------------------------------------------------------------------------------------
struct RankArgs {
  MatchSoA            m;
  const uint32_t     *grp_off;
  PairRow            *rows;
  uint32_t           *p_read;
  int32_t            *p_ir0, *p_ir1;
  uint32_t           *deg;
  unsigned long long *wgt;
};

// match i is the q-th of its group in read order: write its pair row and account its partners to its read
__device__ __forceinline__ void place_row(const RankArgs &a, uint64_t i, uint32_t u, uint32_t q, uint32_t g1) {
  const MatchSoA &m = a.m;
  PairRow         row;
  row.ir0    = m.ir0[i];
  row.ir1    = m.ir1[i];
  row.read   = m.read[i];
  row.score  = m.score[i];
  row.line   = m.line[i];
  row.flags  = m.flags[i];
  row.unitig = u;
  row.orig   = uint32_t(i);
  a.rows[q]   = row;
  a.p_read[q] = row.read;
  a.p_ir0[q]  = row.ir0;
  a.p_ir1[q]  = row.ir1;
  const uint32_t len = g1 - q - 1; // partners = the members behind q
  if (len) {
    atomicAdd(&a.deg[row.read], 1u);
    atomicAdd(&a.wgt[row.read], (unsigned long long)len);
  }
}

// Bitonic sort of n = 2^x keys (padded with ~0) in shared memory by `nthreads` cooperating threads (a warp or a CTA)
template <bool kCta> __device__ __forceinline__ void bitonic_sort_shared(unsigned long long *keys, uint32_t n, uint32_t tid,
                                                                        uint32_t nthreads) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = tid; t < (n >> 1); t += nthreads) {
        const uint32_t i  = ((t & ~(j - 1)) << 1) | (t & (j - 1)); // the pair (i, i + j), bit j of i clear
        const bool     up = (i & k) == 0;
        const unsigned long long x = keys[i], y = keys[i + j];
        if ((x > y) == up) {
          keys[i]     = y;
          keys[i + j] = x;
        }
      }
      if (kCta) __syncthreads(); else __syncwarp();
    }
  }
}

constexpr uint32_t kSortWarpMax = 128;  // groups up to this size: one warp sorts them
constexpr uint32_t kRankChunk   = 256;  // longer groups: one CTA ranks this many members of a group by counting ...
constexpr uint32_t kRankTile    = 4096; // ... against all of the group's read ids, staged this many at a time

// One warp per group: the keys (read id << 32 | row of the match) are sorted in shared memory, then lane by lane the
// matches are placed at their position in read order.  Reads are unique inside a group (MatchMap.cpp:64-73 keeps one
// match per (read, unitig)), so the order is total.  Groups too long for a warp become work items of the CTA kernel.
__global__ void __launch_bounds__(256) rank_rows_kernel(RankArgs a, uint32_t n_groups, uint2 *__restrict__ big_items,
                                                        unsigned int *__restrict__ n_big) {
  pdl_sync();
  __shared__ unsigned long long sm[256 / 32][kSortWarpMax];
  const MatchSoA &m = a.m;
  const unsigned  l = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t  nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  unsigned long long *keys = sm[w];
  for (uint64_t u = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; u < n_groups; u += nw) {
    const uint32_t g0 = a.grp_off[u], g1 = a.grp_off[u + 1], k = g1 - g0;
    if (k == 0) continue;
    if (k > kSortWarpMax) { // one work item per kRankChunk members for rank_rows_cta_kernel
      const uint32_t nc = (k + kRankChunk - 1) / kRankChunk;
      uint32_t       base = 0;
      if (l == 0) base = atomicAdd(n_big, nc);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t c = l; c < nc; c += 32) big_items[base + c] = make_uint2(uint32_t(u), c);
      continue;
    }
    uint32_t n = 2;
    while (n < k) n <<= 1;
    for (uint32_t i = l; i < n; i += 32)
      keys[i] = i < k ? ((unsigned long long)m.read[g0 + i] << 32 | (g0 + i)) : ~0ull;
    __syncwarp();
    bitonic_sort_shared<false>(keys, n, l, 32);
    for (uint32_t j = l; j < k; j += 32) place_row(a, uint32_t(keys[j]), uint32_t(u), g0 + j, g1);
    __syncwarp();
  }
}

// Longer groups, ranked by counting: a work item is (group, chunk of kRankChunk members).  The CTA stages the group's
// read ids in shared memory, kRankTile at a time, and every thread counts for ITS member the smaller ids among them
// (k compares per member served by shared-memory broadcasts, no barrier per step).  The k^2 compares of a long group
// are spread over k / 256 CTAs, so what the kernel costs is the latency of one item, whatever the longest group is
// (round 2 until here: one 1024-thread CTA per group up to 4096 — counting up to 1024, a sorting network beyond — and
// one warp per member over global memory for longer groups: 41 + 7 us at yeast scale, set by the longest groups).
__global__ void __launch_bounds__(kRankChunk) rank_rows_cta_kernel(RankArgs a, const uint2 *__restrict__ big_items,
                                                                  const unsigned int *__restrict__ n_big) {
  pdl_sync();
  __shared__ uint32_t ids[kRankTile];
  const MatchSoA &m = a.m;
  const uint32_t  n = *n_big;
  for (uint32_t x = blockIdx.x; x < n; x += gridDim.x) {
    const uint2    item = big_items[x];
    const uint32_t u = item.x, g0 = a.grp_off[u], g1 = a.grp_off[u + 1], k = g1 - g0;
    const uint32_t j   = item.y * kRankChunk + threadIdx.x; // this thread's member
    const bool     has = j < k;
    const uint32_t my  = has ? m.read[g0 + j] : 0u;
    uint32_t       r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (uint32_t t0 = 0; t0 < k; t0 += kRankTile) {
      const uint32_t tn = k - t0 < kRankTile ? k - t0 : kRankTile;
      __syncthreads(); // the previous tile (or item) is consumed
      for (uint32_t i = threadIdx.x; i < tn; i += kRankChunk) ids[i] = m.read[g0 + t0 + i];
      __syncthreads();
      if (has) {
        uint32_t i = 0;
        for (; i + 4 <= tn; i += 4) {
          r0 += ids[i] < my;
          r1 += ids[i + 1] < my;
          r2 += ids[i + 2] < my;
          r3 += ids[i + 3] < my;
        }
        for (; i < tn; ++i) r0 += ids[i] < my;
      }
    }
    if (has) place_row(a, g0 + j, u, g0 + r0 + r1 + r2 + r3, g1);
  }
}

// pre-pass 2: per read the list of its sources (positions in the pair rows), in unitig order
// (multi-GPU: a device needs the lists of the rows of its own tiles only — `rows_of` names them)
struct RowRange {
  const uint32_t *tile_row; // null: every row
  uint32_t        t_first, t_last, n_reads; // rows tile_row[t_first] .. tile_row[t_last] (inclusive)
  __device__ uint32_t lo() const { return tile_row ? tile_row[t_first] : 0u; }
  __device__ uint32_t hi() const { // exclusive
    if (!tile_row) return n_reads;
    const uint32_t r = tile_row[t_last] + 1u;
    return r < n_reads ? r : n_reads;
  }
};

__global__ void __launch_bounds__(256)
    scatter_sources_kernel(const uint32_t *__restrict__ p_read, const PairRow *__restrict__ rows,
                           const uint32_t *__restrict__ grp_off, uint64_t n, const uint32_t *__restrict__ rd_off,
                           uint32_t *__restrict__ cursor, uint32_t *__restrict__ rd_tmp, RowRange rows_of) {
  pdl_sync();
  const uint32_t r_lo = rows_of.lo(), r_hi = rows_of.hi();
  for (uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < n; q += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t r = p_read[q];
    if (r < r_lo || r >= r_hi) continue;
    if (grp_off[rows[q].unitig + 1] - uint32_t(q) - 1 > 0) rd_tmp[rd_off[r] + atomicAdd(&cursor[r], 1u)] = uint32_t(q);
  }
}

// the atomics above leave a read's list in arbitrary order: rank-sort it (lists are short) so that it is in row order
// = unitig order, which makes the candidate stream of a tile deterministic
__global__ void __launch_bounds__(256)
    order_sources_kernel(const uint32_t *__restrict__ p_read, const uint32_t *__restrict__ rd_off,
                         const uint32_t *__restrict__ rd_tmp, uint64_t n_src, uint32_t *__restrict__ rd_src,
                         RowRange rows_of) {
  pdl_sync();
  const uint64_t x_lo = rd_off[rows_of.lo()], x_hi = rd_off[rows_of.hi()]; // (rd_off has n_reads + 1 entries)
  for (uint64_t x = x_lo + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < x_hi && x < n_src;
       x += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t me = rd_tmp[x];
    uint32_t r  = p_read[me];
    uint32_t lo = rd_off[r], hi = rd_off[r + 1], rank = 0;
    for (uint32_t p = lo; p < hi; ++p) rank += rd_tmp[p] < me;
    rd_src[lo + rank] = me;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// tile planning
// ---------------------------------------------------------------------------------------------------------------------
struct RowPlan {
  const uint32_t           *deg;
  const unsigned long long *wgt;
  __device__ uint64_t weight(uint64_t r) const { // what a light row costs a tile
    uint64_t w = wgt[r], d = uint64_t(deg[r]) * kSrcRatio;
    return w > d ? w : d;
  }
  __device__ bool     heavy(uint64_t r) const { return weight(r) > kHalfCap; }
  __device__ uint32_t sub_cap(uint64_t r) const { return kTileCap - deg[r]; } // see split_heavy_kernel
  __device__ uint32_t n_sub(uint64_t r) const {
    uint64_t t = sub_cap(r);
    uint64_t n = (wgt[r] + t - 1) / t;
    return uint32_t(n ? n : 1);
  }
};

struct ClippedWeight {
  RowPlan p;
  __device__ uint64_t operator()(uint64_t r) const { return p.heavy(r) ? uint64_t(kHalfCap) : p.weight(r); }
};

// tiles that start at row r: 1 where a packed tile begins, n_sub for a heavy row
struct TileCount {
  RowPlan         p;
  const uint64_t *prefix;
  unsigned int   *n_wide; // out: sub-tiles of rows with more sources than the product tile kernel stages (the scan
                          // evaluates every row exactly once, so the count rides along)
  __device__ bool start(uint64_t r) const {
    if (r == 0) return true;
    if (prefix[r] / kHalfCap != prefix[r - 1] / kHalfCap) return true;
    return p.heavy(r - 1);
  }
  __device__ uint32_t operator()(uint64_t r) const {
    if (!p.heavy(r)) return start(r) ? 1u : 0u;
    const uint32_t ns = p.n_sub(r);
    if (p.deg[r] > kSrcCap && p.deg[r] < kSrcCapWide) atomicAdd(n_wide, ns);
    return ns;
  }
};

// slice table entries of a heavy row: (n_sub + 1) boundaries x deg sources
struct SliceCount {
  RowPlan p;
  __device__ uint64_t operator()(uint64_t r) const {
    return p.heavy(r) ? uint64_t(p.n_sub(r) + 1) * p.deg[r] : 0ull;
  }
};

struct WeightOf {
  const unsigned long long *wgt;
  __device__ uint64_t operator()(uint64_t r) const { return wgt[r]; }
};

__global__ void __launch_bounds__(256)
    write_tiles_kernel(TileCount tc, const uint32_t *__restrict__ tile_idx, uint64_t n_reads, uint32_t n_tiles,
                       uint32_t *__restrict__ tile_row, uint32_t *__restrict__ tile_kind,
                       uint32_t *__restrict__ too_many_sources, uint32_t *__restrict__ wide_list,
                       unsigned int *__restrict__ wide_cursor) {
  pdl_sync();
  for (uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n_reads;
       r += uint64_t(gridDim.x) * blockDim.x) {
    if (tc.p.deg[r] >= kSrcCapWide) *too_many_sources = 1; // an edge of such a row could outgrow a tile
    uint32_t t = tile_idx[r];
    if (tc.p.heavy(r)) {
      uint32_t ns = tc.p.n_sub(r);
      for (uint32_t j = 0; j < ns; ++j) {
        tile_row[t + j]  = uint32_t(r);
        tile_kind[t + j] = 1 + j;
      }
      if (tc.p.deg[r] > kSrcCap && tc.p.deg[r] < kSrcCapWide) { // its sub-tiles go to the many-sources instantiation
        const uint32_t w = atomicAdd(wide_cursor, ns);
        for (uint32_t j = 0; j < ns; ++j) wide_list[w + j] = t + j;
      }
    } else if (tc.start(r)) {
      tile_row[t]  = uint32_t(r);
      tile_kind[t] = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tile_row[n_tiles]  = uint32_t(n_reads);
    tile_kind[n_tiles] = 0;
  }
}

// first position in [lo, hi) of the ascending array whose value is >= s
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                    uint32_t s) {
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < s) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// the bisection bracket of a heavy row from the lanes' smallest first / largest last partner ids (all lanes take part)
__device__ __forceinline__ void warp_value_range(uint32_t vmin, uint32_t vmax, uint32_t n_reads, uint32_t *lo, uint32_t *hi) {
  vmin = __reduce_min_sync(0xffffffffu, vmin);
  vmax = __reduce_max_sync(0xffffffffu, vmax);
  if (vmin <= vmax) { // (a row without candidates keeps the full bracket; it has no sub-tile beyond the first anyway)
    *lo = vmin;
    *hi = vmax + 1u < n_reads ? vmax + 1u : n_reads;
  }
}
__device__ __forceinline__ void row_value_range(bool has, const uint32_t *__restrict__ p_read, uint32_t b, uint32_t e,
                                                uint32_t n_reads, uint32_t *lo, uint32_t *hi) {
  uint32_t vmin = 0xffffffffu, vmax = 0;
  if (has) {
    vmin = p_read[b];
    vmax = p_read[e - 1];
  }
  warp_value_range(vmin, vmax, n_reads, lo, hi);
}

// Heavy rows: one warp per sub-tile finds the value s_j that starts the sub-tile, the largest s with
//   C(s) := #{candidates of the row with second < s}  <=  j * T,   T = kTileCap - deg,
// by bisection over s (C is monotone; every source suffix is sorted by read id, so C is a sum of lower bounds).  One
// value of `second` occurs at most once per source, so C jumps by <= deg and sub-tile j = [s_j, s_j+1) holds fewer
// than T + deg = kTileCap candidates.  Written out: for every source the row position where each sub-tile begins.
// Sub-tile 0 starts where the slices start (no search: most heavy rows at yeast scale are ONE sub-tile), and the
// bracket of the others is the row's own value range [smallest first partner, largest last partner + 1], not all
// read ids: C is 0 at its lower end and the row's weight at its upper end.
__global__ void __launch_bounds__(256)
    split_heavy_kernel(const uint32_t *__restrict__ tile_row, const uint32_t *__restrict__ tile_kind, uint32_t t_begin,
                       uint32_t n_tiles, RowPlan plan, const uint64_t *__restrict__ slice_off, const uint32_t *__restrict__ rd_off,
                       const uint32_t *__restrict__ rd_src, const PairRow *__restrict__ rows,
                       const uint32_t *__restrict__ grp_off, const uint32_t *__restrict__ p_read, uint32_t n_reads,
                       uint32_t *__restrict__ slices) {
  pdl_sync();
  const unsigned l  = threadIdx.x & 31;
  uint64_t       wi = t_begin + ((uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (; wi < n_tiles; wi += nw) { // tiles [t_begin, n_tiles)
    const uint32_t kind = tile_kind[wi];
    if (kind == 0) continue;
    const uint32_t r = tile_row[wi], j = kind - 1, d = plan.deg[r], ns = plan.n_sub(r);
    if (d >= kSrcCapWide) continue; // reported by write_tiles_kernel; the build fails with MSGPU_ERR_LIMIT
    const uint32_t q0 = rd_off[r];
    uint32_t      *out = slices + slice_off[r];
    const uint64_t target = uint64_t(j) * plan.sub_cap(r);
    if (d <= 32) {
      // one source per lane.  Every probe of the bisection narrows the lane's window [wlo, whi) of positions where
      // later lower bounds can fall, so the searches get shorter as the bracket [lo, hi] closes.
      const bool has = l < d;
      uint32_t   b = 0, e = 0;
      if (has) {
        const uint32_t q = rd_src[q0 + l];
        b                = q + 1;
        e                = grp_off[rows[q].unitig + 1];
      }
      uint32_t wlo = b, whi = e, lo = 0, hi = n_reads; // invariant: C(lo) <= target, wlo = lower bound of lo
      if (j == 0) {
        hi = 0; // nothing to search: the sub-tile starts at the slices' starts
      } else {
        row_value_range(has && e > b, p_read, b, e, n_reads, &lo, &hi);
      }
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        const uint32_t pos = has ? lower_bound_u32(p_read, wlo, whi, mid) : 0u;
        const uint64_t c   = warp_reduce_sum(uint64_t(pos - b));
        if (c <= target) {
          lo  = mid;
          wlo = pos;
        } else {
          hi  = mid - 1;
          whi = pos;
        }
      }
      if (has) {
        out[uint64_t(j) * d + l] = j == 0 ? b : wlo;
        if (j == ns - 1) out[uint64_t(ns) * d + l] = e;
      }
      continue;
    }
    if (d <= 128) {
      // up to four sources per lane, each with its own shrinking window (the arrays are unrolled into registers)
      uint32_t b[4], e[4], wlo[4], whi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a = l + 32u * i;
        b[i] = e[i] = 0;
        if (a < d) {
          const uint32_t q = rd_src[q0 + a];
          b[i]             = q + 1;
          e[i]             = grp_off[rows[q].unitig + 1];
        }
        wlo[i] = b[i];
        whi[i] = e[i];
      }
      uint32_t lo = 0, hi = n_reads; // invariant: C(lo) <= target, wlo = lower bounds of lo
      if (j == 0) {
        hi = 0;
      } else {
        uint32_t vmin = 0xffffffffu, vmax = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (e[i] > b[i]) {
            const uint32_t x0 = p_read[b[i]], x1 = p_read[e[i] - 1];
            vmin = x0 < vmin ? x0 : vmin;
            vmax = x1 > vmax ? x1 : vmax;
          }
        warp_value_range(vmin, vmax, n_reads, &lo, &hi);
      }
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        uint32_t       pos[4];
        uint64_t       c = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          pos[i] = (l + 32u * i < d) ? lower_bound_u32(p_read, wlo[i], whi[i], mid) : 0u;
          c += pos[i] - b[i];
        }
        c = warp_reduce_sum(c);
        if (c <= target) {
          lo = mid;
#pragma unroll
          for (int i = 0; i < 4; ++i) wlo[i] = pos[i];
        } else {
          hi = mid - 1;
#pragma unroll
          for (int i = 0; i < 4; ++i) whi[i] = pos[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a = l + 32u * i;
        if (a < d) {
          out[uint64_t(j) * d + a] = j == 0 ? b[i] : wlo[i];
          if (j == ns - 1) out[uint64_t(ns) * d + a] = e[i];
        }
      }
      continue;
    }
    // bisection: largest s in [0, n_reads] with C(s) <= target
    uint32_t lo = 0, hi = n_reads; // invariant: C(lo) <= target
    if (j == 0) {
      hi = 0; // (the loop below then writes b for every source: lower_bound is not consulted for j == 0)
    } else {
      uint32_t vmin = 0xffffffffu, vmax = 0;
      for (uint32_t a = l; a < d; a += 32) {
        const uint32_t q = rd_src[q0 + a], b = q + 1, e = grp_off[rows[q].unitig + 1];
        if (e > b) {
          const uint32_t x0 = p_read[b], x1 = p_read[e - 1];
          vmin = x0 < vmin ? x0 : vmin;
          vmax = x1 > vmax ? x1 : vmax;
        }
      }
      warp_value_range(vmin, vmax, n_reads, &lo, &hi);
    }
    while (lo < hi) {
      uint32_t mid = lo + ((hi - lo + 1) >> 1);
      uint64_t c   = 0;
      for (uint32_t a = l; a < d; a += 32) {
        uint32_t q = rd_src[q0 + a], b = q + 1, e = grp_off[rows[q].unitig + 1];
        c += lower_bound_u32(p_read, b, e, mid) - b;
      }
      c = warp_reduce_sum(c);
      if (c <= target) lo = mid; else hi = mid - 1;
    }
    for (uint32_t a = l; a < d; a += 32) {
      uint32_t q = rd_src[q0 + a], b = q + 1, e = grp_off[rows[q].unitig + 1];
      out[uint64_t(j) * d + a] = j == 0 ? b : lower_bound_u32(p_read, b, e, lo);
      if (j == ns - 1) out[uint64_t(ns) * d + a] = e;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K6/K7a: one CTA per tile — enumerate, sort in shared memory, find run heads, write the tile's compact records
// ---------------------------------------------------------------------------------------------------------------------
constexpr int      kSrcBits  = 9;       // kSrcCap == 1 << kSrcBits
static_assert(kSrcCap == (1u << kSrcBits), "staging capacity");
constexpr int      kRecSrcBits = kSrcBitsWide; // a record carries its source in this many bits, whichever kernel wrote it

struct TileArgs {
  const PairRow  *rows;
  const uint32_t *p_read;
  const int32_t  *p_ir0, *p_ir1;
  const uint32_t *grp_off;
  const uint32_t *rd_off, *rd_src;
  const uint32_t *tile_row, *tile_kind;
  const uint64_t *slice_off;
  const uint32_t *slices;
  uint32_t        tile_lo, tile_hi; // this launch processes tiles [tile_lo, tile_hi)
  uint2          *records;          // per tile kTileCap slots: (partner row, source | head << 12 | heads before << 13)
  const uint32_t *wide_list;        // the instantiation for many sources: the tiles it takes (those inside the range)
  uint32_t        n_wide;
  uint2          *tile_cnt;         // per tile: (surviving candidates, run heads)
  unsigned int   *error;
  unsigned int   *ticket;           // zero at launch: hands out the tiles behind the first gridDim.x
  uint32_t        wide; // 0: every sort word fits 32 bits for any tile, so the launch carries the small buffers only
  uint32_t        no_overlap_test; // the scrubber's pairing: every pair of a group is kept (scrubber_bfs.py:109-113)
};

// ---------------------------------------------------------------------------------------------------------------------
// K6/K7a: the tile kernel — candidates ordered by a merge sort that starts from the slices
// ---------------------------------------------------------------------------------------------------------------------
// The candidates of a source are generated in slice order and a slice is sorted by read id, so every slice is a sorted
// run before any sorting happens, and the slices of one row are adjacent.  A word is (second - smin) << 12 | slot:
// words are unique, and an earlier slice holds smaller slots, so comparing whole words makes every merge stable
// (candidates of one read pair stay in slice order = unitig order).  Level v merges, inside every row, runs of 2^v
// slices pairwise.  Each thread produces ceil(n / 512) consecutive outputs of the level: one merge-path search on
// the diagonal of its first output, then a sequential two-way merge — a handful of instructions per element and
// level, log2(slices per row) levels, one barrier per level.

template <int kSB> struct MergeStage { // the sources of the tile, staged in shared memory (up to 1 << kSB of them)
  static constexpr uint32_t kCap = 1u << kSB;
  uint32_t begin[kCap];    // first partner row of the slice
  uint32_t base[kCap + 1]; // exclusive prefix of the slice lengths = first candidate slot of the slice
  int32_t  ir0[kCap], ir1[kCap];
  uint16_t row_first[kCap], row_end[kCap]; // the sources [row_first, row_end) belong to the same read (row)
};

template <int kSB> struct MergeShared { // everything but the two word buffers, which follow it in dynamic shared memory
  static constexpr uint32_t kSrcMask = (1u << kSB) - 1u;
  uint16_t        ent_src[kTileCap]; // slot -> source whose slice owns the slot (| kDropped: failed the overlap test)
  MergeStage<kSB> src;
  uint32_t        scan[33];
  uint32_t        warp_tot[kFusedWarps];
  uint32_t        row_start[kTileCap / 32 + 1]; // bit p: a row's candidates start at position p (positions = slots of the row)
  uint32_t        smin, smax, max_row_src, next_item;
};
constexpr uint16_t kDropped = 0x8000u;
template <int kSB> constexpr size_t merge_off() { return (sizeof(MergeShared<kSB>) + 15) & ~size_t(15); }
template <int kSB> constexpr size_t merge_bytes(bool wide_words) {
  return merge_off<kSB>() + 2 * kTileCap * (wide_words ? sizeof(unsigned long long) : sizeof(uint32_t));
}
constexpr size_t kMergeBytesWide   = merge_bytes<kSrcBits>(true);
constexpr size_t kMergeBytesNarrow = merge_bytes<kSrcBits>(false);

// materialise the candidates in generation order (source by source = unitig order inside a row); returns this
// thread's number of candidates that pass the overlap test
template <class W, int kSB>
__device__ __forceinline__ uint32_t enumerate_candidates(const TileArgs &a, MergeShared<kSB> &s, W *wa, uint32_t n_src,
                                                         uint32_t n_ent) {
  const unsigned t = threadIdx.x, w = t >> 5, l = t & 31;
  const uint32_t smin  = s.smin;
  const uint32_t per   = (n_ent + kFusedBlock - 1) / kFusedBlock;
  const uint32_t wbase = w * per * 32; // warp w owns slots / positions [wbase, wbase + per * 32)
  uint32_t       kept  = 0;
  if (wbase < n_ent) {
    uint32_t j = 0, hi = n_src; // last source with base <= wbase (empty slices share their base with the next one)
    while (hi - j > 1) {
      uint32_t mid = (j + hi) >> 1;
      if (s.src.base[mid] <= wbase) j = mid; else hi = mid;
    }
    for (uint32_t it = 0; it < per; ++it) {
      const uint32_t e = wbase + it * 32 + l;
      if (e < n_ent) {
        while (s.src.base[j + 1] <= e) ++j; // slots grow with `it`, so every lane walks forward only
        const uint32_t ib = s.src.begin[j] + (e - s.src.base[j]);
        const uint32_t sr = a.p_read[ib];
        const int      a0 = s.src.ir0[j], a1 = s.src.ir1[j], b0 = a.p_ir0[ib], b1 = a.p_ir1[ib];
        const int      o0 = a0 > b0 ? a0 : b0, o1 = a1 < b1 ? a1 : b1;
        const bool     keep = (o0 <= o1 && o1 - o0 > kThOverlapF) || a.no_overlap_test; // MatchMap.cpp:192
        wa[e]               = (W(sr - smin) << kIdxBits) | W(e);
        s.ent_src[e]        = uint16_t(j) | (keep ? uint16_t(0) : kDropped);
        kept += keep ? 1u : 0u;
      }
    }
  }
  return kept;
}

// the merge levels; returns the buffer that holds the sorted words (the other one is free afterwards).  Starts and
// ends with a barrier.
#ifndef MSGPU_MERGE_ODD
#define MSGPU_MERGE_ODD 1
#endif
#ifndef MSGPU_MERGE_MULT
#define MSGPU_MERGE_MULT 2 // outputs per thread and level = this x the even share: with 2, half of the threads merge 17
#endif                     // outputs each and the other half idles — one merge-path search per 17 outputs instead of per
                           // 9 outweighs it (0.709 -> 0.699 ms at yeast scale, profiles/r2_ab_heads_rowmask_merge_mult.log)
template <class W, int kSB>
__device__ __forceinline__ W *merge_levels(MergeShared<kSB> &s, W *wa, W *wb, uint32_t n_ent) {
  constexpr uint32_t kSrcMask = MergeShared<kSB>::kSrcMask;
  const unsigned t = threadIdx.x;
  // `per` consecutive outputs per thread.  An ODD count: lane l starts at word l * per, and an odd stride spreads the
  // 32 lanes of a warp over all 32 banks (a stride of 8 puts four lanes on every bank it touches).
  uint32_t per = (n_ent + kFusedBlock - 1) / kFusedBlock * MSGPU_MERGE_MULT;
  if (MSGPU_MERGE_ODD) per |= 1u;
  __syncthreads();
  W *in = wa, *out = wb;
  const uint32_t max_row = s.max_row_src;
  const uint32_t pos0 = t * per;
  const uint32_t oe   = pos0 + per < n_ent ? pos0 + per : n_ent;
  for (uint32_t lvl = 0; (1u << lvl) < max_row; ++lvl) {
    if (pos0 < oe) {
      const uint32_t G = 1u << lvl;
      // the pair of runs whose outputs contain position p: [a0, a1) and [a1, b1), found from the slice that owns slot p
      uint32_t a0, a1, b1;
      auto pair_of = [&](uint32_t p) {
        const uint32_t j  = s.ent_src[p] & kSrcMask;
        const uint32_t rf = s.src.row_first[j], re = s.src.row_end[j];
        const uint32_t s0 = rf + (((j - rf) >> (lvl + 1)) << (lvl + 1));
        const uint32_t s1 = s0 + G < re ? s0 + G : re, s2 = s0 + 2 * G < re ? s0 + 2 * G : re;
        a0 = s.src.base[s0];
        a1 = s.src.base[s1];
        b1 = s.src.base[s2];
      };
      pair_of(pos0);
      if (a1 == b1 && b1 >= oe) { // a run without a partner that covers this thread's outputs: copied
        for (uint32_t p = pos0; p < oe; ++p) out[p] = in[p];
      } else {
        const uint32_t k  = pos0 - a0, na = a1 - a0, nb = b1 - a1;
        uint32_t       lo = k > nb ? k - nb : 0u, hi = k < na ? k : na; // merge path: elements of A among the first k
        while (lo < hi) {                                               // (an empty side leaves lo == hi)
          const uint32_t mid = (lo + hi) >> 1;
          if (in[a0 + mid] < in[a1 + (k - 1 - mid)]) lo = mid + 1; else hi = mid;
        }
        // An exhausted run reads as ~0, which no word equals (the host gives 32-bit words only to read-id ranges
        // below 2^20 - 1, and 64-bit words have their top bits clear), so one compare decides a step.
        uint32_t ia = a0 + lo, ib = a1 + (k - lo);
        W        xa = ia < a1 ? in[ia] : ~W(0), xb = ib < b1 ? in[ib] : ~W(0);
        for (uint32_t p = pos0; p < oe; ++p) { // one loop for all of the thread's outputs: every lane of a warp runs
          if (p == b1) {                       // the same `per` steps; a lane that crosses into the next pair of runs
            pair_of(p);                        // starts it at its beginning, where the partition is trivial
            ia = a0;
            ib = a1;
            xa = ia < a1 ? in[ia] : ~W(0);
            xb = ib < b1 ? in[ib] : ~W(0);
          }
          const bool take_a = xa < xb; // branch-free step: one select, one reload of the side that moved
          out[p]            = take_a ? xa : xb;
          ia += take_a ? 1u : 0u;
          ib += take_a ? 0u : 1u;
          const uint32_t nx = take_a ? ia : ib, lim = take_a ? a1 : b1;
          const W        v  = nx < lim ? in[nx] : ~W(0);
          xa                = take_a ? v : xa;
          xb                = take_a ? xb : v;
        }
      }
    }
    __syncthreads();
    W *tmp = in;
    in     = out;
    out    = tmp;
  }
  return in;
}

template <class W, int kSB>
__device__ __forceinline__ void merge_tail(const TileArgs &a, MergeShared<kSB> &s, W *wa, W *wb, uint32_t st_idx,
                                           uint32_t n_src, uint32_t n_ent) {
  constexpr uint32_t kSrcMask = MergeShared<kSB>::kSrcMask;
  const unsigned t = threadIdx.x, w = t >> 5, l = t & 31;
  const uint32_t per   = (n_ent + kFusedBlock - 1) / kFusedBlock;
  const uint32_t wbase = w * per * 32; // warp w owns slots / positions [wbase, wbase + per * 32)
  enumerate_candidates<W, kSB>(a, s, wa, n_src, n_ent);
  const W *sorted = merge_levels<W, kSB>(s, wa, wb, n_ent);

  // ---- run heads = edges of this tile.  A candidate that failed the overlap test stays in its run (it has the key
  // of its pair) but is skipped; the head of a run is its first surviving candidate.
  W        xs[kItems];          // this lane's sorted words, kept for the record pass
  uint32_t vmask = 0, hmask = 0; // bit `it`: this lane's item of iteration `it` survives / is a head
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    xs[it] = W(0);
    if (uint32_t(it) < per) {
      const uint32_t k = wbase + it * 32 + l;
      bool           ok = false, head = false;
      if (k < n_ent) {
        // the warp's 32 positions share one word of the row-start bits (wbase is a multiple of 32): one broadcast load
        const bool first_of_row = (s.row_start[(wbase >> 5) + it] >> l) & 1u;
        const W    x            = sorted[k];
        xs[it] = x;
        ok     = !(s.ent_src[uint32_t(x) & (kTileCap - 1)] & kDropped);
        if (ok && !first_of_row) { // (position 0 starts a row)
          const W key = x >> kIdxBits;
          W       y   = sorted[k - 1];
          if ((y >> kIdxBits) != key) {
            head = true; // the usual case: the predecessor belongs to another read pair
          } else if (s.ent_src[uint32_t(y) & (kTileCap - 1)] & kDropped) {
            // rare: the predecessor has the key but failed the overlap test — walk back over such candidates to the
            // nearest survivor with the key, the first one with another key, or the start of the row
            head = true;
            for (uint32_t p = k - 1; !((s.row_start[p >> 5] >> (p & 31)) & 1u);) {
              --p;
              y = sorted[p];
              if ((y >> kIdxBits) != key) break;
              if (!(s.ent_src[uint32_t(y) & (kTileCap - 1)] & kDropped)) {
                head = false;
                break;
              }
            }
          }
        } else {
          head = ok;
        }
      }
      vmask |= (ok ? 1u : 0u) << it;
      hmask |= (head ? 1u : 0u) << it;
    }
  }
  const uint32_t n_v = __reduce_add_sync(0xffffffffu, __popc(vmask)), n_h = __reduce_add_sync(0xffffffffu, __popc(hmask));
  if (l == 0) s.warp_tot[w] = n_v | (n_h << 16);
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int ww = 0; ww < kFusedWarps; ++ww) {
    uint32_t c = s.warp_tot[ww];
    before += ww < int(w) ? c : 0u;
    total += c;
  }
  if (t == 0) a.tile_cnt[st_idx] = make_uint2(total & 0xffffu, total >> 16);

  // ---- the tile's records, compacted, in final order
  uint2   *rec = a.records + uint64_t(st_idx) * kTileCap;
  uint32_t pv = before & 0xffffu, ph = before >> 16;
#pragma unroll
  for (int it = 0; it < kItems; ++it) {
    if (uint32_t(it) < per) {
      const bool     ok = (vmask >> it) & 1u, head = (hmask >> it) & 1u;
      const unsigned bv = __ballot_sync(0xffffffffu, ok), bh = __ballot_sync(0xffffffffu, head);
      const unsigned lt = (1u << l) - 1u;
      if (ok) {
        const uint32_t e  = uint32_t(xs[it]) & (kTileCap - 1);
        const uint32_t j  = s.ent_src[e] & kSrcMask;
        const uint32_t ib = s.src.begin[j] + (e - s.src.base[j]);
        rec[pv + __popc(bv & lt)] = make_uint2(ib, j | (head ? 1u << kRecSrcBits : 0u) | ((ph + __popc(bh & lt)) << (kRecSrcBits + 1)));
      }
      pv += __popc(bv);
      ph += __popc(bh);
    }
  }
}

// kSB = kSrcBits: the product — every tile of the range whose sources fit 512 (all of them, unless a read shares more
// unitigs with partners than that).  kSB = kSrcBitsWide: the tiles of `wide_list` (sub-tiles of rows with 513 .. 4095
// sources: a read that crosses that many shared unitigs is rare but legal input), one resident CTA per SM.
template <int kSB>
__global__ void __launch_bounds__(kFusedBlock, kSB == kSrcBits ? kMergeCtas : 1) pair_tiles_merge_kernel(TileArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MergeShared<kSB>  &s = *reinterpret_cast<MergeShared<kSB> *>(smem_raw);
  const unsigned     t = threadIdx.x, l = t & 31;
  constexpr uint32_t kCap    = 1u << kSB;
  constexpr bool     kIsWide = kSB != kSrcBits;
  const uint32_t     n_items = kIsWide ? a.n_wide : a.tile_hi - a.tile_lo;

  // A CTA's first tile is its block index; the next ones come from a ticket counter (tiles differ in candidates and
  // merge levels, and a fixed stride left the average SM idle for the last 6 % of the kernel: ncu sm__cycles_active).
  // The ticket for the NEXT tile is drawn by thread 0 between the two barriers that open a tile, so the atomic's
  // round trip hides behind the staging of this one.  (The instantiation for many sources keeps its stride.)
  uint32_t nxt = 0;
  for (uint32_t item = blockIdx.x; item < n_items; item = nxt) {
    nxt = item + gridDim.x;
    const uint32_t tile = kIsWide ? a.wide_list[item] : a.tile_lo + item;
    if (kIsWide && (tile < a.tile_lo || tile >= a.tile_hi)) continue; // (a chunk takes the listed tiles of its range)
    __syncthreads(); // previous tile fully done with shared memory
    if (t == 0) {
      s.smin        = 0xffffffffu;
      s.smax        = 0;
      s.max_row_src = 0;
      if (!kIsWide && MSGPU_TICKETS) s.next_item = gridDim.x + atomicAdd(a.ticket, 1u);
    }
    if (t < kTileCap / 32 + 1) s.row_start[t] = 0;
    __syncthreads();
    if (!kIsWide && MSGPU_TICKETS) nxt = s.next_item;
    const uint32_t st_idx = tile - a.tile_lo;
    const uint32_t r0 = a.tile_row[tile], kind = a.tile_kind[tile];
    const uint32_t q0 = a.rd_off[r0];
    const uint32_t n_src = (kind ? a.rd_off[r0 + 1] : a.rd_off[a.tile_row[tile + 1]]) - q0;
    if (!kIsWide && n_src > kCap) { // a sub-tile of a row with many sources: the wide instantiation takes it ...
      if (n_src >= kSrcCapWide && t == 0) a.tile_cnt[st_idx] = make_uint2(0, 0); // ... unless the build is refused
      continue;                                                                  // (write_tiles_kernel flags the row)
    }
    if (n_src > kCap) {                     // cannot happen: the planning refuses rows with more sources than that
      if (t == 0) {
        *a.error           = 1;
        a.tile_cnt[st_idx] = make_uint2(0, 0);
      }
      continue;
    }
    // ---- stage the sources, blockDim.x at a time (one pass for the product instantiation)
    uint32_t carry = 0, lmin = 0xffffffffu, lmax = 0, row_src = 0;
#pragma unroll
    for (uint32_t pass = 0; pass < kCap / kFusedBlock; ++pass) { // (exactly one pass in the product instantiation)
      if (pass && pass * kFusedBlock >= n_src) break;
      const uint32_t j   = pass * kFusedBlock + t;
      uint32_t       len = 0;
      if (j < n_src) {
        const uint32_t q   = a.rd_src[q0 + j];
        const PairRow  row = a.rows[q];
        uint32_t       b, e;
        if (kind) {
          const uint32_t *sl = a.slices + a.slice_off[r0] + uint64_t(kind - 1) * n_src;
          b                  = sl[j];
          e                  = sl[n_src + j];
          s.src.row_first[j] = 0;
          s.src.row_end[j]   = uint16_t(n_src);
          row_src            = n_src;
        } else {
          b                  = q + 1;
          e                  = a.grp_off[row.unitig + 1];
          const uint32_t f   = a.rd_off[row.read] - q0, g = a.rd_off[row.read + 1] - q0;
          s.src.row_first[j] = uint16_t(f);
          s.src.row_end[j]   = uint16_t(g);
          row_src            = g - f > row_src ? g - f : row_src;
        }
        len            = e - b;
        s.src.begin[j] = b;
        s.src.ir0[j]   = row.ir0;
        s.src.ir1[j]   = row.ir1;
        if (len) { // the slice is sorted by read id
          const uint32_t x0 = a.p_read[b], x1 = a.p_read[e - 1];
          lmin = x0 < lmin ? x0 : lmin;
          lmax = x1 > lmax ? x1 : lmax;
        }
      }
      uint32_t tot, ex = block_exclusive_scan<uint32_t>(len, s.scan, &tot);
      if (j < n_src) {
        s.src.base[j] = carry + ex;
        // the first source of a row marks where the row's candidates start (a tile that overflows is refused below)
        if (s.src.row_first[j] == j && carry + ex <= kTileCap) atomicOr(&s.row_start[(carry + ex) >> 5], 1u << ((carry + ex) & 31));
      }
      carry += tot;
    }
    if (t == 0) s.src.base[n_src] = carry;
    for (int d = 16; d > 0; d >>= 1) {
      uint32_t omin = __shfl_xor_sync(0xffffffffu, lmin, d), omax = __shfl_xor_sync(0xffffffffu, lmax, d);
      uint32_t orow = __shfl_xor_sync(0xffffffffu, row_src, d);
      lmin    = omin < lmin ? omin : lmin;
      lmax    = omax > lmax ? omax : lmax;
      row_src = orow > row_src ? orow : row_src;
    }
    if (l == 0) {
      if (lmin <= lmax) {
        atomicMin(&s.smin, lmin);
        atomicMax(&s.smax, lmax);
      }
      atomicMax(&s.max_row_src, row_src);
    }
    __syncthreads();
    const uint32_t n_ent = s.src.base[n_src];
    if (n_ent > kTileCap) {
      if (t == 0) {
        *a.error           = 2;
        a.tile_cnt[st_idx] = make_uint2(0, 0);
      }
      continue;
    }
    // 32-bit words when (second - min second) << 12 | slot fits AND can never be all ones (the merge's end marker)
    const uint32_t range = (n_ent && s.smax > s.smin) ? s.smax - s.smin : 0u;
    unsigned char *words = smem_raw + merge_off<kSB>();
    if (range < (1u << (32 - kIdxBits)) - 1u) {
      uint32_t *wa = reinterpret_cast<uint32_t *>(words);
      merge_tail<uint32_t, kSB>(a, s, wa, wa + kTileCap, st_idx, n_src, n_ent);
    } else if (!a.wide) { // cannot happen: the host sizes the launch from the id width
      if (t == 0) {
        *a.error           = 3;
        a.tile_cnt[st_idx] = make_uint2(0, 0);
      }
    } else {
      unsigned long long *wa = reinterpret_cast<unsigned long long *>(words);
      merge_tail<unsigned long long, kSB>(a, s, wa, wa + kTileCap, st_idx, n_src, n_ent);
    }
  }
}

__device__ __forceinline__ PairRow ld_row(const PairRow *p) { // one 256-bit read-only load = one 32-byte sector
  uint4 x, y;
  ld_row32(p, x, y);
  PairRow      r;
  r.ir0    = int32_t(x.x);
  r.ir1    = int32_t(x.y);
  r.read   = x.z;
  r.score  = int32_t(x.w);
  r.line   = y.x;
  r.flags  = y.y;
  r.unitig = y.z;
  r.orig   = y.w;
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// K7b: tile counts -> global offsets of every tile's first edge match and first edge (two single-pass scans)
// ---------------------------------------------------------------------------------------------------------------------
struct TileValid { // surviving candidates of a tile
  const uint2 *cnt;
  __device__ unsigned long long operator()(uint64_t i) const { return cnt[i].x; }
};
struct TileHeads { // run heads (= edges) of a tile
  const uint2 *cnt;
  __device__ unsigned long long operator()(uint64_t i) const { return cnt[i].y; }
};

// ---------------------------------------------------------------------------------------------------------------------
// K7c: every record becomes one edge match (and, for run heads, one edge) at its final position
// ---------------------------------------------------------------------------------------------------------------------
struct EmitArgs {
  const PairRow  *rows;
  const uint32_t *rd_off, *rd_src, *tile_row, *tile_kind;
  uint32_t        tile_lo, n_tiles;
  const uint2    *records, *tile_cnt;
  const unsigned long long *em_base, *edge_base, *totals; // totals: (edge matches, edges) of this launch
  unsigned long long em_value_base; // added to the em_begin VALUES (rows of earlier chunks of the same slice)
  // pipelined emission into ONE resident table (pairing_pipelined): the rows of this launch continue those of the
  // `n_prev` launches before it, whose totals (4 words each) start at prev_totals; only the last launch closes the CSR
  const unsigned long long *prev_totals;
  uint32_t        n_prev, close_csr;
  unsigned int   *ticket; // zero at launch: hands out the tiles behind the first gridDim.x
  EdgeTables      out;
};

// rows (edge matches, edges) written by the earlier launches of a pipelined emission
__device__ __forceinline__ void rows_before(const EmitArgs &a, unsigned long long *em, unsigned long long *ed) {
  unsigned long long x = 0, y = 0;
  for (uint32_t c = 0; c < a.n_prev; ++c) {
    x += a.prev_totals[4 * c];
    y += a.prev_totals[4 * c + 1];
  }
  *em = x;
  *ed = y;
}


#ifndef MSGPU_EMIT_CTAS
#define MSGPU_EMIT_CTAS 3
#endif
#ifndef MSGPU_EMIT_BATCH
#define MSGPU_EMIT_BATCH 4
#endif
constexpr uint32_t kMaxPipeChunks  = 16;   // chunks of a pipelined emission (pairing_pipelined)
constexpr int kEmitBlock = 256;
constexpr int kEmitCtas  = MSGPU_EMIT_CTAS;  // resident CTAs per SM the emission kernels are compiled for
constexpr int kEmitBatch = MSGPU_EMIT_BATCH; // records in flight per thread: the partner gathers are latency-bound
#ifndef MSGPU_EMIT_BATCH_UNDER
#define MSGPU_EMIT_BATCH_UNDER 2
#endif
constexpr int kEmitBatchUnder = MSGPU_EMIT_BATCH_UNDER; // ... of the 64-register instantiation that runs underneath the tile kernel

// One CTA per tile at a time (tiles by tickets).  The tile's sources (the `first` side of its records, <= 512 rows) are staged in shared
// memory once; every record then needs ONE dependent gather (its partner's row) after its own coalesced load.
__global__ void __launch_bounds__(kEmitBlock, kEmitCtas) pair_emit_rows_kernel(EmitArgs a) {
  pdl_sync();
  __shared__ uint4 src_lo[kSrcCap], src_hi[kSrcCap]; // PairRow halves of the tile's sources
  const unsigned   t = threadIdx.x;
  unsigned long long em_before, ed_before;
  rows_before(a, &em_before, &ed_before);
  if (blockIdx.x == 0 && t == 0 && a.close_csr) a.out.em_begin[ed_before + a.totals[1]] = a.em_value_base + em_before + a.totals[0];
  __shared__ uint32_t next_ti; // tiles by tickets, as in the bulk-copy kernel
  uint32_t            nxt = 0;
  for (uint32_t ti = blockIdx.x; ti < a.n_tiles; ti = nxt) {
    const uint32_t           tile    = a.tile_lo + ti;
    const uint32_t           n_valid = a.tile_cnt[ti].x;
    const unsigned long long em0 = em_before + a.em_base[ti], ed0 = ed_before + a.edge_base[ti];
    const uint32_t           r0 = a.tile_row[tile], q0 = a.rd_off[r0];
    const uint32_t n_src = (a.tile_kind[tile] ? a.rd_off[r0 + 1] : a.rd_off[a.tile_row[tile + 1]]) - q0;
    const uint2   *rec   = a.records + uint64_t(ti) * kTileCap;
    uint2          r[kEmitBatch];
#pragma unroll
    for (int u = 0; u < kEmitBatch; ++u) { // the first batch of records travels while the sources are staged
      const uint32_t k = t + u * kEmitBlock;
      r[u]             = k < n_valid ? __ldg(rec + k) : make_uint2(0, 0);
    }
    __syncthreads(); // previous tile done with the staged sources
    if (t == 0) next_ti = MSGPU_TICKETS ? gridDim.x + atomicAdd(a.ticket, 1u) : ti + gridDim.x;
    const bool staged = n_src <= kSrcCap; // (rows with more sources: their rows are read in place, one more gather)
    if (n_valid && staged)
      for (uint32_t j = t; j < n_src; j += kEmitBlock) {
        ld_row32(a.rows + a.rd_src[q0 + j], src_lo[j], src_hi[j]);
      }
    __syncthreads();
    nxt = next_ti;
    for (uint32_t k0 = t; k0 < n_valid; k0 += kEmitBlock * kEmitBatch) {
      PairRow pb[kEmitBatch];
#pragma unroll
      for (int u = 0; u < kEmitBatch; ++u) pb[u] = ld_row(a.rows + r[u].x);
      uint2 cur[kEmitBatch];
#pragma unroll
      for (int u = 0; u < kEmitBatch; ++u) { // next batch of records: in flight during the arithmetic below
        cur[u]           = r[u];
        const uint32_t k = k0 + (kEmitBatch + u) * kEmitBlock;
        r[u]             = k < n_valid ? __ldg(rec + k) : make_uint2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < kEmitBatch; ++u) {
        const uint32_t k = k0 + u * kEmitBlock;
        if (k >= n_valid) continue;
        const uint32_t j  = cur[u].y & ((1u << kRecSrcBits) - 1u);
        uint4          lo, hi;
        if (staged) {
          lo = src_lo[j];
          hi = src_hi[j];
        } else {
          ld_row32(a.rows + a.rd_src[q0 + j], lo, hi);
        }
        const int      a0 = int32_t(lo.x), a1 = int32_t(lo.y), b0 = pb[u].ir0, b1 = pb[u].ir1;
        const int      ov0 = a0 > b0 ? a0 : b0, ov1 = a1 < b1 ? a1 : b1;
        // MatchMap.cpp:195-202: score_x * common / len_x for the outer and the inner match, summed.  IEEE addition
        // is commutative, so which of the two is "outer" (the later line) does not change the sum.  No FMA
        // contraction (explicit _rn intrinsics, and the library is built with --fmad=false).
        const double common = double(ov1 - ov0 + 1);
        const double sa     = double((unsigned long long)(long long)int32_t(lo.w)); // size_t widened from int
        const double sbb    = double((unsigned long long)(long long)pb[u].score);
        const double sc_a   = __ddiv_rn(__dmul_rn(sa, common), double(a1 - a0 + 1));
        const double sc_b   = __ddiv_rn(__dmul_rn(sbb, common), double(b1 - b0 + 1));
        const uint32_t la = hi.x, fa = hi.y;
        const unsigned long long o = em0 + k;
        a.out.em_unitig[o] = hi.z;
        a.out.em_ov0[o]    = ov0;
        a.out.em_ov1[o]    = ov1;
        a.out.em_score[o]  = __dadd_rn(sc_a, sc_b);
        a.out.em_line[o]   = la > pb[u].line ? la : pb[u].line; // outer = the later line (MatchMap.cpp:218)
        a.out.em_flags[o]  = uint8_t((((fa ^ pb[u].flags) & 1u) ? 0u : 1u) | ((fa & pb[u].flags & 2u) ? 2u : 0u));
        if (cur[u].y & (1u << kRecSrcBits)) {
          const unsigned long long eo = ed0 + (cur[u].y >> (kRecSrcBits + 1));
          a.out.first[eo]    = lo.z;
          a.out.second[eo]   = pb[u].read;
          a.out.em_begin[eo] = a.em_value_base + o;
        }
      }
    }
  }
}

// The same emission with the tile's records brought into shared memory by ONE bulk copy (cp.async.bulk — the 1-D form of
// the TMA unit, SASS UBLKCP — completing on an mbarrier) instead of coalesced per-thread loads: a tile's records are one
// contiguous, 16-byte aligned run, the shape bulk copies want.  The north star's "staged via TMA", measured against the
// per-thread form (MSGPU_EMIT=scalar) and adopted: DESIGN.md §4.  (The candidates of the tile kernel are NOT such a shape:
// a source's partner slice starts at an arbitrary 4-byte offset of three separate columns, and staging 4096
// candidates x 12 bytes would take the shared memory that buys the kernel its four resident CTAs.)
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

// kCtas = resident CTAs per SM the kernel is compiled for: kEmitCtas on its own; 4 (at most 64 registers: one such
// CTA takes exactly the registers, and less than the shared memory, that one CTA of the tile kernel leaves behind)
// when it runs underneath the tile kernel of the next chunk (pairing_pipelined).
template <int kCtas, int kBatch> __global__ void __launch_bounds__(kEmitBlock, kCtas) pair_emit_rows_bulk_kernel(EmitArgs a) {
  pdl_sync();
  __shared__ uint4 src_lo[kSrcCap], src_hi[kSrcCap]; // PairRow halves of the tile's sources
  __shared__ __align__(8) unsigned long long bar;
  extern __shared__ __align__(16) unsigned char bulk_raw[];
  uint2           *recs = reinterpret_cast<uint2 *>(bulk_raw); // kTileCap records
  const unsigned   t = threadIdx.x;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  unsigned long long em_before, ed_before;
  rows_before(a, &em_before, &ed_before);
  if (blockIdx.x == 0 && t == 0 && a.close_csr) a.out.em_begin[ed_before + a.totals[1]] = a.em_value_base + em_before + a.totals[0];
  // tiles by tickets, as in the tile kernel (a tile holds 500 .. 4096 records: with a fixed stride the average SM idled
  // for 9 % of the kernel); thread 0 draws the next ticket while the sources of this tile are staged
  __shared__ uint32_t next_ti;
  uint32_t            nxt = 0;
  for (uint32_t ti = blockIdx.x; ti < a.n_tiles; ti = nxt) {
    const uint32_t           tile    = a.tile_lo + ti;
    const uint32_t           n_valid = a.tile_cnt[ti].x;
    const unsigned long long em0 = em_before + a.em_base[ti], ed0 = ed_before + a.edge_base[ti];
    const uint32_t           r0 = a.tile_row[tile], q0 = a.rd_off[r0];
    const uint32_t n_src = (a.tile_kind[tile] ? a.rd_off[r0 + 1] : a.rd_off[a.tile_row[tile + 1]]) - q0;
    const uint2   *rec   = a.records + uint64_t(ti) * kTileCap;
    __syncthreads(); // previous tile done with the records and the staged sources
    if (t == 0) next_ti = MSGPU_TICKETS ? gridDim.x + atomicAdd(a.ticket, 1u) : ti + gridDim.x;
    if (t == 0 && n_valid) {
      const uint32_t bytes = (n_valid * uint32_t(sizeof(uint2)) + 15u) & ~15u; // (a tile's slots are 32 KB: in bounds)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(recs)),
                   "l"(rec), "r"(bytes), "r"(smem_addr(&bar))
                   : "memory");
    }
    const bool staged = n_src <= kSrcCap;
    if (n_valid && staged)
      for (uint32_t j = t; j < n_src; j += kEmitBlock) {
        ld_row32(a.rows + a.rd_src[q0 + j], src_lo[j], src_hi[j]);
      }
    __syncthreads();
    nxt = next_ti;
    if (n_valid) { // every thread waits for the copy (hardware sleep, not a poll)
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smem_addr(&bar)), "r"(phase)
                     : "memory");
      phase ^= 1u;
    }
    for (uint32_t k0 = t; k0 < n_valid; k0 += kEmitBlock * kBatch) {
      uint2   cur[kBatch];
      PairRow pb[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const uint32_t k = k0 + u * kEmitBlock;
        cur[u]           = k < n_valid ? recs[k] : make_uint2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) pb[u] = ld_row(a.rows + cur[u].x);
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const uint32_t k = k0 + u * kEmitBlock;
        if (k >= n_valid) continue;
        const uint32_t j = cur[u].y & ((1u << kRecSrcBits) - 1u);
        uint4          lo, hi;
        if (staged) {
          lo = src_lo[j];
          hi = src_hi[j];
        } else {
          ld_row32(a.rows + a.rd_src[q0 + j], lo, hi);
        }
        const int a0 = int32_t(lo.x), a1 = int32_t(lo.y), b0 = pb[u].ir0, b1 = pb[u].ir1;
        const int ov0 = a0 > b0 ? a0 : b0, ov1 = a1 < b1 ? a1 : b1;
        const double common = double(ov1 - ov0 + 1); // MatchMap.cpp:195-202, as in pair_emit_rows_kernel
        const double sa     = double((unsigned long long)(long long)int32_t(lo.w));
        const double sbb    = double((unsigned long long)(long long)pb[u].score);
        const double sc_a   = __ddiv_rn(__dmul_rn(sa, common), double(a1 - a0 + 1));
        const double sc_b   = __ddiv_rn(__dmul_rn(sbb, common), double(b1 - b0 + 1));
        const uint32_t la = hi.x, fa = hi.y;
        const unsigned long long o = em0 + k;
        a.out.em_unitig[o] = hi.z;
        a.out.em_ov0[o]    = ov0;
        a.out.em_ov1[o]    = ov1;
        a.out.em_score[o]  = __dadd_rn(sc_a, sc_b);
        a.out.em_line[o]   = la > pb[u].line ? la : pb[u].line;
        a.out.em_flags[o]  = uint8_t((((fa ^ pb[u].flags) & 1u) ? 0u : 1u) | ((fa & pb[u].flags & 2u) ? 2u : 0u));
        if (cur[u].y & (1u << kRecSrcBits)) {
          const unsigned long long eo = ed0 + (cur[u].y >> (kRecSrcBits + 1));
          a.out.first[eo]    = lo.z;
          a.out.second[eo]   = pb[u].read;
          a.out.em_begin[eo] = a.em_value_base + o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: planning (everything up to the tile table) and emission of a range of tiles, so that a result can be
// produced whole (resident) or chunk by chunk (results larger than HBM, downloads overlapped with the pairing)
// ---------------------------------------------------------------------------------------------------------------------
void pairing_plan(Arena &arena, cudaStream_t st, const MatchSoA &m, const UnitigTable &unitigs, uint64_t n_reads,
                  uint32_t part, uint32_t n_parts, PairPlan &pl, bool no_overlap_test) {
  pl = PairPlan{};
  pl.no_overlap_test = no_overlap_test;
  const uint64_t M = m.n, R = n_reads;
  if (R >= 0xfffffff0ull || M >= 0xfffffff0ull)
    throw BuildError{MSGPU_ERR_LIMIT, "pairing supports < 2^32 reads and matches per device"};

  // ---- pre-pass: groups ordered by read id, per-read source lists
  PairRow  *rows   = arena.alloc<PairRow>(M);
  uint32_t *p_read = arena.alloc<uint32_t>(M);
  int32_t  *p_ir0 = arena.alloc<int32_t>(M), *p_ir1 = arena.alloc<int32_t>(M);
  // everything that must start at zero lives in one block, cleared by one memset: per-read counters, the flags and
  // counters of the kernels
  const size_t z_deg = 0, z_cursor = z_deg + (R + 1) * sizeof(uint32_t), z_wgt = ((z_cursor + (R + 1) * sizeof(uint32_t)) + 7) & ~size_t(7);
  const size_t z_u32 = z_wgt + (R + 1) * sizeof(unsigned long long), z_u64 = z_u32 + 8 * sizeof(uint32_t);
  const size_t z_end = z_u64 + 4 * sizeof(uint64_t);
  char        *zero  = arena.alloc<char>(z_end);
  MSGPU_CUDA(cudaMemsetAsync(zero, 0, z_end, st));
  uint32_t           *deg    = reinterpret_cast<uint32_t *>(zero + z_deg);
  uint32_t           *cursor = reinterpret_cast<uint32_t *>(zero + z_cursor);
  unsigned long long *wgt    = reinterpret_cast<unsigned long long *>(zero + z_wgt);
  uint32_t           *d_u32  = reinterpret_cast<uint32_t *>(zero + z_u32); // [0] sources [1] tiles [2] too many sources
                                                                           // [3] tile error [4] work items of the longer groups [5] -
                                                                           // [6] tiles with > 512 sources [7] cursor of their list
  uint64_t           *d_u64  = reinterpret_cast<uint64_t *>(zero + z_u64);
  uint32_t           *rd_off = arena.alloc<uint32_t>(R + 1);
  const int gm = grid_for(M, 256), gr = grid_for(R, 256);
  if (M) {
    RankArgs ra{m, unitigs.grp_off, rows, p_read, p_ir0, p_ir1, deg, wgt};
    // groups ordered by read id: a warp sorts a short group, CTAs rank the members of a longer one by counting
    const uint32_t U = uint32_t(unitigs.n);
    uint2         *big_items = arena.alloc<uint2>(uint64_t(U) + M / kRankChunk + 1);
    MSGPU_LAUNCH("rank_rows", st, launch_k(rank_rows_kernel, grid_for(uint64_t(U) * 32, 256), 256, 0, st, ra, U, big_items, d_u32 + 4));
    MSGPU_LAUNCH("rank_rows_cta", st, launch_k(rank_rows_cta_kernel, kNumSMs * 8, kRankChunk, 0, st, ra, big_items, d_u32 + 4));
  }
  // ---- tile planning: four independent scans over the reads in one launch (source offsets, clipped row weights,
  // slice-table offsets, the pair count), then the tile index, which needs the weights' prefix
  RowPlan   plan{deg, wgt};
  uint64_t *prefix = arena.alloc<uint64_t>(R + 1), *slice_off = arena.alloc<uint64_t>(R + 1);
  uint32_t *tile_idx = arena.alloc<uint32_t>(R + 1);
  run_scans(arena, st, scan_job<uint32_t>(arena, st, 0, LoadAs<uint32_t, uint32_t>{deg}, R, rd_off, d_u32 + 0),
            scan_job<uint64_t>(arena, st, 1, ClippedWeight{plan}, R, prefix, (uint64_t *)nullptr),
            scan_job<uint64_t>(arena, st, 2, SliceCount{plan}, R, slice_off, d_u64 + 0),
            scan_job<uint64_t>(arena, st, 3, WeightOf{wgt}, R, (uint64_t *)nullptr, d_u64 + 1));
  MSGPU_CUDA(cudaMemcpyAsync(rd_off + R, d_u32 + 0, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  TileCount tc{plan, prefix, d_u32 + 6};
  exclusive_scan<uint32_t>(arena, st, tc, R, tile_idx, d_u32 + 1);
  // one copy of the adjacent counter words (d_u32[0..7], d_u64[0..3]) into pinned memory
  char *h_words = static_cast<char *>(arena.host_scratch());
  MSGPU_CUDA(cudaMemcpyAsync(h_words, zero + z_u32, z_end - z_u32, cudaMemcpyDeviceToHost, st));
  const uint32_t *h32 = reinterpret_cast<const uint32_t *>(h_words);
  const uint64_t *h64 = reinterpret_cast<const uint64_t *>(h_words + (z_u64 - z_u32));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  const uint32_t n_src = h32[0], n_tiles = h32[1], n_wide = h32[6];
  const uint64_t n_slices = h64[0], P = h64[1]; // P = sum k(k-1)/2: every unordered pair is one suffix member

  uint32_t *tile_row = arena.alloc<uint32_t>(uint64_t(n_tiles) + 1), *tile_kind = arena.alloc<uint32_t>(uint64_t(n_tiles) + 1);
  uint32_t *wide_list = arena.alloc<uint32_t>(uint64_t(n_wide) + 1);
  if (R) MSGPU_LAUNCH("write_tiles", st, launch_k(write_tiles_kernel, gr, 256, 0, st, tc, tile_idx, R, n_tiles, tile_row, tile_kind, d_u32 + 2, wide_list, d_u32 + 7));
  uint32_t *rd_tmp = arena.alloc<uint32_t>(n_src), *rd_src = arena.alloc<uint32_t>(n_src);
  // ---- this device's share: a contiguous range of tiles (tiles weigh about the same)
  if (n_parts == 0) n_parts = 1;
  const uint32_t tile_lo = uint32_t(uint64_t(n_tiles) * part / n_parts);
  const uint32_t tile_hi = uint32_t(uint64_t(n_tiles) * (part + 1) / n_parts);
  const uint32_t my_tiles = tile_hi - tile_lo;
  // source lists: of every row on one device, of the rows of this device's tiles (+ the one after) on several
  RowRange rows_of{nullptr, 0u, 0u, uint32_t(R)};
  if (n_parts > 1) rows_of = RowRange{tile_row, tile_lo, tile_hi < n_tiles ? tile_hi : n_tiles, uint32_t(R)};
  if (M) MSGPU_LAUNCH("scatter_sources", st, launch_k(scatter_sources_kernel, gm, 256, 0, st, p_read, rows, unitigs.grp_off, M, rd_off, cursor, rd_tmp, rows_of));
  if (n_src) MSGPU_LAUNCH("order_sources", st, launch_k(order_sources_kernel, grid_for(n_src, 256), 256, 0, st, p_read, rd_off, rd_tmp, n_src, rd_src, rows_of));
  // heavy rows: every sub-tile computes its START boundary; its end boundary is the start of the next sub-tile, so
  // the range includes one tile past this device's share (a last sub-tile writes the end of its row itself)
  uint32_t *slices = arena.alloc<uint32_t>(n_slices);
  if (n_slices && my_tiles) {
    const uint32_t t_end = tile_hi < n_tiles ? tile_hi + 1 : n_tiles;
    MSGPU_LAUNCH("split_heavy", st, launch_k(split_heavy_kernel, grid_for(uint64_t(t_end - tile_lo) * 32, 256), 256, 0, st, 
                                        tile_row, tile_kind, tile_lo, t_end, plan, slice_off, rd_off, rd_src, rows,
                                        unitigs.grp_off, p_read, uint32_t(R), slices));
  }
  pl.rows      = rows;
  pl.p_read    = p_read;
  pl.p_ir0     = p_ir0;
  pl.p_ir1     = p_ir1;
  pl.grp_off   = unitigs.grp_off;
  pl.rd_off    = rd_off;
  pl.rd_src    = rd_src;
  pl.tile_row  = tile_row;
  pl.tile_kind = tile_kind;
  pl.slice_off = slice_off;
  pl.slices    = slices;
  pl.wide_list = wide_list;
  pl.n_wide    = n_wide;
  pl.flags     = d_u32;
  pl.n_reads   = R;
  pl.n_pairs   = P;
  pl.n_tiles   = n_tiles;
  pl.tile_lo   = tile_lo;
  pl.tile_hi   = tile_hi;
  pl.valid     = true;
}

// scratch of one emission: the compact records of its tiles, their counts and offsets, the totals (device)
void pairing_scratch_alloc(Arena &arena, uint32_t max_tiles, PairScratch &sc) {
  sc.max_tiles = max_tiles;
  sc.records   = arena.alloc<uint2>(uint64_t(max_tiles) * kTileCap);
  sc.tile_cnt  = arena.alloc<uint2>(uint64_t(max_tiles) + 1);
  sc.em_base   = arena.alloc<unsigned long long>(uint64_t(max_tiles) + 1);
  sc.ed_base   = arena.alloc<unsigned long long>(uint64_t(max_tiles) + 1);
  sc.status    = arena.alloc<unsigned long long>(2 * uint64_t(max_tiles) + 2);
  sc.totals    = arena.alloc<unsigned long long>(4 * kMaxPipeChunks); // per chunk of a pipelined emission: [0] edge matches [1] edges
}

uint64_t pairing_rows_bound(const PairPlan &pl, uint32_t n_tiles) { // output rows a range of tiles can produce at most
  const uint64_t cap = uint64_t(n_tiles) * kTileCap;
  return cap < pl.n_pairs ? cap : pl.n_pairs;
}

void pairing_outputs_alloc(Arena &arena, uint64_t cap, EdgeTables &e) {
  e.em_unitig = arena.alloc<uint32_t>(cap);
  e.em_ov0    = arena.alloc<int32_t>(cap);
  e.em_ov1    = arena.alloc<int32_t>(cap);
  e.em_score  = arena.alloc<double>(cap);
  e.em_line   = arena.alloc<uint32_t>(cap);
  e.em_flags  = arena.alloc<uint8_t>(cap);
  e.first     = arena.alloc<uint32_t>(cap);
  e.second    = arena.alloc<uint32_t>(cap);
  e.em_begin  = arena.alloc<uint64_t>(cap + 1);
}

// `off`: where the range's first tile sits in the scratch arrays (0 unless several ranges share one scratch)
static TileArgs tile_args(const PairPlan &pl, uint32_t t_lo, uint32_t t_hi, const PairScratch &sc, uint32_t off = 0) {
  TileArgs a;
  a.rows      = pl.rows;
  a.p_read    = pl.p_read;
  a.p_ir0     = pl.p_ir0;
  a.p_ir1     = pl.p_ir1;
  a.grp_off   = pl.grp_off;
  a.rd_off    = pl.rd_off;
  a.rd_src    = pl.rd_src;
  a.tile_row  = pl.tile_row;
  a.tile_kind = pl.tile_kind;
  a.slice_off = pl.slice_off;
  a.slices    = pl.slices;
  a.tile_lo   = t_lo;
  a.tile_hi   = t_hi;
  a.records   = sc.records + uint64_t(off) * kTileCap;
  a.tile_cnt  = sc.tile_cnt + off;
  a.wide_list = pl.wide_list;
  a.n_wide    = pl.n_wide;
  a.error     = pl.flags + 3;
  a.ticket    = nullptr; // (set by tiles_launch)
  // a merge word is (second - min second) << 12 | slot: 32 bits are enough when read ids have <= 20 bits; otherwise
  // tiles choose per tile and the launch carries the 64-bit buffers
  a.wide = pl.n_reads > (1ull << (32 - kIdxBits)) - 1 ? 1u : 0u; // (then second - min second <= 2^20 - 2: no word is ~0)
  a.no_overlap_test = pl.no_overlap_test ? 1u : 0u;
  return a;
}

// First half of the emission of tiles [t_lo, t_hi): enumerate + sort every tile into compact records, then the two
// scans that turn the tiles' counts into row offsets.  Needs nothing from earlier ranges, so the tiles of the next
// chunk can run while the rows of this one are consumed.  Asynchronous; sc.totals (device) = the range's row counts.
// `totals`: the four zeroed words of this emission — [0] edge matches, [1] edges (written by the scans), [2] / [3] the
// tile tickets of the tile kernel / of the emission kernel
static void tiles_launch(cudaStream_t st, const PairPlan &pl, uint32_t t_lo, uint32_t t_hi, const PairScratch &sc, uint32_t off,
                         unsigned long long *totals) {
  const uint32_t n = t_hi - t_lo;
  TileArgs       a = tile_args(pl, t_lo, t_hi, sc, off);
  a.ticket         = reinterpret_cast<unsigned int *>(totals + 2);
  // (the shared-memory opt-in is an attribute of the function ON THE CURRENT DEVICE: set before every launch — a
  //  context per GPU in one process is supported, and the call costs less than a microsecond)
  MSGPU_CUDA(cudaFuncSetAttribute(pair_tiles_merge_kernel<kSrcBits>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMergeBytesWide)));
  // 32 registers, 51 KB (narrow): four resident CTAs per SM, one wave of them, tiles handed out by tickets
  int grid = kNumSMs * (a.wide ? (kIdxBits <= 12 ? 2 : 1) : kMergeCtas) * (MSGPU_TICKETS ? 1 : 4);
  if (uint32_t(grid) > n) grid = int(n);
  MSGPU_LAUNCH("pair_tiles", st, launch_k(pair_tiles_merge_kernel<kSrcBits>, grid, kFusedBlock, a.wide ? kMergeBytesWide : kMergeBytesNarrow, st, a));
  if (pl.n_wide) { // sub-tiles of rows with 513 .. 4095 sources (skipped above): the same kernel staged for that many
    const size_t bytes = merge_bytes<kSrcBitsWide>(a.wide != 0);
    MSGPU_CUDA(cudaFuncSetAttribute(pair_tiles_merge_kernel<kSrcBitsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(merge_bytes<kSrcBitsWide>(true))));
    int gw = pl.n_wide < uint32_t(kNumSMs) ? int(pl.n_wide) : kNumSMs;
    MSGPU_LAUNCH("pair_tiles_wide", st, launch_k(pair_tiles_merge_kernel<kSrcBitsWide>, gw, kFusedBlock, bytes, st, a));
  }
}
static void tiles_scans(Arena &arena, cudaStream_t st, uint32_t n, const PairScratch &sc, uint32_t off, unsigned long long *totals) {
  run_scans(arena, st, scan_job<unsigned long long>(arena, st, 0, TileValid{sc.tile_cnt + off}, n, sc.em_base + off, totals + 0),
            scan_job<unsigned long long>(arena, st, 1, TileHeads{sc.tile_cnt + off}, n, sc.ed_base + off, totals + 1));
}
void pairing_tiles(Arena &arena, cudaStream_t st, const PairPlan &pl, uint32_t t_lo, uint32_t t_hi, PairScratch &sc) {
  const uint32_t n = t_hi - t_lo;
  if (n > sc.max_tiles) throw BuildError{MSGPU_ERR_INTERNAL, "pairing: emission scratch too small"};
  sc.n_chunks = 1;
  MSGPU_CUDA(cudaMemsetAsync(sc.totals, 0, 4 * sizeof(unsigned long long), st));
  if (!n) return;
  tiles_launch(st, pl, t_lo, t_hi, sc, 0, sc.totals);
  tiles_scans(arena, st, n, sc, 0, sc.totals);
}

// Second half: the records of the range -> the columns of `e` (indexed from the range's first edge match / edge; the
// em_begin VALUES start at em_value_base = the rows of the earlier ranges of the same slice).  Asynchronous.
// `chunk` of a pipelined emission: its rows continue those of the chunks before it (see EmitArgs); `underneath`: the
// launch is meant to run beside the tile kernel of the next chunk.
static void rows_launch(cudaStream_t st, const PairPlan &pl, uint32_t t_lo, uint32_t t_hi, const PairScratch &sc, uint32_t off,
                        const EdgeTables &e, uint64_t em_value_base, uint32_t chunk, bool last, bool underneath) {
  const uint32_t n = t_hi - t_lo;
  EmitArgs ea;
  ea.rows          = pl.rows;
  ea.rd_off        = pl.rd_off;
  ea.rd_src        = pl.rd_src;
  ea.tile_row      = pl.tile_row;
  ea.tile_kind     = pl.tile_kind;
  ea.tile_lo       = t_lo;
  ea.n_tiles       = n;
  ea.records       = sc.records + uint64_t(off) * kTileCap;
  ea.tile_cnt      = sc.tile_cnt + off;
  ea.em_base       = sc.em_base + off;
  ea.edge_base     = sc.ed_base + off;
  ea.totals        = sc.totals + 4 * chunk;
  ea.ticket        = reinterpret_cast<unsigned int *>(sc.totals + 4 * chunk + 3);
  ea.prev_totals   = sc.totals;
  ea.n_prev        = chunk;
  ea.close_csr     = last ? 1u : 0u;
  ea.em_value_base = em_value_base;
  ea.out           = e;
  int grid         = MSGPU_TICKETS ? kNumSMs * (underneath ? 4 : kEmitCtas) : kNumSMs * 8; // one wave; the tiles behind it go by tickets
  if (uint32_t(grid) > n) grid = int(n);
  // (measured and not kept, round 2: the same emission with four consecutive rows per thread and 128-bit column
  //  stores — 0.75 ms instead of 0.49 ms at yeast scale, profiles/r2_ab_emit_vec_vs_scalar_and_fused.log: the kernel is
  //  bound by the latency of its gathers, and a thread that owns a quad serialises four of them behind one another
  //  where four strided records keep four independent chains in flight)
  // The product stages a tile's records with one bulk copy (pair_emit_rows_bulk_kernel: 0.482 ms against 0.506 ms at
  // yeast scale, profiles/r2_ab_emit_bulk_vs_scalar_yeast.log); MSGPU_EMIT=scalar keeps the per-thread loads as the A/B.
  const char *emit = std::getenv("MSGPU_EMIT");
  if (emit && std::string(emit) == "scalar") {
    MSGPU_LAUNCH("pair_emit_rows", st, launch_k(pair_emit_rows_kernel, grid, kEmitBlock, 0, st, ea));
    return;
  }
  const int bytes = int(kTileCap * sizeof(uint2));
  if (underneath) {
    MSGPU_CUDA(cudaFuncSetAttribute(pair_emit_rows_bulk_kernel<4, kEmitBatchUnder>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    MSGPU_LAUNCH("pair_emit_rows", st, launch_k(pair_emit_rows_bulk_kernel<4, kEmitBatchUnder>, grid, kEmitBlock, bytes, st, ea));
  } else {
    MSGPU_CUDA(cudaFuncSetAttribute(pair_emit_rows_bulk_kernel<kEmitCtas, kEmitBatch>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    MSGPU_LAUNCH("pair_emit_rows", st, launch_k(pair_emit_rows_bulk_kernel<kEmitCtas, kEmitBatch>, grid, kEmitBlock, bytes, st, ea));
  }
}
void pairing_rows(Arena &arena, cudaStream_t st, const PairPlan &pl, uint32_t t_lo, uint32_t t_hi, PairScratch &sc,
                  EdgeTables &e, uint64_t em_value_base) {
  (void)arena;
  const uint32_t n = t_hi - t_lo;
  e.em_value_base  = em_value_base;
  if (!n) {
    MSGPU_CUDA(cudaMemcpyAsync(e.em_begin, &e.em_value_base, sizeof(uint64_t), cudaMemcpyHostToDevice, st)); // empty CSR
    return;
  }
  rows_launch(st, pl, t_lo, t_hi, sc, 0, e, em_value_base, 0, true, false);
}

// The whole emission of tiles [t_lo, t_hi) into ONE resident table, software-pipelined over `n_chunks` ranges of tiles
// on two streams: while the tile kernel (bound by instruction issue, 5 % of the DRAM bandwidth) sorts chunk c + 1 on
// `st`, the offsets scan and the emission kernel (bound by DRAM writes and gather latency, a third of the issue slots)
// of chunk c run on the arena's high-priority side stream.  The tile kernel's CTAs of a chunk live for about one tile,
// and an emission CTA compiled for 64 registers fits exactly into what one of them leaves behind (16 384 registers,
// 51 KB of shared memory), so the two kernels share every SM instead of taking turns.  The rows of chunk c start where
// those of the chunks before it end: the emission kernel adds up their totals itself (<= 16 words), nothing returns
// to the host in between.
static void pairing_pipelined(Arena &arena, cudaStream_t st, const PairPlan &pl, uint32_t t_lo, uint32_t t_hi, PairScratch &sc,
                              EdgeTables &e, uint32_t n_chunks) {
  const uint32_t n = t_hi - t_lo;
  e.em_value_base  = 0;
  sc.n_chunks      = n_chunks;
  MSGPU_CUDA(cudaMemsetAsync(sc.totals, 0, 4 * n_chunks * sizeof(unsigned long long), st));
  cudaStream_t side = arena.side_stream();
  for (uint32_t c = 0; c < n_chunks; ++c) {
    const uint32_t lo = t_lo + uint32_t(uint64_t(n) * c / n_chunks), hi = t_lo + uint32_t(uint64_t(n) * (c + 1) / n_chunks);
    if (hi == lo) continue;
    tiles_launch(st, pl, lo, hi, sc, lo - t_lo, sc.totals + 4 * c);
    cudaEvent_t sorted = arena.sync_event(c);
    MSGPU_CUDA(cudaEventRecord(sorted, st));
    MSGPU_CUDA(cudaStreamWaitEvent(side, sorted, 0));
    tiles_scans(arena, side, hi - lo, sc, lo - t_lo, sc.totals + 4 * c);
    rows_launch(side, pl, lo, hi, sc, lo - t_lo, e, 0, c, c + 1 == n_chunks, c + 1 < n_chunks);
  }
  cudaEvent_t done = arena.sync_event(n_chunks);
  MSGPU_CUDA(cudaEventRecord(done, side));
  MSGPU_CUDA(cudaStreamWaitEvent(st, done, 0));
}

// reads the row counts of the last emission and the error flags of the plan (one round trip); throws on the
// shapes the kernels refuse
void pairing_finish(Arena &arena, cudaStream_t st, const PairPlan &pl, const PairScratch &sc, uint64_t *n_em,
                    uint64_t *n_edges) {
  char *h = static_cast<char *>(arena.host_scratch());
  MSGPU_CUDA(cudaMemcpyAsync(h, pl.flags, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  const uint32_t nc = sc.n_chunks ? sc.n_chunks : 1;
  MSGPU_CUDA(cudaMemcpyAsync(h + 64, sc.totals, 4 * nc * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  const uint32_t           *h_flags = reinterpret_cast<const uint32_t *>(h);
  unsigned long long       *h_tot   = reinterpret_cast<unsigned long long *>(h + 64);
  for (uint32_t c = 1; c < nc; ++c) { // a pipelined emission keeps one pair of totals per chunk
    h_tot[0] += h_tot[4 * c];
    h_tot[1] += h_tot[4 * c + 1];
  }
  if (h_flags[2])
    throw BuildError{MSGPU_ERR_LIMIT, "a read shares " + std::to_string(kSrcCapWide) + " or more unitigs with other reads (one "
                                      "edge of it could hold more edge matches than a pairing tile holds candidates)"};
  if (h_flags[3])
    throw BuildError{MSGPU_ERR_INTERNAL, "pairing: tile invariant violated (code " + std::to_string(h_flags[3]) + ")"};
  *n_em    = h_tot[0];
  *n_edges = h_tot[1];
}

void run_pairing_fused(Arena &arena, cudaStream_t st, const MatchSoA &m, const UnitigTable &unitigs, uint64_t n_reads,
                       EdgeTables &e, cudaEvent_t ev_pair_done, uint32_t part, uint32_t n_parts, PairPlan *plan_out,
                       bool no_overlap_test) {
  e = EdgeTables{};
  PairPlan pl;
  pairing_plan(arena, st, m, unitigs, n_reads, part, n_parts, pl, no_overlap_test);
  e.n_pairs = pl.n_pairs;
  const uint32_t my_tiles = pl.tile_hi - pl.tile_lo;
  pairing_outputs_alloc(arena, pairing_rows_bound(pl, my_tiles), e);
  PairScratch sc;
  pairing_scratch_alloc(arena, my_tiles, sc);
  if (ev_pair_done) MSGPU_CUDA(cudaEventRecord(ev_pair_done, st)); // planning done; what follows is "merge" time
  // emission pipelined chunk by chunk underneath the tile kernel (pairing_pipelined):
  // (measured, round 2, and OFF by default: at yeast scale 1.70 ms per construction with 0, 2, 4 or 6 chunks, 1.76 / 1.81 /
  //  1.89 ms with 8 / 12 / 16, profiles/r2_ab_pipelined_emission_yeast.log — the two kernels do share the SMs, and each
  //  loses what the other gains: the tile kernel needs its four resident CTAs to hide its barriers)
  uint32_t n_chunks = 0;
  if (const char *pc = std::getenv("MSGPU_PIPE_CHUNKS")) n_chunks = uint32_t(std::atoi(pc)); // development A/B
  if (n_chunks > kMaxPipeChunks) n_chunks = kMaxPipeChunks;
  if (n_chunks > my_tiles / 64) n_chunks = my_tiles / 64; // (no empty chunks)
  if (n_chunks >= 2) {
    pairing_pipelined(arena, st, pl, pl.tile_lo, pl.tile_hi, sc, e, n_chunks);
  } else {
    pairing_tiles(arena, st, pl, pl.tile_lo, pl.tile_hi, sc);
    pairing_rows(arena, st, pl, pl.tile_lo, pl.tile_hi, sc, e, 0);
  }
  pairing_finish(arena, st, pl, sc, &e.n_em, &e.n_edges);
  e.pair_rows = pl.rows; // kept (in the arena) for the stage that follows: msgpu_chain_overlaps
  e.p_read    = pl.p_read;
  if (plan_out) *plan_out = pl;
}

void run_pairing(Arena &arena, cudaStream_t st, const MatchSoA &m, const UnitigTable &unitigs, uint64_t n_reads,
                 EdgeTables &e, cudaEvent_t ev_pair_done, uint32_t part, uint32_t n_parts) {
  const char *v = std::getenv("MSGPU_VARIANT");
  if (v && std::string(v) == "sorted" && n_parts <= 1)
    run_pairing_sorted(arena, st, m, unitigs, n_reads, e, ev_pair_done);
  else
    run_pairing_fused(arena, st, m, unitigs, n_reads, e, ev_pair_done, part, n_parts, nullptr, false);
}

} // namespace msgpu
