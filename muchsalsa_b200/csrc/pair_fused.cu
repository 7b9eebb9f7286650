// pair_fused.cu — candidate-pair generation AND keyed merge in ONE persistent kernel (K6 + K7, the fused variant).
//
// Replaces MatchMap::calculateEdges / processScaffold (MatchMap.cpp:161-224) + Graph::addEdge (Graph.cpp:212-230,
// 291-311) + MatchMap::addEdgeMatch (MatchMap.cpp:109-134).
//
// The sort-based variant (pair_sort.cu) writes every candidate to HBM in unitig order and then moves it through
// 4-6 radix passes; at yeast scale that is ~250 B of HBM traffic per edge match for 25 B of result.  Here the work is
// turned around so that the output is produced directly in its final order and touched once:
//
//   * the reduction key is (first, second) with first < second.  All candidates of ONE read r as `first` come from
//     the unitig groups r belongs to; so a CTA that owns a run of consecutive reads (a "tile") can enumerate, in
//     shared memory, every candidate those reads own: for each match a of r (its unitig u), every other member b of
//     u's group with read(b) > r and unitig overlap > 100 (MatchMap.cpp:189-192).
//   * the tile's candidates (<= 4096) are sorted in shared memory by (read, second) with a stable LSD radix sort over
//     only the significant key bits; they were generated in unitig order, so rows of one edge come out by unitig.
//   * tiles are handed out in increasing read order by an atomic ticket, and each tile obtains the global offset of
//     its first edge match and of its first edge with a decoupled look-back over per-tile status words (aggregate
//     published as soon as it is known, inclusive prefix after the look-back).  The kernel therefore writes the
//     edge-match columns, the edge table and its CSR offsets exactly once, coalesced, already in (first, second,
//     unitig) order: algorithmic traffic = every match record read once + 25 B per edge match + 16 B per edge.
//
// A small pre-pass builds what the tiles need: per read the list of its matches (unitig order), and tile boundaries
// such that a tile never holds more than kTileCap candidates / kSrcCap matches.  Reads too heavy for one tile
// (thousands of partners: high-multiplicity repeats) make the caller fall back to the sort-based variant.
#include <cstdlib>
#include <string>

#include "../../include/msgpu.h"
#include "internal.h"
#include "scan.cuh"

namespace msgpu {

constexpr int kThOverlapF = 100; // MatchMap.cpp:41

constexpr int      kFusedBlock = 256;
constexpr int      kFusedWarps = kFusedBlock / 32;
constexpr int      kTileCap    = 4096; // candidates per tile
constexpr int      kSrcCap     = 1024; // matches ("sources") per tile
constexpr int      kSrcRatio   = (kTileCap / 2) / (kSrcCap / 2);
constexpr uint32_t kMinWeight  = (kTileCap / 2) / 128; // caps a tile at 256 reads
constexpr uint32_t kHalfCap    = kTileCap / 2;

// ---------------------------------------------------------------------------------------------------------------------
// pre-pass
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    read_stats_kernel(const uint32_t *__restrict__ unitig, const uint32_t *__restrict__ read,
                      const uint32_t *__restrict__ grp_off, uint64_t n, uint32_t *__restrict__ deg,
                      unsigned long long *__restrict__ ub) {
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t u = unitig[i];
    uint32_t k = grp_off[u + 1] - grp_off[u];
    if (k >= 2) { // a match alone in its group has no partner (MatchMap.cpp:185: the outer loop starts at the 2nd)
      atomicAdd(&deg[read[i]], 1u);
      atomicAdd(&ub[read[i]], (unsigned long long)(k - 1));
    }
  }
}

__global__ void __launch_bounds__(256)
    scatter_sources_kernel(const uint32_t *__restrict__ unitig, const uint32_t *__restrict__ read,
                           const uint32_t *__restrict__ grp_off, uint64_t n, const uint32_t *__restrict__ rd_off,
                           uint32_t *__restrict__ cursor, uint32_t *__restrict__ rd_tmp) {
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t u = unitig[i];
    if (grp_off[u + 1] - grp_off[u] >= 2) {
      uint32_t r               = read[i];
      rd_tmp[rd_off[r] + atomicAdd(&cursor[r], 1u)] = uint32_t(i);
    }
  }
}

// the atomics above leave a read's list in arbitrary order: rank-sort it (lists are short) so that it is in match
// order = unitig order, which makes the candidate stream of a tile deterministic
__global__ void __launch_bounds__(256)
    order_sources_kernel(const uint32_t *__restrict__ read, const uint32_t *__restrict__ rd_off,
                         const uint32_t *__restrict__ rd_tmp, uint64_t n_src, uint32_t *__restrict__ rd_src) {
  for (uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < n_src;
       q += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t me = rd_tmp[q];
    uint32_t r  = read[me];
    uint32_t lo = rd_off[r], hi = rd_off[r + 1], rank = 0;
    for (uint32_t p = lo; p < hi; ++p) rank += rd_tmp[p] < me;
    rd_src[lo + rank] = me;
  }
}

struct TileWeight { // weight of a read when tiles are formed; clipped so that any run of reads < one tile
  const uint32_t           *deg;
  const unsigned long long *ub;
  __device__ uint64_t raw(uint64_t r) const {
    uint64_t w = ub[r], d = uint64_t(deg[r]) * kSrcRatio;
    w          = w > d ? w : d;
    return w > kMinWeight ? w : kMinWeight;
  }
  __device__ uint64_t operator()(uint64_t r) const {
    uint64_t w = raw(r);
    return w > kHalfCap ? kHalfCap : w;
  }
};

struct TileStart { // 1 where a tile begins: quotient of the weight prefix changes, or a heavy read is involved
  TileWeight      w;
  const uint64_t *prefix;
  __device__ uint32_t operator()(uint64_t r) const {
    if (r == 0) return 1u;
    if (prefix[r] / kHalfCap != prefix[r - 1] / kHalfCap) return 1u;
    return (w.raw(r) > kHalfCap || w.raw(r - 1) > kHalfCap) ? 1u : 0u;
  }
};

__global__ void __launch_bounds__(256)
    write_tiles_kernel(TileStart start, const uint32_t *__restrict__ tile_rank, uint64_t n_reads,
                       const uint32_t *n_tiles, uint32_t *__restrict__ tile_first, uint32_t *__restrict__ n_heavy) {
  for (uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n_reads;
       r += uint64_t(gridDim.x) * blockDim.x) {
    if (start(r)) tile_first[tile_rank[r]] = uint32_t(r);
    if (start.w.raw(r) > kHalfCap) atomicAdd(n_heavy, 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) tile_first[*n_tiles] = uint32_t(n_reads);
}

// ---------------------------------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------------------------------
constexpr unsigned long long kStAggregate = 1ull << 62, kStPrefix = 2ull << 62, kStMask = (1ull << 62) - 1;

struct FusedArgs {
  MatchSoA        m;
  const uint32_t *grp_off;
  const uint32_t *rd_off, *rd_src, *tile_first;
  uint32_t        n_tiles;
  unsigned int   *ticket;
  volatile unsigned long long *st_em, *st_edge; // per tile: flag (2 bits) | value
  unsigned long long *totals;                   // [0] = edge matches, [1] = edges, [2] = error flag
  EdgeTables      out;
};

struct FusedSmem {
  uint32_t keyA[kTileCap], keyB[kTileCap];
  uint32_t ent_ib[kTileCap];
  uint16_t idxA[kTileCap], idxB[kTileCap];
  uint16_t ent_j[kTileCap];
  uint32_t src_ia[kSrcCap], src_base[kSrcCap];
  uint32_t warp_cnt[kFusedWarps][256];
  uint32_t bin_base[256], chunk_base[256];
  uint32_t scan[33];
  uint32_t tile, n_ent, n_heads, smin, smax;
  unsigned long long em_base, edge_base;
};

__device__ __forceinline__ bool overlaps(int a0, int a1, int b0, int b1) {
  int o0 = a0 > b0 ? a0 : b0, o1 = a1 < b1 ? a1 : b1;
  return o0 <= o1 && o1 - o0 > kThOverlapF; // MatchMap.cpp:192
}

// stable LSD radix sort of n (key, idx) pairs in shared memory over key bits [0, nbits); returns true if the result
// is in the B buffers.  blockDim.x == 256.
__device__ bool cta_radix_sort(FusedSmem &s, uint32_t n, int nbits) {
  const unsigned t = threadIdx.x, w = t >> 5, l = t & 31;
  uint32_t      *kin = s.keyA, *kout = s.keyB;
  uint16_t      *iin = s.idxA, *iout = s.idxB;
  bool           in_b = false;
  for (int shift = 0; shift < nbits; shift += 8) {
    const unsigned mask = nbits - shift >= 8 ? 255u : ((1u << (nbits - shift)) - 1u);
    s.bin_base[t] = 0;
    __syncthreads();
    for (uint32_t e = t; e < n; e += kFusedBlock) atomicAdd(&s.bin_base[(kin[e] >> shift) & mask], 1u);
    __syncthreads();
    uint32_t total;
    uint32_t ex   = block_exclusive_scan<uint32_t>(s.bin_base[t], s.scan, &total);
    s.bin_base[t] = ex;
    for (uint32_t c0 = 0; c0 < n; c0 += kFusedBlock) {
#pragma unroll
      for (int ww = 0; ww < kFusedWarps; ++ww) s.warp_cnt[ww][t] = 0;
      __syncthreads();
      uint32_t e     = c0 + t;
      bool     valid = e < n;
      uint32_t key   = valid ? kin[e] : 0u;
      uint16_t idx   = valid ? iin[e] : uint16_t(0);
      unsigned d     = valid ? ((key >> shift) & mask) : 256u;
      unsigned peers = __match_any_sync(0xffffffffu, d);
      unsigned rank  = __popc(peers & ((1u << l) - 1u));
      if (valid && rank == 0) s.warp_cnt[w][d] = __popc(peers);
      __syncthreads();
      {
        uint32_t sum = 0;
#pragma unroll
        for (int ww = 0; ww < kFusedWarps; ++ww) {
          uint32_t c        = s.warp_cnt[ww][t];
          s.warp_cnt[ww][t] = sum;
          sum += c;
        }
        s.chunk_base[t] = s.bin_base[t];
        s.bin_base[t] += sum;
      }
      __syncthreads();
      if (valid) {
        uint32_t dst = s.chunk_base[d] + s.warp_cnt[w][d] + rank;
        kout[dst]    = key;
        iout[dst]    = idx;
      }
      __syncthreads();
    }
    uint32_t *tk = kin;
    kin          = kout;
    kout         = tk;
    uint16_t *ti = iin;
    iin          = iout;
    iout         = ti;
    in_b         = !in_b;
  }
  return in_b;
}

__device__ __forceinline__ void publish(volatile unsigned long long *st, uint32_t tile, unsigned long long flag,
                                        unsigned long long v) {
  st[tile] = flag | v;
}

// decoupled look-back: exclusive prefix of this tile's value; then the inclusive prefix is published
__device__ unsigned long long look_back(volatile unsigned long long *st, uint32_t tile, unsigned long long mine) {
  unsigned long long excl = 0;
  for (int64_t t = int64_t(tile) - 1; t >= 0; --t) {
    unsigned long long v;
    do {
      v = st[t];
    } while ((v >> 62) == 0);
    excl += v & kStMask;
    if ((v >> 62) == 2) break;
  }
  __threadfence();
  st[tile] = kStPrefix | (excl + mine);
  return excl;
}

__global__ void __launch_bounds__(kFusedBlock, 2) pair_fused_kernel(FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FusedSmem     &s = *reinterpret_cast<FusedSmem *>(smem_raw);
  const unsigned t = threadIdx.x, w = t >> 5, l = t & 31;
  const MatchSoA &m = a.m;

  for (;;) {
    __syncthreads(); // previous tile fully done with shared memory
    if (t == 0) {
      s.tile = atomicAdd(a.ticket, 1u);
      s.smin = 0xffffffffu;
      s.smax = 0;
    }
    __syncthreads();
    const uint32_t tile = s.tile;
    if (tile >= a.n_tiles) break;
    const uint32_t r0 = a.tile_first[tile], r1 = a.tile_first[tile + 1];
    const uint32_t q0 = a.rd_off[r0], n_src = a.rd_off[r1] - q0;
    // a tile that violates its capacity invariant (cannot happen: tile weights) reports the error, contributes
    // nothing, keeps the look-back chain alive and never touches memory out of bounds
    auto bail = [&](unsigned long long code) {
      if (t == 0) {
        a.totals[2] = code;
        publish(a.st_em, tile, tile == 0 ? kStPrefix : kStAggregate, 0);
        publish(a.st_edge, tile, tile == 0 ? kStPrefix : kStAggregate, 0);
        if (tile) {
          look_back(a.st_em, tile, 0);
          look_back(a.st_edge, tile, 0);
        }
      }
    };
    if (n_src > uint32_t(kSrcCap)) {
      bail(1);
      continue;
    }
    for (uint32_t q = t; q < n_src; q += kFusedBlock) s.src_ia[q] = a.rd_src[q0 + q];
    __syncthreads();

    // ---- phase 1: how many candidates does each source own?  (one warp per source, lanes over the group)
    for (uint32_t j = w; j < n_src; j += kFusedWarps) {
      uint32_t ia = s.src_ia[j];
      uint32_t u  = m.unitig[ia], r = m.read[ia];
      int      a0 = m.ir0[ia], a1 = m.ir1[ia];
      uint32_t g0 = a.grp_off[u], g1 = a.grp_off[u + 1], cnt = 0;
      for (uint32_t b = g0; b < g1; b += 32) {
        uint32_t ib   = b + l;
        bool     keep = ib < g1 && m.read[ib] > r && overlaps(a0, a1, m.ir0[ib], m.ir1[ib]);
        cnt += __popc(__ballot_sync(0xffffffffu, keep));
      }
      if (l == 0) s.src_base[j] = cnt;
    }
    __syncthreads();
    uint32_t carry = 0;
    for (uint32_t c0 = 0; c0 < n_src; c0 += kFusedBlock) { // exclusive scan of the per-source counts
      uint32_t j = c0 + t, v = j < n_src ? s.src_base[j] : 0u, total;
      uint32_t ex = block_exclusive_scan<uint32_t>(v, s.scan, &total);
      if (j < n_src) s.src_base[j] = carry + ex;
      carry += total;
    }
    const uint32_t n_ent = carry;
    if (n_ent > uint32_t(kTileCap)) {
      bail(2);
      continue;
    }
    if (t == 0) publish(a.st_em, tile, tile == 0 ? kStPrefix : kStAggregate, n_ent);
    __syncthreads();

    // ---- phase 2: materialise the candidate list (second read, partner row, source) in generation order
    {
      uint32_t lmin = 0xffffffffu, lmax = 0;
      for (uint32_t j = w; j < n_src; j += kFusedWarps) {
        uint32_t ia = s.src_ia[j];
        uint32_t u  = m.unitig[ia], r = m.read[ia];
        int      a0 = m.ir0[ia], a1 = m.ir1[ia];
        uint32_t g0 = a.grp_off[u], g1 = a.grp_off[u + 1], base = s.src_base[j];
        for (uint32_t b = g0; b < g1; b += 32) {
          uint32_t ib   = b + l;
          uint32_t sr   = ib < g1 ? m.read[ib] : 0u;
          bool     keep = ib < g1 && sr > r && overlaps(a0, a1, m.ir0[ib], m.ir1[ib]);
          unsigned mk   = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            uint32_t e  = base + __popc(mk & ((1u << l) - 1u));
            s.keyA[e]   = sr;
            s.ent_ib[e] = ib;
            s.ent_j[e]  = uint16_t(j);
            lmin        = sr < lmin ? sr : lmin;
            lmax        = sr > lmax ? sr : lmax;
          }
          base += __popc(mk);
        }
      }
      for (int d = 16; d > 0; d >>= 1) {
        uint32_t omin = __shfl_xor_sync(0xffffffffu, lmin, d), omax = __shfl_xor_sync(0xffffffffu, lmax, d);
        lmin = omin < lmin ? omin : lmin;
        lmax = omax > lmax ? omax : lmax;
      }
      if (l == 0 && lmin <= lmax) {
        atomicMin(&s.smin, lmin);
        atomicMax(&s.smax, lmax);
      }
    }
    __syncthreads();

    // ---- sort by (read, second); generation order (= unitig order) breaks ties because the sort is stable
    bool     in_b = false;
    int      sb   = 0;
    uint32_t smin = s.smin;
    if (n_ent) {
      uint32_t range = s.smax - smin;
      sb             = range ? 32 - __clz(range) : 0;
      int rb         = (r1 - r0 > 1) ? 32 - __clz(r1 - r0 - 1) : 0;
      if (sb + rb > 32) { // > 2^24 reads with a 256-read tile: outside the supported shape
        if (t == 0) a.totals[2] = 3;
        sb = 32 - rb;
      }
      for (uint32_t e = t; e < n_ent; e += kFusedBlock) {
        uint32_t rl = m.read[s.src_ia[s.ent_j[e]]] - r0;
        s.keyA[e]   = (sb < 32 ? (rl << sb) : 0u) | (s.keyA[e] - smin);
        s.idxA[e]   = uint16_t(e);
      }
      __syncthreads();
      in_b = cta_radix_sort(s, n_ent, sb + rb);
    }
    uint32_t *key  = in_b ? s.keyB : s.keyA;
    uint16_t *idx  = in_b ? s.idxB : s.idxA;
    uint32_t *rank = in_b ? s.keyA : s.keyB; // the free key buffer: edge rank of every run head

    // ---- run heads = edges of this tile
    carry = 0;
    for (uint32_t c0 = 0; c0 < n_ent; c0 += kFusedBlock) {
      uint32_t k    = c0 + t;
      uint32_t head = (k < n_ent && (k == 0 || key[k] != key[k - 1])) ? 1u : 0u, total;
      uint32_t ex   = block_exclusive_scan<uint32_t>(head, s.scan, &total);
      if (k < n_ent) rank[k] = head ? carry + ex : 0xffffffffu;
      carry += total;
    }
    const uint32_t n_heads = carry;
    if (t == 0) publish(a.st_edge, tile, tile == 0 ? kStPrefix : kStAggregate, n_heads);
    if (t == 0) s.em_base = tile ? look_back(a.st_em, tile, n_ent) : 0ull;
    if (t == 32) s.edge_base = tile ? look_back(a.st_edge, tile, n_heads) : 0ull;
    __syncthreads();
    const unsigned long long em_base = s.em_base, edge_base = s.edge_base;

    // ---- materialise and write each row once, in final order
    for (uint32_t k = t; k < n_ent; k += kFusedBlock) {
      uint32_t e  = idx[k];
      uint32_t ib = s.ent_ib[e], ia = s.src_ia[s.ent_j[e]];
      int      a0 = m.ir0[ia], a1 = m.ir1[ia], b0 = m.ir0[ib], b1 = m.ir1[ib];
      int      ov0 = a0 > b0 ? a0 : b0, ov1 = a1 < b1 ? a1 : b1;
      // MatchMap.cpp:195-202: score_x * common / len_x, summed; the sum is commutative so outer/inner need not be told
      // apart for the score.  No FMA contraction (explicit _rn intrinsics).
      double  common = double(ov1 - ov0 + 1);
      double  sa     = double((unsigned long long)(long long)m.score[ia]);
      double  sbb    = double((unsigned long long)(long long)m.score[ib]);
      double  sc_a   = __ddiv_rn(__dmul_rn(sa, common), double(a1 - a0 + 1));
      double  sc_b   = __ddiv_rn(__dmul_rn(sbb, common), double(b1 - b0 + 1));
      uint8_t fa = m.flags[ia], fb = m.flags[ib];
      // the outer match is the one on the later line = the later row of the group (rows are in line order)
      double             score = ia > ib ? __dadd_rn(sc_a, sc_b) : __dadd_rn(sc_b, sc_a);
      unsigned long long o     = em_base + k;
      a.out.em_unitig[o]       = m.unitig[ia];
      a.out.em_ov0[o]          = ov0;
      a.out.em_ov1[o]          = ov1;
      a.out.em_score[o]        = score;
      a.out.em_line[o]         = m.line[ia > ib ? ia : ib]; // MatchMap.cpp:218
      a.out.em_flags[o]        = (((fa ^ fb) & 1) ? 0 : 1) | ((fa & fb & 2) ? 2 : 0);
      uint32_t er              = rank[k];
      if (er != 0xffffffffu) {
        unsigned long long eo = edge_base + er;
        uint32_t           kk = key[k];
        a.out.first[eo]       = r0 + (sb < 32 ? (kk >> sb) : 0u);
        a.out.second[eo]      = smin + (sb < 32 ? (kk & ((1u << sb) - 1u)) : kk);
        a.out.em_begin[eo]    = o;
      }
    }
    if (tile == a.n_tiles - 1 && t == 0) { // the last tile closes the CSR and reports the totals
      a.out.em_begin[edge_base + n_heads] = em_base + n_ent;
      a.totals[0]                         = em_base + n_ent;
      a.totals[1]                         = edge_base + n_heads;
    }
  }
}

template <class T> static T read_back(cudaStream_t st, const T *d) {
  T h;
  MSGPU_CUDA(cudaMemcpyAsync(&h, d, sizeof(T), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  return h;
}

bool run_pairing_fused(Arena &arena, cudaStream_t st, const MatchSoA &m, const UnitigTable &unitigs,
                       uint64_t n_reads, EdgeTables &e, cudaEvent_t ev_pair_done) {
  e = EdgeTables{};
  const uint64_t M = m.n, R = n_reads;
  if (R >= (1ull << 24)) return false; // key packing (8 + 24 bits); the sort-based variant has no such limit

  // ---- pre-pass: per-read source lists
  uint32_t           *deg    = arena.alloc<uint32_t>(R + 1);
  unsigned long long *ub     = arena.alloc<unsigned long long>(R + 1);
  uint32_t           *cursor = arena.alloc<uint32_t>(R + 1);
  uint32_t           *rd_off = arena.alloc<uint32_t>(R + 1);
  uint32_t           *d_u32  = arena.alloc<uint32_t>(4);
  uint64_t           *d_u64  = arena.alloc<uint64_t>(2);
  MSGPU_CUDA(cudaMemsetAsync(deg, 0, (R + 1) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(ub, 0, (R + 1) * sizeof(unsigned long long), st));
  MSGPU_CUDA(cudaMemsetAsync(cursor, 0, (R + 1) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(d_u32, 0, 4 * sizeof(uint32_t), st));
  const int gm = grid_for(M, 256), gr = grid_for(R, 256);
  if (M) MSGPU_LAUNCH("read_stats", st, read_stats_kernel<<<gm, 256, 0, st>>>(m.unitig, m.read, unitigs.grp_off, M, deg, ub));
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{deg}, R, rd_off, d_u32 + 0);
  MSGPU_CUDA(cudaMemcpyAsync(rd_off + R, d_u32 + 0, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  // weights -> tile boundaries
  TileWeight wgt{deg, ub};
  uint64_t  *prefix = arena.alloc<uint64_t>(R + 1);
  exclusive_scan<uint64_t>(arena, st, wgt, R, prefix, d_u64 + 0);
  uint32_t *tile_rank = arena.alloc<uint32_t>(R + 1);
  TileStart start{wgt, prefix};
  exclusive_scan<uint32_t>(arena, st, start, R, tile_rank, d_u32 + 1);
  // sum of the per-read partner counts = 2 x the candidate pairs (each unordered pair is seen from both reads)
  exclusive_scan<uint64_t>(arena, st, LoadAs<unsigned long long, uint64_t>{ub}, R, (uint64_t *)nullptr, d_u64 + 1);
  uint32_t n_src   = read_back(st, d_u32 + 0);
  uint32_t n_tiles = read_back(st, d_u32 + 1);
  uint64_t P       = read_back(st, d_u64 + 1) / 2;
  e.n_pairs        = P;
  uint32_t *tile_first = arena.alloc<uint32_t>(uint64_t(n_tiles) + 1);
  if (R) MSGPU_LAUNCH("write_tiles", st, write_tiles_kernel<<<gr, 256, 0, st>>>(start, tile_rank, R, d_u32 + 1, tile_first, d_u32 + 2));
  if (read_back(st, d_u32 + 2) != 0) return false; // heavy reads: let the sort-based variant handle this input
  uint32_t *rd_tmp = arena.alloc<uint32_t>(n_src), *rd_src = arena.alloc<uint32_t>(n_src);
  if (M) MSGPU_LAUNCH("scatter_sources", st, scatter_sources_kernel<<<gm, 256, 0, st>>>(m.unitig, m.read, unitigs.grp_off, M, rd_off, cursor, rd_tmp));
  if (n_src) MSGPU_LAUNCH("order_sources", st, order_sources_kernel<<<grid_for(n_src, 256), 256, 0, st>>>(m.read, rd_off, rd_tmp, n_src, rd_src));

  // ---- outputs: E_m <= P and E <= E_m, so P rows of capacity are always enough
  e.em_unitig = arena.alloc<uint32_t>(P);
  e.em_ov0    = arena.alloc<int32_t>(P);
  e.em_ov1    = arena.alloc<int32_t>(P);
  e.em_score  = arena.alloc<double>(P);
  e.em_line   = arena.alloc<uint32_t>(P);
  e.em_flags  = arena.alloc<uint8_t>(P);
  e.first     = arena.alloc<uint32_t>(P);
  e.second    = arena.alloc<uint32_t>(P);
  e.em_begin  = arena.alloc<uint64_t>(P + 1);

  FusedArgs a;
  a.m          = m;
  a.grp_off    = unitigs.grp_off;
  a.rd_off     = rd_off;
  a.rd_src     = rd_src;
  a.tile_first = tile_first;
  a.n_tiles    = n_tiles;
  a.ticket     = arena.alloc<unsigned int>(1);
  a.st_em      = arena.alloc<unsigned long long>(uint64_t(n_tiles) + 1);
  a.st_edge    = arena.alloc<unsigned long long>(uint64_t(n_tiles) + 1);
  a.totals     = arena.alloc<unsigned long long>(4);
  a.out        = e;
  MSGPU_CUDA(cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), st));
  MSGPU_CUDA(cudaMemsetAsync((void *)a.st_em, 0, (uint64_t(n_tiles) + 1) * 8, st));
  MSGPU_CUDA(cudaMemsetAsync((void *)a.st_edge, 0, (uint64_t(n_tiles) + 1) * 8, st));
  MSGPU_CUDA(cudaMemsetAsync(a.totals, 0, 4 * 8, st));
  MSGPU_CUDA(cudaMemsetAsync(e.em_begin, 0, 8, st)); // em_begin[0] when there is no edge at all
  if (n_tiles) {
    static bool attr_set = false;
    if (!attr_set) {
      MSGPU_CUDA(cudaFuncSetAttribute(pair_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(sizeof(FusedSmem))));
      attr_set = true;
    }
    int grid = kNumSMs * 2;
    if (uint32_t(grid) > n_tiles) grid = int(n_tiles);
    MSGPU_LAUNCH("pair_fused", st, pair_fused_kernel<<<grid, kFusedBlock, sizeof(FusedSmem), st>>>(a));
  }
  if (ev_pair_done) MSGPU_CUDA(cudaEventRecord(ev_pair_done, st));
  unsigned long long totals[4];
  MSGPU_CUDA(cudaMemcpyAsync(totals, a.totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  if (totals[2] != 0)
    throw BuildError{MSGPU_ERR_INTERNAL, "fused pairing: tile invariant violated (code " + std::to_string(totals[2]) + ")"};
  e.n_em    = totals[0];
  e.n_edges = totals[1];
  return true;
}

} // namespace msgpu

namespace msgpu {

void run_pairing(Arena &arena, cudaStream_t st, const MatchSoA &m, const UnitigTable &unitigs, uint64_t n_reads,
                 EdgeTables &e, cudaEvent_t ev_pair_done) {
  const char *v = std::getenv("MSGPU_VARIANT");
  if (!(v && std::string(v) == "sorted") && run_pairing_fused(arena, st, m, unitigs, n_reads, e, ev_pair_done)) return;
  run_pairing_sorted(arena, st, m, unitigs, n_reads, e, ev_pair_done);
}

} // namespace msgpu
