// pair_sort.cu — candidate-pair generation per unitig + keyed reduction by a global radix sort (K6 + K7, the
// sort-based variant).
//
// Replaces MatchMap::calculateEdges / processScaffold (MatchMap.cpp:161-224) together with the merge it performs
// through Graph::addEdge (Graph.cpp:212-230,291-311) and MatchMap::addEdgeMatch (MatchMap.cpp:109-134):
//   K6  the k(k-1)/2 (outer, inner) pairs of every unitig group are flattened into ONE index space so that the work
//       is balanced no matter how skewed the group sizes are; a pair is kept iff its unitig overlap is > 100
//       (MatchMap.cpp:189-192).  Count pass -> scan -> emit pass, so the candidate stream is deterministic:
//       ordered by (unitig, outer line, inner line).
//   K7  the candidates are sorted (stable LSD radix, significant id bits only) by the 64-bit key
//       (first << 32 | second), first = the read with the smaller id = the vertex with the smaller lineIdx
//       (MatchMap.cpp:204-213; read ids are ranks of exactly those line indices).  Because the sort is stable and
//       the stream is in unitig order, rows of one edge end up ordered by unitig.  Run heads give the edge table.
//
// The score is the reference's expression, evaluated in IEEE-754 double with the same association and no FMA
// contraction:  double(score_o) * common / len_o + double(score_i) * common / len_i   (MatchMap.cpp:195-202).
#include "../../include/msgpu.h"
#include "internal.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace msgpu {

constexpr int kThOverlap = 100; // MatchMap.cpp:41

struct One {
  __device__ uint32_t operator()(uint64_t) const { return 1u; }
};

struct PairCount {
  const uint32_t *grp_off;
  __device__ uint64_t operator()(uint64_t u) const {
    uint64_t k = grp_off[u + 1] - grp_off[u];
    return k * (k - (k ? 1 : 0)) / 2;
  }
};

// pair index p (global) -> (outer row, inner row): group by binary search in the pair prefix, then the triangular
// index t = outer*(outer-1)/2 + inner with 0 <= inner < outer
__device__ __forceinline__ void locate_pair(uint64_t p, const uint64_t *__restrict__ pair_off, uint32_t n_unitigs,
                                            const uint32_t *__restrict__ grp_off, uint32_t *outer, uint32_t *inner,
                                            uint32_t *unitig) {
  uint32_t lo = 0, hi = n_unitigs; // find the last u with pair_off[u] <= p
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (pair_off[mid] <= p) lo = mid; else hi = mid;
  }
  uint64_t t = p - pair_off[lo];
  uint64_t o = uint64_t((1.0 + sqrt(1.0 + 8.0 * double(t))) * 0.5);
  while (o * (o - 1) / 2 > t) --o;
  while ((o + 1) * o / 2 <= t) ++o;
  uint32_t g0 = grp_off[lo];
  *outer      = g0 + uint32_t(o);
  *inner      = g0 + uint32_t(t - o * (o - 1) / 2);
  *unitig     = lo;
}

__device__ __forceinline__ bool pair_overlaps(const MatchSoA &m, uint32_t o, uint32_t i, int *ov0, int *ov1) {
  int a0 = m.ir0[o], a1 = m.ir1[o], b0 = m.ir0[i], b1 = m.ir1[i];
  *ov0 = a0 > b0 ? a0 : b0;
  *ov1 = a1 < b1 ? a1 : b1;
  return *ov0 <= *ov1 && *ov1 - *ov0 > kThOverlap; // MatchMap.cpp:192
}

constexpr int kPairBlock = 256;

// each block owns a contiguous slice of the pair index space; both passes walk it in chunks of one pair per thread
__global__ void __launch_bounds__(kPairBlock)
    pair_count_kernel(MatchSoA m, const uint64_t *__restrict__ pair_off, const uint32_t *__restrict__ grp_off,
                      uint32_t n_unitigs, uint64_t n_pairs, uint64_t per_block, uint32_t *__restrict__ block_cnt) {
  pdl_sync();
  __shared__ uint32_t sm[33];
  uint64_t            begin = uint64_t(blockIdx.x) * per_block;
  uint64_t            end   = begin + per_block < n_pairs ? begin + per_block : n_pairs;
  uint32_t            c     = 0;
  for (uint64_t p = begin + threadIdx.x; p < end; p += kPairBlock) {
    uint32_t o, i, u;
    int      ov0, ov1;
    locate_pair(p, pair_off, n_unitigs, grp_off, &o, &i, &u);
    c += pair_overlaps(m, o, i, &ov0, &ov1);
  }
  uint32_t total;
  block_exclusive_scan<uint32_t>(c, sm, &total);
  if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

struct Candidates { // candidate stream, in (unitig, outer, inner) order
  unsigned long long *key;
  uint32_t           *unitig;
  int32_t            *ov0, *ov1;
  double             *score;
  uint32_t           *line;
  uint8_t            *flags;
};

__global__ void __launch_bounds__(kPairBlock)
    pair_emit_kernel(MatchSoA m, const uint64_t *__restrict__ pair_off, const uint32_t *__restrict__ grp_off,
                     uint32_t n_unitigs, uint64_t n_pairs, uint64_t per_block, const uint32_t *__restrict__ block_off,
                     Candidates c) {
  pdl_sync();
  __shared__ uint32_t sm[33];
  uint64_t            begin = uint64_t(blockIdx.x) * per_block;
  uint64_t            end   = begin + per_block < n_pairs ? begin + per_block : n_pairs;
  uint32_t            base  = block_off[blockIdx.x];
  for (uint64_t p0 = begin; p0 < end; p0 += kPairBlock) {
    uint64_t p    = p0 + threadIdx.x;
    bool     keep = false;
    uint32_t o = 0, i = 0, u = 0;
    int      ov0 = 0, ov1 = 0;
    if (p < end) {
      locate_pair(p, pair_off, n_unitigs, grp_off, &o, &i, &u);
      keep = pair_overlaps(m, o, i, &ov0, &ov1);
    }
    uint32_t total;
    uint32_t pos = base + block_exclusive_scan<uint32_t>(keep ? 1u : 0u, sm, &total);
    base += total;
    if (keep) {
      uint8_t fo = m.flags[o], fi = m.flags[i];
      // MatchMap.cpp:193-202, evaluated exactly as written (no FMA contraction: explicit _rn intrinsics)
      double outer_len = double(m.ir1[o] - m.ir0[o] + 1);
      double inner_len = double(m.ir1[i] - m.ir0[i] + 1);
      double common    = double(ov1 - ov0 + 1);
      double so        = double((unsigned long long)(long long)m.score[o]); // size_t score widened from int
      double si        = double((unsigned long long)(long long)m.score[i]);
      double outer_sc  = __ddiv_rn(__dmul_rn(so, common), outer_len);
      double inner_sc  = __ddiv_rn(__dmul_rn(si, common), inner_len);
      uint32_t ro = m.read[o], ri = m.read[i];
      uint32_t first = ro < ri ? ro : ri, second = ro < ri ? ri : ro;
      c.key[pos]    = (unsigned long long)first << 32 | second;
      c.unitig[pos] = u;
      c.ov0[pos]    = ov0;
      c.ov1[pos]    = ov1;
      c.score[pos]  = __dadd_rn(outer_sc, inner_sc);
      c.line[pos]   = m.line[o];                                                // MatchMap.cpp:218
      c.flags[pos]  = (((fo ^ fi) & 1) ? 0 : 1) | ((fo & fi & 2) ? 2 : 0);      // dir: same strand; prim: both
    }
  }
}

// after the sort: one thread per row gathers its payload; run heads are flagged for the edge table
__global__ void __launch_bounds__(256)
    gather_sorted_kernel(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ idx, uint64_t n,
                         Candidates c, EdgeTables e, uint8_t *__restrict__ head) {
  pdl_sync();
  for (uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t s     = idx[r];
    e.em_unitig[r] = c.unitig[s];
    e.em_ov0[r]    = c.ov0[s];
    e.em_ov1[r]    = c.ov1[s];
    e.em_score[r]  = c.score[s];
    e.em_line[r]   = c.line[s];
    e.em_flags[r]  = c.flags[s];
    head[r]        = r == 0 || key[r] != key[r - 1];
  }
}

__global__ void __launch_bounds__(256)
    write_edges_kernel(const unsigned long long *__restrict__ key, const uint8_t *__restrict__ head,
                       const uint32_t *__restrict__ edge_rank, uint64_t n, EdgeTables e, const uint32_t *n_edges) {
  pdl_sync();
  for (uint64_t r = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += uint64_t(gridDim.x) * blockDim.x) {
    if (head[r]) {
      uint32_t k    = edge_rank[r];
      e.first[k]    = uint32_t(key[r] >> 32);
      e.second[k]   = uint32_t(key[r]);
      e.em_begin[k] = r;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) e.em_begin[*n_edges] = n;
}

template <class T> static T read_back(cudaStream_t st, const T *d) {
  T h;
  MSGPU_CUDA(cudaMemcpyAsync(&h, d, sizeof(T), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  return h;
}

void run_pairing_sorted(Arena &arena, cudaStream_t st, const MatchSoA &m, const UnitigTable &unitigs,
                        uint64_t n_reads, EdgeTables &e, cudaEvent_t ev_pair_done) {
  e = EdgeTables{};
  uint32_t U = uint32_t(unitigs.n);

  // ---- K6: flatten the pair space
  uint64_t *pair_off = arena.alloc<uint64_t>(uint64_t(U) + 1);
  uint64_t *d_total  = arena.alloc<uint64_t>(1);
  exclusive_scan<uint64_t>(arena, st, PairCount{unitigs.grp_off}, U, pair_off, d_total);
  MSGPU_CUDA(cudaMemcpyAsync(pair_off + U, d_total, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  uint64_t P = read_back(st, d_total);
  e.n_pairs  = P;
  if (P >= 0xffffffffull)
    throw BuildError{MSGPU_ERR_LIMIT, "sort-based pairing supports < 2^32 candidate pairs per device"};

  int      n_blocks  = grid_for(P, kPairBlock, 4);
  uint64_t per_block = (P + n_blocks - 1) / n_blocks;
  per_block          = (per_block + kPairBlock - 1) / kPairBlock * kPairBlock;
  uint32_t *block_cnt = arena.alloc<uint32_t>(n_blocks);
  uint32_t *d_nem     = arena.alloc<uint32_t>(2);
  if (P) {
    MSGPU_LAUNCH("pair_count", st, pair_count_kernel<<<n_blocks, kPairBlock, 0, st>>>(m, pair_off, unitigs.grp_off, U, P, per_block, block_cnt));
  } else {
    MSGPU_CUDA(cudaMemsetAsync(block_cnt, 0, n_blocks * sizeof(uint32_t), st));
  }
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{block_cnt}, n_blocks, block_cnt, d_nem);
  uint64_t n_em = read_back(st, d_nem);
  e.n_em        = n_em;

  Candidates c;
  c.key    = arena.alloc<unsigned long long>(n_em);
  c.unitig = arena.alloc<uint32_t>(n_em);
  c.ov0    = arena.alloc<int32_t>(n_em);
  c.ov1    = arena.alloc<int32_t>(n_em);
  c.score  = arena.alloc<double>(n_em);
  c.line   = arena.alloc<uint32_t>(n_em);
  c.flags  = arena.alloc<uint8_t>(n_em);
  if (P) {
    MSGPU_LAUNCH("pair_emit", st, pair_emit_kernel<<<n_blocks, kPairBlock, 0, st>>>(m, pair_off, unitigs.grp_off, U, P, per_block, block_cnt, c));
  }
  if (ev_pair_done) MSGPU_CUDA(cudaEventRecord(ev_pair_done, st));

  // ---- K7: sort by (first, second), gather, run-length encode
  unsigned long long *key2 = arena.alloc<unsigned long long>(n_em);
  uint32_t           *idx = arena.alloc<uint32_t>(n_em), *idx2 = arena.alloc<uint32_t>(n_em);
  int                 g   = grid_for(n_em, 256);
  unsigned long long *key = c.key;
  if (n_em) {
    exclusive_scan<uint32_t>(arena, st, One{}, n_em, idx, (uint32_t *)nullptr); // idx = 0, 1, 2, ...
    int rb = bits_for(n_reads ? n_reads : 1);
    radix_sort_pairs<unsigned long long, uint32_t>(arena, st, key, idx, key2, idx2, n_em, {0, rb, 32, 32 + rb});
  }
  e.em_unitig = arena.alloc<uint32_t>(n_em);
  e.em_ov0    = arena.alloc<int32_t>(n_em);
  e.em_ov1    = arena.alloc<int32_t>(n_em);
  e.em_score  = arena.alloc<double>(n_em);
  e.em_line   = arena.alloc<uint32_t>(n_em);
  e.em_flags  = arena.alloc<uint8_t>(n_em);
  uint8_t  *head      = arena.alloc<uint8_t>(n_em);
  uint32_t *edge_rank = arena.alloc<uint32_t>(n_em);
  if (n_em) {
    MSGPU_LAUNCH("gather_sorted", st, gather_sorted_kernel<<<g, 256, 0, st>>>(key, idx, n_em, c, e, head));
  }
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint8_t, uint32_t>{head}, n_em, edge_rank, d_nem + 1);
  e.n_edges  = read_back(st, d_nem + 1);
  e.first    = arena.alloc<uint32_t>(e.n_edges);
  e.second   = arena.alloc<uint32_t>(e.n_edges);
  e.em_begin = arena.alloc<uint64_t>(e.n_edges + 1);
  MSGPU_LAUNCH("write_edges", st, write_edges_kernel<<<g, 256, 0, st>>>(key, head, edge_rank, n_em, e, d_nem + 1));
}

} // namespace msgpu
