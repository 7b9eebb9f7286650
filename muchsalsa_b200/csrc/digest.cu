// digest.cu — an order-independent checksum of the edge / edge-match tables plus their structural invariants, computed
// where the tables are: in HBM.
//
// The reference has no counterpart (it never serialises its Graph / MatchMap); this is the "checksum of checksums"
// with which results too large to download or to rebuild on the CPU are compared:
//   * the slices of a multi-GPU build: the sums of the ranks' digests equal the digest of the one-GPU build and the
//     digest of the reference's tables (tests/golden/digests.json), and the slices' key ranges must not overlap;
//   * the chunks of a chunked build (results larger than HBM): the digests of the chunks add up the same way.
// Every edge match contributes a hash of (first, second, unitig, ov0, ov1, score bits, line, flags) — the fields
// MatchMap::addEdgeMatch stores (MatchMap.h:68-74) under the key Graph::addEdge gave the pair (MatchMap.cpp:204-218) —
// and every edge a hash of (first, second, number of edge matches).  Sums are mod 2^64, so they commute.
#include "../../include/msgpu.h"
#include "internal.h"

namespace msgpu {

// the hash of include/msgpu.h ("msgpu_digest"); the tests mirror it in numpy on the checker's tables
__device__ __forceinline__ uint64_t digest_em(uint32_t first, uint32_t second, uint32_t unitig, int32_t ov0, int32_t ov1,
                                              uint64_t score_bits, uint32_t line, uint8_t flags) {
  uint64_t h = mix64(uint64_t(flags) + 0x9e3779b97f4a7c15ull);
  h          = mix64(h ^ score_bits);
  h          = mix64(h ^ (uint64_t(uint32_t(ov1)) << 32 | line));
  h          = mix64(h ^ (uint64_t(unitig) << 32 | uint32_t(ov0)));
  return mix64(h ^ (uint64_t(first) << 32 | second));
}
__device__ __forceinline__ uint64_t digest_edge(uint32_t first, uint32_t second, uint64_t n_em) {
  return mix64(mix64(n_em + 0xd1b54a32d192ed03ull) ^ (uint64_t(first) << 32 | second));
}

struct DigestAcc { // device-resident accumulator; zeroed by the caller
  unsigned long long edge_sum, em_sum, violations;
};

__global__ void __launch_bounds__(256) digest_kernel(EdgeTables e, DigestAcc *acc) {
  pdl_sync();
  unsigned long long es = 0, ms = 0, bad = 0;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < e.n_edges; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t f = e.first[i], s = e.second[i];
    const uint64_t b0 = e.em_begin[i] - e.em_value_base, b1 = e.em_begin[i + 1] - e.em_value_base; // column indices
    if (!(f < s)) ++bad;                                  // Edge::getVertices(): first = the smaller line index = id
    if (!(b1 > b0) || b1 > e.n_em) ++bad;                 // an edge exists iff it has >= 1 edge match
    if (i > 0) {                                          // strictly increasing (first, second): every pair once
      const uint64_t prev = uint64_t(e.first[i - 1]) << 32 | e.second[i - 1], cur = uint64_t(f) << 32 | s;
      if (!(cur > prev)) ++bad;
    } else if (b0 != 0) {
      ++bad;
    }
    if (i + 1 == e.n_edges && b1 != e.n_em) ++bad;
    es += digest_edge(f, s, b1 - b0);
    if (b1 > e.n_em) continue;
    for (uint64_t k = b0; k < b1; ++k) {
      const uint32_t u   = e.em_unitig[k];
      const int32_t  ov0 = e.em_ov0[k], ov1 = e.em_ov1[k];
      if (k > b0 && !(u > e.em_unitig[k - 1])) ++bad;     // one edge match per (edge, unitig), in unitig order
      if (!(ov0 <= ov1 && ov1 - ov0 > 100)) ++bad;        // MatchMap.cpp:192
      ms += digest_em(f, s, u, ov0, ov1, (uint64_t)__double_as_longlong(e.em_score[k]), e.em_line[k], e.em_flags[k]);
    }
  }
  es  = warp_reduce_sum(es);
  ms  = warp_reduce_sum(ms);
  bad = warp_reduce_sum(bad);
  if (lane_id() == 0) {
    if (es) atomicAdd(&acc->edge_sum, es);
    if (ms) atomicAdd(&acc->em_sum, ms);
    if (bad) atomicAdd(&acc->violations, bad);
  }
}

void run_digest(Arena &arena, cudaStream_t st, const EdgeTables &e, msgpu_digest *out) {
  DigestAcc *d_acc = arena.alloc<DigestAcc>(1);
  MSGPU_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(DigestAcc), st));
  if (e.n_edges) MSGPU_LAUNCH("digest", st, launch_k(digest_kernel, grid_for(e.n_edges, 256, 16), 256, 0, st, e, d_acc));
  struct Tail {
    DigestAcc acc;
    uint32_t  f0, s0, f1, s1;
  };
  static_assert(sizeof(Tail) <= 256, "host scratch");
  Tail *h = static_cast<Tail *>(arena.host_scratch());
  MSGPU_CUDA(cudaMemcpyAsync(&h->acc, d_acc, sizeof(DigestAcc), cudaMemcpyDeviceToHost, st));
  if (e.n_edges) {
    MSGPU_CUDA(cudaMemcpyAsync(&h->f0, e.first, 4, cudaMemcpyDeviceToHost, st));
    MSGPU_CUDA(cudaMemcpyAsync(&h->s0, e.second, 4, cudaMemcpyDeviceToHost, st));
    MSGPU_CUDA(cudaMemcpyAsync(&h->f1, e.first + e.n_edges - 1, 4, cudaMemcpyDeviceToHost, st));
    MSGPU_CUDA(cudaMemcpyAsync(&h->s1, e.second + e.n_edges - 1, 4, cudaMemcpyDeviceToHost, st));
  }
  MSGPU_CUDA(cudaStreamSynchronize(st));
  out->n_edges        = e.n_edges;
  out->n_edge_matches = e.n_em;
  out->edge_sum       = h->acc.edge_sum;
  out->em_sum         = h->acc.em_sum;
  out->violations     = h->acc.violations;
  out->first_key      = e.n_edges ? (uint64_t(h->f0) << 32 | h->s0) : ~0ull;
  out->last_key       = e.n_edges ? (uint64_t(h->f1) << 32 | h->s1) : 0ull;
}

} // namespace msgpu
