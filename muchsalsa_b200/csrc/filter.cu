// filter.cu — the per-edge passes that follow chaining while the edge / order tables are still in HBM
// (SURVEY.md §8 f3).
//
// Replaces the job bodies
//   findContractionEdges   src/main.cpp:416-460   (+ sanityCheck, libms/src/kernel/sc.cpp:29-88)
//   findDeletableEdges     src/main.cpp:534-550   (Edge::replaceOrders)
//   computeBitweight       src/main.cpp:552-573   (Edge::setWeight / setConsensusDirection)
// Each of them is a pure function of one edge, its orders and — for the contraction test — the orders of the edges
// around its start vertex, so the three run on every edge of the last chaining in two launches.  What stays on the
// host is the bookkeeping between them (findContractionTargets / findDeletableVertices / contract, main.cpp:462-531):
// it walks an unordered_map keyed by Edge pointers under one mutex and its outcome depends on that iteration order.
// Deleting vertices in between only removes edges; it does not change the verdict on an edge that survives, so the
// host applies the per-edge results to the edges it still has.
//
// Graph lookups of the reference and what serves them here (edges are sorted by (first, second), first < second):
//   getNeighbors(v)         -> the edges whose first vertex is v (a contiguous range, `row_off`) and the edges whose
//                              second vertex is v (`low_off` / `low_list`, built with one counting pass).  The
//                              reference walks the neighbours in id order and stops at the first failure; the
//                              verdict is the conjunction over all neighbours, so the order does not matter.
//   hasEdge / getEdge(a, b) -> binary search for max(a, b) in the row of min(a, b)
//   Edge::getEdgeOrders()   -> the order slots of the edge (msgpu_orders, include/msgpu.h)
#include <string>

#include "../../include/msgpu.h"
#include "internal.h"
#include "scan.cuh"

namespace msgpu {

constexpr uint8_t kOStartFirst = 1, kOContained = 2, kODirection = 4, kOPrimary = 8;

// row_off[v] = first edge whose first vertex is >= v   (v in [0, n_reads])
__global__ void __launch_bounds__(256)
    edge_rows_kernel(const uint32_t *__restrict__ first, uint64_t n_edges, uint32_t n_reads, uint32_t *__restrict__ row_off) {
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e <= n_edges; e += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t lo = e == 0 ? 0u : first[e - 1] + 1u;        // vertices in [lo, hi] start at edge e
    const uint32_t hi = e == n_edges ? n_reads : first[e];
    for (uint64_t v = lo; v <= hi; ++v) row_off[v] = uint32_t(e);
  }
}

__global__ void __launch_bounds__(256)
    lower_degree_kernel(const uint32_t *__restrict__ second, uint64_t n_edges, uint32_t *__restrict__ deg) {
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_edges; e += uint64_t(gridDim.x) * blockDim.x)
    atomicAdd(&deg[second[e]], 1u);
}

__global__ void __launch_bounds__(256)
    lower_scatter_kernel(const uint32_t *__restrict__ second, uint64_t n_edges, const uint32_t *__restrict__ off,
                         uint32_t *__restrict__ cursor, uint32_t *__restrict__ list) {
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_edges; e += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t v                        = second[e];
    list[off[v] + atomicAdd(&cursor[v], 1u)] = uint32_t(e);
  }
}

struct FilterArgs {
  EdgeTables      e;
  OrderTables     o;
  FilterTables    out;
  const uint32_t *row_off, *low_off, *low_list;
  uint64_t        wiggle;
  uint32_t       *cand;   // edges that hold a contained primary order
  unsigned int   *n_cand;
};

// findDeletableEdges + computeBitweight for every edge; edges with a contained primary order are listed for the
// contraction test
__global__ void __launch_bounds__(256) edge_verdicts_kernel(FilterArgs a) {
  for (uint64_t base = uint64_t(blockIdx.x) * blockDim.x; base < a.e.n_edges; base += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t e    = base + threadIdx.x;
    const bool     live = e < a.e.n_edges;
    bool           candidate = false;
    if (live) {
      const uint64_t k0 = a.e.em_begin[e];
      const uint32_t n  = a.o.edge_n_orders[e];
      uint32_t       kept = 0;
      uint64_t       w0 = 0;
      bool           d0 = false, other = false;
      for (uint32_t i = 0; i < n; ++i) {
        const uint8_t f = a.o.flags[k0 + i];
        candidate |= (f & kOContained) && (f & kOPrimary);
        if (f & kOContained) continue; // std::remove_copy_if(..., isContained)   main.cpp:536-538
        const bool d = (f & kODirection) != 0;
        if (kept == 0) {
          w0 = a.o.score[k0 + i];
          d0 = d;
        } else {
          other |= d != d0;
        }
        ++kept;
      }
      a.out.n_kept[e]    = kept;
      a.out.deletable[e] = kept == 0 ? 1 : 0; // main.cpp:540-545
      // computeBitweight on the filtered orders (main.cpp:552-573); a new Edge has weight 0 and e_NONE (Edge.cpp:32-33)
      uint64_t weight = 0;
      uint8_t  cons   = 'c';
      if (kept) {
        if (a.o.edge_shadow[e]) {
          if (!other) cons = d0 ? 'a' : 'b';
        } else {
          weight = w0;
          cons   = d0 ? 'a' : 'b';
        }
      }
      a.out.weight[e]      = weight;
      a.out.consensus[e]   = cons;
      a.out.contraction[e] = -1;
    }
    // warp-aggregated append
    const unsigned mask = __ballot_sync(0xffffffffu, candidate);
    if (mask) {
      const unsigned l = threadIdx.x & 31, lead = __ffs(mask) - 1;
      unsigned       b = 0;
      if (l == lead) b = atomicAdd(a.n_cand, unsigned(__popc(mask)));
      b = __shfl_sync(0xffffffffu, b, lead);
      if (candidate) a.cand[b + __popc(mask & ((1u << l) - 1u))] = uint32_t(e);
    }
  }
}

struct Ord { // one EdgeOrder as sanityCheck reads it
  double  left, right;
  uint8_t f;
  __device__ bool dir() const { return (f & kODirection) != 0; }
  __device__ bool contained() const { return (f & kOContained) != 0; }
  __device__ bool start_first() const { return (f & kOStartFirst) != 0; }
};

__device__ __forceinline__ Ord load_ord(const FilterArgs &a, uint64_t slot) {
  Ord o;
  o.left  = a.o.left[slot];
  o.right = a.o.right[slot];
  o.f     = a.o.flags[slot];
  return o;
}

// index of edge (x, y) or 0xffffffff     (GraphBase::hasEdge / getEdge, Graph.cpp:62-116)
__device__ __forceinline__ uint32_t find_edge(const FilterArgs &a, uint32_t x, uint32_t y) {
  const uint32_t f = x < y ? x : y, s = x < y ? y : x;
  uint32_t       lo = a.row_off[f], hi = a.row_off[f + 1];
  const uint32_t end = hi;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.e.second[mid] < s) lo = mid + 1; else hi = mid;
  }
  return (lo < end && a.e.second[lo] == s) ? lo : 0xffffffffu;
}

// sanityCheck(graph, subnode, node, target, order, wiggleRoom)   sc.cpp:29-88
//   e_on  = edge (node, target), e_for = edge (subnode, target); `order` belongs to edge (subnode, node) whose first
//   vertex is `order_first`.  baseVertex of an order is the first vertex of its edge (ol.cpp:78-97).
__device__ bool sanity_check(const FilterArgs &a, uint32_t e_on, uint32_t e_for, uint32_t node, uint32_t target,
                             const Ord &order) {
  const uint64_t on0 = a.e.em_begin[e_on], for0 = a.e.em_begin[e_for];
  const uint32_t n_on = a.o.edge_n_orders[e_on], n_for = a.o.edge_n_orders[e_for];
  const uint32_t on_first = a.e.first[e_on], on_second = a.e.second[e_on];
  const uint32_t for_first = a.e.first[e_for], for_second = a.e.second[e_for];
  const double   wiggle = double(a.wiggle);
  for (uint32_t i = 0; i < n_on; ++i) {
    const Ord      on       = load_ord(a, on0 + i);
    const uint32_t on_start = on.start_first() ? on_first : on_second, on_end = on.start_first() ? on_second : on_first;
    for (uint32_t j = 0; j < n_for; ++j) {
      const Ord      fo        = load_ord(a, for0 + j);
      const uint32_t for_start = fo.start_first() ? for_first : for_second;
      const uint32_t for_end   = fo.start_first() ? for_second : for_first;
      bool           sane      = ((order.dir() == on.dir()) == fo.dir()); // Toggle::operator* is equality (Toggle.h:153)
      if (fo.contained() && on.contained()) {
        sane = sane && (for_start == target || for_end == target) && on_start == target;
      } else if (fo.contained() && !on.contained()) {
        if (for_end != target) {
          bool l1 = false, l2 = false, l3 = false;
          // both arms of the reference's condition reduce to checkOnOrder.endVertex == &target (sc.cpp:50-52)
          (void)node;
          if (on_end == target) {
            if (!order.dir()) l2 = true;
          } else {
            l1 = true;
            l3 = true;
            if (order.dir()) l2 = true;
          }
          // order.baseVertex != order.endVertex  <=>  the order starts at its edge's first vertex
          if (!order.dir() && order.start_first()) l1 = !l1;
          if (!fo.dir() && fo.start_first()) l2 = !l2;
          const double d1 = l1 ? order.left : order.right;
          const double d2 = l2 ? fo.left : fo.right;
          const double d3 = l3 ? on.left : on.right;
          sane = sane && (__dadd_rn(__dadd_rn(d1, d2), d3) < wiggle);
        }
      } else if (!fo.contained() && on.contained()) {
        sane = sane && on_start == target;
      } else {
        bool d1 = for_start == target, d2 = on_start == target;
        if (!fo.dir() && for_first == target) d1 = !d1;
        if (!on.dir() && on_first == target) d2 = !d2;
        if (!order.dir()) d1 = !d1;
        sane = sane && d1 == d2;
      }
      if (sane) return true;
    }
  }
  return false;
}

// findContractionEdges: one warp per listed edge; the lanes share the neighbours of the order's start vertex
__global__ void __launch_bounds__(256) contraction_edges_kernel(FilterArgs a) {
  const unsigned l  = threadIdx.x & 31;
  const uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5, n = *a.n_cand;
  for (uint64_t x = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; x < n; x += nw) {
    const uint32_t e  = a.cand[x];
    const uint64_t k0 = a.e.em_begin[e];
    const uint32_t no = a.o.edge_n_orders[e], v1 = a.e.first[e], v2 = a.e.second[e];
    int            chosen = -1;
    for (uint32_t oi = 0; oi < no && chosen < 0; ++oi) {
      const Ord order = load_ord(a, k0 + oi);
      if (!(order.contained() && (order.f & kOPrimary))) continue; // main.cpp:422
      const uint32_t start = order.start_first() ? v1 : v2, end = order.start_first() ? v2 : v1;
      const uint32_t up0 = a.row_off[start], n_up = a.row_off[start + 1] - up0;
      const uint32_t lw0 = a.low_off[start], n_lw = a.low_off[start + 1] - lw0;
      bool           sane = true;
      for (uint32_t i0 = 0; i0 < n_up + n_lw && sane; i0 += 32) {
        const uint32_t i    = i0 + l;
        bool           fail = false;
        if (i < n_up + n_lw) {
          const uint32_t sub    = i < n_up ? up0 + i : a.low_list[lw0 + (i - n_up)]; // edge (start, target)
          const uint32_t target = i < n_up ? a.e.second[sub] : a.e.first[sub];
          if (target != end && !a.o.edge_shadow[sub]) { // main.cpp:432-434
            const uint32_t on = find_edge(a, end, target); // main.cpp:436
            fail              = on == 0xffffffffu || !sanity_check(a, on, sub, end, target, order); // main.cpp:441-443
          }
        }
        sane = !__any_sync(0xffffffffu, fail);
      }
      if (sane) chosen = int(oi); // main.cpp:450-457: the first sane order is recorded
    }
    if (l == 0) a.out.contraction[e] = chosen;
  }
}

void run_edge_filters(Arena &arena, cudaStream_t st, uint64_t n_reads, const EdgeTables &e, const OrderTables &o,
                      uint64_t wiggle_room, FilterTables &out, float *ms_out) {
  out = FilterTables{};
  if (!o.valid) throw BuildError{MSGPU_ERR_NO_RESULT, "the edge filters need the orders of msgpu_chain_overlaps"};
  const uint64_t E = e.n_edges, R = n_reads;
  if (E >= 0xffffffffull) throw BuildError{MSGPU_ERR_LIMIT, "the edge filters support < 2^32 edges per device"};
  cudaEvent_t ev0, ev1;
  MSGPU_CUDA(cudaEventCreate(&ev0));
  MSGPU_CUDA(cudaEventCreate(&ev1));
  MSGPU_CUDA(cudaEventRecord(ev0, st));

  uint32_t *row_off = arena.alloc<uint32_t>(R + 2), *low_deg = arena.alloc<uint32_t>(R + 2);
  uint32_t *low_off = arena.alloc<uint32_t>(R + 2), *cursor = arena.alloc<uint32_t>(R + 2);
  uint32_t *low_list = arena.alloc<uint32_t>(E), *cand = arena.alloc<uint32_t>(E);
  unsigned int *d_n = arena.alloc<unsigned int>(2);
  MSGPU_CUDA(cudaMemsetAsync(low_deg, 0, (R + 2) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(cursor, 0, (R + 2) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(d_n, 0, 2 * sizeof(unsigned int), st));
  const int ge = grid_for(E + 1, 256, 32);
  MSGPU_LAUNCH("edge_rows", st, edge_rows_kernel<<<ge, 256, 0, st>>>(e.first, E, uint32_t(R), row_off));
  if (E) MSGPU_LAUNCH("lower_degree", st, lower_degree_kernel<<<ge, 256, 0, st>>>(e.second, E, low_deg));
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{low_deg}, R + 1, low_off, (uint32_t *)nullptr);
  if (E) MSGPU_LAUNCH("lower_scatter", st, lower_scatter_kernel<<<ge, 256, 0, st>>>(e.second, E, low_off, cursor, low_list));

  out.contraction = arena.alloc<int32_t>(E);
  out.deletable   = arena.alloc<uint8_t>(E);
  out.n_kept      = arena.alloc<uint32_t>(E);
  out.weight      = arena.alloc<uint64_t>(E);
  out.consensus   = arena.alloc<uint8_t>(E);
  FilterArgs a;
  a.e        = e;
  a.o        = o;
  a.out      = out;
  a.row_off  = row_off;
  a.low_off  = low_off;
  a.low_list = low_list;
  a.wiggle   = wiggle_room;
  a.cand     = cand;
  a.n_cand   = d_n;
  if (E) {
    MSGPU_LAUNCH("edge_verdicts", st, edge_verdicts_kernel<<<grid_for(E, 256, 32), 256, 0, st>>>(a));
    MSGPU_LAUNCH("contraction_edges", st, contraction_edges_kernel<<<grid_for(E * 32, 256, 8), 256, 0, st>>>(a));
  }
  MSGPU_CUDA(cudaEventRecord(ev1, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  if (ms_out) MSGPU_CUDA(cudaEventElapsedTime(ms_out, ev0, ev1));
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  out.n_edges = E;
  out.valid   = true;
}

} // namespace msgpu
