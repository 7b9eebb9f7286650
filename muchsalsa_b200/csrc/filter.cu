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
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e <= n_edges; e += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t lo = e == 0 ? 0u : first[e - 1] + 1u;        // vertices in [lo, hi] start at edge e
    const uint32_t hi = e == n_edges ? n_reads : first[e];
    for (uint64_t v = lo; v <= hi; ++v) row_off[v] = uint32_t(e);
  }
}

__global__ void __launch_bounds__(256)
    lower_degree_kernel(const uint32_t *__restrict__ second, uint64_t n_edges, uint32_t *__restrict__ deg) {
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_edges; e += uint64_t(gridDim.x) * blockDim.x)
    atomicAdd(&deg[second[e]], 1u);
}

__global__ void __launch_bounds__(256)
    lower_scatter_kernel(const uint32_t *__restrict__ second, uint64_t n_edges, const uint32_t *__restrict__ off,
                         uint32_t *__restrict__ cursor, uint32_t *__restrict__ list) {
  pdl_sync();
  for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n_edges; e += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t v                        = second[e];
    list[off[v] + atomicAdd(&cursor[v], 1u)] = uint32_t(e);
  }
}

// What the contraction test reads of ANY edge of the graph: on one device the tables of the build itself; on one slice
// of a multi-GPU build the concatenation of every rank's exported view (run_edge_filters_sharded).
struct EdgeView {
  uint64_t        n_edges = 0;
  const uint32_t *first = nullptr, *second = nullptr, *n_orders = nullptr;
  const uint8_t  *shadow = nullptr;
  const uint64_t *ord_off = nullptr; // first order slot of an edge
  const double   *left = nullptr, *right = nullptr;
  const uint8_t  *flags = nullptr;
};

struct FilterArgs {
  EdgeTables      e; // this device's edges (all of them, or its slice) ...
  OrderTables     o; // ... and their orders
  FilterTables    out;
  EdgeView        g;         // every edge of the graph
  uint64_t        edge_base; // index of this device's first edge in `g`
  const uint32_t *row_off, *low_off, *low_list;
  uint64_t        wiggle;
  uint32_t       *cand;   // edges (indices in `g`) that hold a contained primary order
  unsigned int   *n_cand;
};

// findDeletableEdges + computeBitweight for every edge; edges with a contained primary order are listed for the
// contraction test
__global__ void __launch_bounds__(256) edge_verdicts_kernel(FilterArgs a) {
  pdl_sync();
  for (uint64_t base = uint64_t(blockIdx.x) * blockDim.x; base < a.e.n_edges; base += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t e    = base + threadIdx.x;
    const bool     live = e < a.e.n_edges;
    bool           candidate = false;
    if (live) {
      const uint64_t k0 = a.e.em_begin[e] - a.e.em_value_base; // (order slots are indexed like this device's edge matches)
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
      if (candidate) a.cand[b + __popc(mask & ((1u << l) - 1u))] = uint32_t(a.edge_base + e);
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
  o.left  = a.g.left[slot];
  o.right = a.g.right[slot];
  o.f     = a.g.flags[slot];
  return o;
}

// index of edge (x, y) or 0xffffffff     (GraphBase::hasEdge / getEdge, Graph.cpp:62-116)
__device__ __forceinline__ uint32_t find_edge(const FilterArgs &a, uint32_t x, uint32_t y) {
  const uint32_t f = x < y ? x : y, s = x < y ? y : x;
  uint32_t       lo = a.row_off[f], hi = a.row_off[f + 1];
  const uint32_t end = hi;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a.g.second[mid] < s) lo = mid + 1; else hi = mid;
  }
  return (lo < end && a.g.second[lo] == s) ? lo : 0xffffffffu;
}

// sanityCheck(graph, subnode, node, target, order, wiggleRoom)   sc.cpp:29-88
//   e_on  = edge (node, target), e_for = edge (subnode, target); `order` belongs to edge (subnode, node) whose first
//   vertex is `order_first`.  baseVertex of an order is the first vertex of its edge (ol.cpp:78-97).
__device__ bool sanity_check(const FilterArgs &a, uint32_t e_on, uint32_t e_for, uint32_t node, uint32_t target,
                             const Ord &order) {
  const uint64_t on0 = a.g.ord_off[e_on], for0 = a.g.ord_off[e_for];
  const uint32_t n_on = a.g.n_orders[e_on], n_for = a.g.n_orders[e_for];
  const uint32_t on_first = a.g.first[e_on], on_second = a.g.second[e_on];
  const uint32_t for_first = a.g.first[e_for], for_second = a.g.second[e_for];
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
  pdl_sync();
  const unsigned l  = threadIdx.x & 31;
  const uint64_t nw = (uint64_t(gridDim.x) * blockDim.x) >> 5, n = *a.n_cand;
  for (uint64_t x = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; x < n; x += nw) {
    const uint32_t e  = a.cand[x];
    const uint64_t k0 = a.g.ord_off[e];
    const uint32_t no = a.g.n_orders[e], v1 = a.g.first[e], v2 = a.g.second[e];
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
          const uint32_t target = i < n_up ? a.g.second[sub] : a.g.first[sub];
          if (target != end && !a.g.shadow[sub]) { // main.cpp:432-434
            const uint32_t on = find_edge(a, end, target); // main.cpp:436
            fail              = on == 0xffffffffu || !sanity_check(a, on, sub, end, target, order); // main.cpp:441-443
          }
        }
        sane = !__any_sync(0xffffffffu, fail);
      }
      if (sane) chosen = int(oi); // main.cpp:450-457: the first sane order is recorded
    }
    if (l == 0) a.out.contraction[e - a.edge_base] = chosen;
  }
}

// the three passes over this device's edges `e` / orders `o`, with the graph around them given by `g`
static void filters_core(Arena &arena, cudaStream_t st, uint64_t n_reads, const EdgeTables &e, const OrderTables &o,
                         const EdgeView &g, uint64_t edge_base, uint64_t wiggle_room, FilterTables &out) {
  const uint64_t E = e.n_edges, G = g.n_edges, R = n_reads;
  if (G >= 0xffffffffull) throw BuildError{MSGPU_ERR_LIMIT, "the edge filters support < 2^32 edges in the whole graph"};
  uint32_t *row_off = arena.alloc<uint32_t>(R + 2), *low_deg = arena.alloc<uint32_t>(R + 2);
  uint32_t *low_off = arena.alloc<uint32_t>(R + 2), *cursor = arena.alloc<uint32_t>(R + 2);
  uint32_t *low_list = arena.alloc<uint32_t>(G), *cand = arena.alloc<uint32_t>(E);
  unsigned int *d_n = arena.alloc<unsigned int>(2);
  MSGPU_CUDA(cudaMemsetAsync(low_deg, 0, (R + 2) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(cursor, 0, (R + 2) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(d_n, 0, 2 * sizeof(unsigned int), st));
  const int gg = grid_for(G + 1, 256, 32);
  MSGPU_LAUNCH("edge_rows", st, launch_k(edge_rows_kernel, gg, 256, 0, st, g.first, G, uint32_t(R), row_off));
  if (G) MSGPU_LAUNCH("lower_degree", st, launch_k(lower_degree_kernel, gg, 256, 0, st, g.second, G, low_deg));
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{low_deg}, R + 1, low_off, (uint32_t *)nullptr);
  if (G) MSGPU_LAUNCH("lower_scatter", st, launch_k(lower_scatter_kernel, gg, 256, 0, st, g.second, G, low_off, cursor, low_list));

  out.contraction = arena.alloc<int32_t>(E);
  out.deletable   = arena.alloc<uint8_t>(E);
  out.n_kept      = arena.alloc<uint32_t>(E);
  out.weight      = arena.alloc<uint64_t>(E);
  out.consensus   = arena.alloc<uint8_t>(E);
  FilterArgs a;
  a.e         = e;
  a.o         = o;
  a.out       = out;
  a.g         = g;
  a.edge_base = edge_base;
  a.row_off   = row_off;
  a.low_off   = low_off;
  a.low_list  = low_list;
  a.wiggle    = wiggle_room;
  a.cand      = cand;
  a.n_cand    = d_n;
  if (E) {
    MSGPU_LAUNCH("edge_verdicts", st, launch_k(edge_verdicts_kernel, grid_for(E, 256, 32), 256, 0, st, a));
    MSGPU_LAUNCH("contraction_edges", st, launch_k(contraction_edges_kernel, grid_for(E * 32, 256, 8), 256, 0, st, a));
  }
  out.n_edges = E;
  out.valid   = true;
}

void run_edge_filters(Arena &arena, cudaStream_t st, uint64_t n_reads, const EdgeTables &e, const OrderTables &o,
                      uint64_t wiggle_room, FilterTables &out, float *ms_out) {
  out = FilterTables{};
  if (!o.valid) throw BuildError{MSGPU_ERR_NO_RESULT, "the edge filters need the orders of msgpu_chain_overlaps"};
  cudaEvent_t ev0, ev1;
  MSGPU_CUDA(cudaEventCreate(&ev0));
  MSGPU_CUDA(cudaEventCreate(&ev1));
  MSGPU_CUDA(cudaEventRecord(ev0, st));
  EdgeView g;
  g.n_edges  = e.n_edges;
  g.first    = e.first;
  g.second   = e.second;
  g.n_orders = o.edge_n_orders;
  g.shadow   = o.edge_shadow;
  g.ord_off  = e.em_begin; // the orders of an edge sit in the slots of its edge matches
  g.left     = o.left;
  g.right    = o.right;
  g.flags    = o.flags;
  filters_core(arena, st, n_reads, e, o, g, 0, wiggle_room, out);
  MSGPU_CUDA(cudaEventRecord(ev1, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  if (ms_out) MSGPU_CUDA(cudaEventElapsedTime(ms_out, ev0, ev1));
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-GPU: every rank holds a slice of the edge table, but the contraction test walks the edges AROUND an order's start
// vertex and the third edge of every triangle (main.cpp:432-443) — rows that live on other ranks.  Each rank therefore
// exports a compact VIEW of its slice — per edge (first, second, number of orders, shadow flag), per order (left,
// right, flags), the orders packed densely — the views travel (one all-gather), and every rank runs the passes on its
// own edges against the concatenation of all views, which is the view of the whole graph because the slices are
// consecutive ranges of the (first, second)-sorted table.
// ---------------------------------------------------------------------------------------------------------------------
struct ViewLayout { // byte offsets of the sections of a view blob, each 16-byte aligned
  uint64_t first, second, n_orders, shadow, left, right, flags, total;
};
static ViewLayout view_layout(uint64_t E, uint64_t O) {
  auto       up = [](uint64_t x) { return (x + 15) & ~uint64_t(15); };
  ViewLayout l;
  uint64_t   at = 0;
  l.first    = at; at += up(E * 4);
  l.second   = at; at += up(E * 4);
  l.n_orders = at; at += up(E * 4);
  l.shadow   = at; at += up(E);
  l.left     = at; at += up(O * 8);
  l.right    = at; at += up(O * 8);
  l.flags    = at; at += up(O);
  l.total    = at;
  return l;
}

__global__ void __launch_bounds__(256)
    pack_orders_kernel(EdgeTables e, OrderTables o, const uint64_t *__restrict__ ord_prefix, double *__restrict__ left,
                       double *__restrict__ right, uint8_t *__restrict__ flags) {
  pdl_sync();
  for (uint64_t x = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < e.n_edges; x += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t k0 = e.em_begin[x] - e.em_value_base, w0 = ord_prefix[x];
    const uint32_t n  = o.edge_n_orders[x];
    for (uint32_t i = 0; i < n; ++i) {
      left[w0 + i]  = o.left[k0 + i];
      right[w0 + i] = o.right[k0 + i];
      flags[w0 + i] = o.flags[k0 + i];
    }
  }
}

// sizes of this device's view (one round trip: the number of orders is a sum over the edges)
void filter_view_describe(Arena &arena, cudaStream_t st, const EdgeTables &e, const OrderTables &o, FilterView &v) {
  if (!o.valid) throw BuildError{MSGPU_ERR_NO_RESULT, "the edge filters need the orders of msgpu_chain_overlaps"};
  v            = FilterView{};
  v.ord_prefix = arena.alloc<uint64_t>(e.n_edges + 1);
  uint64_t *d_total = arena.alloc<uint64_t>(1);
  exclusive_scan<uint64_t>(arena, st, LoadAs<uint32_t, uint64_t>{o.edge_n_orders}, e.n_edges, v.ord_prefix, d_total);
  uint64_t *h = static_cast<uint64_t *>(arena.host_scratch());
  MSGPU_CUDA(cudaMemcpyAsync(h, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  v.n_edges  = e.n_edges;
  v.n_orders = h[0];
  v.bytes    = view_layout(v.n_edges, v.n_orders).total;
  v.valid    = true;
}

void filter_view_export(cudaStream_t st, const EdgeTables &e, const OrderTables &o, const FilterView &v, void *dev_dst) {
  const ViewLayout l = view_layout(v.n_edges, v.n_orders);
  char            *d = static_cast<char *>(dev_dst);
  const uint64_t   E = v.n_edges;
  if (E) {
    MSGPU_CUDA(cudaMemcpyAsync(d + l.first, e.first, E * 4, cudaMemcpyDeviceToDevice, st));
    MSGPU_CUDA(cudaMemcpyAsync(d + l.second, e.second, E * 4, cudaMemcpyDeviceToDevice, st));
    MSGPU_CUDA(cudaMemcpyAsync(d + l.n_orders, o.edge_n_orders, E * 4, cudaMemcpyDeviceToDevice, st));
    MSGPU_CUDA(cudaMemcpyAsync(d + l.shadow, o.edge_shadow, E, cudaMemcpyDeviceToDevice, st));
    MSGPU_LAUNCH("pack_orders", st, launch_k(pack_orders_kernel, grid_for(E, 256, 32), 256, 0, st, 
                                        e, o, v.ord_prefix, reinterpret_cast<double *>(d + l.left),
                                        reinterpret_cast<double *>(d + l.right), reinterpret_cast<uint8_t *>(d + l.flags)));
  }
}

void run_edge_filters_sharded(Arena &arena, cudaStream_t st, uint64_t n_reads, const EdgeTables &e, const OrderTables &o,
                              uint64_t wiggle_room, uint32_t n_parts, uint32_t part, const msgpu_filter_view_header *headers,
                              const void *const *dev_blobs, FilterTables &out, float *ms_out) {
  out = FilterTables{};
  if (!o.valid) throw BuildError{MSGPU_ERR_NO_RESULT, "the edge filters need the orders of msgpu_chain_overlaps"};
  if (part >= n_parts) throw BuildError{MSGPU_ERR_INVALID_ARG, "part >= n_parts"};
  uint64_t G = 0, O = 0, base = 0;
  for (uint32_t r = 0; r < n_parts; ++r) {
    if (r == part) base = G;
    G += headers[r].n_edges;
    O += headers[r].n_orders;
  }
  if (headers[part].n_edges != e.n_edges)
    throw BuildError{MSGPU_ERR_INVALID_ARG, "the view header of this rank does not describe the context's edge tables"};
  cudaEvent_t ev0, ev1;
  MSGPU_CUDA(cudaEventCreate(&ev0));
  MSGPU_CUDA(cudaEventCreate(&ev1));
  MSGPU_CUDA(cudaEventRecord(ev0, st));
  uint32_t *g_first = arena.alloc<uint32_t>(G), *g_second = arena.alloc<uint32_t>(G), *g_no = arena.alloc<uint32_t>(G);
  uint8_t  *g_shadow = arena.alloc<uint8_t>(G), *g_flags = arena.alloc<uint8_t>(O);
  double   *g_left = arena.alloc<double>(O), *g_right = arena.alloc<double>(O);
  uint64_t *g_off = arena.alloc<uint64_t>(G + 1);
  uint64_t  ge = 0, go = 0;
  for (uint32_t r = 0; r < n_parts; ++r) { // the slices follow each other in key order: concatenation = the whole graph
    const uint64_t   Er = headers[r].n_edges, Or = headers[r].n_orders;
    const ViewLayout l  = view_layout(Er, Or);
    const char      *b  = static_cast<const char *>(dev_blobs[r]);
    auto copy = [&](void *dst, uint64_t at, uint64_t bytes) {
      if (bytes) MSGPU_CUDA(cudaMemcpyAsync(dst, b + at, bytes, cudaMemcpyDeviceToDevice, st));
    };
    copy(g_first + ge, l.first, Er * 4);
    copy(g_second + ge, l.second, Er * 4);
    copy(g_no + ge, l.n_orders, Er * 4);
    copy(g_shadow + ge, l.shadow, Er);
    copy(g_left + go, l.left, Or * 8);
    copy(g_right + go, l.right, Or * 8);
    copy(g_flags + go, l.flags, Or);
    ge += Er;
    go += Or;
  }
  uint64_t *d_total = arena.alloc<uint64_t>(1);
  exclusive_scan<uint64_t>(arena, st, LoadAs<uint32_t, uint64_t>{g_no}, G, g_off, d_total);
  if (G) MSGPU_CUDA(cudaMemcpyAsync(g_off + G, d_total, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  EdgeView g;
  g.n_edges  = G;
  g.first    = g_first;
  g.second   = g_second;
  g.n_orders = g_no;
  g.shadow   = g_shadow;
  g.ord_off  = g_off;
  g.left     = g_left;
  g.right    = g_right;
  g.flags    = g_flags;
  filters_core(arena, st, n_reads, e, o, g, base, wiggle_room, out);
  MSGPU_CUDA(cudaEventRecord(ev1, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  if (ms_out) MSGPU_CUDA(cudaEventElapsedTime(ms_out, ev0, ev1));
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
}

} // namespace msgpu
