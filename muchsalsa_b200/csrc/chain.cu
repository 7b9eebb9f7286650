// chain.cu — the stage that consumes the edge / edge-match CSR while it is still in HBM: per edge, chain its edge
// matches into maximal compatible paths and derive the overlap geometry.
//
// Replaces, per edge, the job body chainingAndOverlaps (src/main.cpp:328-414) and what it calls:
//   getMaxPairwisePaths   libms/src/kernel/mpp.cpp:145-305   (checkCompatibility: mpp.cpp:38-142)
//   getOverlap            libms/src/kernel/ol.cpp:53-101     (getOverhangs: ol.cpp:31-50)
// and yields what it leaves on the Edge: the shadow flag (Edge.h:241) and the EdgeOrder list (Edge.h:49-60,257-259):
// start / end vertex, leftOffset / rightOffset (the "relative shifts"), containment, the integer score, the unitig ids
// of the path, direction and primary flag.
//
// One thread per edge (edges are independent; most hold 1-4 edge matches, the dynamic programme is O(n^2) in the edge
// matches of one direction).  All arithmetic is the reference's, in IEEE double with the same operation order
// (explicit _rn intrinsics; the library is built with --fmad=false), so offsets and scores are bit-identical.
//
// What the reference reads through MatchMap lookups is read here from the tables of the construction:
//   getEdgeMatch(edge, unitig)     -> the edge's rows of the edge-match table (CSR)
//   getVertexMatch(read, unitig)   -> binary search for `read` in the unitig's group of the pairing index (groups are
//                                     ordered by read id), done ONCE per edge match by chain_rows_kernel (every edge
//                                     in parallel, no other state), which leaves the two positions next to the edge
//                                     match; the fields come from a 32-byte row (one L2 sector) in pairing order
//   getVertexMatches(read)         -> a per-read list of unitig ids ordered by (nanoporeRange, unitig), built here
//
// Five launches share the per-edge programme.  90 % of the edges at yeast scale hold ONE edge match: for them the
// dynamic programme is the identity and chain_single_kernel does the whole job in registers.  Edges with 2..4 edge
// matches go to the programme instantiated with 4-entry state arrays, those with 5..8 to the 8-entry one, the rest to
// the 96-entry instantiation (and beyond 96 of one direction to a fallback over a global scratch slab); the lists are
// filled by warp-aggregated appends, and every result is written at a position that depends on the edge alone, so the
// output does not depend on the order of the lists.  (Measured and not kept, round 2: the state arrays in shared
// memory instead of local memory — the few threads per SM that fit no longer cover the programme's global loads.)
#include <string>

#include "../../include/msgpu.h"
#include "internal.h"
#include "scan.cuh"

namespace msgpu {

constexpr int kChainMax   = 96; // edge matches of ONE direction on one edge that the per-thread programme holds
constexpr int kChainSmall = 8;  // edges with at most this many edge matches run the small instantiation
constexpr int kChainTiny  = 4;  // ... and these the smallest one: its state stays in L1 / L2

// the fields of one VertexMatch the stage reads, in pairing order (grouped by unitig, by read id inside): 32 bytes
struct __align__(16) VmRow {
  int32_t  nr0, nr1, ir0, ir1;
  double   rratio;
  uint32_t flags, unitig;
};

__global__ void __launch_bounds__(256)
    build_vm_rows_kernel(MatchSoA m, const PairRow *__restrict__ rows, VmRow *__restrict__ vm) {
  pdl_sync();
  for (uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < m.n; q += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t i = rows[q].orig;
    VmRow          v;
    v.nr0    = m.nr0[i];
    v.nr1    = m.nr1[i];
    v.ir0    = m.ir0[i];
    v.ir1    = m.ir1[i];
    v.rratio = m.rratio[i];
    v.flags  = m.flags[i];
    v.unitig = m.unitig[i];
    vm[q]    = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-read match lists ordered by (nanoporeRange.first, nanoporeRange.second, unitig)   [mpp.cpp:253-273]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    read_degree_kernel(const uint32_t *__restrict__ read, uint64_t n, uint32_t *__restrict__ deg) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
    atomicAdd(&deg[read[i]], 1u);
}

__global__ void __launch_bounds__(256)
    read_scatter_kernel(const uint32_t *__restrict__ read, uint64_t n, const uint32_t *__restrict__ off,
                        uint32_t *__restrict__ cursor, uint32_t *__restrict__ tmp) {
  pdl_sync();
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t r                             = read[i];
    tmp[off[r] + atomicAdd(&cursor[r], 1u)] = uint32_t(i);
  }
}

__device__ __forceinline__ bool nano_less(const MatchSoA &m, uint32_t a, uint32_t b) { // tuple<pair<int,int>, id> order
  if (m.nr0[a] != m.nr0[b]) return m.nr0[a] < m.nr0[b];
  if (m.nr1[a] != m.nr1[b]) return m.nr1[a] < m.nr1[b];
  return m.unitig[a] < m.unitig[b];
}

__global__ void __launch_bounds__(256)
    read_order_kernel(MatchSoA m, const uint32_t *__restrict__ off, const uint32_t *__restrict__ tmp,
                      uint32_t *__restrict__ list_unitig) {
  pdl_sync();
  for (uint64_t x = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < m.n; x += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t me = tmp[x], r = m.read[me], lo = off[r], hi = off[r + 1];
    uint32_t       rank = 0;
    for (uint32_t p = lo; p < hi; ++p) rank += nano_less(m, tmp[p], me) ? 1u : 0u;
    list_unitig[lo + rank] = m.unitig[me]; // (unitig ids are unique inside a read, so the order is total)
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the per-edge programme
// ---------------------------------------------------------------------------------------------------------------------
struct ChainArgs {
  const int32_t  *read_nlen;
  const uint32_t *grp_off;
  const uint32_t *p_read;
  const VmRow    *vm;               // match fields in pairing order
  const uint32_t *em_row1, *em_row2; // per edge match: position of the first / second vertex's match in `vm`
  const uint32_t *rm_off, *rm_unitig; // per read: its unitig ids ordered by (nanoporeRange, unitig)
  EdgeTables      e;
  uint64_t        wiggle;
  OrderTables     out;
  uint32_t       *list_in, *list_out; // edges handed from one launch to the next
  unsigned int   *n_in, *n_out;
  // edges with more edge matches of one direction than the largest fast instantiation holds: listed by chain_large,
  // done by chain_huge with their state in a global scratch slab
  uint32_t           *huge_list;
  unsigned int       *n_huge;
  unsigned long long *huge_em, *huge_cursor;
  char               *huge_scratch;
};

struct Vm { // the fields of one VertexMatch the stage reads
  int    nr0, nr1, ir0, ir1;
  double rratio;
  bool   dir;
};

__device__ __forceinline__ Vm load_vm(const ChainArgs &a, uint32_t pos) { // one 256-bit load = one 32-byte sector
  uint4 x, y;
  ld_row32(a.vm + pos, x, y);
  Vm           v;
  v.nr0    = int32_t(x.x);
  v.nr1    = int32_t(x.y);
  v.ir0    = int32_t(x.z);
  v.ir1    = int32_t(x.w);
  v.rratio = __hiloint2double(int(y.y), int(y.x));
  v.dir    = (y.z & 1u) != 0;
  return v;
}

// getVertexMatch(read, unitig) for both vertices of every edge match: one thread per edge, nothing but the searches
// (the group of a unitig is ordered by read id and first < second, so the second search starts behind the first hit)
__global__ void __launch_bounds__(256)
    chain_rows_kernel(EdgeTables e, const uint32_t *__restrict__ grp_off, const uint32_t *__restrict__ p_read,
                      uint32_t *__restrict__ em_row1, uint32_t *__restrict__ em_row2) {
  pdl_sync();
  for (uint64_t x = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < e.n_edges;
       x += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t v1 = e.first[x], v2 = e.second[x];
    const uint64_t k0 = e.em_begin[x], k1 = e.em_begin[x + 1];
    for (uint64_t k = k0; k < k1; ++k) {
      const uint32_t u  = e.em_unitig[k];
      const uint32_t g1 = grp_off[u + 1];
      uint32_t       lo = grp_off[u], hi = g1;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p_read[mid] < v1) lo = mid + 1; else hi = mid;
      }
      em_row1[k] = lo;
      ++lo;
      hi = g1;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p_read[mid] < v2) lo = mid + 1; else hi = mid;
      }
      em_row2[k] = lo;
    }
  }
}

// appends `x` to a list for the lanes that ask for it: one atomic per warp
__device__ __forceinline__ void warp_append(bool want, uint32_t x, uint32_t *list, unsigned int *n) {
  const unsigned act  = __activemask();
  const unsigned mask = __ballot_sync(act, want);
  if (!mask) return;
  const unsigned l = threadIdx.x & 31, lead = __ffs(mask) - 1;
  unsigned       base = 0;
  if (l == lead) base = atomicAdd(n, unsigned(__popc(mask)));
  base = __shfl_sync(act, base, lead);
  if (want) list[base + __popc(mask & ((1u << l) - 1u))] = x;
}

// corrected nanopore interval of a vertex match under an edge match (mpp.cpp:50-69)
__device__ __forceinline__ void corrected_nano(const Vm &v, int ov0, int ov1, double *c0, double *c1) {
  double ncl = __ddiv_rn(double(ov0 - v.ir0), v.rratio);
  double ncr = __ddiv_rn(double(v.ir1 - ov1), v.rratio);
  if (!v.dir) {
    double t = ncl;
    ncl      = ncr;
    ncr      = t;
  }
  *c0 = __dadd_rn(double(v.nr0), ncl);
  *c1 = __dsub_rn(double(v.nr1), ncr);
}

// nanoCheck of checkCompatibility (mpp.cpp:40-113) for one vertex; returns the `abort` flag
__device__ __forceinline__ bool nano_check(double a0, double a1, double b0, double b1, int ra0, int ra1, int rb0,
                                           int rb1, int *orientation, double *diff) {
  *orientation = 0;
  *diff        = 0.0;
  if (a0 <= b1 && b0 <= a1) {
    if (a0 < b0 && a1 < b1) {
      *orientation = 2;
      *diff        = __dadd_rn(__dsub_rn(a1, b0), 1.0);
    }
    if (a0 > b0 && a1 > b1) {
      *orientation = -2;
      *diff        = __dadd_rn(__dsub_rn(b1, a0), 1.0);
    }
  } else if (a0 < b0) {
    *orientation = 1;
    *diff        = __dadd_rn(__dsub_rn(b0, a1), 1.0);
  } else {
    *orientation = -1;
    *diff        = __dadd_rn(__dsub_rn(a0, b1), 1.0);
  }
  if (ra0 <= rb1 && rb0 <= ra1) {
    int uco = 0;
    if (ra0 < rb0 && ra1 < rb1) uco = 2;
    if (ra0 > rb0 && ra1 > rb1) uco = -2;
    if ((*orientation < 0 && uco >= 0) || (*orientation > 0 && uco <= 0)) return true;
  }
  return false;
}

// double -> std::size_t as the reference's build does it (path_t's second member, mpp.cpp:34,217,238).  Out of range
// the conversion is undefined behaviour in C++; x86-64 code (cvttsd2si on d, or on d - 2^63 with the top bit flipped)
// yields 0 for d >= 2^64 and 2^63 for NaN, where CUDA's cast saturates — only reachable with absurd inputs (a
// negative match count widened to size_t), but the tables are compared bit for bit.
__device__ __forceinline__ unsigned long long double_to_size_t(double d) {
  const double two63 = 9223372036854775808.0;
  if (d < two63) return d != d ? 0x8000000000000000ull : (unsigned long long)(long long)d;
  const double t = d - two63;
  if (t < two63) return (unsigned long long)(long long)t ^ 0x8000000000000000ull;
  return 0ull;
}

struct EmState { // per edge match of the direction being chained
  double c1a, c1b, c2a, c2b; // corrected nanopore intervals on the first / second vertex
  int    r1a, r1b, r2a, r2b; // nanoporeRange on the first / second vertex
  double score;
  uint32_t unitig;
  uint32_t k; // row of the edge match inside its edge
  bool   prim;
};

struct PathRes { // one entry of getMaxPairwisePaths' result
  int    tail;   // population entry whose path it is
  double score;
  bool   prim;
};

// getOverhangs (ol.cpp:31-50)
__device__ __forceinline__ void overhangs(const ChainArgs &a, uint32_t vertex, uint32_t pos, int ov0, int ov1,
                                          double *left, double *right) {
  const Vm v   = load_vm(a, pos);
  double   ncl = __ddiv_rn(double(ov0 - v.ir0), v.rratio);
  double   ncr = __ddiv_rn(double(v.ir1 - ov1), v.rratio);
  if (!v.dir) {
    double t = ncl;
    ncl      = ncr;
    ncr      = t;
  }
  *left  = __dadd_rn(double(v.nr0), ncl);
  *right = __dadd_rn(double(a.read_nlen[vertex] - v.nr1), ncr);
}

// State of the per-edge programme.  kCap > 0: arrays of kCap entries per direction in the thread's local memory (the
// fast instantiations).  kCap == 0: the same arrays carved out of a global scratch slab, sized by the edge itself —
// the fallback for edges with more edge matches of one direction than the largest fast instantiation holds (two long
// reads that share hundreds of unitigs), so that no input the reference accepts is refused.
template <int kCap> struct ChainStore {
  PathRes            res[2 * kCap];
  EmState            st[kCap];
  double             pop[kCap];
  int                pred[kCap];
  uint32_t           path_rows[2 * kCap];
  int                path_off[2 * (kCap + 1)];
  bool               keep[2 * kCap];
  unsigned long long used[(kCap + 63) / 64];
  __device__ __forceinline__ int  cap() const { return kCap; }
  __device__ __forceinline__ void bind(char *, int) {}
};
template <> struct ChainStore<0> {
  PathRes            *res;
  EmState            *st;
  double             *pop;
  int                *pred;
  uint32_t           *path_rows;
  int                *path_off;
  bool               *keep;
  unsigned long long *used;
  int                 n;
  __device__ __forceinline__ int cap() const { return n; }
  __host__ __device__ static size_t bytes(size_t cap) { // every section 8-byte aligned
    auto up = [](size_t x) { return (x + 7) & ~size_t(7); };
    return up(2 * cap * sizeof(PathRes)) + up(cap * sizeof(EmState)) + up(cap * sizeof(double)) + up(cap * sizeof(int)) +
           up(2 * cap * sizeof(uint32_t)) + up(2 * (cap + 1) * sizeof(int)) + up(2 * cap) + up(((cap + 63) / 64) * 8);
  }
  __device__ __forceinline__ void bind(char *p, int cap) {
    auto up = [](size_t x) { return (x + 7) & ~size_t(7); };
    n   = cap;
    res = reinterpret_cast<PathRes *>(p);
    p += up(2 * size_t(cap) * sizeof(PathRes));
    st = reinterpret_cast<EmState *>(p);
    p += up(size_t(cap) * sizeof(EmState));
    pop = reinterpret_cast<double *>(p);
    p += up(size_t(cap) * sizeof(double));
    pred = reinterpret_cast<int *>(p);
    p += up(size_t(cap) * sizeof(int));
    path_rows = reinterpret_cast<uint32_t *>(p);
    p += up(2 * size_t(cap) * sizeof(uint32_t));
    path_off = reinterpret_cast<int *>(p);
    p += up(2 * (size_t(cap) + 1) * sizeof(int));
    keep = reinterpret_cast<bool *>(p);
    p += up(2 * size_t(cap));
    used = reinterpret_cast<unsigned long long *>(p);
  }
};

// the whole job of one edge, with state arrays of kCap entries per direction (kCap == 0: of as many as the edge has
// edge matches, in `scratch`)
template <int kCap> __device__ __forceinline__ void chain_one_edge(const ChainArgs &a, uint64_t e, char *scratch = nullptr) {
  {
    const uint32_t v1 = a.e.first[e], v2 = a.e.second[e];
    const uint64_t k0 = a.e.em_begin[e], k1 = a.e.em_begin[e + 1];

    ChainStore<kCap> S;
    S.bind(scratch, int(k1 - k0));
    const int cap = S.cap();
    // results of the two directions: [0] minus (direction false), [1] plus; main.cpp:352-353
    PathRes            *res  = S.res;       // res[d * cap + r]
    int                 n_res[2] = {0, 0};
    EmState            *st   = S.st;
    double             *pop  = S.pop;
    int                *pred = S.pred;
    // path members per (direction, population entry) are re-derived by walking `pred`; after the first direction is
    // done its paths are frozen as lists of edge-match rows
    uint32_t           *path_rows = S.path_rows; // [d * cap + i]: concatenated paths of a direction, in result order
    int                *path_off  = S.path_off;  // [d * (cap + 1) + r]
    unsigned long long *used      = S.used;
    bool                overflow  = false;

    for (int d = 0; d < 2; ++d) {
      const bool direction = d == 1;
      // ---- the edge matches of this direction, ordered by (nanoporeRange of the first vertex, unitig): vStart
      int n = 0;
      for (uint64_t k = k0; k < k1; ++k) {
        if (((a.e.em_flags[k] & 1) != 0) != direction) continue;
        if (n == cap) {
          overflow = true;
          break;
        }
        EmState s;
        s.k      = uint32_t(k - k0);
        s.unitig = a.e.em_unitig[k];
        s.score  = a.e.em_score[k];
        s.prim   = (a.e.em_flags[k] & 2) != 0;
        const Vm  w1 = load_vm(a, a.em_row1[k]), w2 = load_vm(a, a.em_row2[k]);
        const int ov0 = a.e.em_ov0[k], ov1 = a.e.em_ov1[k];
        corrected_nano(w1, ov0, ov1, &s.c1a, &s.c1b);
        corrected_nano(w2, ov0, ov1, &s.c2a, &s.c2b);
        s.r1a = w1.nr0;
        s.r1b = w1.nr1;
        s.r2a = w2.nr0;
        s.r2b = w2.nr1;
        // insertion by (r1a, r1b, unitig): std::sort of tuple<pair<int,int>, unsigned> (mpp.cpp:172)
        int p = n;
        while (p > 0) {
          const EmState &q = st[p - 1];
          const bool     less = s.r1a != q.r1a ? s.r1a < q.r1a : (s.r1b != q.r1b ? s.r1b < q.r1b : s.unitig < q.unitig);
          if (!less) break;
          st[p] = q;
          --p;
        }
        st[p] = s;
        ++n;
      }
      path_off[d * (cap + 1)] = 0;
      if (overflow) break;
      if (n == 0) continue; // mpp.cpp:151-153

      // ---- the dynamic programme (mpp.cpp:177-199)
      for (int i = 0; i < n; ++i) {
        pop[i]  = st[i].score;
        pred[i] = -1;
      }
      for (int k = 0; k + 1 < n; ++k) {
        for (int l = k + 1; l < n; ++l) {
          const double score = __dadd_rn(pop[k], st[l].score);
          if (!(score > pop[l])) continue; // evaluated second in the reference; it has no side effects
          // checkCompatibility(vStart[k], vStart[l]) — mpp.cpp:38-142
          int    o1, o2;
          double d1, d2;
          bool   abort = nano_check(st[k].c1a, st[k].c1b, st[l].c1a, st[l].c1b, st[k].r1a, st[k].r1b, st[l].r1a,
                                    st[l].r1b, &o1, &d1);
          abort |= nano_check(st[k].c2a, st[k].c2b, st[l].c2a, st[l].c2b, st[k].r2a, st[k].r2b, st[l].r2a, st[l].r2b,
                              &o2, &d2);
          if (abort) continue;
          if (!direction) o2 = -o2; // exchange_if(orientation2, -orientation2, !edgeMatch1->direction)
          bool matching = false;
          if (o1 == o2 && o1 != 0) {
            const double mx = d1 > d2 ? d1 : d2, mn = d1 < d2 ? d1 : d2; // std::max / std::min
            const double df = __dsub_rn(mx, mn);
            matching = (df <= double(a.wiggle)) || (__ddiv_rn(__dmul_rn(df, 100.0), mx) <= 15.0);
          } else if ((o1 < 0 && o2 < 0) || (o1 > 0 && o2 > 0)) {
            matching = __dadd_rn(d1, d2) <= double(a.wiggle);
          }
          if (matching) {
            pop[l]  = score;
            pred[l] = k;
          }
        }
      }

      // ---- best path (mpp.cpp:201-217), then the disjoint runners-up above 75 % (mpp.cpp:219-243)
      double max_val = 0.0;
      int    max_it  = 0;
      for (int i = 0; i < n; ++i)
        if (pop[i] > max_val) {
          max_it  = i;
          max_val = pop[i];
        }
      for (int i = 0; i < (n + 63) / 64; ++i) used[i] = 0ull; // population entries on any accepted path
      int nr = 0;
      {
        bool prim = false;
        int  len  = 0;
        for (int i = max_it; i >= 0; i = pred[i]) {
          prim |= st[i].prim;
          used[i >> 6] |= 1ull << (i & 63);
          ++len;
        }
        res[d * cap + nr++] = PathRes{max_it, max_val, prim || len > 2};
      }
      const double threshold = __dmul_rn(max_val, 0.75);
      for (int j = 0; j < n; ++j) {
        if (!(pop[j] > threshold)) continue;
        bool disjoint = true, prim = false;
        for (int i = j; i >= 0; i = pred[i]) {
          disjoint &= !((used[i >> 6] >> (i & 63)) & 1ull);
          prim |= st[i].prim;
        }
        if (!disjoint) continue;
        for (int i = j; i >= 0; i = pred[i]) used[i >> 6] |= 1ull << (i & 63);
        res[d * cap + nr++] = PathRes{j, pop[j], prim};
      }

      // ---- a single primary path may still be a shadow (mpp.cpp:245-302)
      if (nr == 1 && res[d * cap].prim) {
        const uint32_t s0 = a.rm_off[v1], s1 = a.rm_off[v1 + 1], e0 = a.rm_off[v2], e1 = a.rm_off[v2 + 1];
        auto start_id = [&](uint32_t i) { return a.rm_unitig[s0 + i]; };
        auto end_id   = [&](uint32_t i) { // vIDsEnd, reversed when !direction (mpp.cpp:273)
          return a.rm_unitig[direction ? e0 + i : e1 - 1 - i];
        };
        const uint32_t ns = s1 - s0, ne = e1 - e0;
        // the path in order: walk back from the tail, then read it forwards
        int plen = 0;
        for (int i = res[d * cap].tail; i >= 0; i = pred[i]) ++plen;
        uint32_t front_id = 0, back_id = st[res[d * cap].tail].unitig;
        for (int i = res[d * cap].tail; i >= 0; i = pred[i]) front_id = st[i].unitig;
        if ((start_id(0) != front_id && end_id(0) != front_id) ||
            (start_id(ns - 1) != back_id && end_id(ne - 1) != back_id)) {
          res[d * cap].prim = false;
        } else {
          uint32_t i = 0, j = 0;
          bool     is_shadow = false;
          for (int step = 0; step < plen && !is_shadow; ++step) {
            int node = res[d * cap].tail; // the step-th element of the path
            for (int back = plen - 1 - step; back > 0; --back) node = pred[node];
            const uint32_t id = st[node].unitig;
            uint32_t       ps = i;
            while (ps < ns && start_id(ps) != id) ++ps; // std::find_if from begin + i
            bool inter = ps > i;
            i          = ps + 1;
            uint32_t pe = j;
            while (pe < ne && end_id(pe) != id) ++pe;
            inter &= pe > j;
            j = pe + 1;
            is_shadow = inter;
          }
          if (is_shadow) res[d * cap].prim = false;
        }
      }
      n_res[d] = nr;

      // ---- freeze the paths of this direction as edge-match rows (the state arrays are reused by the other one)
      int w = 0;
      for (int r = 0; r < nr; ++r) {
        int len = 0;
        for (int i = res[d * cap + r].tail; i >= 0; i = pred[i]) ++len;
        int pos = w + len;
        for (int i = res[d * cap + r].tail; i >= 0; i = pred[i]) path_rows[d * cap + --pos] = st[i].k;
        w += len;
        path_off[d * (cap + 1) + r + 1] = w;
        res[d * cap + r].tail           = len; // from here on: the path length
      }
    }
    if (overflow) { // more edge matches of one direction than this instantiation holds: the fallback takes the edge
      a.huge_list[atomicAdd(a.n_huge, 1u)] = uint32_t(e);
      atomicAdd(a.huge_em, (unsigned long long)(k1 - k0));
      a.out.edge_n_orders[e] = 0;
      a.out.edge_shadow[e]   = 0;
      return;
    }

    // ---- chainingAndOverlaps' selection (main.cpp:355-397); keep[] marks the paths that survive
    bool *keep = S.keep; // keep[d * cap + r]
    bool  has_primary = false, has_multi = false;
    for (int d = 0; d < 2; ++d)
      for (int r = 0; r < n_res[d]; ++r) {
        keep[d * cap + r] = true;
        has_primary |= res[d * cap + r].prim;
      }
    if (has_primary)
      for (int d = 0; d < 2; ++d)
        for (int r = 0; r < n_res[d]; ++r) keep[d * cap + r] = res[d * cap + r].prim;
    for (int d = 0; d < 2; ++d)
      for (int r = 0; r < n_res[d]; ++r) has_multi |= keep[d * cap + r] && res[d * cap + r].tail > 1;
    if (has_multi)
      for (int d = 0; d < 2; ++d)
        for (int r = 0; r < n_res[d]; ++r) keep[d * cap + r] = keep[d * cap + r] && res[d * cap + r].tail > 1;
    int  combined = 0;
    bool first_prim = false, have_first = false;
    for (int d = 0; d < 2; ++d) // minus first: `!minusPaths.empty() ? minusPaths[0] : plusPaths[0]`
      for (int r = 0; r < n_res[d]; ++r)
        if (keep[d * cap + r]) {
          if (!have_first) {
            first_prim = res[d * cap + r].prim;
            have_first = true;
          }
          ++combined;
        }
    a.out.edge_shadow[e] = combined > 1 ? 1 : (first_prim ? 0 : 1);

    // ---- getOverlap for the minus paths, then the plus paths (main.cpp:399-413, ol.cpp:53-101)
    uint32_t n_ord = 0, id_w = 0;
    for (int d = 0; d < 2; ++d) {
      const bool direction = d == 1;
      for (int r = 0; r < n_res[d]; ++r) {
        if (!keep[d * cap + r]) continue;
        const int      len = res[d * cap + r].tail;
        const uint32_t *pr = path_rows + d * cap + path_off[d * (cap + 1) + r];
        const uint64_t kf = k0 + pr[0], kl = k0 + pr[len - 1]; // first / last id of the path
        double lf1, rf1, ll1, rl1, lf2, rf2, ll2, rl2; // overhangs of (first id | last id) on (first | second vertex)
        overhangs(a, v1, a.em_row1[kf], a.e.em_ov0[kf], a.e.em_ov1[kf], &lf1, &rf1);
        overhangs(a, v1, a.em_row1[kl], a.e.em_ov0[kl], a.e.em_ov1[kl], &ll1, &rl1);
        overhangs(a, v2, a.em_row2[kf], a.e.em_ov0[kf], a.e.em_ov1[kf], &lf2, &rf2);
        overhangs(a, v2, a.em_row2[kl], a.e.em_ov0[kl], a.e.em_ov1[kl], &ll2, &rl2);
        const double lo1 = lf1, ro1 = rl1;
        double       lo2 = lf2, ro2 = rl2;
        if (!direction) {
          lo2 = rf2;
          ro2 = ll2;
        }
        bool   have = true, start_first = true, contained = true;
        double left = 0.0, right = 0.0;
        if (lo1 <= lo2 && ro1 <= ro2) {
          left  = __dsub_rn(lo2, lo1);
          right = __dsub_rn(ro2, ro1);
        } else if (lo1 >= lo2 && ro1 >= ro2) {
          start_first = false;
          left        = __dsub_rn(lo1, lo2);
          right       = __dsub_rn(ro1, ro2);
        } else if (lo1 > lo2 && ro1 < ro2) {
          contained = false;
          left      = __dsub_rn(lo1, lo2);
          right     = __dsub_rn(ro2, ro1);
        } else if (lo1 < lo2 && ro1 > ro2) {
          start_first = false;
          contained   = false;
          left        = __dsub_rn(lo2, lo1);
          right       = __dsub_rn(ro1, ro2);
        } else {
          have = false; // NaNs only: the reference's optional stays empty
        }
        if (!have) continue;
        const uint64_t slot = k0 + n_ord;
        a.out.left[slot]  = left;
        a.out.right[slot] = right;
        a.out.score[slot] = double_to_size_t(res[d * cap + r].score);
        a.out.n_ids[slot] = uint32_t(len);
        a.out.flags[slot] = uint8_t((start_first ? 1 : 0) | (contained ? 2 : 0) | (direction ? 4 : 0) |
                                    (res[d * cap + r].prim ? 8 : 0));
        for (int i = 0; i < len; ++i) a.out.ids[k0 + id_w + i] = a.e.em_unitig[k0 + pr[i]];
        id_w += uint32_t(len);
        ++n_ord;
      }
    }
    a.out.edge_n_orders[e] = n_ord;
  }
}

// Edges with ONE edge match (the common case).  With a single candidate getMaxPairwisePaths' programme is the
// identity: the population is the match's own score (mpp.cpp:177-181), the maximum search starts from 0.0 so a
// non-positive score yields 0 (mpp.cpp:201-208), the only runner-up shares the entry and is dropped (mpp.cpp:219-243),
// the shadow test walks a path of length one (mpp.cpp:245-302), chainingAndOverlaps' filters keep the single path
// (main.cpp:355-397) and the order comes from one getOverlap (ol.cpp:53-101).  Everything else is handed on.
__global__ void __launch_bounds__(256) chain_single_kernel(ChainArgs a) {
  pdl_sync();
  for (uint64_t base = uint64_t(blockIdx.x) * blockDim.x; base < a.e.n_edges; base += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t e    = base + threadIdx.x;
    const bool     live = e < a.e.n_edges;
    uint64_t       k0 = 0, k1 = 0;
    if (live) {
      k0 = a.e.em_begin[e];
      k1 = a.e.em_begin[e + 1];
    }
    warp_append(live && k1 - k0 != 1, uint32_t(e), a.list_out, a.n_out);
    if (!live || k1 - k0 != 1) continue;
    const uint32_t v1 = a.e.first[e], v2 = a.e.second[e];
    const uint8_t  fl = a.e.em_flags[k0];
    const bool     direction = (fl & 1) != 0;
    bool           prim      = (fl & 2) != 0;
    const uint32_t id        = a.e.em_unitig[k0];
    const double   sc        = a.e.em_score[k0];
    const double   max_val   = sc > 0.0 ? sc : 0.0;
    if (prim) { // mpp.cpp:245-302 for a path of one id
      const uint32_t s0 = a.rm_off[v1], ns = a.rm_off[v1 + 1] - s0, e0 = a.rm_off[v2], ne = a.rm_off[v2 + 1] - e0;
      auto start_id = [&](uint32_t i) { return a.rm_unitig[s0 + i]; };
      auto end_id   = [&](uint32_t i) { return a.rm_unitig[direction ? e0 + i : e0 + ne - 1 - i]; };
      if ((start_id(0) != id && end_id(0) != id) || (start_id(ns - 1) != id && end_id(ne - 1) != id)) {
        prim = false;
      } else {
        uint32_t ps = 0, pe = 0;
        while (ps < ns && start_id(ps) != id) ++ps;
        while (pe < ne && end_id(pe) != id) ++pe;
        if (ps > 0 && pe > 0) prim = false;
      }
    }
    a.out.edge_shadow[e] = prim ? 0 : 1;
    const int ov0 = a.e.em_ov0[k0], ov1 = a.e.em_ov1[k0];
    double    l1, r1, l2, r2;
    overhangs(a, v1, a.em_row1[k0], ov0, ov1, &l1, &r1);
    overhangs(a, v2, a.em_row2[k0], ov0, ov1, &l2, &r2);
    const double lo1 = l1, ro1 = r1, lo2 = direction ? l2 : r2, ro2 = direction ? r2 : l2;
    bool   have = true, start_first = true, contained = true;
    double left = 0.0, right = 0.0;
    if (lo1 <= lo2 && ro1 <= ro2) {
      left  = __dsub_rn(lo2, lo1);
      right = __dsub_rn(ro2, ro1);
    } else if (lo1 >= lo2 && ro1 >= ro2) {
      start_first = false;
      left        = __dsub_rn(lo1, lo2);
      right       = __dsub_rn(ro1, ro2);
    } else if (lo1 > lo2 && ro1 < ro2) {
      contained = false;
      left      = __dsub_rn(lo1, lo2);
      right     = __dsub_rn(ro2, ro1);
    } else if (lo1 < lo2 && ro1 > ro2) {
      start_first = false;
      contained   = false;
      left        = __dsub_rn(lo2, lo1);
      right       = __dsub_rn(ro1, ro2);
    } else {
      have = false; // NaNs only
    }
    if (have) {
      a.out.left[k0]  = left;
      a.out.right[k0] = right;
      a.out.score[k0] = double_to_size_t(max_val);
      a.out.n_ids[k0] = 1;
      a.out.flags[k0] = uint8_t((start_first ? 1 : 0) | (contained ? 2 : 0) | (direction ? 4 : 0) | (prim ? 8 : 0));
      a.out.ids[k0]   = id;
    }
    a.out.edge_n_orders[e] = have ? 1u : 0u;
  }
}

// the edges of a list through the programme with kCap-entry arrays; longer edges are handed on (kHandOn)
template <int kCap, bool kHandOn> __global__ void __launch_bounds__(128) chain_list_kernel(ChainArgs a) {
  pdl_sync();
  const uint64_t n = *a.n_in;
  for (uint64_t base = uint64_t(blockIdx.x) * blockDim.x; base < n; base += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t x    = base + threadIdx.x;
    const bool     live = x < n;
    const uint32_t e    = live ? a.list_in[x] : 0u;
    bool           mine = live;
    if (kHandOn) {
      const bool big = live && a.e.em_begin[e + 1] - a.e.em_begin[e] > uint64_t(kCap);
      warp_append(big, e, a.list_out, a.n_out);
      mine = live && !big;
    }
    if (mine) chain_one_edge<kCap>(a, e);
  }
}

// the fallback: one thread per listed edge, state arrays sized by the edge in the scratch slab (slow — the dynamic
// programme is quadratic in the edge matches of a direction and runs on one thread — but nothing is refused)
__global__ void __launch_bounds__(64) chain_huge_kernel(ChainArgs a, uint32_t n_huge) {
  pdl_sync();
  for (uint64_t x = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; x < n_huge; x += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t e   = a.huge_list[x];
    const uint64_t n   = a.e.em_begin[e + 1] - a.e.em_begin[e];
    const uint64_t off = atomicAdd(a.huge_cursor, (unsigned long long)ChainStore<0>::bytes(n));
    chain_one_edge<0>(a, e, a.huge_scratch + off);
  }
}

void run_chain_overlaps(Arena &arena, cudaStream_t st, const ReadTable &reads, const UnitigTable &unitigs,
                        const MatchSoA &m, const EdgeTables &e, uint64_t wiggle_room, OrderTables &out, float *ms_out) {
  out = OrderTables{};
  if (!e.pair_rows || !reads.nlen)
    throw BuildError{MSGPU_ERR_NO_RESULT, "chaining needs a construction from PAF text by the product pairing variant"};
  const uint64_t M = m.n, R = reads.n, E = e.n_edges, EM = e.n_em;
  cudaEvent_t    ev0, ev1;
  MSGPU_CUDA(cudaEventCreate(&ev0));
  MSGPU_CUDA(cudaEventCreate(&ev1));
  MSGPU_CUDA(cudaEventRecord(ev0, st));

  // per-read ordered match lists
  uint32_t *deg = arena.alloc<uint32_t>(R + 1), *cursor = arena.alloc<uint32_t>(R + 1);
  uint32_t *rm_off = arena.alloc<uint32_t>(R + 1), *tmp = arena.alloc<uint32_t>(M), *list = arena.alloc<uint32_t>(M);
  uint32_t *d_tot = arena.alloc<uint32_t>(2);
  MSGPU_CUDA(cudaMemsetAsync(deg, 0, (R + 1) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(cursor, 0, (R + 1) * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(d_tot, 0, 2 * sizeof(uint32_t), st));
  const int gm = grid_for(M, 256);
  if (M) MSGPU_LAUNCH("read_degree", st, launch_k(read_degree_kernel, gm, 256, 0, st, m.read, M, deg));
  exclusive_scan<uint32_t>(arena, st, LoadAs<uint32_t, uint32_t>{deg}, R + 1, rm_off, d_tot);
  if (M) {
    MSGPU_LAUNCH("read_scatter", st, launch_k(read_scatter_kernel, gm, 256, 0, st, m.read, M, rm_off, cursor, tmp));
    MSGPU_LAUNCH("read_order", st, launch_k(read_order_kernel, gm, 256, 0, st, m, rm_off, tmp, list));
  }

  out.edge_shadow   = arena.alloc<uint8_t>(E);
  out.edge_n_orders = arena.alloc<uint32_t>(E);
  out.left          = arena.alloc<double>(EM);
  out.right         = arena.alloc<double>(EM);
  out.score         = arena.alloc<uint64_t>(EM);
  out.n_ids         = arena.alloc<uint32_t>(EM);
  out.ids           = arena.alloc<uint32_t>(EM);
  out.flags         = arena.alloc<uint8_t>(EM);
  // slots no order lands in stay zero, so that the host copy is deterministic
  MSGPU_CUDA(cudaMemsetAsync(out.left, 0, EM * sizeof(double), st));
  MSGPU_CUDA(cudaMemsetAsync(out.right, 0, EM * sizeof(double), st));
  MSGPU_CUDA(cudaMemsetAsync(out.score, 0, EM * sizeof(uint64_t), st));
  MSGPU_CUDA(cudaMemsetAsync(out.n_ids, 0, EM * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(out.ids, 0, EM * sizeof(uint32_t), st));
  MSGPU_CUDA(cudaMemsetAsync(out.flags, 0, EM, st));

  // match fields in pairing order, and the position of both vertices' matches for every edge match
  VmRow    *vm      = arena.alloc<VmRow>(M);
  uint32_t *em_row1 = arena.alloc<uint32_t>(EM), *em_row2 = arena.alloc<uint32_t>(EM);
  if (M) MSGPU_LAUNCH("build_vm_rows", st, launch_k(build_vm_rows_kernel, gm, 256, 0, st, m, e.pair_rows, vm));
  if (E) MSGPU_LAUNCH("chain_rows", st, launch_k(chain_rows_kernel, grid_for(E, 256, 64), 256, 0, st, e, unitigs.grp_off, e.p_read, em_row1, em_row2));

  if (E >= 0xffffffffull) throw BuildError{MSGPU_ERR_LIMIT, "chaining supports < 2^32 edges per device"};
  uint32_t     *multi = arena.alloc<uint32_t>(E), *big = arena.alloc<uint32_t>(E);
  unsigned int *d_cnt = arena.alloc<unsigned int>(4);
  MSGPU_CUDA(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned int), st));
  // an edge reaches the fallback only with more than kChainMax edge matches, so at most EM / (kChainMax + 1) edges do
  uint32_t           *huge_list = arena.alloc<uint32_t>(EM / (kChainMax + 1) + 1);
  unsigned long long *d_huge    = arena.alloc<unsigned long long>(4); // [0] edges listed [1] their edge matches [2] cursor
  MSGPU_CUDA(cudaMemsetAsync(d_huge, 0, 4 * sizeof(unsigned long long), st));

  ChainArgs a;
  a.read_nlen = reads.nlen;
  a.grp_off   = unitigs.grp_off;
  a.p_read    = e.p_read;
  a.vm        = vm;
  a.em_row1   = em_row1;
  a.em_row2   = em_row2;
  a.rm_off    = rm_off;
  a.rm_unitig = list;
  a.e         = e;
  a.wiggle    = wiggle_room;
  a.out       = out;
  a.huge_list    = huge_list;
  a.n_huge       = reinterpret_cast<unsigned int *>(d_huge);
  a.huge_em      = d_huge + 1;
  a.huge_cursor  = d_huge + 2;
  a.huge_scratch = nullptr;
  if (E) {
    a.list_in  = nullptr;
    a.n_in     = nullptr;
    a.list_out = multi;
    a.n_out    = d_cnt + 0;
    MSGPU_LAUNCH("chain_single", st, launch_k(chain_single_kernel, grid_for(E, 256, 64), 256, 0, st, a));
    // the lists are sized on the device; the grids are bounded by E and stride over what is there
    a.list_in  = multi;
    a.n_in     = d_cnt + 0;
    a.list_out = big;
    a.n_out    = d_cnt + 1;
    MSGPU_LAUNCH("chain_tiny", st, (launch_k(chain_list_kernel<kChainTiny, true>, grid_for(E, 128, 16), 128, 0, st, a)));
    a.list_in  = big;
    a.n_in     = d_cnt + 1;
    a.list_out = multi;
    a.n_out    = d_cnt + 2;
    MSGPU_LAUNCH("chain_small", st, (launch_k(chain_list_kernel<kChainSmall, true>, grid_for(E, 128, 16), 128, 0, st, a)));
    a.list_in  = multi;
    a.n_in     = d_cnt + 2;
    a.list_out = nullptr;
    a.n_out    = nullptr;
    MSGPU_LAUNCH("chain_large", st, (launch_k(chain_list_kernel<kChainMax, false>, grid_for(E, 128, 4), 128, 0, st, a)));
  }
  unsigned long long *h_huge = static_cast<unsigned long long *>(arena.host_scratch());
  MSGPU_CUDA(cudaMemcpyAsync(h_huge, d_huge, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  const uint32_t n_huge = uint32_t(h_huge[0] & 0xffffffffull);
  if (n_huge) { // rare: edges beyond the fast instantiations, one thread each over a scratch slab sized by the edges
    const uint64_t huge_em = h_huge[1];
    a.huge_scratch = arena.alloc<char>(152 * huge_em + 128 * uint64_t(n_huge));
    MSGPU_LAUNCH("chain_huge", st, launch_k(chain_huge_kernel, grid_for(n_huge, 64, 8), 64, 0, st, a, n_huge));
  }
  MSGPU_CUDA(cudaEventRecord(ev1, st));
  MSGPU_CUDA(cudaStreamSynchronize(st));
  if (ms_out) MSGPU_CUDA(cudaEventElapsedTime(ms_out, ev0, ev1));
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  out.valid = true;
}

} // namespace msgpu
