// ref_driver.cpp — runs the REFERENCE's own graph-construction path and dumps it.
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled (by oracle/Makefile) together
// with the reference's unmodified sources taken where they lie under
// /root/reference into oracle/_ref/libmsref.so and oracle/_ref/msref_cli.  No
// reference source is copied into this repository; this driver only *calls* the
// reference's public classes exactly the way its main() does
// (src/main.cpp:143-157 for construction, src/main.cpp:161-163 for the sequence
// index) and flattens what they hold into ms_tables.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load
// the result.

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <string>
#include <string_view>
#include <tuple>
#include <unordered_map>
#include <thread>
#include <vector>

#include "BlastFileAccessor.h"
#include "BlastFileReader.h"
#include "Kernel.h"
#include "Registry.h"
#include "SequenceAccessor.h"
#include "SequenceUtils.h"
#include "graph/Edge.h"
#include "graph/Graph.h"
#include "graph/Vertex.h"
#include "matching/MatchMap.h"
#include "threading/ThreadPool.h"
#include "types/Toggle.h"

#include "ms_tables.h"

#ifdef MSREF_WITH_GPU_SHIM
#include "ms_gpu_shim.h"
#endif

namespace {

using clk = std::chrono::steady_clock;
double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

template <class T> T *dup(std::vector<T> const &v) {
  T *p = static_cast<T *>(std::malloc(std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

struct NameTable {
  std::vector<std::string> names;
  void set(unsigned id, std::string const &s) {
    if (id >= names.size()) names.resize(id + 1);
    if (names[id].empty()) names[id] = s;
  }
  void flatten(uint64_t n, uint64_t **off, char **blob) const {
    std::vector<uint64_t> o(n + 1, 0);
    std::string           b;
    for (uint64_t i = 0; i < n; ++i) {
      o[i] = b.size();
      if (i < names.size()) b += names[i];
    }
    o[n]  = b.size();
    *off  = dup(o);
    *blob = static_cast<char *>(std::malloc(std::max<size_t>(1, b.size())));
    std::memcpy(*blob, b.data(), b.size());
  }
};

} // namespace

extern "C" {

void ms_tables_free(ms_tables *t) {
  if (!t) return;
  void *ptrs[] = {t->read_name_off, t->read_names,  t->read_nlen,  t->read_line,  t->unitig_name_off, t->unitig_names,
                  t->vm_read,       t->vm_unitig,   t->vm_nr0,     t->vm_nr1,     t->vm_ir0,          t->vm_ir1,
                  t->vm_rratio,     t->vm_dir,      t->vm_prim,    t->vm_score,   t->vm_line,         t->edge_first,
                  t->edge_second,   t->edge_em_begin, t->em_unitig, t->em_ov0,    t->em_ov1,          t->em_score,
                  t->em_dir,        t->em_prim,     t->em_line};
  for (void *p : ptrs) std::free(p);
  std::memset(t, 0, sizeof(*t));
}

// Flattens what the reference's objects hold into `out` (canonical orders of ms_tables.h).  Shared by the two ways of
// filling them: the reference's own read() + calculateEdges(), and the GPU path replayed through ms_gpu_shim.
static void dump_tables(const char *paf_path, std::size_t line_count, muchsalsa::graph::Graph &graph,
                        muchsalsa::matching::MatchMap &matchMap, muchsalsa::Registry &registryNanopore,
                        muchsalsa::Registry &registryIllumina, ms_tables *out) {
  {
    // ---- vertices, by id
    auto vertices = graph.getVertices();
    std::sort(vertices.begin(), vertices.end(), [](auto *a, auto *b) { return a->getId() < b->getId(); });
    std::vector<uint64_t> nlen, vline;
    for (auto *pV : vertices) {
      nlen.push_back(pV->getNanoporeLength());
      vline.push_back(pV->template getMetaDatum<std::size_t>(0));
    }
    out->read_nlen = dup(nlen);
    out->read_line = dup(vline);

    // ---- vertex matches, by (read, unitig)
    std::vector<uint32_t> vmRead, vmUnitig;
    std::vector<int32_t>  nr0, nr1, ir0, ir1;
    std::vector<double>   rr;
    std::vector<uint8_t>  vdir, vprim;
    std::vector<uint64_t> vscore, vmline;
    unsigned              maxUnitig = 0;
    bool                  anyVm     = false;
    for (auto *pV : vertices) {
      auto const *pVm = matchMap.getVertexMatches(pV->getId());
      if (!pVm) continue;
      std::vector<unsigned> ids;
      for (auto const &kv : *pVm) ids.push_back(kv.first);
      std::sort(ids.begin(), ids.end());
      for (auto u : ids) {
        auto const &m = *pVm->at(u);
        vmRead.push_back(pV->getId());
        vmUnitig.push_back(u);
        nr0.push_back(m.nanoporeRange.first);
        nr1.push_back(m.nanoporeRange.second);
        ir0.push_back(m.illuminaRange.first);
        ir1.push_back(m.illuminaRange.second);
        rr.push_back(m.rRatio);
        vdir.push_back(static_cast<bool>(m.direction) ? 1 : 0);
        vprim.push_back(static_cast<bool>(m.isPrimary) ? 1 : 0);
        vscore.push_back(m.score);
        vmline.push_back(m.lineNumber);
        maxUnitig = std::max(maxUnitig, u);
        anyVm     = true;
      }
    }
    out->n_vm      = vmRead.size();
    out->vm_read   = dup(vmRead);
    out->vm_unitig = dup(vmUnitig);
    out->vm_nr0    = dup(nr0);
    out->vm_nr1    = dup(nr1);
    out->vm_ir0    = dup(ir0);
    out->vm_ir1    = dup(ir1);
    out->vm_rratio = dup(rr);
    out->vm_dir    = dup(vdir);
    out->vm_prim   = dup(vprim);
    out->vm_score  = dup(vscore);
    out->vm_line   = dup(vmline);
    out->n_unitigs = anyVm ? maxUnitig + 1 : 0;

    // ---- names: the Registry has no reverse lookup, so ask it for the id of
    // the names on every accepted line (operator[] on a known name only reads).
    {
      NameTable     readNames, unitigNames;
      std::ifstream in(paf_path);
      std::string   line;
      std::vector<uint8_t> accepted(line_count, 0);
      for (auto l : vmline) accepted[l] = 1;
      for (std::size_t li = 0; std::getline(in, line); ++li) {
        if (li >= accepted.size() || !accepted[li]) continue;
        std::vector<std::string> tok;
        std::size_t              p = 0;
        while (tok.size() < 6) {
          auto q = line.find('\t', p);
          tok.push_back(line.substr(p, q == std::string::npos ? q : q - p));
          if (q == std::string::npos) break;
          p = q + 1;
        }
        if (tok.size() < 6) continue;
        unitigNames.set(registryIllumina[tok[0]], tok[0]);
        readNames.set(registryNanopore[tok[5]], tok[5]);
      }
      readNames.flatten(out->n_reads, &out->read_name_off, &out->read_names);
      unitigNames.flatten(out->n_unitigs, &out->unitig_name_off, &out->unitig_names);
    }

    // ---- edges by (first, second) and their edge matches by unitig
    auto edges = graph.getEdges();
    std::sort(edges.begin(), edges.end(), [](auto *a, auto *b) {
      auto va = a->getVertices(), vb = b->getVertices();
      return std::make_pair(va.first->getId(), va.second->getId()) <
             std::make_pair(vb.first->getId(), vb.second->getId());
    });
    std::vector<uint32_t> ef, es, emU;
    std::vector<uint64_t> eb, emLine;
    std::vector<int32_t>  ov0, ov1;
    std::vector<double>   emScore;
    std::vector<uint8_t>  emDir, emPrim;
    for (auto *pE : edges) {
      auto v = pE->getVertices();
      ef.push_back(v.first->getId());
      es.push_back(v.second->getId());
      eb.push_back(emU.size());
      auto const *pEm = matchMap.getEdgeMatches(pE);
      if (!pEm) continue;
      std::vector<unsigned> ids;
      for (auto const &kv : *pEm) ids.push_back(kv.first);
      std::sort(ids.begin(), ids.end());
      for (auto u : ids) {
        auto const &m = *pEm->at(u);
        emU.push_back(u);
        ov0.push_back(m.overlap.first);
        ov1.push_back(m.overlap.second);
        emScore.push_back(m.score);
        emDir.push_back(static_cast<bool>(m.direction) ? 1 : 0);
        emPrim.push_back(static_cast<bool>(m.isPrimary) ? 1 : 0);
        emLine.push_back(m.lineNumber);
      }
    }
    eb.push_back(emU.size());
    out->edge_first    = dup(ef);
    out->edge_second   = dup(es);
    out->edge_em_begin = dup(eb);
    out->n_em          = emU.size();
    out->em_unitig     = dup(emU);
    out->em_ov0        = dup(ov0);
    out->em_ov1        = dup(ov1);
    out->em_score      = dup(emScore);
    out->em_dir        = dup(emDir);
    out->em_prim       = dup(emPrim);
    out->em_line       = dup(emLine);
  }
}

// Runs BlastFileReader::read() + MatchMap::calculateEdges() on `paf_path` with a
// `threads`-sized pool (1 = the canonical, deterministic mode) and fills `out`.
// do_dump = 0 only measures the phase times (used by the CPU baseline).
int msref_build(const char *paf_path, int threads, int do_dump, ms_tables *out) {
  using namespace muchsalsa;
  std::memset(out, 0, sizeof(*out));
  try {
    auto pool             = threading::ThreadPool(static_cast<std::size_t>(threads));
    auto graph            = graph::Graph();
    auto matchMap         = matching::MatchMap(&pool, &graph);
    auto registryNanopore = Registry();
    auto registryIllumina = Registry();

    auto t0         = clk::now();
    auto bfAccessor = BlastFileAccessor(paf_path);
    auto t1         = clk::now();
    auto reader     = BlastFileReader(&pool, &bfAccessor, &graph, &matchMap, &registryNanopore, &registryIllumina);
    reader.read();
    auto t2 = clk::now();
    matchMap.calculateEdges();
    auto t3      = clk::now();
    out->t_index = secs(t0, t1);
    out->t_read  = secs(t1, t2);
    out->t_edges = secs(t2, t3);

    out->n_reads = graph.getOrder();
    out->n_edges = graph.getSize();
    if (!do_dump) {
      // still report the row counts the baseline rates are quoted on
      for (auto *pE : graph.getEdges()) {
        auto const *pEm = matchMap.getEdgeMatches(pE);
        if (pEm) out->n_em += pEm->size();
      }
      for (auto *pV : graph.getVertices()) {
        auto const *pVm = matchMap.getVertexMatches(pV->getId());
        if (pVm) out->n_vm += pVm->size();
      }
      return 0;
    }

    dump_tables(paf_path, bfAccessor.getLineCount(), graph, matchMap, registryNanopore, registryIllumina, out);
    return 0;
  } catch (std::exception const &e) {
    std::fprintf(stderr, "msref_build: %s\n", e.what());
    return 1;
  }
}

#ifdef MSREF_WITH_GPU_SHIM
// The drop-in check: the SAME reference objects (Graph, MatchMap, two Registry), filled by the GPU path through the
// host shim a maintainer would add to libms (muchsalsa_b200/host/ms_gpu_shim.cpp) instead of by read() +
// calculateEdges(), then dumped by the same code.  Returns 0, 1 (exception; message in err) or 2 (no GPU).
int msref_build_gpu(const char *paf_path, ms_tables *out, char *err, uint64_t err_cap) {
  using namespace muchsalsa;
  std::memset(out, 0, sizeof(*out));
  auto fail = [&](const char *m, int rc) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", m);
    return rc;
  };
  try {
    std::unique_ptr<gpu::Context> ctx;
    try {
      ctx = std::make_unique<gpu::Context>(0);
    } catch (std::exception const &e) {
      return fail(e.what(), 2);
    }
    auto pool             = threading::ThreadPool(1);
    auto graph            = graph::Graph();
    auto matchMap         = matching::MatchMap(&pool, &graph);
    auto registryNanopore = Registry();
    auto registryIllumina = Registry();
    auto t0               = clk::now();
    gpu::constructGraph(*ctx, paf_path, graph, matchMap, registryNanopore, registryIllumina);
    out->t_edges = secs(t0, clk::now());
    out->n_reads = graph.getOrder();
    out->n_edges = graph.getSize();
    msgpu_counts c{};
    msgpu_result_counts(ctx->get(), &c);
    dump_tables(paf_path, c.n_lines, graph, matchMap, registryNanopore, registryIllumina, out);
    return 0;
  } catch (std::exception const &e) {
    return fail(e.what(), 1);
  }
}

namespace {
void dump_orders(std::vector<muchsalsa::graph::Edge *> const &edges, ms_orders *out); // defined with msref_chain
}
// construction AND chaining through the shim (constructGraph + chainAndOverlaps), dumped like msref_chain
int msref_chain_gpu(const char *paf_path, uint64_t wiggle_room, ms_orders *out, char *err, uint64_t err_cap) {
  using namespace muchsalsa;
  std::memset(out, 0, sizeof(*out));
  auto fail = [&](const char *m, int rc) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", m);
    return rc;
  };
  try {
    std::unique_ptr<gpu::Context> ctx;
    try {
      ctx = std::make_unique<gpu::Context>(0);
    } catch (std::exception const &e) {
      return fail(e.what(), 2);
    }
    auto pool             = threading::ThreadPool(1);
    auto graph            = graph::Graph();
    auto matchMap         = matching::MatchMap(&pool, &graph);
    auto registryNanopore = Registry();
    auto registryIllumina = Registry();
    gpu::constructGraph(*ctx, paf_path, graph, matchMap, registryNanopore, registryIllumina);
    auto t0 = clk::now();
    gpu::chainAndOverlaps(*ctx, graph, static_cast<std::size_t>(wiggle_room));
    out->t_chain = secs(t0, clk::now());
    auto edges   = graph.getEdges();
    std::sort(edges.begin(), edges.end(), [](auto *a, auto *b) {
      auto va = a->getVertices(), vb = b->getVertices();
      return std::make_pair(va.first->getId(), va.second->getId()) <
             std::make_pair(vb.first->getId(), vb.second->getId());
    });
    dump_orders(edges, out);
    return 0;
  } catch (std::exception const &e) {
    return fail(e.what(), 1);
  }
}

// Construction, chaining AND the per-edge jobs by the GPU path through the host shim into the reference's own objects;
// dumped from the Edge objects like msref_filters does.
int msref_filters_gpu(const char *paf_path, uint64_t wiggle_room, ms_filters *out, char *err, uint64_t err_cap) {
  using namespace muchsalsa;
  std::memset(out, 0, sizeof(*out));
  auto fail = [&](const char *m, int rc) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", m);
    return rc;
  };
  try {
    std::unique_ptr<gpu::Context> ctx;
    try {
      ctx = std::make_unique<gpu::Context>(0);
    } catch (std::exception const &e) {
      return fail(e.what(), 2);
    }
    auto pool             = threading::ThreadPool(1);
    auto graph            = graph::Graph();
    auto matchMap         = matching::MatchMap(&pool, &graph);
    auto registryNanopore = Registry();
    auto registryIllumina = Registry();
    gpu::constructGraph(*ctx, paf_path, graph, matchMap, registryNanopore, registryIllumina);
    gpu::chainAndOverlaps(*ctx, graph, static_cast<std::size_t>(wiggle_room));
    std::unordered_map<graph::Edge const *, graph::EdgeOrder const *> contractionEdges;
    gpu::findContractionEdges(*ctx, graph, static_cast<std::size_t>(wiggle_room), contractionEdges);
    auto edges = graph.getEdges();
    std::sort(edges.begin(), edges.end(), [](auto *a, auto *b) {
      auto va = a->getVertices(), vb = b->getVertices();
      return std::make_pair(va.first->getId(), va.second->getId()) <
             std::make_pair(vb.first->getId(), vb.second->getId());
    });
    std::vector<int32_t> contraction;
    for (auto *pE : edges) { // before the orders are replaced: the map points into the unfiltered lists
      auto it = contractionEdges.find(pE);
      contraction.push_back(it == contractionEdges.end() ? -1 : static_cast<int32_t>(it->second - pE->getEdgeOrders().data()));
    }
    std::set<graph::Edge const *> deletableEdges;
    gpu::findDeletableEdges(*ctx, graph, deletableEdges);
    gpu::computeBitweights(*ctx, graph);
    std::vector<uint8_t>  deletable, consensus;
    std::vector<uint32_t> kept;
    std::vector<uint64_t> weight;
    for (auto *pE : edges) {
      deletable.push_back(deletableEdges.count(pE) ? 1 : 0);
      kept.push_back(static_cast<uint32_t>(pE->getEdgeOrders().size()));
      weight.push_back(pE->getWeight());
      consensus.push_back(static_cast<uint8_t>(pE->getConsensusDirection()));
    }
    out->n_edges           = edges.size();
    out->contraction_order = dup(contraction);
    out->deletable         = dup(deletable);
    out->n_kept_orders     = dup(kept);
    out->weight            = dup(weight);
    out->consensus         = dup(consensus);
    return 0;
  } catch (std::exception const &e) {
    return fail(e.what(), 1);
  }
}

// get{Illumina,Nanopore}Sequence through the shim, on a fresh context; returns length, -1 (reference-style exception)
// or -2 (no GPU / load failure)
int64_t msref_seq_fetch_gpu(const char *unitigs_path, const char *nanopore_path, int which, const char *name, int l,
                            int r, int dir, char *out, uint64_t cap) {
  using namespace muchsalsa;
  try {
    gpu::Context ctx(0);
    gpu::loadSequences(ctx, unitigs_path, nanopore_path);
    uint32_t id = 0;
    if (msgpu_seq_id(ctx.get(), static_cast<uint32_t>(which), name, &id) != MSGPU_OK) return -2;
    try {
      auto s = which ? gpu::getNanoporeSequence(ctx, id, l, r, Toggle(dir != 0))
                     : gpu::getIlluminaSequence(ctx, id, l, r, Toggle(dir != 0));
      if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
      return static_cast<int64_t>(s.size());
    } catch (std::out_of_range const &) {
      return -1;
    }
  } catch (std::exception const &) {
    return -2;
  }
}
#endif


// ---------------------------------------------------------------------------------------------------------------------
// The stage after construction: chainingAndOverlaps (src/main.cpp:328-414) for every edge.  The job body lives in the
// reference's main.cpp (which cannot be linked: it defines main()), so its control flow is restated here line by line
// around the reference's OWN getMaxPairwisePaths (mpp.cpp:145-305) and getOverlap (ol.cpp:53-101), which do all the
// arithmetic.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
void chain_one_edge(muchsalsa::matching::MatchMap &matchMap, muchsalsa::graph::Edge *pEdge, std::size_t wiggleRoom) {
  using namespace muchsalsa;
  std::vector<unsigned int> plusIDs, minusIDs;
  auto const *const         pEdgeMatches = matchMap.getEdgeMatches(pEdge); // main.cpp:334
  if (!pEdgeMatches) return;
  for (auto const &[illuminaId, edgeMatch] : *pEdgeMatches) { // main.cpp:341-349
    if (edgeMatch->direction) plusIDs.push_back(illuminaId); else minusIDs.push_back(illuminaId);
  }
  auto minusPaths = getMaxPairwisePaths(matchMap, *pEdge, minusIDs, false, wiggleRoom); // main.cpp:352-353
  auto plusPaths  = getMaxPairwisePaths(matchMap, *pEdge, plusIDs, true, wiggleRoom);
  auto isPrim  = [](auto const &p) { return std::get<2>(p); };
  auto isMulti = [](auto const &p) { return std::get<0>(p).size() > 1; };
  bool hasPrimary = std::any_of(plusPaths.begin(), plusPaths.end(), isPrim) || // main.cpp:355-361
                    std::any_of(minusPaths.begin(), minusPaths.end(), isPrim);
  if (hasPrimary) { // main.cpp:363-371
    plusPaths.erase(std::remove_if(plusPaths.begin(), plusPaths.end(), [&](auto const &p) { return !isPrim(p); }), plusPaths.end());
    minusPaths.erase(std::remove_if(minusPaths.begin(), minusPaths.end(), [&](auto const &p) { return !isPrim(p); }), minusPaths.end());
  }
  bool hasMulti = std::any_of(plusPaths.begin(), plusPaths.end(), isMulti) || // main.cpp:373-379
                  std::any_of(minusPaths.begin(), minusPaths.end(), isMulti);
  if (hasMulti) { // main.cpp:381-389
    plusPaths.erase(std::remove_if(plusPaths.begin(), plusPaths.end(), [&](auto const &p) { return !isMulti(p); }), plusPaths.end());
    minusPaths.erase(std::remove_if(minusPaths.begin(), minusPaths.end(), [&](auto const &p) { return !isMulti(p); }), minusPaths.end());
  }
  if (plusPaths.size() + minusPaths.size() > 1) { // main.cpp:391-397
    pEdge->setShadow(true);
  } else {
    auto const &path = !minusPaths.empty() ? minusPaths[0] : plusPaths[0];
    pEdge->setShadow(!std::get<2>(path));
  }
  for (auto &p : minusPaths) { // main.cpp:399-405
    auto o = getOverlap(matchMap, std::get<0>(p), *pEdge, false, std::get<1>(p), std::get<2>(p));
    if (o.has_value()) pEdge->appendOrder(std::move(o.value()));
  }
  for (auto &p : plusPaths) { // main.cpp:407-413
    auto o = getOverlap(matchMap, std::get<0>(p), *pEdge, true, std::get<1>(p), std::get<2>(p));
    if (o.has_value()) pEdge->appendOrder(std::move(o.value()));
  }
}
} // namespace

namespace {
// flattens the shadow flags and EdgeOrder lists of `edges` (already in (first, second) order) into `out`
void dump_orders(std::vector<muchsalsa::graph::Edge *> const &edges, ms_orders *out) {
    std::vector<uint8_t>  shadow, sif, cont, dir, prim;
  std::vector<uint64_t> ob, score, ib;
  std::vector<double>   left, right;
  std::vector<uint32_t> ids;
  for (auto *pE : edges) {
    auto v = pE->getVertices();
    shadow.push_back(pE->isShadow() ? 1 : 0);
    ob.push_back(left.size());
    for (auto const &o : pE->getEdgeOrders()) {
      if (o.baseVertex != v.first) throw std::runtime_error("baseVertex is not the first vertex");
      sif.push_back(o.startVertex == v.first ? 1 : 0);
      left.push_back(o.leftOffset);
      right.push_back(o.rightOffset);
      cont.push_back(static_cast<bool>(o.isContained) ? 1 : 0);
      dir.push_back(static_cast<bool>(o.direction) ? 1 : 0);
      prim.push_back(static_cast<bool>(o.isPrimary) ? 1 : 0);
      score.push_back(o.score);
      ib.push_back(ids.size());
      for (auto u : o.ids) ids.push_back(u);
    }
  }
  ob.push_back(left.size());
  ib.push_back(ids.size());
  out->n_edges            = edges.size();
  out->edge_shadow        = dup(shadow);
  out->edge_order_begin   = dup(ob);
  out->n_orders           = left.size();
  out->ord_start_is_first = dup(sif);
  out->ord_left           = dup(left);
  out->ord_right          = dup(right);
  out->ord_contained      = dup(cont);
  out->ord_direction      = dup(dir);
  out->ord_primary        = dup(prim);
  out->ord_score          = dup(score);
  out->ord_ids_begin      = dup(ib);
  out->n_ids              = ids.size();
  out->ord_ids            = dup(ids);
}
} // namespace

void ms_orders_free(ms_orders *o) {
  if (!o) return;
  void *ptrs[] = {o->edge_shadow, o->edge_order_begin, o->ord_start_is_first, o->ord_left,  o->ord_right, o->ord_contained,
                  o->ord_direction, o->ord_primary,    o->ord_score,          o->ord_ids_begin, o->ord_ids};
  for (void *p : ptrs) std::free(p);
  std::memset(o, 0, sizeof(*o));
}

// construction with one thread (the canonical mode) + chainingAndOverlaps on every edge; edges in (first, second) order
int msref_chain(const char *paf_path, uint64_t wiggle_room, ms_orders *out) {
  using namespace muchsalsa;
  std::memset(out, 0, sizeof(*out));
  try {
    auto pool             = threading::ThreadPool(1);
    auto graph            = graph::Graph();
    auto matchMap         = matching::MatchMap(&pool, &graph);
    auto registryNanopore = Registry();
    auto registryIllumina = Registry();
    auto bfAccessor       = BlastFileAccessor(paf_path);
    auto reader = BlastFileReader(&pool, &bfAccessor, &graph, &matchMap, &registryNanopore, &registryIllumina);
    reader.read();
    matchMap.calculateEdges();
    auto edges = graph.getEdges();
    std::sort(edges.begin(), edges.end(), [](auto *a, auto *b) {
      auto va = a->getVertices(), vb = b->getVertices();
      return std::make_pair(va.first->getId(), va.second->getId()) <
             std::make_pair(vb.first->getId(), vb.second->getId());
    });
    auto t0 = clk::now();
    for (auto *pE : edges) chain_one_edge(matchMap, pE, static_cast<std::size_t>(wiggle_room));
    out->t_chain = secs(t0, clk::now());

    dump_orders(edges, out);
    return 0;
  } catch (std::exception const &e) {
    std::fprintf(stderr, "msref_chain: %s\n", e.what());
    return 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The per-edge jobs after chaining: findContractionEdges (src/main.cpp:416-460), findDeletableEdges (main.cpp:534-550)
// and computeBitweight (main.cpp:552-573).  Like chainingAndOverlaps their bodies live in main.cpp, so the control flow
// is restated around the reference's OWN sanityCheck (sc.cpp:29-88), Graph::getNeighbors / hasEdge / getVertex and
// Edge::replaceOrders / setWeight / setConsensusDirection.  The jobs in between (contraction bookkeeping and vertex
// deletion, main.cpp:462-531) are not run: they only remove edges and do not change the verdict on the others.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
int contraction_order_of(muchsalsa::graph::Graph &graph, muchsalsa::graph::Edge const *pEdge, std::size_t wiggleRoom) {
  using namespace muchsalsa;
  auto const &edgeOrders = pEdge->getEdgeOrders();
  for (auto const &order : edgeOrders) { // main.cpp:421
    if (order.isContained && order.isPrimary) {
      auto       isSane    = true;
      auto const neighbors = graph.getNeighbors(order.startVertex);
      auto const edges     = std::map<unsigned int, graph::Edge *>(std::begin(neighbors), std::end(neighbors));
      for (auto const &[targetId, pSubedge] : edges) { // main.cpp:427-448
        auto const *const pTargetVertex = graph.getVertex(targetId);
        if (targetId == order.endVertex->getId() || pSubedge->isShadow()) continue;
        isSane &= graph.hasEdge(std::make_pair(order.endVertex, pTargetVertex));
        if (!isSane) break;
        isSane &= sanityCheck(graph, *order.startVertex, *order.endVertex, *pTargetVertex, order, wiggleRoom);
        if (!isSane) break;
      }
      if (isSane) return static_cast<int>(&order - edgeOrders.data()); // main.cpp:450-457 stores &order
    }
  }
  return -1;
}
} // namespace

void ms_filters_free(ms_filters *f) {
  if (!f) return;
  void *ptrs[] = {f->contraction_order, f->deletable, f->n_kept_orders, f->weight, f->consensus};
  for (void *p : ptrs) std::free(p);
  std::memset(f, 0, sizeof(*f));
}

// construction (one thread) + chainingAndOverlaps + the three per-edge jobs; edges in (first, second) order
int msref_filters(const char *paf_path, uint64_t wiggle_room, ms_filters *out) {
  using namespace muchsalsa;
  std::memset(out, 0, sizeof(*out));
  try {
    auto pool             = threading::ThreadPool(1);
    auto graph            = graph::Graph();
    auto matchMap         = matching::MatchMap(&pool, &graph);
    auto registryNanopore = Registry();
    auto registryIllumina = Registry();
    auto bfAccessor       = BlastFileAccessor(paf_path);
    auto reader = BlastFileReader(&pool, &bfAccessor, &graph, &matchMap, &registryNanopore, &registryIllumina);
    reader.read();
    matchMap.calculateEdges();
    auto edges = graph.getEdges();
    std::sort(edges.begin(), edges.end(), [](auto *a, auto *b) {
      auto va = a->getVertices(), vb = b->getVertices();
      return std::make_pair(va.first->getId(), va.second->getId()) <
             std::make_pair(vb.first->getId(), vb.second->getId());
    });
    auto const wiggle = static_cast<std::size_t>(wiggle_room);
    auto       tc     = clk::now();
    for (auto *pE : edges) chain_one_edge(matchMap, pE, wiggle);
    out->t_chain = secs(tc, clk::now());

    std::vector<int32_t>  contraction;
    std::vector<uint8_t>  deletable, consensus;
    std::vector<uint32_t> kept;
    std::vector<uint64_t> weight;
    auto                  t0 = clk::now();
    for (auto *pE : edges) contraction.push_back(contraction_order_of(graph, pE, wiggle));
    for (auto *pE : edges) { // findDeletableEdges, main.cpp:534-550
      std::vector<graph::EdgeOrder> filteredOrders;
      std::remove_copy_if(std::begin(pE->getEdgeOrders()), std::end(pE->getEdgeOrders()),
                          std::back_inserter(filteredOrders), [](auto const &order) { return order.isContained; });
      deletable.push_back(filteredOrders.empty() ? 1 : 0);
      pE->replaceOrders(std::move(filteredOrders));
    }
    for (auto *pE : edges) { // computeBitweight, main.cpp:552-573
      auto const &orders = pE->getEdgeOrders();
      if (!orders.empty()) {
        if (pE->isShadow()) {
          auto const initial   = orders[0].direction;
          auto const findOther = std::find_if(orders.begin(), orders.end(), [&](auto const &order) {
                                   return order.direction != initial;
                                 }) != orders.end();
          if (!findOther) pE->setConsensusDirection(orders[0].direction);
        } else {
          pE->setWeight(orders[0].score);
          pE->setConsensusDirection(orders[0].direction);
        }
      }
      kept.push_back(static_cast<uint32_t>(orders.size()));
      weight.push_back(pE->getWeight());
      consensus.push_back(static_cast<uint8_t>(pE->getConsensusDirection()));
    }
    out->t_filters         = secs(t0, clk::now());
    out->n_edges           = edges.size();
    out->contraction_order = dup(contraction);
    out->deletable         = dup(deletable);
    out->n_kept_orders     = dup(kept);
    out->weight            = dup(weight);
    out->consensus         = dup(consensus);
    return 0;
  } catch (std::exception const &e) {
    std::fprintf(stderr, "msref_filters: %s\n", e.what());
    return 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sequence path: strSlice / getReverseComplement (SequenceUtils.cpp:27-61) and the FASTA/FASTQ store
// (SequenceAccessor.cpp:54-231) + get{Illumina,Nanopore}Sequence (SequenceUtils.cpp:63-85).
// ---------------------------------------------------------------------------------------------------------------------

// returns the slice length, or -1 if the reference threw (std::out_of_range from substr)
int64_t msref_strslice(const char *s, uint64_t n, int i, int j, char *out) {
  try {
    auto v = muchsalsa::strSlice(std::string_view(s, n), i, j);
    std::memcpy(out, v.data(), v.size());
    return static_cast<int64_t>(v.size());
  } catch (std::exception const &) {
    return -1;
  }
}

void msref_revcomp(const char *s, uint64_t n, char *out) {
  auto r = muchsalsa::getReverseComplement(std::string(s, n));
  std::memcpy(out, r.data(), r.size());
}

struct msref_seq {
  std::unique_ptr<muchsalsa::threading::ThreadPool> pool;
  muchsalsa::Registry                               regNanopore, regIllumina;
  std::unique_ptr<muchsalsa::SequenceAccessor>      acc;
};

// Opens the two sequence files and builds the offset index like src/main.cpp:161-163.  When `prime_*` name lists are
// given (newline separated) they are registered first, in order — that is what ingest has done to the registries by
// the time main() reaches the sequence files.
msref_seq *msref_seq_open(const char *nanopore_path, const char *unitigs_path, const char *prime_reads,
                          const char *prime_unitigs) {
  try {
    auto h  = std::make_unique<msref_seq>();
    h->pool = std::make_unique<muchsalsa::threading::ThreadPool>(1);
    auto prime = [](muchsalsa::Registry &reg, const char *names) {
      if (!names) return;
      std::string_view sv(names);
      while (!sv.empty()) {
        auto q = sv.find('\n');
        auto t = sv.substr(0, q);
        if (!t.empty()) reg[std::string(t)];
        if (q == std::string_view::npos) break;
        sv.remove_prefix(q + 1);
      }
    };
    prime(h->regNanopore, prime_reads);
    prime(h->regIllumina, prime_unitigs);
    h->acc = std::make_unique<muchsalsa::SequenceAccessor>(h->pool.get(), nanopore_path, unitigs_path, &h->regNanopore,
                                                           &h->regIllumina);
    h->acc->buildIndex();
    return h.release();
  } catch (std::exception const &e) {
    std::fprintf(stderr, "msref_seq_open: %s\n", e.what());
    return nullptr;
  }
}

void msref_seq_close(msref_seq *h) { delete h; }

// id the registry holds for `name` (which: 0 = unitig/illumina, 1 = nanopore)
uint32_t msref_seq_id(msref_seq *h, int which, const char *name) {
  return which ? h->regNanopore[name] : h->regIllumina[name];
}

// whole record (whitespace stripped); returns length or -1 when the reference threw
int64_t msref_seq_full(msref_seq *h, int which, uint32_t id, char *out, uint64_t cap) {
  try {
    auto s = which ? h->acc->getNanoporeSequence(id) : h->acc->getIlluminaSequence(id);
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return static_cast<int64_t>(s.size());
  } catch (std::exception const &) {
    return -1;
  }
}

// get{Illumina,Nanopore}Sequence(acc, id, l, r, dir); returns length or -1 when the reference threw
int64_t msref_seq_fetch(msref_seq *h, int which, uint32_t id, int l, int r, int dir, char *out, uint64_t cap) {
  try {
    auto s = which ? muchsalsa::getNanoporeSequence(*h->acc, id, l, r, muchsalsa::Toggle(dir != 0))
                   : muchsalsa::getIlluminaSequence(*h->acc, id, l, r, muchsalsa::Toggle(dir != 0));
    if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
    return static_cast<int64_t>(s.size());
  } catch (std::exception const &) {
    return -1;
  }
}

// The reference's own fetch loop, timed: request i = (which, id, left, right, direction) as five int32; `threads`
// std::threads split the batch the way main()'s pool threads call get*Sequence concurrently (src/main.cpp:645-653;
// the accessor serialises them on one mutex per file, SequenceAccessor.cpp:56,131-137).  Returns the bases fetched.
uint64_t msref_seq_bench(msref_seq *h, const int32_t *reqs, uint64_t n, int threads, double *seconds) {
  if (threads < 1) threads = 1;
  std::vector<uint64_t>    bases(size_t(threads), 0);
  std::vector<std::thread> pool;
  auto                     t0 = clk::now();
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t] {
      uint64_t b = 0;
      for (uint64_t i = uint64_t(t); i < n; i += uint64_t(threads)) {
        const int32_t *q = reqs + 5 * i;
        try {
          auto s = q[0] ? muchsalsa::getNanoporeSequence(*h->acc, uint32_t(q[1]), q[2], q[3], muchsalsa::Toggle(q[4] != 0))
                        : muchsalsa::getIlluminaSequence(*h->acc, uint32_t(q[1]), q[2], q[3], muchsalsa::Toggle(q[4] != 0));
          b += s.size();
        } catch (std::exception const &) {
        }
      }
      bases[size_t(t)] = b;
    });
  }
  for (auto &th : pool) th.join();
  *seconds = secs(t0, clk::now());
  uint64_t total = 0;
  for (auto b : bases) total += b;
  return total;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// msref_cli: `msref_cli <paf> <threads>` prints one line of phase timings + counts.  It is a separate PROCESS on
// purpose: with threads >= 2 the reference races on GraphBase::m_adjacencyList (Graph.cpp:96-116 vs 212-230) and can
// segfault; bench.py reports that exit status faithfully instead of dying with it.
// ---------------------------------------------------------------------------------------------------------------------
#ifdef MSREF_CLI
int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s <paf> <threads>\n", argv[0]);
    return 2;
  }
  if (argc >= 6 && std::string(argv[1]) == "seq") { // `msref_cli seq <unitigs.fa> <nanopore.fx> <requests.bin> <threads>`
    msref_seq *h = msref_seq_open(argv[3], argv[2], nullptr, nullptr);
    if (!h) return 3;
    std::FILE *f = std::fopen(argv[4], "rb");
    if (!f) return 4;
    std::vector<int32_t> reqs;
    int32_t              buf[5 * 1024];
    size_t               got;
    while ((got = std::fread(buf, sizeof(int32_t) * 5, 1024, f)) > 0) reqs.insert(reqs.end(), buf, buf + 5 * got);
    std::fclose(f);
    double   t     = 0;
    uint64_t bases = msref_seq_bench(h, reqs.data(), reqs.size() / 5, std::atoi(argv[5]), &t);
    std::printf("{\"threads\": %d, \"requests\": %llu, \"bases\": %llu, \"seconds\": %.6f}\n", std::atoi(argv[5]),
                (unsigned long long)(reqs.size() / 5), (unsigned long long)bases, t);
    msref_seq_close(h);
    return 0;
  }
  if (argc >= 4 && std::string(argv[3]) == "chain") { // `msref_cli <paf> 1 chain`: time the stages after construction
    ms_filters f;
    int        rc = msref_filters(argv[1], 300, &f);
    if (rc) return rc;
    std::printf("{\"threads\": 1, \"t_chain\": %.6f, \"t_filters\": %.6f, \"n_edges\": %llu}\n", f.t_chain, f.t_filters,
                (unsigned long long)f.n_edges);
    return 0;
  }
  ms_tables t;
  int       rc = msref_build(argv[1], std::atoi(argv[2]), 0, &t);
  if (rc) return rc;
  std::printf("{\"threads\": %d, \"t_index\": %.6f, \"t_read\": %.6f, \"t_edges\": %.6f, \"n_reads\": %llu, "
              "\"n_vm\": %llu, \"n_edges\": %llu, \"n_em\": %llu}\n",
              std::atoi(argv[2]), t.t_index, t.t_read, t.t_edges, (unsigned long long)t.n_reads,
              (unsigned long long)t.n_vm, (unsigned long long)t.n_edges, (unsigned long long)t.n_em);
  return 0;
}
#endif
