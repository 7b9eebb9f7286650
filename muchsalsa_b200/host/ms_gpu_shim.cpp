// ms_gpu_shim.cpp — see ms_gpu_shim.h.  Replays the flat result of libmsgpu into the reference's objects.
#include "ms_gpu_shim.h"

#include <memory>
#include <stdexcept>
#include <vector>

#include "graph/Edge.h"
#include "graph/Vertex.h"

namespace muchsalsa::gpu {

namespace {

[[noreturn]] void raise(msgpu_ctx *pCtx, int status) {
  std::string msg = msgpu_last_error(pCtx);
  switch (status) {
  case MSGPU_ERR_IO:
    throw std::runtime_error(msg.empty() ? "Can't open blast file." : msg); // BlastFileAccessor.cpp:44
  case MSGPU_ERR_INVALID_PAF:
    throw std::runtime_error("Invalid BLAST file."); // BlastFileReader.cpp:98
  case MSGPU_ERR_OUT_OF_RANGE:
    throw std::out_of_range("basic_string_view::substr"); // what strSlice's substr throws, SequenceUtils.cpp:37
  default:
    throw std::runtime_error("msgpu: " + msg + " (status " + std::to_string(status) + ")");
  }
}

std::string fetchOne(Context &ctx, uint32_t which, unsigned int id, int left, int right, Toggle const direction) {
  msgpu_seq_req req{which, id, left, right, static_cast<bool>(direction) ? 1u : 0u};
  uint64_t      bound = 0;
  if (int rc = msgpu_seq_batch_bound(ctx.get(), &req, 1, &bound); rc != MSGPU_OK) raise(ctx.get(), rc);
  std::string out(bound, '\0');
  uint64_t    off[2] = {0, 0};
  uint8_t     status = 0;
  if (int rc = msgpu_seq_fetch_batch(ctx.get(), &req, 1, out.data(), bound, off, &status); rc != MSGPU_OK)
    raise(ctx.get(), rc);
  if (status != 0) raise(ctx.get(), status == 1 ? MSGPU_ERR_IO : MSGPU_ERR_OUT_OF_RANGE);
  out.resize(off[1] - off[0]);
  return out;
}

} // namespace

Context::Context(int cudaDevice) {
  if (msgpu_create(cudaDevice, &m_pCtx) != MSGPU_OK || !m_pCtx)
    throw std::runtime_error("msgpu: no usable CUDA device (the GPU construction path has no CPU fallback)");
}

Context::~Context() { msgpu_destroy(m_pCtx); }

void constructGraph(Context &ctx, std::string const &pafPath, graph::Graph &graph, matching::MatchMap &matchMap,
                    Registry &registryNanopore, Registry &registryIllumina) {
  if (int rc = msgpu_build_from_paf(ctx.get(), pafPath.c_str()); rc != MSGPU_OK) raise(ctx.get(), rc);
  msgpu_counts c{};
  msgpu_tables t{};
  if (int rc = msgpu_result_counts(ctx.get(), &c); rc != MSGPU_OK) raise(ctx.get(), rc);
  if (int rc = msgpu_result_host(ctx.get(), &t); rc != MSGPU_OK) raise(ctx.get(), rc);

  // Registries: ids are ranks of first appearance, so registering the names in id order reproduces them
  // (Registry.cpp:36-45 hands out 0, 1, 2, ... in call order).
  for (uint64_t i = 0; i < c.n_reads; ++i)
    registryNanopore[std::string(t.read_names + t.read_name_off[i], t.read_names + t.read_name_off[i + 1])];
  for (uint64_t i = 0; i < c.n_unitigs; ++i)
    registryIllumina[std::string(t.unitig_names + t.unitig_name_off[i], t.unitig_names + t.unitig_name_off[i + 1])];

  // Vertices: Vertex(id, nanoporeLength, lineIdx) exactly as BlastFileReader.cpp:113-114 builds them.
  std::vector<graph::Vertex const *> vertices(c.n_reads, nullptr);
  for (uint64_t i = 0; i < c.n_reads; ++i) {
    graph.addVertex(std::make_shared<graph::Vertex>(static_cast<unsigned int>(i),
                                                    static_cast<std::size_t>(t.read_nlen[i]),
                                                    static_cast<std::size_t>(t.read_line[i])));
    vertices[i] = graph.getVertex(static_cast<unsigned int>(i));
  }

  // Vertex matches (BlastFileReader.cpp:116-126, MatchMap.cpp:53-81).
  for (uint64_t k = 0; k < c.n_matches; ++k) {
    auto spMatch = std::shared_ptr<matching::VertexMatch>(new matching::VertexMatch{
        {t.m_nr0[k], t.m_nr1[k]},
        {t.m_ir0[k], t.m_ir1[k]},
        t.m_rratio[k],
        Toggle((t.m_flags[k] & 1) != 0),
        static_cast<std::size_t>(t.m_score[k]),
        Toggle((t.m_flags[k] & 2) != 0),
        static_cast<std::size_t>(t.m_line[k])});
    matchMap.addVertexMatch(t.m_read[k], t.m_unitig[k], spMatch);
  }

  // Edges + edge matches (MatchMap.cpp:204-218): edge.first is the vertex with the smaller line index = smaller id.
  for (uint64_t e = 0; e < c.n_edges; ++e) {
    auto pair = std::make_pair(gsl::make_not_null(vertices[t.edge_first[e]]),
                               gsl::make_not_null(vertices[t.edge_second[e]]));
    graph.addEdge(std::pair<gsl::not_null<graph::Vertex const *> const, gsl::not_null<graph::Vertex const *> const>(
        pair.first, pair.second));
    auto const *pEdge = graph.getEdge(pair);
    for (uint64_t k = t.edge_em_begin[e]; k < t.edge_em_begin[e + 1]; ++k) {
      auto spMatch = std::shared_ptr<matching::EdgeMatch>(new matching::EdgeMatch{
          {t.em_ov0[k], t.em_ov1[k]},
          Toggle((t.em_flags[k] & 1) != 0),
          t.em_score[k],
          Toggle((t.em_flags[k] & 2) != 0),
          static_cast<std::size_t>(t.em_line[k])});
      matchMap.addEdgeMatch(pEdge, t.em_unitig[k], spMatch);
    }
  }
}

void chainAndOverlaps(Context &ctx, graph::Graph &graph, std::size_t wiggleRoom) {
  if (int rc = msgpu_chain_overlaps(ctx.get(), wiggleRoom); rc != MSGPU_OK) raise(ctx.get(), rc);
  msgpu_counts c{};
  msgpu_tables t{};
  msgpu_orders o{};
  if (int rc = msgpu_result_counts(ctx.get(), &c); rc != MSGPU_OK) raise(ctx.get(), rc);
  if (int rc = msgpu_result_host(ctx.get(), &t); rc != MSGPU_OK) raise(ctx.get(), rc);
  if (int rc = msgpu_orders_host(ctx.get(), &o); rc != MSGPU_OK) raise(ctx.get(), rc);
  for (uint64_t e = 0; e < c.n_edges; ++e) {
    auto const *const pFirst  = graph.getVertex(t.edge_first[e]);
    auto const *const pSecond = graph.getVertex(t.edge_second[e]);
    auto *const       pEdge =
        graph.getEdge(std::make_pair(gsl::make_not_null(pFirst), gsl::make_not_null(pSecond)));
    pEdge->setShadow(o.edge_shadow[e] != 0);
    uint64_t idAt = t.edge_em_begin[e]; // the ids of the edge's orders are concatenated from its first slot on
    for (uint32_t i = 0; i < o.edge_n_orders[e]; ++i) {
      uint64_t const slot        = t.edge_em_begin[e] + i;
      bool const     startIsFirst = (o.ord_flags[slot] & 1) != 0;
      std::vector<unsigned int> ids(o.ord_ids + idAt, o.ord_ids + idAt + o.ord_n_ids[slot]);
      idAt += o.ord_n_ids[slot];
      pEdge->appendOrder(graph::EdgeOrder{startIsFirst ? pFirst : pSecond,
                                          startIsFirst ? pSecond : pFirst,
                                          o.ord_left[slot],
                                          o.ord_right[slot],
                                          Toggle((o.ord_flags[slot] & 2) != 0),
                                          pFirst, // baseVertex is always the edge's first vertex (ol.cpp:81-97)
                                          static_cast<std::size_t>(o.ord_score[slot]),
                                          std::move(ids),
                                          Toggle((o.ord_flags[slot] & 4) != 0),
                                          Toggle((o.ord_flags[slot] & 8) != 0)});
    }
  }
}

namespace {
struct FilterView {
  msgpu_counts             c{};
  msgpu_tables             t{};
  msgpu_edge_filter_tables f{};
};

FilterView filterView(Context &ctx) {
  FilterView v;
  if (int rc = msgpu_result_counts(ctx.get(), &v.c); rc != MSGPU_OK) raise(ctx.get(), rc);
  if (int rc = msgpu_result_host(ctx.get(), &v.t); rc != MSGPU_OK) raise(ctx.get(), rc);
  if (int rc = msgpu_edge_filters_host(ctx.get(), &v.f); rc != MSGPU_OK) raise(ctx.get(), rc);
  return v;
}

// the edge of table row e if both vertices and the edge are still in the graph
graph::Edge *liveEdge(graph::Graph &graph, msgpu_tables const &t, uint64_t e) {
  auto const *const pFirst  = graph.getVertex(t.edge_first[e]);
  auto const *const pSecond = graph.getVertex(t.edge_second[e]);
  if (pFirst == nullptr || pSecond == nullptr) return nullptr;
  return graph.getEdge(std::make_pair(gsl::make_not_null(pFirst), gsl::make_not_null(pSecond)));
}
} // namespace

void findContractionEdges(Context &ctx, graph::Graph &graph, std::size_t wiggleRoom,
                          std::unordered_map<graph::Edge const *, graph::EdgeOrder const *> &contractionEdges) {
  if (int rc = msgpu_edge_filters(ctx.get(), wiggleRoom); rc != MSGPU_OK) raise(ctx.get(), rc);
  auto const v = filterView(ctx);
  for (uint64_t e = 0; e < v.c.n_edges; ++e) {
    if (v.f.contraction_order[e] < 0) continue;
    auto const *const pEdge = liveEdge(graph, v.t, e);
    if (pEdge == nullptr) continue;
    contractionEdges.insert({pEdge, &pEdge->getEdgeOrders()[static_cast<std::size_t>(v.f.contraction_order[e])]});
  }
}

void findDeletableEdges(Context &ctx, graph::Graph &graph, std::set<graph::Edge const *> &deletableEdges) {
  auto const v = filterView(ctx);
  for (uint64_t e = 0; e < v.c.n_edges; ++e) {
    auto *const pEdge = liveEdge(graph, v.t, e);
    if (pEdge == nullptr) continue;
    std::vector<graph::EdgeOrder> kept; // the slots with isContained clear, in their old order (main.cpp:536-538)
    kept.reserve(v.f.n_kept_orders[e]);
    for (auto const &order : pEdge->getEdgeOrders())
      if (!order.isContained) kept.push_back(order);
    if (v.f.deletable[e] != 0) deletableEdges.insert(pEdge);
    pEdge->replaceOrders(std::move(kept));
  }
}

void computeBitweights(Context &ctx, graph::Graph &graph) {
  auto const v = filterView(ctx);
  for (uint64_t e = 0; e < v.c.n_edges; ++e) {
    auto *const pEdge = liveEdge(graph, v.t, e);
    if (pEdge == nullptr) continue;
    if (v.f.consensus[e] == 'c') continue;          // untouched: weight 0, e_NONE (Edge.cpp:32-33)
    if (!pEdge->isShadow()) pEdge->setWeight(static_cast<std::size_t>(v.f.weight[e])); // main.cpp:568
    pEdge->setConsensusDirection(v.f.consensus[e] == 'a');
  }
}

void loadSequences(Context &ctx, std::string const &unitigsPath, std::string const &nanoporePath) {
  if (int rc = msgpu_seq_load(ctx.get(), unitigsPath.c_str(), nanoporePath.c_str()); rc != MSGPU_OK)
    raise(ctx.get(), rc);
}

std::string getIlluminaSequence(Context &ctx, unsigned int illuminaId, int illuminaOverlapLeft,
                                int illuminaOverlapRight, Toggle const direction) {
  return fetchOne(ctx, 0, illuminaId, illuminaOverlapLeft, illuminaOverlapRight, direction);
}

std::string getNanoporeSequence(Context &ctx, unsigned int nanoporeId, int nanoporeRegionLeft,
                                int nanoporeRegionRight, Toggle const direction) {
  return fetchOne(ctx, 1, nanoporeId, nanoporeRegionLeft, nanoporeRegionRight, direction);
}

// ---------------------------------------------------------------------------------------------------------------------
// SequenceBatch
// ---------------------------------------------------------------------------------------------------------------------
namespace {
thread_local SequenceBatch *t_pActiveBatch = nullptr;
}

std::string SequenceBatch::keyOf(msgpu_seq_req const &req) {
  return std::string(reinterpret_cast<char const *>(&req), sizeof(req)); // five 32-bit words, no padding
}

SequenceBatch *SequenceBatch::active() { return t_pActiveBatch; }

SequenceBatch::~SequenceBatch() {
  if (t_pActiveBatch == this) t_pActiveBatch = nullptr;
}

void SequenceBatch::begin(Mode mode) {
  if (t_pActiveBatch != nullptr) throw std::logic_error("SequenceBatch: another batch is active on this thread");
  m_mode   = mode;
  m_cursor = 0;
  if (mode == Mode::e_RECORD) m_requests.clear();
  t_pActiveBatch = this;
}

void SequenceBatch::end() {
  t_pActiveBatch = nullptr;
  m_mode         = Mode::e_IDLE;
}

void SequenceBatch::fetch() {
  auto const n = m_requests.size();
  m_offsets.assign(n + 1, 0);
  m_status.assign(n, 0);
  uint64_t bound = 0;
  if (int rc = msgpu_seq_batch_bound(m_ctx.get(), m_requests.data(), n, &bound); rc != MSGPU_OK) raise(m_ctx.get(), rc);
  m_data.assign(bound, '\0');
  m_index.clear();
  for (std::size_t i = 0; i < n; ++i) m_index.emplace(keyOf(m_requests[i]), i);
  if (n == 0) return;
  if (int rc = msgpu_seq_fetch_batch(m_ctx.get(), m_requests.data(), n, m_data.data(), bound, m_offsets.data(),
                                     m_status.data());
      rc != MSGPU_OK)
    raise(m_ctx.get(), rc);
}

std::string SequenceBatch::get(unsigned int which, unsigned int id, int left, int right, Toggle const direction) {
  msgpu_seq_req const req{which, id, left, right, static_cast<bool>(direction) ? 1u : 0u};
  if (m_mode == Mode::e_RECORD) {
    // the length the reference's strSlice(record, left, right + 1) would have (SequenceUtils.cpp:27-38), from the index
    uint64_t bytes = 0;
    if (int rc = msgpu_seq_batch_bound(m_ctx.get(), &req, 1, &bytes); rc != MSGPU_OK) raise(m_ctx.get(), rc);
    m_requests.push_back(req);
    return std::string(bytes, 'N');
  }
  if (m_mode != Mode::e_REPLAY) throw std::logic_error("SequenceBatch: get() outside record / replay");
  // Served by KEY, not by position: assemblePath walks containers keyed by pointers to objects it allocates itself
  // (its anchor graph), so two passes over the same path can ask for the same sequences in a different order.
  auto const it = m_index.find(keyOf(req));
  if (it == m_index.end()) { // not seen while recording (cannot happen for a pure function of the lengths): fetch it now
    ++m_misses;
    return fetchOne(m_ctx, which, id, left, right, direction);
  }
  auto const i = it->second;
  if (m_status[i] != 0) raise(m_ctx.get(), m_status[i] == 1 ? MSGPU_ERR_IO : MSGPU_ERR_OUT_OF_RANGE);
  return m_data.substr(m_offsets[i], m_offsets[i + 1] - m_offsets[i]);
}

} // namespace muchsalsa::gpu
