// ms_gpu_shim.h — the host side of the drop-in: what a MuCHSALSA maintainer adds to libms so that the existing
// `muchsalsa` binary, pipeline.sh and everything downstream of construction keep working unchanged.
//
// It replaces two statements of main() (src/main.cpp:156-157)
//
//     blastFileReader.read();          ->   muchsalsa::gpu::constructGraph(ctx, pafPath, graph, matchMap,
//     matchMap.calculateEdges();                                            registryNanopore, registryIllumina);
//
// and backs the two sequence getters of SequenceUtils.h:41-45 with the resident sequence store.  Everything is done
// through the reference's own PUBLIC API (Registry::operator[], Graph::addVertex / addEdge / getEdge / getVertex,
// MatchMap::addVertexMatch / addEdgeMatch), so observers, ownership and all downstream kernels see exactly the
// objects the CPU path would have built.  Errors come back as the reference's own exceptions (std::runtime_error with
// the reference's messages: BlastFileAccessor.cpp:44, BlastFileReader.cpp:98, SequenceAccessor.cpp:61,110).
//
// Compiles against the reference's headers (include/ms) and include/msgpu.h; links libmsgpu.so.  C++20 like libms.
#pragma once

#include <set>
#include <string>
#include <unordered_map>

#include "Registry.h"
#include "graph/Graph.h"
#include "matching/MatchMap.h"
#include "types/Toggle.h"

#include "msgpu.h"

namespace muchsalsa::gpu {

// RAII owner of one msgpu context (= one GPU).  Throws std::runtime_error if no CUDA device is usable: there is no
// CPU fallback behind this class — a caller that wants one keeps calling BlastFileReader::read() itself.
class Context {
public:
  explicit Context(int cudaDevice = 0);
  ~Context();
  Context(Context const &)            = delete;
  Context &operator=(Context const &) = delete;
  msgpu_ctx *get() const { return m_pCtx; }

private:
  msgpu_ctx *m_pCtx = nullptr;
};

// BlastFileAccessor(path) + BlastFileReader::read() + MatchMap::calculateEdges() (src/main.cpp:153-157).
void constructGraph(Context &ctx, std::string const &pafPath, muchsalsa::graph::Graph &graph,
                    muchsalsa::matching::MatchMap &matchMap, muchsalsa::Registry &registryNanopore,
                    muchsalsa::Registry &registryIllumina);

// The stage right after construction, for every edge of the graph built by constructGraph on `ctx`: replaces the
// ThreadPool loop over chainingAndOverlaps jobs (src/main.cpp:170-178, 328-414).  Sets Edge::setShadow and appends the
// EdgeOrders exactly as getMaxPairwisePaths (mpp.cpp) + getOverlap (ol.cpp) would.
void chainAndOverlaps(Context &ctx, muchsalsa::graph::Graph &graph, std::size_t wiggleRoom);

// The per-edge jobs that follow, computed on the GPU in one call from the orders chainAndOverlaps left in HBM, then
// applied at the three places main() runs them:
//   findContractionEdges   replaces the job loop of src/main.cpp:181-191 (job body 416-460, sanityCheck of sc.cpp):
//                          fills `contractionEdges` with {edge, &order} exactly like the reference's map.  Call it
//                          right after chainAndOverlaps, while the graph is complete; it runs msgpu_edge_filters.
//   findDeletableEdges     replaces src/main.cpp:247-256 (job body 534-550): drops the contained orders of every edge
//                          still in the graph (Edge::replaceOrders) and collects the edges left without orders.
//   computeBitweights      replaces src/main.cpp:263-270 (job body 552-573): Edge::setWeight / setConsensusDirection
//                          for every edge still in the graph.
// Vertices and edges the host deleted in between (main.cpp:193-245, 258-260) are skipped: the verdict on an edge
// depends on its own orders only.
void findContractionEdges(Context &ctx, muchsalsa::graph::Graph &graph, std::size_t wiggleRoom,
                          std::unordered_map<muchsalsa::graph::Edge const *, muchsalsa::graph::EdgeOrder const *> &contractionEdges);
void findDeletableEdges(Context &ctx, muchsalsa::graph::Graph &graph,
                        std::set<muchsalsa::graph::Edge const *> &deletableEdges);
void computeBitweights(Context &ctx, muchsalsa::graph::Graph &graph);

// SequenceAccessor(...) + buildIndex() (src/main.cpp:161-163): both files indexed and kept resident in HBM; ids
// continue the numbering of the construction that ran on `ctx`.
void loadSequences(Context &ctx, std::string const &unitigsPath, std::string const &nanoporePath);

// getIlluminaSequence / getNanoporeSequence (SequenceUtils.cpp:63-85); thread-safe like the reference's.
std::string getIlluminaSequence(Context &ctx, unsigned int illuminaId, int illuminaOverlapLeft,
                                int illuminaOverlapRight, muchsalsa::Toggle const direction);
std::string getNanoporeSequence(Context &ctx, unsigned int nanoporeId, int nanoporeRegionLeft,
                                int nanoporeRegionRight, muchsalsa::Toggle const direction);

} // namespace muchsalsa::gpu
