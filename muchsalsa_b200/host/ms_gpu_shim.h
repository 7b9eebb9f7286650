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
#include <vector>

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

// The sequence fetches of ONE assemblePath call as ONE batch.  assemblePath issues its fetches one
// by one (ap.cpp:352-579: getSequenceLeft/RightOfAnchor, getAnchorSequence, getSequenceBetweenAnchors; :1288-1301 the
// contain elements), and what it does with a fetched sequence depends on its LENGTH only (append, slice, reverse
// complement, write) — never on its bytes.  So the call is run twice:
//   record  with the getters answering placeholder strings of exactly the length the reference would return (the
//           lengths come from the index, no sequence byte is touched) while every request is logged; the pass writes
//           to a null OutputWriter and a scratch Id2OverlapMap, its other arguments are const;
//   fetch   ONE msgpu_seq_fetch_batch for everything the path needs;
//   replay  with the getters handing out the fetched sequences, looked up by request (assemblePath iterates containers
//           keyed by pointers to objects it allocates itself, so two passes may ask in a different order) — this pass
//           produces the output.
// While a batch is active on a thread (thread-local, like the reference's per-job state), the two getters of
// SequenceUtils.cpp:63-85 forward to it: the two-line change INTEGRATION.md §6 shows.
class SequenceBatch {
public:
  explicit SequenceBatch(Context &ctx) : m_ctx(ctx) {}
  ~SequenceBatch();
  SequenceBatch(SequenceBatch const &)            = delete;
  SequenceBatch &operator=(SequenceBatch const &) = delete;

  static SequenceBatch *active(); // the batch the calling thread is recording / replaying, or nullptr
  template <class Fn> void record(Fn &&pass) {
    begin(Mode::e_RECORD);
    try {
      pass();
    } catch (...) {
      end();
      throw;
    }
    end();
  }
  void fetch(); // one msgpu_seq_fetch_batch for the recorded requests
  template <class Fn> void replay(Fn &&pass) {
    begin(Mode::e_REPLAY);
    try {
      pass();
    } catch (...) {
      end();
      throw;
    }
    end();
  }
  // what the getters call while the batch is active (which: 0 = unitig / illumina, 1 = nanopore)
  std::string get(unsigned int which, unsigned int id, int left, int right, muchsalsa::Toggle const direction);
  std::size_t size() const { return m_requests.size(); }
  std::size_t misses() const { return m_misses; } // replayed requests that had not been recorded (fetched singly)

private:
  enum class Mode { e_IDLE, e_RECORD, e_REPLAY };
  void begin(Mode mode);
  void end();

  Context                  &m_ctx;
  Mode                      m_mode = Mode::e_IDLE;
  std::vector<msgpu_seq_req> m_requests;
  std::vector<uint64_t>     m_offsets;
  std::vector<uint8_t>      m_status;
  std::string               m_data;
  std::size_t               m_cursor = 0, m_misses = 0;
  std::unordered_map<std::string, std::size_t> m_index; // request -> position in the batch
  static std::string keyOf(msgpu_seq_req const &req);
};

} // namespace muchsalsa::gpu
