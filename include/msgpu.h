/* msgpu.h — C ABI of the B200 graph-construction path for MuCHSALSA.
 *
 * This is the drop-in boundary.  The reference (MuCHSALSA v0.0.2, a CPU-only C++20 program) has no plugin / FFI
 * interface; the seam is two statements of its main() plus four free functions.  Every entry point below names the
 * reference interface it replaces (file:line relative to the reference tree):
 *
 *   construction   src/main.cpp:156-157      blastFileReader.read(); matchMap.calculateEdges();
 *                  = BlastFileAccessor.cpp:40-91, BlastFileReader.cpp:72-130, Registry.cpp:36-45,
 *                    Graph.cpp:141-149,212-230,291-311, MatchMap.cpp:53-81,109-134,161-224
 *   sequences      include/ms/SequenceUtils.h:37-45 (strSlice, getReverseComplement, getIlluminaSequence,
 *                  getNanoporeSequence) over SequenceAccessor.cpp:54-85,114-139,143-231
 *
 * INTEGRATION.md shows the ~60-line C++ shim that replays a msgpu_host_result into the reference's own
 * Graph / MatchMap / Registry objects through their public API, so everything downstream is untouched.
 *
 * Conventions: plain pointers and sizes only; every function returns an msgpu_status (0 = ok) and never throws;
 * msgpu_last_error() gives the message for the last failing call on that context.  A context is bound to ONE CUDA
 * device and is used by one thread at a time for construction; msgpu_seq_fetch_batch() may be called concurrently
 * from many threads (the reference fetches sequences from pool threads, src/main.cpp:645-653).
 *
 * Determinism contract (= the reference run with one worker thread, its only deterministic mode):
 *   - the last line of the PAF is never parsed                                   (BlastFileReader.cpp:76)
 *   - a line is accepted iff matches >= 400 and c3 - c2 >= 400                   (BlastFileReader.cpp:106-107)
 *   - read ids / unitig ids = rank of the first accepted line naming them        (Registry.cpp:36-45)
 *   - a read's vertex carries the line index and read length of that first line  (Graph.cpp:148)
 *   - per (read, unitig) the accepted line with the smallest index is kept       (MatchMap.cpp:64-73)
 *   - per unitig every pair of kept matches whose unitig overlap is > 100 gives one edge match; an edge exists iff
 *     it has >= 1 edge match; edge.first is the read with the smaller id        (MatchMap.cpp:185-218)
 */
#ifndef MSGPU_H
#define MSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSGPU_ABI_VERSION 1

typedef enum msgpu_status {
  MSGPU_OK               = 0,
  MSGPU_ERR_INVALID_ARG  = 1,
  MSGPU_ERR_CUDA         = 2, /* a CUDA runtime call failed; message in msgpu_last_error */
  MSGPU_ERR_IO           = 3, /* "Can't open blast file." / "Can't open sequence file(s)." */
  MSGPU_ERR_INVALID_PAF  = 4, /* "Invalid BLAST file." (fewer than 10 fields) or a field std::stoi would reject */
  MSGPU_ERR_NOMEM        = 5,
  MSGPU_ERR_LIMIT        = 6, /* > 2^32-2 lines in one PAF buffer, or a shape the kernels do not support */
  MSGPU_ERR_NO_RESULT    = 7, /* result accessor called before a successful build */
  MSGPU_ERR_OUT_OF_RANGE = 8, /* sequence fetch: unknown id, or strSlice would throw std::out_of_range */
  MSGPU_ERR_INTERNAL     = 9
} msgpu_status;

typedef struct msgpu_ctx msgpu_ctx;

/* ---- context ------------------------------------------------------------------------------------------------------
 * replaces: ThreadPool construction (src/main.cpp:143-144) — the GPU is the worker pool of this path. */
int         msgpu_create(int cuda_device, msgpu_ctx **out_ctx);
void        msgpu_destroy(msgpu_ctx *ctx);
const char *msgpu_last_error(const msgpu_ctx *ctx);
int         msgpu_abi_version(void);
/* All work of the context is enqueued on `cuda_stream` (a cudaStream_t; NULL = the context's own non-blocking
 * stream — note that NULL is also the handle of the legacy default stream, which therefore cannot be selected: a
 * caller that wants its own work ordered with the context's passes an explicit stream). */
int         msgpu_set_stream(msgpu_ctx *ctx, void *cuda_stream);

/* ---- construction -------------------------------------------------------------------------------------------------
 * Each build replaces the context's previous result.
 *
 * msgpu_build_from_paf        : replaces BlastFileAccessor(path) + read() + calculateEdges() (src/main.cpp:153-157).
 * msgpu_build_from_paf_buffer : same, PAF text already in host memory (pinned memory avoids a staging copy).
 * msgpu_build_from_paf_device : same, PAF text already resident in HBM; nothing is copied to the host.
 */
int msgpu_build_from_paf(msgpu_ctx *ctx, const char *paf_path);
int msgpu_build_from_paf_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes);
int msgpu_build_from_paf_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes);

/* ---- the scrubber's anchor graph ---------------------------------------------------------------------
 * msgpu_anchor_graph_from_paf_*  replaces the loop of pipeline/scrubber_bfs.py:57-114 — the pre-filter that decides
 * which nanopore reads reach the assembler — on the same PAF columns: the same per-unitig read pairing as
 * calculateEdges, with the scrubber's rules instead of BlastFileReader's:
 *   - a line is line.rstrip().split("\t"); a line with ONE field is skipped; 2..8 fields, or a field 1 2 3 6 7 8 that
 *     Python's int() rejects, are MSGPU_ERR_INVALID_PAF (the script dies with IndexError / ValueError); digit-group
 *     underscores and values beyond 32 bits, which int() accepts, are reported as invalid too
 *   - EVERY line is read (no last-line drop); a line is accepted iff c3 - c2 >= 500 (scrubber_bfs.py:85)
 *   - reads (graph nodes) and anchors are numbered by first accepted appearance; read_nlen = field 6 of that line
 *   - per (read, anchor) the first accepted line is kept (scrubber_bfs.py:97-98); the match table holds these rows in
 *     LINE order with RAW ranges: m_nr0 / m_nr1 = (s_2, e_2) as the script stores them, m_ir0 / m_ir1 = (s_1, e_1),
 *     m_score = field 1 (the anchor's length), m_flags bit0 = strand '+'
 *   - an edge joins two reads kept in the same BLOCK of consecutive kept rows naming one anchor
 *     (scrubber_bfs.py:103-113); there is no overlap test
 * Result: the usual tables.  edge_first / edge_second are the graph's edges (each pair once, first < second).  The
 * edge-match rows list, per edge, the blocks that produced it (em_unitig = the block's anchor; the other columns carry
 * what the pairing kernels compute and mean nothing to the scrubber).  msgpu_chain_overlaps / msgpu_edge_filters do
 * not apply to this graph (MSGPU_ERR_NO_RESULT). */
int msgpu_anchor_graph_from_paf_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes);
int msgpu_anchor_graph_from_paf_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes);

/* ---- the unitig filter's coverage profiles -----------------------------------------------------------
 * Replaces the two passes of pipeline/unitig_filter.py over its PAF (the pre-filter that removes over-covered unitigs
 * and cuts the peaks out of the rest).  Rules of the script, not BlastFileReader's: every line counts (no accept rule,
 * no skipped line, no last-line drop); a line is line.rstrip().split("\t") and needs fields 0..6 with 1, 2, 3, 6
 * integers for Python's int() (else MSGPU_ERR_INVALID_PAF, as for positions outside the unitig: IndexError).
 *   msgpu_unitig_coverage_from_paf_*   pass 1 (unitig_filter.py:34-75).  Per BLOCK of consecutive lines naming one
 *                  unitig, in file order: block_unitig (id = rank of first appearance; names through
 *                  msgpu_result_host: unitig_names), block_n_reads (distinct reads), block_len (the unitig length of the
 *                  block's first line) and block_max_cov = the maximum over the unitig's positions of the number of
 *                  reads covering it, a read counting ONCE per block (its first line there; c2 .. c3 - 1 inclusive).
 *   msgpu_unitig_coverage_cut          pass 2 (unitig_filter.py:102-176) for the cut-offs the caller derives from pass 1
 *                  (:83-97): block_outlier[b] = 1 iff the maximum of the LAST block of b's unitig is > cutoff; such a
 *                  block is profiled again with ALL its lines and cut into the maximal runs of positions with coverage
 *                  <= q3 that are >= 500 long: fragments (frag_block, frag_start, frag_end inclusive), in file order.
 * Pointers are pinned host memory, valid until the next build on the context. */
typedef struct msgpu_coverage {
  uint64_t        n_blocks;
  const uint32_t *block_unitig, *block_n_reads, *block_len, *block_max_cov;
} msgpu_coverage;
typedef struct msgpu_fragments {
  uint64_t        n_blocks, n_fragments;
  const uint8_t  *block_outlier;
  const uint32_t *frag_block;
  const int32_t  *frag_start, *frag_end;
} msgpu_fragments;
int msgpu_unitig_coverage_from_paf_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes, msgpu_coverage *out);
int msgpu_unitig_coverage_from_paf_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes, msgpu_coverage *out);
int msgpu_unitig_coverage_cut(msgpu_ctx *ctx, double cutoff, double q3, msgpu_fragments *out);

/* Packed structure-of-arrays match records (the result of ingest, = the VertexMatch rows of MatchMap.h:51-59 plus
 * their ids), grouped by unitig and ordered by line inside a unitig.  Device pointers.
 * flags: bit0 = direction ('+'), bit1 = isPrimary. */
typedef struct msgpu_match_soa {
  uint64_t        n;
  const uint32_t *unitig; /* non-decreasing */
  const uint32_t *read;
  const int32_t  *ir0, *ir1; /* illuminaRange (inclusive) */
  const int32_t  *nr0, *nr1; /* nanoporeRange (inclusive) */
  const int32_t  *score;     /* number of matches (PAF column 10) */
  const uint32_t *line;      /* PAF line index */
  const uint8_t  *flags;
} msgpu_match_soa;

/* msgpu_build_from_records : pairing + keyed merge only (MatchMap::calculateEdges, MatchMap.cpp:161-224) on records
 * that are already packed in HBM.  n_reads / n_unitigs bound the ids.  The read table of the result is left empty. */
int msgpu_build_from_records(msgpu_ctx *ctx, const msgpu_match_soa *dev_records, uint32_t n_reads,
                             uint32_t n_unitigs);

/* ---- multi-GPU construction (one context = one GPU = one process; the caller moves bytes between ranks) ------------
 * The reference has no distributed mode (its only parallelism is one in-process ThreadPool, ThreadPool.cpp:38-91);
 * these entry points produce the reference's single-threaded result on the WHOLE PAF from N GPUs:
 *
 *   1. msgpu_shard_ingest_*   rank r parses the r-th contiguous chunk of the PAF (split at line boundaries; only the
 *                             last chunk drops its last line, BlastFileReader.cpp:76) -> chunk-local tables
 *   2. msgpu_shard_describe / msgpu_shard_export   the chunk-local tables as ONE flat device blob; the ranks
 *                             all-gather headers and blobs (NCCL over NVLink: torch.distributed.all_gather_into_tensor),
 *                             or msgpu_shard_export_peers writes the blob straight into every rank's buffer
 *   3. msgpu_build_from_shards   every rank merges all blobs into the global id space (ids = rank of first accepted
 *                             line over the whole file, keep-min-line across chunks, Registry.cpp:36-45,
 *                             MatchMap.cpp:64-73) and pairs ITS share `part` of `n_parts` of the edge rows.
 * After step 3 the read / unitig / match tables of the result are global (identical on every rank) and the edge /
 * edge-match tables are the rank's contiguous slice of the global (first, second)-sorted tables; concatenating the
 * slices in rank order gives exactly the single-GPU result (edge_em_begin is relative to the slice).
 * Only match records cross NVLink; candidate pairs (two orders of magnitude more) never leave their GPU. */
typedef struct msgpu_shard_header {
  uint64_t magic, n_lines, n_reads, n_unitigs, n_matches, read_name_bytes, unitig_name_bytes, blob_bytes;
} msgpu_shard_header;

int msgpu_shard_ingest_buffer(msgpu_ctx *ctx, const char *host_paf, size_t n_bytes, int is_last_shard);
int msgpu_shard_ingest_device(msgpu_ctx *ctx, const char *dev_paf, size_t n_bytes, int is_last_shard);
int msgpu_shard_describe(const msgpu_ctx *ctx, msgpu_shard_header *out);
/* writes blob_bytes bytes (see msgpu_shard_describe) to dev_blob, on the context's stream */
int msgpu_shard_export(msgpu_ctx *ctx, void *dev_blob, uint64_t capacity);
/* Export and all-gather in one kernel: writes the blob into dev_slots[0..n_slots) — this rank's slot of every rank's
 * receive buffer, device pointers that may be PEER-mapped memory of other GPUs (CUDA IPC / symmetric memory; the
 * stores to them travel over NVLink) — each `capacity` bytes, 16-byte aligned, n_slots <= 16.  Every unit of the blob
 * is read once and stored n_slots times.  The header (the blob's first 64 bytes) is always written; the sections only
 * if blob_bytes <= capacity, so after the caller's barrier every rank sees every size and all ranks reach the same
 * verdict about a slot that was too small.  Asynchronous on the context's stream; ordering ACROSS ranks (all exports
 * done before any rank reads its buffer; a buffer not overwritten while a peer still reads it) is the caller's:
 * muchsalsa_b200/multi_gpu.py uses two buffer sets and one barrier per step. */
int msgpu_shard_export_peers(msgpu_ctx *ctx, void *const *dev_slots, uint32_t n_slots, uint64_t capacity);
int msgpu_build_from_shards(msgpu_ctx *ctx, uint32_t n_shards, const msgpu_shard_header *headers,
                            const void *const *dev_blobs, uint32_t part, uint32_t n_parts);

typedef struct msgpu_counts {
  uint64_t n_lines;   /* lines in the buffer (BlastFileAccessor::getLineCount) */
  uint64_t n_reads;   /* graph order */
  uint64_t n_unitigs;
  uint64_t n_matches; /* kept VertexMatch rows */
  uint64_t n_pairs;   /* candidate pairs examined, sum over unitigs of k(k-1)/2 */
  uint64_t n_edges;   /* graph size */
  uint64_t n_edge_matches;
} msgpu_counts;

int msgpu_result_counts(const msgpu_ctx *ctx, msgpu_counts *out);

/* The result as flat arrays.  Row orders:
 *   reads / unitigs : by id
 *   matches         : grouped by unitig (ascending id), by line inside a unitig
 *   edges           : by (first, second), first < second
 *   edge matches    : by (first, second, unitig); edge e owns rows [edge_em_begin[e], edge_em_begin[e+1])
 * In msgpu_device_result every pointer is a device pointer, in msgpu_host_result a pointer into pinned host memory;
 * both stay valid until the next build on the context or msgpu_destroy. */
typedef struct msgpu_tables {
  /* Vertex(id, nanoporeLength, lineIdx) — Vertex.h:185-188 */
  const int32_t  *read_nlen;     /* as parsed by std::stoi; the reference widens it to size_t */
  const uint32_t *read_line;
  const uint64_t *read_name_off; /* n_reads + 1, into read_names */
  const char     *read_names;
  const uint64_t *unitig_name_off; /* n_unitigs + 1 */
  const char     *unitig_names;
  /* VertexMatch — MatchMap.h:51-59 */
  const uint32_t *m_unitig, *m_read;
  const int32_t  *m_nr0, *m_nr1, *m_ir0, *m_ir1;
  const double   *m_rratio;
  const int32_t  *m_score;  /* the reference widens it to size_t */
  const uint32_t *m_line;
  const uint8_t  *m_flags;  /* bit0 direction, bit1 isPrimary */
  /* Edge — Edge::getVertices() order */
  const uint32_t *edge_first, *edge_second;
  const uint64_t *edge_em_begin; /* n_edges + 1 */
  /* EdgeMatch — MatchMap.h:68-74 */
  const uint32_t *em_unitig;
  const int32_t  *em_ov0, *em_ov1;
  const double   *em_score;
  const uint32_t *em_line;
  const uint8_t  *em_flags; /* bit0 direction, bit1 isPrimary */
} msgpu_tables;

int msgpu_result_device(const msgpu_ctx *ctx, msgpu_tables *out);
/* Copies the result to pinned host memory (once per build; later calls return the same pointers). */
int msgpu_result_host(msgpu_ctx *ctx, msgpu_tables *out);

/* ---- digest: an order-independent checksum of the edge / edge-match tables, computed in HBM -------------------------
 * The reference has no counterpart (it never serialises Graph / MatchMap); this is how results too large to download
 * are compared: the digests of the slices of a multi-GPU build (or of the chunks of a chunked build) ADD UP, mod 2^64,
 * to the digest of the one-GPU build and to the digest of the reference's tables (tests/golden/digests.json).
 * With mix(x) = splitmix64's finaliser (x ^= x >> 30; x *= 0xbf58476d1ce4e5b9; x ^= x >> 27; x *= 0x94d049bb133111eb;
 * x ^= x >> 31), for every edge match of edge (first, second)
 *   h = mix(flags + 0x9e3779b97f4a7c15); h = mix(h ^ bits(score)); h = mix(h ^ (u32(ov1) << 32 | line));
 *   h = mix(h ^ (unitig << 32 | u32(ov0))); em_sum += mix(h ^ (first << 32 | second))
 * and for every edge   edge_sum += mix(mix(n_edge_matches_of_the_edge + 0xd1b54a32d192ed03) ^ (first << 32 | second)).
 * violations counts rows that break an invariant of the tables: first < second, (first, second) strictly increasing,
 * edge_em_begin strictly increasing from 0 to n_edge_matches, unitigs strictly increasing inside an edge,
 * ov1 - ov0 > 100 (MatchMap.cpp:192).  first_key / last_key = (first << 32 | second) of the first / last edge of the
 * slice (~0 / 0 when empty): consecutive slices of a multi-GPU build must satisfy last_key(r) < first_key(r + 1). */
typedef struct msgpu_digest {
  uint64_t n_edges, n_edge_matches, edge_sum, em_sum, first_key, last_key, violations;
} msgpu_digest;
int msgpu_result_digest(msgpu_ctx *ctx, msgpu_digest *out);

/* ---- chunked construction: the edge tables produced piece by piece -------------------------------------------------
 * For results larger than HBM (config 5: 10^11 edge matches do not fit eight GPUs) and for consumers that stream.
 * The reference materialises everything in two hash maps (Graph::m_adjacencyList, MatchMap::m_edgeMatches); nothing
 * there corresponds to this, it is what lets the same tables exist at sizes the reference cannot reach.
 *
 * msgpu_set_chunk_tiles(ctx, n)   n > 0: the builds that follow stop after planning the pairing; the edge / edge-match
 *                      tables are then produced by msgpu_chunk_next, `n` pairing tiles (<= 4096 candidate pairs each)
 *                      at a time, in (first, second) order, into two alternating sets of device buffers.  0 (default):
 *                      the whole result is materialised by the build.
 * msgpu_chunk_next     the next chunk.  Its pointers stay valid until the NEXT call of msgpu_chunk_next (the chunk after
 *                      it is already being computed into the other buffer set while the caller works on this one;
 *                      work the caller queues on the context's stream is ordered before that buffer is reused).  MSGPU_ERR_NO_RESULT when every chunk has
 *                      been delivered.  edge_em_begin holds n_edges + 1 values that continue the slice: [0] == em_base;
 *                      the edge-match columns are indexed from the chunk's own first row (value - em_base).
 *                      msgpu_result_counts reports the rows delivered so far.
 * msgpu_chunk_digest   msgpu_digest of the chunk last delivered; the chunks' sums add up to the digest of the whole.
 * msgpu_result_host    with chunking on and no chunk taken yet, it drains the chunks itself and downloads each one on a
 *                      second stream while the next is computed: the same host tables as without chunking, with the
 *                      copy overlapping the pairing.  chain / filter stages need the resident form. */
typedef struct msgpu_chunk {
  uint64_t        index;
  uint64_t        n_edges, n_edge_matches; /* rows of this chunk */
  uint64_t        edge_base, em_base;      /* rows of the slice before this chunk */
  int             last;                    /* non-zero: no chunk follows */
  const uint32_t *edge_first, *edge_second;
  const uint64_t *edge_em_begin;
  const uint32_t *em_unitig;
  const int32_t  *em_ov0, *em_ov1;
  const double   *em_score;
  const uint32_t *em_line;
  const uint8_t  *em_flags;
} msgpu_chunk;
int msgpu_set_chunk_tiles(msgpu_ctx *ctx, uint32_t tiles_per_chunk);
int msgpu_chunk_next(msgpu_ctx *ctx, msgpu_chunk *out);
int msgpu_chunk_digest(msgpu_ctx *ctx, msgpu_digest *out);

/* ---- chaining + overlaps: the stage that consumes the construction (optional; the tables stay in HBM for it) --------
 * msgpu_chain_overlaps  replaces, for every edge of the last build, the job chainingAndOverlaps (src/main.cpp:170-178,
 *                       328-414) = getMaxPairwisePaths (mpp.cpp:145-305, checkCompatibility mpp.cpp:38-142) +
 *                       getOverlap (ol.cpp:53-101, getOverhangs ol.cpp:31-50): the shadow flag of the edge
 *                       (Edge::setShadow) and its EdgeOrder list (Edge::appendOrder, Edge.h:49-60) — start / end vertex,
 *                       leftOffset / rightOffset, containment, integer score, unitig ids of the path, direction, primary.
 *                       wiggle_room = Application::getWiggleRoom() (default 300, Application.h:132).
 * Layout: the orders of edge e are edge_n_orders[e] consecutive slots starting at slot edge_em_begin[e] (an edge has
 * at most as many orders as edge matches), in the order the reference appends them (minus paths, then plus paths).
 * The unitig ids of the orders of edge e are concatenated, order after order, in ord_ids[edge_em_begin[e] ...];
 * ord_n_ids[slot] is the path length of the order in `slot`.  baseVertex is always the edge's first vertex.
 * ord_flags: bit0 startVertex is the edge's first vertex (else the second), bit1 isContained, bit2 direction,
 * bit3 isPrimary.  Unused slots are zero. */
typedef struct msgpu_orders {
  uint64_t        n_edges, n_slots; /* n_slots = n_edge_matches */
  const uint8_t  *edge_shadow;
  const uint32_t *edge_n_orders;
  const double   *ord_left, *ord_right;
  const uint64_t *ord_score;
  const uint32_t *ord_n_ids;
  const uint8_t  *ord_flags;
  const uint32_t *ord_ids;
} msgpu_orders;

int msgpu_chain_overlaps(msgpu_ctx *ctx, uint64_t wiggle_room);
int msgpu_orders_device(const msgpu_ctx *ctx, msgpu_orders *out);
int msgpu_orders_host(msgpu_ctx *ctx, msgpu_orders *out); /* pinned host copy, valid until the next build / chain */
/* device time of the last msgpu_chain_overlaps, milliseconds */
int msgpu_chain_ms(const msgpu_ctx *ctx, float *out);

/* ---- per-edge passes after chaining (optional; they read the orders of the last msgpu_chain_overlaps in HBM) ---------
 * msgpu_edge_filters  replaces, for every edge, the job bodies
 *   findContractionEdges (src/main.cpp:416-460, with sanityCheck of libms/src/kernel/sc.cpp:29-88):
 *       contraction_order[e] = index, among the orders of edge e, of the first contained primary order whose start
 *       vertex passes the sanity test against every non-shadow neighbour edge; -1 if none (the edge is not a
 *       contraction edge).  The value the reference stores is &order, i.e. orders[index].
 *   findDeletableEdges (src/main.cpp:534-550): the contained orders are dropped (Edge::replaceOrders);
 *       n_kept_orders[e] = orders left, in their old relative order (they are the slots with ord_flags bit1 clear);
 *       deletable[e] = 1 iff none is left.
 *   computeBitweight (src/main.cpp:552-573) on the orders that are left: weight[e] (Edge::setWeight, 0 if untouched)
 *       and consensus[e] = 'a' e_POS / 'b' e_NEG / 'c' e_NONE (Direction.h:41).
 * The bookkeeping between these jobs (findContractionTargets, findDeletableVertices, contract: main.cpp:462-531) stays
 * on the host: it depends on the iteration order of a pointer-keyed unordered_map.  Deleting vertices in between
 * removes edges but does not change the verdict on an edge that survives. */
typedef struct msgpu_edge_filter_tables {
  uint64_t        n_edges;
  const int32_t  *contraction_order;
  const uint8_t  *deletable;
  const uint32_t *n_kept_orders;
  const uint64_t *weight;
  const uint8_t  *consensus;
} msgpu_edge_filter_tables;

int msgpu_edge_filters(msgpu_ctx *ctx, uint64_t wiggle_room);
int msgpu_edge_filters_device(const msgpu_ctx *ctx, msgpu_edge_filter_tables *out);
int msgpu_edge_filters_host(msgpu_ctx *ctx, msgpu_edge_filter_tables *out); /* pinned host copy */
int msgpu_edge_filters_ms(const msgpu_ctx *ctx, float *out); /* device time of the last msgpu_edge_filters, ms */

/* The same passes on ONE SLICE of a multi-GPU build (msgpu_build_from_shards with n_parts > 1).  findContractionEdges
 * walks the edges around an order's start vertex and the third edge of every triangle (src/main.cpp:432-443,
 * Graph::getNeighbors / hasEdge / getEdge) — rows that live on other ranks — so every rank first publishes a compact
 * VIEW of its slice: per edge (first, second, number of orders, shadow flag), per order (leftOffset, rightOffset,
 * flags), orders packed densely.
 *   msgpu_filter_view_describe : sizes of this rank's view (after msgpu_chain_overlaps); blob_bytes is what
 *                                msgpu_filter_view_export writes
 *   msgpu_filter_view_export   : the view into a device buffer of at least blob_bytes (16-byte aligned)
 *   msgpu_edge_filters_sharded : the caller has exchanged headers and views (an all-gather: host code's business);
 *                                headers[r] / dev_blobs[r] = rank r's, device pointers, rank order.  Verdicts are
 *                                produced for THIS rank's edges (msgpu_edge_filters_device / _host as usual) and are
 *                                the ones msgpu_edge_filters gives on one device for the same edges.
 * msgpu_edge_filters itself refuses a slice (MSGPU_ERR_LIMIT) rather than answer without the neighbours. */
typedef struct msgpu_filter_view_header {
  uint64_t n_edges, n_orders, blob_bytes, reserved;
} msgpu_filter_view_header;
int msgpu_filter_view_describe(msgpu_ctx *ctx, msgpu_filter_view_header *out);
int msgpu_filter_view_export(msgpu_ctx *ctx, void *dev_dst, uint64_t capacity);
int msgpu_edge_filters_sharded(msgpu_ctx *ctx, uint32_t n_parts, uint32_t part, const msgpu_filter_view_header *headers,
                               const void *const *dev_blobs, uint64_t wiggle_room);

/* Device time of the stages of the last build, milliseconds (CUDA events on the context's stream). */
typedef struct msgpu_timings {
  float h2d_ms, ingest_ms, pair_ms, merge_ms, d2h_ms, total_ms;
} msgpu_timings;
int msgpu_result_timings(const msgpu_ctx *ctx, msgpu_timings *out);

/* Per-kernel device time of the last build (CUDA events around every launch, on the context's stream) — what
 * bench.py uses for the roofline of the dominant kernel.  Off by default: the events cost host time per launch. */
typedef struct msgpu_kernel_stat {
  char     name[48];
  uint32_t launches;
  float    total_ms;
} msgpu_kernel_stat;
int      msgpu_profile_enable(msgpu_ctx *ctx, int enable);
int      msgpu_profile_read(msgpu_ctx *ctx, msgpu_kernel_stat *out, size_t capacity, size_t *n_out);
/* kernels this context has launched since it was created (the library's own count; bench.py's `gpu_launches`) */
uint64_t msgpu_launch_count(const msgpu_ctx *ctx);

/* ---- sequences ----------------------------------------------------------------------------------------------------
 * msgpu_seq_load   replaces SequenceAccessor(...) + buildIndex() (src/main.cpp:161-163): both files are indexed,
 *                  whitespace-stripped and kept resident in HBM.  Ids continue the numbering of the context's last
 *                  build (the reference passes the same two Registry objects, src/main.cpp:161-162); names the PAF
 *                  never mentioned get the next free ids in file order.  A file whose extension is not .fa/.fasta
 *                  (case-insensitive) is FASTQ (SequenceAccessor.cpp:71-78); the unitig file is always FASTA.
 * msgpu_seq_fetch_batch  replaces getIlluminaSequence / getNanoporeSequence (SequenceUtils.cpp:63-85) for a batch:
 *                  request i yields strSlice(record, left, right + 1), reverse-complemented iff !direction, written
 *                  at out + out_off[i] (out_off has n + 1 entries, filled by the call).  Requests the reference
 *                  would answer with an exception get status[i] != 0 and an empty slice.
 */
typedef struct msgpu_seq_req {
  uint32_t which; /* 0 = unitig (getIlluminaSequence), 1 = nanopore read (getNanoporeSequence) */
  uint32_t id;
  int32_t  left, right;
  uint32_t direction; /* Toggle: non-zero = forward, zero = reverse complement */
} msgpu_seq_req;

int msgpu_seq_load(msgpu_ctx *ctx, const char *unitigs_fasta_path, const char *nanopore_path);
int msgpu_seq_load_buffers(msgpu_ctx *ctx, const char *unitigs_fasta, size_t n_unitigs_bytes,
                           const char *nanopore, size_t n_nanopore_bytes, int nanopore_is_fastq);
/* id of a sequence name (which as above); MSGPU_ERR_OUT_OF_RANGE if unknown */
int msgpu_seq_id(const msgpu_ctx *ctx, uint32_t which, const char *name, uint32_t *out_id);
/* upper bound of the bytes a batch needs, so the caller can size `out` */
int msgpu_seq_batch_bound(const msgpu_ctx *ctx, const msgpu_seq_req *reqs, size_t n, uint64_t *out_bytes);
int msgpu_seq_fetch_batch(msgpu_ctx *ctx, const msgpu_seq_req *reqs, size_t n, char *out, uint64_t out_capacity,
                          uint64_t *out_off, uint8_t *status);
/* same with every buffer resident in HBM (requests, output, offsets, status) — the path bench.py times */
int msgpu_seq_fetch_batch_device(msgpu_ctx *ctx, const msgpu_seq_req *dev_reqs, size_t n, char *dev_out,
                                 uint64_t out_capacity, uint64_t *dev_out_off, uint8_t *dev_status);
/* strSlice / getReverseComplement on caller-provided text (SequenceUtils.cpp:27-61), batched, device pointers:
 * text i = text[text_off[i] .. text_off[i+1]) ; writes like msgpu_seq_fetch_batch_device. */
int msgpu_slice_revcomp_device(msgpu_ctx *ctx, const char *dev_text, const uint64_t *dev_text_off,
                               const int32_t *dev_left, const int32_t *dev_right, const uint8_t *dev_revcomp, size_t n,
                               char *dev_out, uint64_t out_capacity, uint64_t *dev_out_off, uint8_t *dev_status);

#ifdef __cplusplus
}
#endif
#endif /* MSGPU_H */
