// internal.h — structures shared by the translation units of libmsgpu (not part of the ABI).
#pragma once

#include <cstdint>
#include <string>

#include "../../include/msgpu.h"
#include "common.cuh"

namespace msgpu {

// VertexMatch rows (MatchMap.h:51-59) packed as structure-of-arrays, grouped by unitig, by line inside a unitig
struct MatchSoA {
  uint64_t  n      = 0;
  uint32_t *unitig = nullptr, *read = nullptr;
  int32_t  *ir0 = nullptr, *ir1 = nullptr, *nr0 = nullptr, *nr1 = nullptr, *score = nullptr;
  uint32_t *line   = nullptr;
  uint8_t  *flags  = nullptr; // bit0 direction, bit1 isPrimary
  double   *rratio = nullptr;
};

struct NameTable {
  uint64_t  n        = 0;
  uint64_t *name_off = nullptr; // n + 1
  char     *names    = nullptr;
  uint64_t  n_bytes  = 0;
};

struct ReadTable : NameTable {
  int32_t  *nlen = nullptr;
  uint32_t *line = nullptr;
};

struct UnitigTable : NameTable {
  uint32_t *grp_off = nullptr; // n + 1 : matches of unitig u are rows [grp_off[u], grp_off[u+1])
};

// One match, as the pairing kernels gather it (32 bytes = one L2 sector).  Rows are grouped by unitig and ordered by
// read id inside a unitig; `orig` is the row of the same match in the public match table (MatchSoA).
struct __align__(16) PairRow {
  int32_t  ir0, ir1; // illuminaRange (inclusive)
  uint32_t read;
  int32_t  score;
  uint32_t line, flags, unitig, orig;
};

// EdgeOrder rows (Edge.h:49-60) + the per-edge shadow flag, as msgpu_chain_overlaps leaves them (see include/msgpu.h)
struct OrderTables {
  bool      valid = false;
  uint8_t  *edge_shadow = nullptr;
  uint32_t *edge_n_orders = nullptr;
  double   *left = nullptr, *right = nullptr;
  uint64_t *score = nullptr;
  uint32_t *n_ids = nullptr, *ids = nullptr;
  uint8_t  *flags = nullptr;
  uint64_t  n_orders = 0;
};

// per-edge verdicts of the passes that follow chaining (filter.cu; see msgpu_edge_filters in include/msgpu.h)
struct FilterTables {
  bool      valid   = false;
  uint64_t  n_edges = 0;
  int32_t  *contraction = nullptr;
  uint8_t  *deletable   = nullptr;
  uint32_t *n_kept      = nullptr;
  uint64_t *weight      = nullptr;
  uint8_t  *consensus   = nullptr;
};

struct EdgeTables {
  uint64_t  n_pairs = 0;
  uint64_t  n_edges = 0;
  uint32_t *first = nullptr, *second = nullptr;
  uint64_t *em_begin = nullptr; // n_edges + 1
  uint64_t  n_em     = 0;
  uint32_t *em_unitig = nullptr;
  int32_t  *em_ov0 = nullptr, *em_ov1 = nullptr;
  double   *em_score = nullptr;
  uint32_t *em_line  = nullptr;
  uint8_t  *em_flags = nullptr;
  // the pairing's index of the matches (grouped by unitig, by read inside): what the chaining stage looks matches up in
  const PairRow  *pair_rows = nullptr;
  const uint32_t *p_read    = nullptr;
  // chunked emission: the em_begin VALUES of a chunk continue the slice (first = rows of the earlier chunks), while
  // the columns are indexed from the chunk's own first row
  uint64_t em_value_base = 0;
};

// What the planning of a pairing leaves behind (all pointers into the build arena): the matches ordered for pairing,
// the per-read source lists and the tile table.  Tiles [tile_lo, tile_hi) are this device's share.
struct PairPlan {
  bool            valid = false;
  const PairRow  *rows = nullptr;
  const uint32_t *p_read = nullptr;
  const int32_t  *p_ir0 = nullptr, *p_ir1 = nullptr;
  const uint32_t *grp_off = nullptr, *rd_off = nullptr, *rd_src = nullptr, *tile_row = nullptr, *tile_kind = nullptr;
  const uint64_t *slice_off = nullptr;
  const uint32_t *slices = nullptr;
  const uint32_t *wide_list = nullptr; // tiles (sub-tiles of rows with 513 .. 4095 sources) for the wide instantiation
  uint32_t        n_wide = 0;
  uint32_t       *flags = nullptr; // device: [2] a read with too many sources, [3] tile invariant violated
  bool            no_overlap_test = false; // the scrubber's pairing keeps every pair of a group
  uint64_t        n_reads = 0, n_pairs = 0;
  uint32_t        n_tiles = 0, tile_lo = 0, tile_hi = 0;
};

// scratch of one emission (a range of tiles): compact records, per-tile counts and offsets, totals
struct PairScratch {
  uint32_t            max_tiles = 0, n_chunks = 1; // n_chunks: pairs of totals the last emission left (pipelined: one per chunk)
  uint2              *records = nullptr, *tile_cnt = nullptr;
  unsigned long long *em_base = nullptr, *ed_base = nullptr, *status = nullptr, *totals = nullptr;
};

struct BuildError {
  int         status;
  std::string message;
};

// ingest.cu : PAF text in HBM -> read table, unitig table (+ group offsets), match SoA   (K1-K5)
void run_ingest(Arena &arena, cudaStream_t st, const char *d_paf, uint64_t n_bytes, uint64_t *n_lines_out,
                ReadTable &reads, UnitigTable &unitigs, MatchSoA &matches, bool drop_last_line = true,
                int mode = 0 /* 0 BlastFileReader's rules, 1 scrubber_bfs.py's, 2 unitig_filter.py's */);
// scrubber mode: block id of every record (records in line order) + the anchor of every block; returns the block count
uint64_t run_scrubber_blocks(Arena &arena, cudaStream_t st, const MatchSoA &matches, uint32_t **block_out,
                             uint32_t **block_unitig_out);
void alloc_match_soa(Arena &arena, MatchSoA &s, uint64_t n);
// stable sort of the rows by unitig id unless they are grouped already (known_unsorted: 1 / 0 = the caller has
// checked, -1 = check here, which costs a round trip to the host)
void group_by_unitig(Arena &arena, cudaStream_t st, MatchSoA &matches, uint64_t n_unitigs, int known_unsorted = -1);
// group offsets for records that arrive already packed (msgpu_build_from_records)
void build_group_offsets(Arena &arena, cudaStream_t st, const MatchSoA &matches, UnitigTable &unitigs);

// pair_sort.cu : per-unitig pair enumeration + global radix sort + run-length edges   (K6, K7; sort-based variant)
void run_pairing_sorted(Arena &arena, cudaStream_t st, const MatchSoA &matches, const UnitigTable &unitigs,
                        uint64_t n_reads, EdgeTables &edges, cudaEvent_t ev_pair_done);

// pair_fused.cu : the product variant — per-row tiles sorted in shared memory (pair_tiles), offsets by two scans,
// tables written once in final order (pair_emit_rows).  `part` of `n_parts`: this device handles that contiguous share
// of the tiles, and its tables are that slice of the global result (multi-GPU); 0 of 1 = everything.
void run_pairing_fused(Arena &arena, cudaStream_t st, const MatchSoA &matches, const UnitigTable &unitigs,
                       uint64_t n_reads, EdgeTables &edges, cudaEvent_t ev_pair_done, uint32_t part, uint32_t n_parts,
                       PairPlan *plan_out, bool no_overlap_test = false);
// the same in pieces, for chunked emission (msgpu_chunk_next): plan once, then emit ranges of tiles
void     pairing_plan(Arena &arena, cudaStream_t st, const MatchSoA &matches, const UnitigTable &unitigs, uint64_t n_reads,
                      uint32_t part, uint32_t n_parts, PairPlan &plan, bool no_overlap_test = false);
void     pairing_scratch_alloc(Arena &arena, uint32_t max_tiles, PairScratch &sc);
uint64_t pairing_rows_bound(const PairPlan &plan, uint32_t n_tiles);
void     pairing_outputs_alloc(Arena &arena, uint64_t cap, EdgeTables &e);
void     pairing_tiles(Arena &arena, cudaStream_t st, const PairPlan &plan, uint32_t t_lo, uint32_t t_hi, PairScratch &sc);
void     pairing_rows(Arena &arena, cudaStream_t st, const PairPlan &plan, uint32_t t_lo, uint32_t t_hi, PairScratch &sc,
                      EdgeTables &e, uint64_t em_value_base);
void     pairing_finish(Arena &arena, cudaStream_t st, const PairPlan &plan, const PairScratch &sc, uint64_t *n_em,
                        uint64_t *n_edges);

// picks the variant: fused unless MSGPU_VARIANT=sorted (kept as an independent cross-check of the fused kernel)
void run_pairing(Arena &arena, cudaStream_t st, const MatchSoA &matches, const UnitigTable &unitigs, uint64_t n_reads,
                 EdgeTables &edges, cudaEvent_t ev_pair_done, uint32_t part = 0, uint32_t n_parts = 1);

// shard.cu : multi-GPU construction.  One rank's ingest result travels as a flat blob (header + the arrays below, each
// section 16-byte aligned); run_shard_merge turns all ranks' blobs into the global read / unitig / match tables.
struct ShardLayout { // byte offsets of the sections of a blob
  uint64_t read_nlen, read_line, read_name_off, read_names, unitig_name_off, unitig_names;
  uint64_t m_unitig, m_read, m_ir0, m_ir1, m_nr0, m_nr1, m_score, m_line, m_rratio, m_flags, total;
};
ShardLayout        shard_layout(const msgpu_shard_header &h);
msgpu_shard_header shard_describe(uint64_t n_lines, const ReadTable &reads, const UnitigTable &unitigs,
                                  const MatchSoA &matches);
void shard_export(cudaStream_t st, const msgpu_shard_header &h, const ReadTable &reads, const UnitigTable &unitigs,
                  const MatchSoA &matches, void *dev_blob);
constexpr uint32_t kMaxPeers = 16; // ranks one export kernel writes to (one node: 8 GPUs today)
void shard_export_peers(cudaStream_t st, const msgpu_shard_header &h, const ReadTable &reads, const UnitigTable &unitigs,
                        const MatchSoA &m, void *const *dev_slots, uint32_t n_slots, bool with_sections);
void run_shard_merge(Arena &arena, cudaStream_t st, uint32_t n_shards, const msgpu_shard_header *headers,
                     const void *const *dev_blobs, uint64_t *n_lines_out, ReadTable &reads, UnitigTable &unitigs,
                     MatchSoA &matches);

// digest.cu : order-independent checksum + structural invariants of the edge tables, in HBM
void run_digest(Arena &arena, cudaStream_t st, const EdgeTables &edges, msgpu_digest *out);

// coverage.cu : the per-unitig coverage profiles of the pipeline's unitig filter (pipeline/unitig_filter.py)
struct CoverageTables {
  bool      valid = false, cut_valid = false;
  uint64_t  n_unitigs = 0, n_blocks = 0, n_positions = 0;
  MatchSoA  m;                       // the records (every PAF line), in line order
  uint32_t *rec_block = nullptr;     // block of every record
  uint32_t *block_unitig = nullptr, *block_len = nullptr, *block_n_reads = nullptr, *block_max = nullptr;
  uint64_t *block_off = nullptr;     // first counter slot of the block (length + 1 slots each)
  uint32_t *starts = nullptr, *ends = nullptr, *slot = nullptr;
  struct {
    unsigned long long *keys = nullptr;
    uint32_t           *minline = nullptr;
    uint64_t            mask = 0;
  } first_hit;
  // pass 2
  uint8_t  *block_outlier = nullptr;
  uint64_t  n_frags = 0;
  uint32_t *frag_block = nullptr;
  int32_t  *frag_start = nullptr, *frag_end = nullptr;
};
void run_unitig_coverage(Arena &arena, cudaStream_t st, const MatchSoA &matches, uint64_t n_unitigs, CoverageTables &out);
void run_unitig_cut(Arena &arena, cudaStream_t st, CoverageTables &cov, double cutoff, double q3);

// chain.cu : the stage after construction — per edge chaining of its edge matches and the overlap geometry
// (src/main.cpp:328-414, mpp.cpp:38-305, ol.cpp:31-101) -> shadow flag + EdgeOrder rows
void run_chain_overlaps(Arena &arena, cudaStream_t st, const ReadTable &reads, const UnitigTable &unitigs,
                        const MatchSoA &matches, const EdgeTables &edges, uint64_t wiggle_room, OrderTables &out,
                        float *ms_out);

// filter.cu : the per-edge passes after chaining — findContractionEdges (+ sanityCheck), findDeletableEdges,
// computeBitweight (src/main.cpp:416-460, 534-573, sc.cpp:29-88)
void run_edge_filters(Arena &arena, cudaStream_t st, uint64_t n_reads, const EdgeTables &edges, const OrderTables &orders,
                      uint64_t wiggle_room, FilterTables &out, float *ms_out);
// the same on one slice of a multi-GPU build: every rank describes and exports a compact view of its edges and orders
// (what the contraction test reads of a neighbour edge), the views travel, and the passes run on this rank's edges
// against the concatenation of all views
struct FilterView {
  bool      valid = false;
  uint64_t  n_edges = 0, n_orders = 0, bytes = 0;
  uint64_t *ord_prefix = nullptr; // first packed order slot of every edge of this device
};
void filter_view_describe(Arena &arena, cudaStream_t st, const EdgeTables &edges, const OrderTables &orders, FilterView &view);
void filter_view_export(cudaStream_t st, const EdgeTables &edges, const OrderTables &orders, const FilterView &view,
                        void *dev_dst);
void run_edge_filters_sharded(Arena &arena, cudaStream_t st, uint64_t n_reads, const EdgeTables &edges,
                              const OrderTables &orders, uint64_t wiggle_room, uint32_t n_parts, uint32_t part,
                              const msgpu_filter_view_header *headers, const void *const *dev_blobs, FilterTables &out,
                              float *ms_out);

} // namespace msgpu
