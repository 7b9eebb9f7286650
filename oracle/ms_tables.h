/* ms_tables.h — flat, canonical dump of the graph-construction result.
 *
 * TEST INFRASTRUCTURE ONLY.  Shared by the two checkers under oracle/:
 *   - oracle/ms_oracle.c      (plain-C restatement of the reference algorithm)
 *   - oracle/ref_driver.cpp   (thin driver around the reference's own classes,
 *                              compiled from /root/reference into oracle/_ref/)
 * The product (muchsalsa_b200/csrc, include/msgpu.h) never includes this file.
 *
 * Everything is indexed by the ids the reference assigns when it runs with ONE
 * worker thread (the only deterministic mode, SURVEY.md §8a rules i-vii):
 * read id / unitig id = rank of the first accepted PAF line naming it
 * (Registry.cpp:36-45 reached from BlastFileReader.cpp:110-111).
 *
 * Sort orders (so two dumps can be compared with memcmp / array_equal):
 *   vertex matches : (read, unitig)
 *   edges          : (first, second)
 *   edge matches   : (first, second, unitig); edge e owns rows
 *                    [edge_em_begin[e], edge_em_begin[e+1])
 */
#ifndef MS_ORACLE_TABLES_H
#define MS_ORACLE_TABLES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ms_tables {
  /* reads (graph vertices): Vertex(id, nanoporeLength, lineIdx), Vertex.h:185-188 */
  uint64_t  n_reads;
  uint64_t *read_name_off; /* n_reads + 1 */
  char     *read_names;    /* concatenated, no terminators */
  uint64_t *read_nlen;     /* Vertex::getNanoporeLength() */
  uint64_t *read_line;     /* Vertex::getMetaDatum<size_t>(0) */

  /* unitigs ("illumina ids") */
  uint64_t  n_unitigs;
  uint64_t *unitig_name_off; /* n_unitigs + 1 */
  char     *unitig_names;

  /* VertexMatch rows, MatchMap.h:51-59 */
  uint64_t  n_vm;
  uint32_t *vm_read;
  uint32_t *vm_unitig;
  int32_t  *vm_nr0, *vm_nr1; /* nanoporeRange */
  int32_t  *vm_ir0, *vm_ir1; /* illuminaRange */
  double   *vm_rratio;
  uint8_t  *vm_dir;
  uint8_t  *vm_prim;
  uint64_t *vm_score;
  uint64_t *vm_line;

  /* edges, Edge::getVertices() order (Edge.h:235-237) */
  uint64_t  n_edges;
  uint32_t *edge_first;
  uint32_t *edge_second;
  uint64_t *edge_em_begin; /* n_edges + 1 */

  /* EdgeMatch rows, MatchMap.h:68-74 */
  uint64_t  n_em;
  uint32_t *em_unitig;
  int32_t  *em_ov0, *em_ov1;
  double   *em_score;
  uint8_t  *em_dir;
  uint8_t  *em_prim;
  uint64_t *em_line;

  /* wall-clock seconds of the phases, for the CPU baseline */
  double t_index, t_read, t_edges;
} ms_tables;

void ms_tables_free(ms_tables *t);

/* Result of the stage right after construction (src/main.cpp:170-178,328-414: chainingAndOverlaps ->
 * getMaxPairwisePaths, mpp.cpp:145-305 -> getOverlap, ol.cpp:53-101): per edge its shadow flag and its EdgeOrder
 * list (Edge.h:49-60), edges in (first, second) order, orders in the order the reference appends them (minus paths,
 * then plus paths).  baseVertex is always the edge's first vertex and is not dumped. */
typedef struct ms_orders {
  uint64_t  n_edges;
  uint8_t  *edge_shadow;
  uint64_t *edge_order_begin; /* n_edges + 1 */
  uint64_t  n_orders;
  uint8_t  *ord_start_is_first; /* startVertex == Edge::getVertices().first */
  double   *ord_left, *ord_right;
  uint8_t  *ord_contained, *ord_direction, *ord_primary;
  uint64_t *ord_score;
  uint64_t *ord_ids_begin; /* n_orders + 1 */
  uint64_t  n_ids;
  uint32_t *ord_ids;
  double    t_chain; /* wall-clock seconds of the stage */
} ms_orders;

void ms_orders_free(ms_orders *o);

/* Per-edge verdicts of the jobs that follow chaining (src/main.cpp:416-460, 534-573), edges in (first, second) order:
 * the order findContractionEdges selects (index into the edge's order list BEFORE filtering, -1 none), whether
 * findDeletableEdges marks the edge, the orders it leaves, and what computeBitweight then sets. */
typedef struct ms_filters {
  uint64_t  n_edges;
  int32_t  *contraction_order;
  uint8_t  *deletable;
  uint32_t *n_kept_orders;
  uint64_t *weight;
  uint8_t  *consensus; /* Direction::Enum as char: 'a' e_POS, 'b' e_NEG, 'c' e_NONE */
  double    t_filters; /* wall-clock seconds of the three passes */
  double    t_chain;   /* wall-clock seconds of chainingAndOverlaps before them */
} ms_filters;

void ms_filters_free(ms_filters *f);

#ifdef __cplusplus
}
#endif
#endif
