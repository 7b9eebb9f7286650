"""CPU restatement of the read-pairing loop of the pipeline's scrubber (pipeline/scrubber_bfs.py:57-114) — TEST
INFRASTRUCTURE, never imported by the product.  tests/golden/make_scrubber_golden.py checks it against the UNMODIFIED
script (its first 118 lines executed with a stand-in for Bio.SeqIO.index_db, which those lines never call) and records
golden vectors; tests/test_oracle.py pins it on them.

The loop builds the scrubber's "anchor graph": one node per nanopore read (first line that passes the length
threshold), per read the first hit of every anchor (unitig), and one edge per pair of reads that hit the same anchor
inside one BLOCK of consecutive processed lines naming that anchor — the same per-unitig read pairing as
MatchMap::calculateEdges, without the overlap test and with other accept rules:

  scrubber_bfs.py:75-77   line.rstrip().split("\\t"); a line with ONE field (empty, no tab) is skipped; 2..8 fields
                          raise IndexError, a non-integer in fields 1, 2, 3, 6, 7, 8 raises ValueError
  :85-86                  accept iff e_1 - s_1 >= 500 (unitig span; no match-count threshold, no last-line drop)
  :89-93                  node per read, created by its first accepted line, with that line's read length
  :97-100                 a second hit of the same (read, anchor) is ignored — before the block logic sees it
  :103-105                a new block starts when the anchor NAME differs from the previous processed line's
  :109-113                edges between the read and every earlier read of the block
"""
from __future__ import annotations


class InvalidPaf(ValueError):
    pass


def anchor_graph(paf: bytes, threshold_length: int = 500):
    """Returns (nodes, edges): nodes = {read: (length, {anchor: (s_2, e_2)})} in creation order, edges = set of
    frozenset({read_a, read_b})."""
    nodes: dict = {}
    edges: set = set()
    prev, chunk = None, []
    text = paf.decode("latin-1")
    for line in _lines(text):
        f = line.rstrip().split("\t")
        if len(f) == 1:
            continue
        if len(f) < 9:
            raise InvalidPaf("fewer than nine fields")
        try:
            int(f[1])
            len_2 = int(f[6])
            s_1, s_2 = int(f[2]), int(f[7])
            e_1, e_2 = int(f[3]), int(f[8])
        except ValueError as e:
            raise InvalidPaf(str(e))
        id_1, id_2 = f[0], f[5]
        if e_1 - s_1 < threshold_length:
            continue
        if id_2 not in nodes:
            nodes[id_2] = (len_2, {})
        if id_1 in nodes[id_2][1]:
            continue
        nodes[id_2][1][id_1] = (s_2, e_2)
        if id_1 != prev:
            chunk, prev = [], id_1
        for other in chunk:
            if other != id_2:
                edges.add(frozenset((other, id_2)))
        chunk.append(id_2)
    return nodes, edges


def _lines(text: str):
    """What `for line in open(path)` yields (universal newlines are OFF for our purposes: the GPU path and this
    restatement split on '\\n' only; a '\\r' before it is removed by rstrip like any other trailing white space)."""
    start = 0
    while start < len(text):
        nl = text.find("\n", start)
        if nl < 0:
            yield text[start:]
            return
        yield text[start:nl + 1]
        start = nl + 1
