"""Host-side mirrors of the pipeline's two Python pre-filters over the GPU entry points.

The reference scripts are Python programs that take file paths on the command line (pipeline.sh:163-175); these
functions take the same arguments in the same order and write the same files, with the per-line loops replaced by
libmsgpu calls:

    unitig_filter(paf, unitigs, stats, out)        pipeline/unitig_filter.py        msgpu_unitig_coverage_*
    anchor_graph(paf)                              pipeline/scrubber_bfs.py:57-114  msgpu_anchor_graph_from_paf_*

What stays on the host is what the scripts do outside their PAF loops: the two percentiles (numpy, like the script),
reading the unitig FASTA and writing FASTA records wrapped at 60 columns.  Biopython is not needed: the scripts use it
only as an indexed FASTA reader.
"""
from __future__ import annotations

import numpy as np

from .construct import GraphBuilder


def _read_fasta(path):
    """{id: (description, sequence)} — id = the header up to the first white space, like SeqIO's record.id."""
    recs, name, desc, parts = {}, None, None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    recs[name] = (desc, "".join(parts))
                desc = line[1:].rstrip("\n")
                name, parts = (desc.split() or [""])[0], []
            else:
                parts.append(line.strip())
    if name is not None:
        recs[name] = (desc, "".join(parts))
    return recs


def _write_sequence(f, seq):
    for i in range(0, len(seq), 60):  # unitig_filter.py:25-27
        f.write(seq[i:i + 60] + "\n")


def unitig_filter(paf_file, unitigs, output_stats, output_file, builder: GraphBuilder | None = None) -> dict:
    """pipeline/unitig_filter.py with its two passes over the PAF on the GPU.  Same argv order, same two output files
    (`output_stats` is appended to, like the script does).  Returns the counters the script reports."""
    own = builder is None
    gb = builder or GraphBuilder(0)
    try:
        paf = open(paf_file, "rb").read()
        cov = gb.unitig_coverage_from_buffer(paf)                         # pass 1 (unitig_filter.py:34-75)
        names = gb.result().unitig_names
        # mapping_freq is a dict keyed by the unitig: the LAST block of a unitig is what the cut-off sees
        last = {}
        for b, u in enumerate(cov["block_unitig"]):
            last[int(u)] = int(cov["block_max_cov"][b])
        blocks_maxcov = list(last.values())
        q1, q3 = np.percentile(blocks_maxcov, 25), np.percentile(blocks_maxcov, 75)   # unitig_filter.py:83-97
        upper_whisker = q3 + 1.5 * (q3 - q1)
        cut = gb.unitig_coverage_cut(float(upper_whisker), float(q3))     # pass 2 (unitig_filter.py:102-176)
    finally:
        if own:
            gb.close()
    fasta = _read_fasta(unitigs)
    frags_of = {}
    for b, s, e in zip(cut["frag_block"], cut["frag_start"], cut["frag_end"]):
        frags_of.setdefault(int(b), []).append((int(s), int(e)))
    outliers = rescued = 0
    with open(output_file, "w") as out:
        for b, u in enumerate(cov["block_unitig"]):
            uid = names[int(u)]
            if cut["block_outlier"][b]:
                outliers += 1
                frs = frags_of.get(b, [])
                rescued += 1 if frs else 0
                for k, (s, e) in enumerate(frs):                          # cut_peaks, unitig_filter.py:113-126
                    out.write(">" + uid + "_" + str(k) + " " + str(e - s + 1) + " " + str(s) + " " + str(e) + "\n")
                    _write_sequence(out, fasta[uid][1][s:e + 1])
            else:
                out.write(">" + fasta[uid][0] + "\n")                     # unitig_filter.py:167-169
                _write_sequence(out, fasta[uid][1])
    with open(output_stats, "a") as st:
        st.write(">>> unitig filter \n")
        st.write("upper_outlier: " + str(upper_whisker) + "\n")
        st.write("Q3: " + str(q3) + "\n")
        st.write("#all unitigs: " + str(len(cov["block_unitig"])) + "\n")
        st.write("#outliers: " + str(outliers) + "\n")
        st.write("#rescued outliers: " + str(rescued) + "\n")
    return dict(blocks=len(cov["block_unitig"]), outliers=outliers, rescued=rescued, cutoff=float(upper_whisker),
                q3=float(q3))


def anchor_graph(paf: bytes, builder: GraphBuilder | None = None):
    """The graph pipeline/scrubber_bfs.py:57-114 builds, in the script's terms: ({read: (length, {anchor: (s_2, e_2)})}
    in node-creation order, set of frozenset({read, read}) edges)."""
    own = builder is None
    gb = builder or GraphBuilder(0)
    try:
        r = gb.anchor_graph_from_buffer(paf).result()
    finally:
        if own:
            gb.close()
    m = r.matches
    nodes = {nm: (int(r.read_nlen[i]), {}) for i, nm in enumerate(r.read_names)}
    for k in range(m["read"].size):
        nodes[r.read_names[int(m["read"][k])]][1][r.unitig_names[int(m["unitig"][k])]] = (int(m["nr0"][k]), int(m["nr1"][k]))
    edges = {frozenset((r.read_names[int(a)], r.read_names[int(b)])) for a, b in zip(r.edge_first, r.edge_second)}
    return nodes, edges
