"""Records the order-independent digests (include/msgpu.h: msgpu_digest; numpy restatement in oracle/pyoracle.py:
tables_sum_digest) of the synthetic workloads bench.py and the multi-GPU tests build, so that a result that is never
downloaded — the slices of an N-GPU build, the chunks of a chunked build — can still be compared with the CPU checker:

    python tests/golden/make_sum_digests.py [--big]      (build container only; --big adds yeast x4 / x8, ~10 min)
    python tests/golden/make_sum_digests.py --filters [--yeast]   (digests of the per-edge filter verdicts; --yeast adds
                                                                   yeast x1 / x2: ~10 / ~25 min of the reference's CPU time)

    tests/golden/sum_digests.json   { "<workload>x<N>": {n_edges, n_edge_matches, edge_sum, em_sum, source},
                                      "filters_<workload>x<N>": {n_edges, sum, contraction_edges, deletable_edges} }

The filter digests (bench.py: filters_sum_digest) are computed on the verdict arrays of the REFERENCE's own jobs
(findContractionEdges + sanityCheck, findDeletableEdges, computeBitweight run by oracle/_ref/libmsref.so on the
reference's Graph); bench.py compares the sums over the ranks of a multi-GPU run with them.

x1 and x2 come from the REFERENCE ITSELF (oracle/_ref/libmsref.so); x4 / x8 from the plain-C restatement
(oracle/libms_oracle.so, which tests/test_oracle.py pins on the reference): the reference's pointer-keyed hash maps need
more memory than this container has for 2.4e8 edges.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from muchsalsa_b200 import synth  # noqa: E402
from oracle.pyoracle import Oracle, Reference, tables_sum_digest  # noqa: E402

OUT = os.path.join(HERE, "sum_digests.json")


def record_filters(out):
    from bench import filters_sum_digest
    jobs = [("small", 1), ("small", 2), ("small", 3), ("ecoli", 1), ("ecoli", 2)]
    if "--yeast" in sys.argv:
        jobs += [("yeast", 1), ("yeast", 2)]
    ref = Reference()
    for wl, n in jobs:
        key = f"filters_{wl}x{n}"
        if key in out and "--force" not in sys.argv:
            continue
        paf = b"".join(synth.rank_chunk(wl, r, n) for r in range(n))
        d = filters_sum_digest(ref.filters(paf, 300))
        out[key] = dict(d, source="reference", wiggle_room=300)
        print(key, out[key], flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


def main():
    if "--filters" in sys.argv:
        record_filters(json.load(open(OUT)) if os.path.exists(OUT) else {})
        return
    big = "--big" in sys.argv
    jobs = [("small", 1), ("small", 2), ("small", 3), ("ecoli", 1), ("ecoli", 2), ("yeast", 1), ("yeast", 2)]
    if big:
        jobs += [("yeast", 4), ("yeast", 8)]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    ref, orc = Reference(), Oracle()
    for wl, n in jobs:
        key = f"{wl}x{n}"
        if key in out and "--force" not in sys.argv:
            continue
        paf = b"".join(synth.rank_chunk(wl, r, n) for r in range(n))
        src = "reference" if (wl != "yeast" or n <= 2) else "oracle"
        t = (ref if src == "reference" else orc).build(paf)
        d = tables_sum_digest(t)
        out[key] = dict(n_edges=d["n_edges"], n_edge_matches=d["n_edge_matches"], edge_sum=d["edge_sum"],
                        em_sum=d["em_sum"], n_matches=t.counts["vm"], n_reads=t.counts["reads"], source=src)
        print(key, out[key], flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
