"""Records golden vectors for the scrubber's read pairing from the UNMODIFIED pipeline script — build container only:

    python tests/golden/make_scrubber_golden.py      ->  tests/golden/scrubber.json

The first 118 lines of /root/reference/pipeline/scrubber_bfs.py (everything up to and including the loop that builds
the anchor graph, scrubber_bfs.py:57-114, and the two prints after it) are executed as they are, with sys.argv set and
a stand-in for `Bio.SeqIO.index_db` (Biopython is not installed here; those lines only bind the index to a name and
never read it).  The graph the script builds is dumped per fixture, and oracle/scrubber_oracle.py is checked
against it on every fixture.
"""
import io
import json
import os
import sys
import tempfile
import types
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from muchsalsa_b200 import synth  # noqa: E402
from oracle import scrubber_oracle  # noqa: E402
import util  # noqa: E402

SCRIPT = "/root/reference/pipeline/scrubber_bfs.py"


def run_script(paf: bytes):
    """nodes, edges as the unmodified script's own networkx graph holds them after its pairing loop."""
    src = open(SCRIPT).read().splitlines(keepends=True)
    cut = next(i for i, ln in enumerate(src) if ln.startswith("def print_corrected_read"))
    prefix = "".join(src[:cut])
    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.index_db = lambda *a, **k: {}
    bio.SeqIO = seqio
    sys.modules["Bio"], sys.modules["Bio.SeqIO"] = bio, seqio
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "anchors.paf")
        open(p, "wb").write(paf)
        argv = sys.argv
        sys.argv = [SCRIPT, p, os.path.join(td, "reads.fa"), os.path.join(td, "out.fa"), td]
        ns = {"__name__": "scrubber_prefix"}
        try:
            with redirect_stdout(io.StringIO()):
                exec(compile(prefix, SCRIPT, "exec"), ns)
        finally:
            sys.argv = argv
            ns.get("output") and ns["output"].close()
    G = ns["G"]
    nodes = {n: (G.nodes[n]["length"], dict(G.nodes[n]["illu_to_ranges"])) for n in G.nodes()}
    edges = {frozenset(e) for e in G.edges()}
    return nodes, edges


def fixtures() -> dict:
    L = util.paf_line
    cases = {"kat8": util.kat8_paf(), "small": synth.generate_paf("small")}
    cases.update({k: v for k, v in util.edge_case_pafs().items()
                  if k in ("duplicates_ungrouped", "heavy_group", "no_trailing_newline", "single_match", "long_tags")})
    cases["scrubber_rules"] = b"".join([
        L("a1", 9000, 0, 499, "+", "r1", 20000, 0, 499, 400),        # span 499: rejected, r1 not yet a node
        L("a1", 9000, 0, 500, "+", "r2", 21000, 10, 510, 1),         # span 500: accepted whatever the match count
        b"\n",                                                       # empty line: one field, skipped
        b"no tabs at all\n",                                         # one field, skipped
        L("a1", 9000, 100, 900, "-", "r1", 20000, 5, 805, 2),        # r1's node is created here
        L("a1", 9000, 100, 900, "-", "r1", 20000, 7000, 7800, 2),    # second hit of (r1, a1): ignored
        L("a2", 9000, 0, 700, "+", "r3", 22000, 0, 700, 3),
        L("a1", 9000, 0, 800, "+", "r3", 22000, 900, 1700, 3),       # a NEW block of a1: r3 pairs with nobody ...
        L("a1", 9000, 0, 800, "+", "r4", 23000, 0, 800, 3),          # ... but r4 pairs with r3
        L("a1", 9000, 0, 800, "+", "r2", 21000, 3000, 3800, 3),      # (r2, a1) known: ignored, block not touched
        L("a1", 9000, 0, 800, "+", "r5", 24000, 0, 800, 3)[:-1] + b" \t \r\n",  # trailing white space is stripped
        L("a3", 9000, 0, 800, "+", "r5", 24000, 0, 800, 3)[:-1],     # last line without a newline IS processed
    ])
    return cases


def dump(nodes, edges):
    return {"nodes": [[n, ln, sorted([a, s, e] for a, (s, e) in hits.items())] for n, (ln, hits) in nodes.items()],
            "edges": sorted(sorted(e) for e in edges)}


def main():
    out = {}
    for name, paf in fixtures().items():
        nodes, edges = run_script(paf)
        o_nodes, o_edges = scrubber_oracle.anchor_graph(paf)
        assert dump(nodes, edges) == dump(o_nodes, o_edges), f"restatement differs from the script on {name}"
        out[name] = dump(nodes, edges)
        print(name, len(nodes), "nodes", len(edges), "edges", flush=True)
    json.dump(out, open(os.path.join(HERE, "scrubber.json"), "w"))


if __name__ == "__main__":
    main()
