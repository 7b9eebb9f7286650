"""The scrubber's read pairing (pipeline/scrubber_bfs.py:57-114).

CPU: oracle/scrubber_oracle.py against the golden vectors tests/golden/make_scrubber_golden.py recorded from the
UNMODIFIED script.  GPU: msgpu_anchor_graph_from_paf_* against those vectors and against the restatement."""
import numpy as np
import pytest

import util
from muchsalsa_b200 import synth
from oracle import scrubber_oracle


def _canon(nodes, edges):
    return {"nodes": [[n, ln, sorted([a, s, e] for a, (s, e) in hits.items())] for n, (ln, hits) in nodes.items()],
            "edges": sorted(sorted(e) for e in edges)}


def _fixture_pafs():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_scrubber_golden",
                                                  os.path.join(util.GOLDEN, "make_scrubber_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.fixtures()


@pytest.mark.parametrize("name", sorted(util.load_golden("scrubber.json")))
def test_restatement_reproduces_the_vectors_recorded_from_the_script(name):
    paf = _fixture_pafs()[name]
    assert _canon(*scrubber_oracle.anchor_graph(paf)) == util.load_golden("scrubber.json")[name]


def test_restatement_rejects_what_the_script_dies_on():
    good = util.paf_line("a", 900, 0, 600, "+", "r", 1000, 0, 600, 1)
    for bad in (b"a\tb\n", b"a\t900\t0\t600\t+\tr\t1000\t0\n", good.replace(b"\t600\t+", b"\t6x0\t+")):
        with pytest.raises(scrubber_oracle.InvalidPaf):
            scrubber_oracle.anchor_graph(good + bad)


def _gpu_graph(builder, paf):
    """The GPU result in the script's terms: {read: (length, {anchor: (s_2, e_2)})}, set of name pairs."""
    r = builder.anchor_graph_from_buffer(paf).result()
    m = r.matches
    nodes = {nm: (int(r.read_nlen[i]), {}) for i, nm in enumerate(r.read_names)}
    for k in range(m["read"].size):
        nodes[r.read_names[int(m["read"][k])]][1][r.unitig_names[int(m["unitig"][k])]] = (int(m["nr0"][k]), int(m["nr1"][k]))
    edges = {frozenset((r.read_names[int(a)], r.read_names[int(b)])) for a, b in zip(r.edge_first, r.edge_second)}
    assert len(edges) == r.edge_first.size and np.all(r.edge_first < r.edge_second)
    return nodes, edges


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(util.load_golden("scrubber.json")))
def test_gpu_anchor_graph_equals_the_vectors_recorded_from_the_script(builder, name):
    paf = _fixture_pafs()[name]
    got = _canon(*_gpu_graph(builder, paf))
    want = util.load_golden("scrubber.json")[name]
    assert got["nodes"] == want["nodes"]      # same reads, in creation order, same hits
    assert got["edges"] == want["edges"]


@pytest.mark.gpu
def test_gpu_anchor_graph_at_ecoli_scale_and_on_ungrouped_input(builder):
    import dataclasses
    for cfg in (synth.CONFIGS["ecoli"], dataclasses.replace(synth.CONFIGS["small"], shuffle_groups=True)):
        paf = synth.generate_paf(cfg)
        assert _canon(*_gpu_graph(builder, paf)) == _canon(*scrubber_oracle.anchor_graph(paf))


@pytest.mark.gpu
def test_gpu_anchor_graph_rejects_what_the_script_dies_on(builder):
    from muchsalsa_b200.construct import MsGpuError
    good = util.paf_line("a", 900, 0, 600, "+", "r", 1000, 0, 600, 1)
    for bad in (b"a\tb\n", b"a\t900\t0\t600\t+\tr\t1000\t0\n", good.replace(b"\t600\t+", b"\t6x0\t+")):
        with pytest.raises(MsGpuError) as e:
            builder.anchor_graph_from_buffer(good + bad)
        assert e.value.status == 4
    # the stages after construction do not apply to this graph
    builder.anchor_graph_from_buffer(good + good.replace(b"\tr\t", b"\tq\t"))
    with pytest.raises(MsGpuError) as e:
        builder.chain_overlaps(300)
    assert e.value.status == 7
