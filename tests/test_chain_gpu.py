"""GPU parity of the stage after construction: per-edge chaining + overlap geometry.

The checker is the REFERENCE ITSELF: oracle/_ref/libmsref.so runs the reference's own construction, then its own
getMaxPairwisePaths (mpp.cpp) and getOverlap (ol.cpp) on every edge in the control flow of chainingAndOverlaps
(src/main.cpp:328-414), and dumps shadow flags and EdgeOrder rows.  The CUDA path must reproduce them bit for bit:
offsets are doubles, scores integers, paths lists of unitig ids."""
import dataclasses

import numpy as np
import pytest

import util
from muchsalsa_b200 import synth
from oracle.pyoracle import Orders, orders_equal

pytestmark = pytest.mark.gpu


def _gpu_orders(builder, paf, wiggle=300) -> Orders:
    o = builder.build_from_buffer(paf).chain_overlaps(wiggle).orders()
    return Orders(**o)


@pytest.mark.parametrize("name", sorted(util.edge_case_pafs()))
def test_chain_edge_cases_vs_reference(builder, reference, name):
    paf = util.edge_case_pafs()[name]
    assert orders_equal(_gpu_orders(builder, paf), reference.chain(paf)) == []


def test_chain_kat8_known_answer(builder):
    """SURVEY.md §8c known-answer PAF; values recorded from the reference."""
    o = _gpu_orders(builder, util.kat8_paf())
    assert o.edge_shadow.tolist() == [0, 1, 1]
    assert o.left.tolist() == [10401.0, 2600.0, 16001.0, 900.0] and o.right.tolist() == [15401.0, 600.0, 13001.0, 2100.0]
    assert o.score.tolist() == [3800, 573, 1670, 3215] and o.ids.tolist() == [0, 0, 0, 1]


@pytest.mark.parametrize("name", ["small", "small_shuffled", "ecoli"])
def test_chain_synthetic_vs_reference(builder, reference, name):
    cfg = dataclasses.replace(synth.CONFIGS[name.split("_")[0]], shuffle_groups=name.endswith("_shuffled"))
    paf = synth.generate_paf(cfg)
    got, want = _gpu_orders(builder, paf), reference.chain(paf)
    assert orders_equal(got, want) == []
    assert got.left.size == want.left.size > 0
    g = util.load_golden("digests.json").get(name, {})
    if "orders_digest" in g:  # recorded from the reference by tests/golden/make_golden.py
        import hashlib
        h = hashlib.sha256()
        for f in ("edge_shadow", "edge_order_begin", "start_is_first", "left", "right", "contained", "direction",
                  "primary", "score", "ids_begin", "ids"):
            h.update(np.ascontiguousarray(getattr(got, f)).tobytes())
        assert h.hexdigest() == g["orders_digest"] and got.left.size == g["orders"]


@pytest.mark.parametrize("wiggle", [0, 50, 5000])
def test_chain_wiggle_room(builder, reference, wiggle):
    paf = synth.generate_paf("small")
    assert orders_equal(_gpu_orders(builder, paf, wiggle), reference.chain(paf, wiggle)) == []


def test_chain_needs_a_construction(reference):
    from muchsalsa_b200.construct import GraphBuilder, MsGpuError
    with GraphBuilder(0) as gb:
        with pytest.raises(MsGpuError) as e:
            gb.chain_overlaps()
        assert e.value.status == 7


def long_shared_reads_paf(n_unitigs=150):
    """Two long reads that share `n_unitigs` unitigs (a third shares most of them on the other strand): one edge holds
    more edge matches of ONE direction than the largest fast instantiation of the per-edge programme (96)."""
    L = util.paf_line
    rows = []
    for u in range(n_unitigs):
        a = u * 2600
        rows.append(L(f"u{u}", 2400, 0, 2200, "+", "longA", 500000, a + 100, a + 2300, 2100))
        rows.append(L(f"u{u}", 2400, 100, 2400, "+", "longB", 480000, a + 40, a + 2340, 2150))
        if u % 5 != 0:
            rows.append(L(f"u{u}", 2400, 50, 2300, "-", "longC", 470000, 460000 - a, 462250 - a, 2000))
    return b"".join(rows) + L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)


def test_chain_edge_with_more_matches_than_the_fast_paths_hold(builder, reference):
    """The reference chains any number of shared unitigs (mpp.cpp:145-305 works on vectors); edges beyond the 96-entry
    instantiation go to the fallback whose state lives in a global scratch slab — same orders, bit for bit."""
    paf = long_shared_reads_paf()
    got, want = _gpu_orders(builder, paf), reference.chain(paf)
    assert orders_equal(got, want) == []
    assert int(np.diff(got.ids_begin.astype(np.int64)).max()) > 96  # a path longer than the old limit
