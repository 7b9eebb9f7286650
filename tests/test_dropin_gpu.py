"""The drop-in boundary at the level the reference's downstream code sees it: the reference's OWN Graph, MatchMap and
Registry objects (compiled from the reference sources into oracle/_ref/libmsref.so) are filled by the GPU path through
the host shim a maintainer would add to libms (muchsalsa_b200/host/ms_gpu_shim.cpp), then dumped by the same code
that dumps them after the reference's own read() + calculateEdges().  The two dumps must be identical."""
import os

import pytest

import util
from muchsalsa_b200 import synth
from oracle.pyoracle import orders_equal, tables_digest, tables_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(reference):
    if not reference.has_gpu_shim:
        pytest.skip("oracle/_ref/libmsref.so was built without the GPU shim")
    return reference


@pytest.mark.parametrize("name", ["kat8", "duplicates_ungrouped", "thresholds", "stoi_crlf_tags", "heavy_group", "long_tags"])
def test_reference_objects_filled_by_gpu_equal_reference(ref, name):
    paf = util.edge_case_pafs()[name]
    assert tables_equal(ref.build_gpu(paf), ref.build(paf)) == []


def test_reference_objects_filled_by_gpu_small_and_ecoli(ref, oracle):
    paf = synth.generate_paf("small")
    assert tables_equal(ref.build_gpu(paf), ref.build(paf)) == []
    paf = synth.generate_paf("ecoli")
    got = ref.build_gpu(paf)
    g = util.load_golden("digests.json")["ecoli"]
    assert got.counts == g["counts"] and tables_digest(got) == g["digest"]


def test_shim_translates_errors_to_the_reference_exceptions(ref):
    with pytest.raises(RuntimeError, match="Invalid BLAST file."):
        ref.build_gpu(util.invalid_pafs()["too_few_fields"])
    with pytest.raises(RuntimeError, match="Can't open blast file."):
        ref.build_gpu("/nonexistent/x.paf")


def test_sequence_getters_through_the_shim(ref, tmp_path):
    ufa, nfa = synth.generate_sequences("small")
    up, np_ = tmp_path / "unitigs.fa", tmp_path / "reads.fa"
    up.write_bytes(ufa[:200000])
    np_.write_bytes(nfa[:200000])
    h = ref.seq_open(np_, up)
    try:
        for which, name in ((0, "u0"), (0, "u3")):
            for l, r, d in ((0, 10, True), (5, 60, False), (-20, -1, True), (3, 100000, False)):
                want = h.fetch(which, h.id(which, name), l, r, d)
                assert ref.seq_fetch_gpu(up, np_, which, name, l, r, d) == want
    finally:
        h.close()


@pytest.mark.parametrize("name", ["kat8", "duplicates_ungrouped", "heavy_group", "long_tags"])
def test_reference_edges_chained_by_gpu_equal_reference(ref, name):
    """Edge::isShadow() and Edge::getEdgeOrders() of the reference's own Edge objects after the shim's
    chainAndOverlaps, against the reference's own chainingAndOverlaps (src/main.cpp:328-414)."""
    paf = util.edge_case_pafs()[name]
    assert orders_equal(ref.chain_gpu(paf), ref.chain(paf)) == []


def test_reference_edges_chained_by_gpu_small(ref):
    paf = synth.generate_paf("small")
    assert orders_equal(ref.chain_gpu(paf), ref.chain(paf)) == []


@pytest.mark.parametrize("name", ["kat8", "duplicates_ungrouped", "heavy_group", "small"])
def test_reference_edges_filtered_by_gpu_equal_reference(ref, name):
    """The contraction-edge map, Edge::getEdgeOrders() after replaceOrders, Edge::getWeight() and
    getConsensusDirection() of the reference's own objects after the shim's findContractionEdges / findDeletableEdges /
    computeBitweights, against the reference's own job bodies (src/main.cpp:416-460, 534-573)."""
    import numpy as np
    from oracle.pyoracle import FILTER_FIELDS
    paf = synth.generate_paf("small") if name == "small" else util.edge_case_pafs()[name]
    got, want = ref.filters_gpu(paf), ref.filters(paf)
    assert [k for k in FILTER_FIELDS if not np.array_equal(got[k], want[k])] == []
