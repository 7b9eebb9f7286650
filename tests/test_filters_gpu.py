"""GPU parity of the per-edge jobs after chaining: findContractionEdges (+ sanityCheck),
findDeletableEdges and computeBitweight.

The checker is the REFERENCE ITSELF: oracle/_ref/libmsref.so runs the reference's construction and chaining, then the
three job bodies of src/main.cpp:416-460, 534-573 around the reference's own sanityCheck (sc.cpp), Graph and Edge
methods.  Integers, flags and the selected order must be identical for every edge."""
import dataclasses
import hashlib

import numpy as np
import pytest

import util
from muchsalsa_b200 import synth
from oracle.pyoracle import FILTER_FIELDS

pytestmark = pytest.mark.gpu


def _gpu(builder, paf, wiggle=300):
    return builder.build_from_buffer(paf).chain_overlaps(wiggle).edge_filters(wiggle).filters()


def _diff(got, want):
    return [k for k in FILTER_FIELDS if got[k].shape != want[k].shape or not np.array_equal(got[k], want[k])]


@pytest.mark.parametrize("name", sorted(util.edge_case_pafs()))
def test_filters_edge_cases_vs_reference(builder, reference, name):
    paf = util.edge_case_pafs()[name]
    assert _diff(_gpu(builder, paf), reference.filters(paf)) == []


def test_filters_kat8_known_answer(builder):
    """SURVEY.md §8c known-answer PAF; values recorded from the reference (edges r1-r2, r1-r3, r2-r3)."""
    f = _gpu(builder, util.kat8_paf())
    g = util.load_golden("filters_kat8.json")
    for k in FILTER_FIELDS:
        assert f[k].tolist() == g[k], k


@pytest.mark.parametrize("name", ["small", "small_shuffled", "ecoli"])
def test_filters_synthetic_vs_reference(builder, reference, name):
    cfg = dataclasses.replace(synth.CONFIGS[name.split("_")[0]], shuffle_groups=name.endswith("_shuffled"))
    paf = synth.generate_paf(cfg)
    got, want = _gpu(builder, paf), reference.filters(paf)
    assert _diff(got, want) == []
    assert got["deletable"].size > 0
    g = util.load_golden("digests.json").get(name, {})
    if "filters_digest" in g:  # recorded from the reference by tests/golden/make_golden.py
        h = hashlib.sha256()
        for k in FILTER_FIELDS:
            h.update(np.ascontiguousarray(got[k]).tobytes())
        assert h.hexdigest() == g["filters_digest"]
        assert int((got["contraction_order"] >= 0).sum()) == g["contraction_edges"]
        assert int(got["deletable"].sum()) == g["deletable_edges"]


@pytest.mark.parametrize("wiggle", [0, 50, 5000])
def test_filters_wiggle_room(builder, reference, wiggle):
    paf = synth.generate_paf("small")
    assert _diff(_gpu(builder, paf, wiggle), reference.filters(paf, wiggle)) == []


def test_filters_need_the_orders(builder):
    from muchsalsa_b200.construct import MsGpuError
    builder.build_from_buffer(util.kat8_paf())
    with pytest.raises(MsGpuError) as e:
        builder.edge_filters()
    assert e.value.status == 7


def test_chain_and_filters_at_yeast_scale_match_the_reference_digests(builder):
    """BASELINE.json's metric configuration at full size (3.0e7 edges): the shadow flags, the EdgeOrder rows and the
    per-edge verdicts hash to the digests recorded from the reference (tests/golden/make_golden.py --yeast; about ten
    minutes of the reference's CPU time, so it is recorded once and not re-run here)."""
    g = util.load_golden("digests.json").get("yeast", {})
    if "orders_digest" not in g or "filters_digest" not in g:
        pytest.skip("yeast-scale digests of the stages after construction not recorded")
    from oracle.pyoracle import Orders
    paf = synth.generate_paf("yeast")
    builder.build_from_buffer(paf).chain_overlaps(300).edge_filters(300)
    o = Orders(**builder.orders())
    h = hashlib.sha256()
    for f in ("edge_shadow", "edge_order_begin", "start_is_first", "left", "right", "contained", "direction",
              "primary", "score", "ids_begin", "ids"):
        h.update(np.ascontiguousarray(getattr(o, f)).tobytes())
    assert h.hexdigest() == g["orders_digest"] and o.left.size == g["orders"]
    got = builder.filters()
    h = hashlib.sha256()
    for k in FILTER_FIELDS:
        h.update(np.ascontiguousarray(got[k]).tobytes())
    assert h.hexdigest() == g["filters_digest"]
    assert int((got["contraction_order"] >= 0).sum()) == g["contraction_edges"]
