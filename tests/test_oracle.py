"""CPU tests: the oracle (oracle/ms_oracle.c) against the golden vectors recorded from the reference, and — wherever
oracle/_ref/libmsref.so is present — against the reference's own code run live."""
import dataclasses
import json
import os

import numpy as np
import pytest

import util
from muchsalsa_b200 import synth
from oracle.pyoracle import Tables, tables_digest, tables_equal


def _golden_tables(d) -> Tables:
    f64 = lambda v: np.array(v, np.uint64).view(np.float64)
    vm = {k: (f64(v) if k == "rratio" else np.array(v)) for k, v in d["vm"].items()}
    em = {k: (f64(v) if k == "score" else np.array(v)) for k, v in d["em"].items()}
    cast = dict(read=np.uint32, unitig=np.uint32, nr0=np.int32, nr1=np.int32, ir0=np.int32, ir1=np.int32,
                dir=np.uint8, prim=np.uint8, score=np.uint64, line=np.uint64, ov0=np.int32, ov1=np.int32)
    vm = {k: (v if v.dtype == np.float64 else v.astype(cast[k])) for k, v in vm.items()}
    em = {k: (v if v.dtype == np.float64 else v.astype(cast[k])) for k, v in em.items()}
    return Tables(d["read_names"], np.array(d["read_nlen"], np.uint64), np.array(d["read_line"], np.uint64),
                  d["unitig_names"], vm, np.array(d["edge_first"], np.uint32), np.array(d["edge_second"], np.uint32),
                  np.array(d["edge_em_begin"], np.uint64), em, {})


GOLDEN_CASES = util.load_golden("cases.json")


def test_kat8_matches_survey(oracle):
    """The known-answer vector of SURVEY.md §8c, spelled out."""
    t = oracle.build(util.kat8_paf())
    assert t.read_names == ["r1", "r2", "r3"] and t.unitig_names == ["u1", "u2"]
    assert t.read_line.tolist() == [0, 1, 2] and t.read_nlen.tolist() == [20000, 15000, 18000]
    assert list(zip(t.edge_first.tolist(), t.edge_second.tolist())) == [(0, 1), (0, 2), (1, 2)]
    assert t.edge_em_begin.tolist() == [0, 1, 2, 4]
    assert t.em["unitig"].tolist() == [0, 0, 0, 1]
    assert t.em["ov0"].tolist() == [1000, 2500, 2500, 100] and t.em["ov1"].tolist() == [2999, 2999, 3999, 1799]
    assert t.em["dir"].tolist() == [0, 1, 0, 1] and t.em["prim"].tolist() == [1, 0, 0, 1]
    assert t.em["line"].tolist() == [1, 2, 2, 4]
    assert t.em["score"].tolist() == [3800.0, 573.33333333333326, 1670.0, 3215.0]


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_vs_golden_cases(oracle, name):
    got = oracle.build(util.edge_case_pafs()[name])
    assert tables_equal(got, _golden_tables(GOLDEN_CASES[name])) == []


@pytest.mark.parametrize("name", ["small", "small_shuffled", "ecoli", "ecoli_shuffled"])
def test_oracle_vs_golden_digests(oracle, name):
    g = util.load_golden("digests.json")[name]
    cfg = dataclasses.replace(synth.CONFIGS[name.split("_")[0]], shuffle_groups=name.endswith("_shuffled"))
    t = oracle.build(synth.generate_paf(cfg))
    assert t.counts == g["counts"]
    assert tables_digest(t) == g["digest"]


@pytest.mark.parametrize("name", sorted(util.edge_case_pafs()))
def test_oracle_vs_reference_live(oracle, reference, name):
    paf = util.edge_case_pafs()[name]
    assert tables_equal(oracle.build(paf), reference.build(paf)) == []


def test_oracle_vs_reference_live_small(oracle, reference):
    paf = synth.generate_paf("small")
    assert tables_equal(oracle.build(paf), reference.build(paf)) == []


@pytest.mark.parametrize("name", sorted(util.invalid_pafs()))
def test_oracle_rejects_invalid(oracle, name):
    with pytest.raises(RuntimeError):
        oracle.build(util.invalid_pafs()[name])


def test_empty_and_tiny_inputs(oracle):
    for paf in (b"", b"\n", util.paf_line("u", 900, 0, 800, "+", "r", 1000, 0, 800, 700)):
        t = oracle.build(paf)
        assert t.counts == dict(reads=0, unitigs=0, vm=0, edges=0, em=0)


# ---- sequences -------------------------------------------------------------------------------------------------------
def test_slice_and_revcomp_kat(oracle):
    kat = util.load_golden("slice_kat.json")
    s = kat["text"].encode()
    assert oracle.revcomp(s).decode() == kat["revcomp"] == "TTAAGGCCYRtgcaNACGT"
    for i, j, want in kat["slices"]:
        got = oracle.strslice(s, i, j)
        assert (None if got is None else got.decode()) == want, (i, j)
    # the probes recorded in SURVEY.md §8a row a15
    assert oracle.strslice(s, 0, 3) == b"ACGT" and oracle.strslice(s, 5, 1) == b"a"
    assert oracle.strslice(s, -4, -1) == b"TTAA" and oracle.strslice(s, 19, 19) == b""
    assert oracle.strslice(s, -30, 2) == s and oracle.strslice(s, 20, 25) is None


def _oracle_store(oracle, data: bytes, fastq: bool):
    """name -> whole record, first record of a name wins (m_idx*.emplace, SequenceAccessor.cpp:176,217)."""
    out = {}
    for name, off, length in oracle.seq_index(data, fastq):
        out.setdefault(name, oracle.seq_extract(data, off, length))
    return out


@pytest.mark.parametrize("label", ["ref_fasta_both", "ref_fasta_fastq", "odd", "odd_fasta_reads", "synth_small"])
def test_oracle_sequences_vs_golden(oracle, label):
    import importlib.util as iu
    spec = iu.spec_from_file_location("make_golden", os.path.join(util.GOLDEN, "make_golden.py"))
    mg = iu.module_from_spec(spec)
    spec.loader.exec_module(mg)
    case = {c[0]: c for c in mg.seq_cases()}[label]
    g = util.load_golden("seq.json")[label]
    stores = [_oracle_store(oracle, case[1], False), _oracle_store(oracle, case[2], g["fastq"])]
    by_id = [{}, {}]
    for which in (0, 1):
        for name, id_, full in g["records"][which]:
            if full is None:  # the reference has no such record (e.g. a header swallowed by the FASTQ '@' rule)
                assert name not in stores[which], (label, which, name)
                continue
            assert stores[which][name].decode("latin-1") == full, (label, which, name)
            by_id[which][id_] = stores[which][name]
    for which, id_, l, r, direction, want in g["fetch"]:
        got = oracle.seq_fetch(by_id[which][id_], l, r, direction)
        assert (None if got is None else got.decode("latin-1")) == want, (label, which, id_, l, r, direction)


def test_reference_chain_known_answer(reference):
    """chainingAndOverlaps of the reference on the SURVEY §8c known-answer PAF (values recorded here; they pin the
    driver that the GPU chaining tests compare against)."""
    o = reference.chain(util.kat8_paf())
    assert o.edge_shadow.tolist() == [0, 1, 1] and o.edge_order_begin.tolist() == [0, 1, 2, 4]
    assert o.left.tolist() == [10401.0, 2600.0, 16001.0, 900.0] and o.right.tolist() == [15401.0, 600.0, 13001.0, 2100.0]
    assert o.start_is_first.tolist() == [0, 1, 0, 1] and o.contained.tolist() == [0, 0, 0, 1]
    assert o.score.tolist() == [3800, 573, 1670, 3215] and o.primary.tolist() == [1, 0, 0, 0]
    assert o.ids.tolist() == [0, 0, 0, 1] and o.direction.tolist() == [0, 1, 0, 1]


def test_reference_chain_small_digest(reference):
    import hashlib
    from muchsalsa_b200 import synth
    o = reference.chain(synth.generate_paf("small"))
    h = hashlib.sha256()
    for f in ("edge_shadow", "edge_order_begin", "start_is_first", "left", "right", "contained", "direction", "primary",
              "score", "ids_begin", "ids"):
        h.update(np.ascontiguousarray(getattr(o, f)).tobytes())
    g = util.load_golden("digests.json")["small"]
    assert o.left.size == g["orders"] == 1465 and int(o.edge_shadow.sum()) == g["shadow_edges"] == 39
    assert h.hexdigest() == g["orders_digest"]


def test_reference_filters_known_answer(reference):
    """findContractionEdges / findDeletableEdges / computeBitweight of the reference on the SURVEY §8c known-answer
    PAF against the vector committed under tests/golden (pins the GPU test of the same name)."""
    from oracle.pyoracle import FILTER_FIELDS
    f = reference.filters(util.kat8_paf())
    g = util.load_golden("filters_kat8.json")
    for k in FILTER_FIELDS:
        assert f[k].tolist() == g[k], k
    # r1-r2 is the only non-shadow edge: weight = its order's score, direction negative; r1-r3 / r2-r3 are shadows
    assert f["weight"].tolist() == [3800, 0, 0] and bytes(f["consensus"].tolist()) == b"bab"


def test_reference_filters_small_digest(reference):
    import hashlib
    from oracle.pyoracle import FILTER_FIELDS
    f = reference.filters(synth.generate_paf("small"))
    g = util.load_golden("digests.json")["small"]
    h = hashlib.sha256()
    for k in FILTER_FIELDS:
        h.update(np.ascontiguousarray(f[k]).tobytes())
    assert h.hexdigest() == g["filters_digest"]
    assert int((f["contraction_order"] >= 0).sum()) == g["contraction_edges"] > 0
