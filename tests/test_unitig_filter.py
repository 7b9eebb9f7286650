"""The pipeline's unitig filter (pipeline/unitig_filter.py).

CPU: oracle/unitig_filter_oracle.py against the vectors tests/golden/make_prefilter_golden.py recorded from the
UNMODIFIED script.  GPU: msgpu_unitig_coverage_* against those vectors, and the host mirror
muchsalsa_b200.prefilter.unitig_filter against the SHA-256 of the two files the script wrote."""
import hashlib
import importlib.util
import os

import numpy as np
import pytest

import util
from oracle import unitig_filter_oracle as ufo

GOLD = util.load_golden("unitig_filter.json")


def _fixtures():
    spec = importlib.util.spec_from_file_location("make_prefilter_golden",
                                                  os.path.join(util.GOLDEN, "make_prefilter_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.fixtures()


@pytest.fixture(scope="module")
def fixtures():
    return _fixtures()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_restatement_reproduces_the_vectors_recorded_from_the_script(fixtures, name):
    paf, _ = fixtures[name]
    g = GOLD[name]
    blocks = ufo.blocks_pass1(paf)
    assert [[u, n, ln, mx] for u, n, ln, mx in blocks] == g["blocks"]
    cutoff, q3, last = ufo.cutoffs(blocks)
    assert cutoff == g["cutoff"] and q3 == g["q3"]
    p2 = ufo.blocks_pass2(paf, last, cutoff, q3)
    assert [[u, None if fr is None else [list(x) for x in fr]] for u, fr in p2] == g["pass2"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_coverage_blocks_and_fragments_equal_the_script(builder, fixtures, name):
    paf, _ = fixtures[name]
    g = GOLD[name]
    cov = builder.unitig_coverage_from_buffer(paf)
    names = builder.result().unitig_names
    got = [[names[int(u)], int(n), int(ln), int(mx)] for u, n, ln, mx in
           zip(cov["block_unitig"], cov["block_n_reads"], cov["block_len"], cov["block_max_cov"])]
    assert got == g["blocks"]
    cut = builder.unitig_coverage_cut(g["cutoff"], g["q3"])
    frags = {}
    for b, s, e in zip(cut["frag_block"], cut["frag_start"], cut["frag_end"]):
        frags.setdefault(int(b), []).append([int(s), int(e)])
    got2 = [[names[int(u)], (frags.get(b, []) if cut["block_outlier"][b] else None)]
            for b, u in enumerate(cov["block_unitig"])]
    assert got2 == g["pass2"]
    assert int(cut["block_outlier"].sum()) == g["outliers"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_host_mirror_writes_the_files_the_script_writes(builder, fixtures, tmp_path, name):
    """muchsalsa_b200.prefilter.unitig_filter(paf, unitigs, stats, out): byte-identical to the unmodified script's
    output file and report (SHA-256 recorded in the build container)."""
    from muchsalsa_b200 import prefilter
    paf, ufa = fixtures[name]
    (tmp_path / "in.paf").write_bytes(paf)
    (tmp_path / "unitigs.fa").write_bytes(ufa)
    res = prefilter.unitig_filter(str(tmp_path / "in.paf"), str(tmp_path / "unitigs.fa"), str(tmp_path / "report.txt"),
                                  str(tmp_path / "out.fa"), builder)
    g = GOLD[name]
    assert res["outliers"] == g["outliers"] and res["rescued"] == g["rescued"]
    assert hashlib.sha256((tmp_path / "out.fa").read_bytes()).hexdigest() == g["output_sha256"]
    assert hashlib.sha256((tmp_path / "report.txt").read_bytes()).hexdigest() == g["stats_sha256"]


@pytest.mark.gpu
def test_gpu_coverage_rejects_what_the_script_dies_on(builder):
    from muchsalsa_b200.construct import MsGpuError
    good = util.paf_line("a", 900, 0, 600, "+", "r", 1000, 0, 600, 1)
    for bad in (b"\n", b"a\t900\t0\t600\t+\tr\n", good.replace(b"\t600\t+", b"\t6x0\t+"),
                util.paf_line("a", 900, 0, 901, "+", "q", 1000, 0, 600, 1)):   # beyond the unitig: IndexError
        with pytest.raises(MsGpuError) as e:
            builder.unitig_coverage_from_buffer(good + bad)
        assert e.value.status == 4
    cov = builder.unitig_coverage_from_buffer(good * 3 + good.replace(b"\tr\t", b"\tq\t"))
    assert cov["block_n_reads"].tolist() == [2] and cov["block_max_cov"].tolist() == [2]
