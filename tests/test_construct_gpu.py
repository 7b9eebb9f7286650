"""GPU parity tests: the CUDA construction path, called through the C ABI, against the oracle on the same inputs
(bit-exact, doubles compared as bit patterns) and against the golden vectors recorded from the reference."""
import dataclasses

import numpy as np
import pytest

import util
from muchsalsa_b200 import synth
from oracle.pyoracle import tables_digest, tables_equal

pytestmark = pytest.mark.gpu


def _same(builder, oracle, paf):
    got = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    want = oracle.build(paf)
    assert tables_equal(got, want) == []
    return got


@pytest.mark.parametrize("name", sorted(util.edge_case_pafs()))
def test_edge_cases_vs_oracle(builder, oracle, name):
    got = _same(builder, oracle, util.edge_case_pafs()[name])
    if name == "kat8":
        assert got.em["score"].tolist() == [3800.0, 573.33333333333326, 1670.0, 3215.0]


@pytest.mark.parametrize("paf", [b"", b"\n", util.paf_line("u", 900, 0, 800, "+", "r", 1000, 0, 800, 700)])
def test_empty_inputs(builder, paf):
    c = builder.build_from_buffer(paf).counts()
    assert (c["n_reads"], c["n_matches"], c["n_edges"], c["n_edge_matches"]) == (0, 0, 0, 0)


@pytest.mark.parametrize("name", sorted(util.invalid_pafs()))
def test_invalid_paf_is_an_error(builder, name):
    from muchsalsa_b200.construct import MsGpuError
    with pytest.raises(MsGpuError) as e:
        builder.build_from_buffer(util.invalid_pafs()[name])
    assert e.value.status == 4 and "Invalid BLAST file." in e.value.message


@pytest.mark.parametrize("name", ["small", "small_shuffled", "ecoli", "ecoli_shuffled"])
def test_synthetic_vs_oracle_and_reference_digest(builder, oracle, name):
    cfg = dataclasses.replace(synth.CONFIGS[name.split("_")[0]], shuffle_groups=name.endswith("_shuffled"))
    paf = synth.generate_paf(cfg)
    got = _same(builder, oracle, paf)
    g = util.load_golden("digests.json")[name]
    assert got.counts == g["counts"] and tables_digest(got) == g["digest"]  # recorded from the reference itself


def test_unaligned_device_buffer_and_file_entry(builder, oracle, tmp_path):
    """The three entry points agree: file path, host buffer, device buffer at an odd address."""
    import torch
    paf = synth.generate_paf("small")
    want = oracle.build(paf)
    p = tmp_path / "small.paf"
    p.write_bytes(paf)
    assert tables_equal(util.gpu_to_tables(builder.build_from_paf(p).result()), want) == []
    dev = torch.empty(len(paf) + 64, dtype=torch.uint8, device="cuda")
    for shift in (0, 1, 7, 13):
        dev[shift:shift + len(paf)] = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
        torch.cuda.synchronize()
        builder.build_from_device(dev.data_ptr() + shift, len(paf))
        assert tables_equal(util.gpu_to_tables(builder.result()), want) == []


def test_missing_file(builder):
    from muchsalsa_b200.construct import MsGpuError
    with pytest.raises(MsGpuError) as e:
        builder.build_from_paf("/nonexistent/x.paf")
    assert e.value.status == 3 and "Can't open blast file." in e.value.message


def test_yeast_scale_invariants_and_digest(builder):
    """Full-size check without the (slow) oracle: structural invariants + the digest recorded from the reference."""
    paf = synth.generate_paf("yeast")
    res = builder.build_from_buffer(paf).result()
    util.check_invariants(res)
    g = util.load_golden("digests.json").get("yeast")
    if g:
        t = util.gpu_to_tables(res)
        assert t.counts == g["counts"] and tables_digest(t) == g["digest"]


def test_records_entry_matches_paf_entry(builder):
    """msgpu_build_from_records (packed SoA already in HBM) gives the same edges as the PAF entry."""
    import torch
    paf = synth.generate_paf("small")
    full = builder.build_from_buffer(paf).result()
    m = full.matches
    rec = {k: torch.from_numpy(m[k].copy()).cuda() for k in
           ("unitig", "read", "ir0", "ir1", "nr0", "nr1", "score", "line", "flags")}
    res = builder.build_from_records(rec, full.counts["n_reads"], full.counts["n_unitigs"]).result()
    assert np.array_equal(res.edge_first, full.edge_first) and np.array_equal(res.edge_second, full.edge_second)
    assert np.array_equal(res.edge_em_begin, full.edge_em_begin)
    for k in full.edge_matches:
        a, b = res.edge_matches[k], full.edge_matches[k]
        assert np.array_equal(a.view(np.uint64) if a.dtype == np.float64 else a,
                              b.view(np.uint64) if b.dtype == np.float64 else b), k


def test_heavy_rows_vs_oracle(builder, oracle):
    """Rows with thousands of candidates (reads crossing a repeat that thousands of reads cross): the pairing kernel
    splits them into sub-tiles by value ranges of the partner id; the result must not show the seams."""
    paf = util.heavy_rows_paf()
    got = _same(builder, oracle, paf)
    assert got.counts["em"] > 4_000_000


def test_group_longer_than_the_cta_sort_vs_oracle(builder, oracle):
    """A unitig matched by 4 500 reads: its group is ranked by counting, 256 members per CTA against the group's read
    ids staged 4096 at a time (two passes here); next to it groups of every other size class."""
    paf = util.heavy_rows_paf(k0=4500, k1=700, shared_from=4200, seed=5, small_groups=60)
    got = _same(builder, oracle, paf)
    assert got.counts["em"] > 2_000_000


def test_group_sizes_around_the_ranking_boundaries_vs_oracle(builder, oracle):
    """Groups of exactly the sizes where rank_rows changes its method: one warp sorts up to 128 members; longer groups
    become (group, 256-member chunk) work items of the counting kernel."""
    paf = util.group_sizes_paf([1, 2, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1024, 1025])
    _same(builder, oracle, paf)


def test_pipelined_emission_agrees_with_the_product(builder, monkeypatch):
    """MSGPU_PIPE_CHUNKS=n emits the tiles in n ranges, the emission kernel of one range on a second stream underneath
    the tile kernel of the next (measured and off by default).  Same tables, bit for bit."""
    paf = util.heavy_rows_paf()
    want = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    for n in ("2", "5"):
        monkeypatch.setenv("MSGPU_PIPE_CHUNKS", n)
        got = util.gpu_to_tables(builder.build_from_buffer(paf).result())
        monkeypatch.delenv("MSGPU_PIPE_CHUNKS")
        assert tables_equal(want, got) == []


def test_sort_based_variant_agrees_with_fused(builder, monkeypatch):
    """The sort-based pairing (global radix sort of the candidate stream) is an independent implementation of the
    same tables: both variants must agree bit for bit."""
    paf = synth.generate_paf("ecoli")
    fused = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    monkeypatch.setenv("MSGPU_VARIANT", "sorted")
    srt = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    monkeypatch.delenv("MSGPU_VARIANT")
    assert tables_equal(fused, srt) == []


def test_per_thread_emission_variant_agrees_with_the_product(builder, monkeypatch):
    """The product brings a tile's records into shared memory with one cp.async.bulk (the TMA unit's 1-D copy) on an
    mbarrier; MSGPU_EMIT=scalar loads them per thread — the A/B of DESIGN.md §4.  Same tables, bit for bit."""
    for paf in (synth.generate_paf("ecoli"), util.heavy_rows_paf()):
        want = util.gpu_to_tables(builder.build_from_buffer(paf).result())
        monkeypatch.setenv("MSGPU_EMIT", "scalar")
        got = util.gpu_to_tables(builder.build_from_buffer(paf).result())
        monkeypatch.delenv("MSGPU_EMIT")
        assert tables_equal(got, want) == []


def test_hash_collision_restarts_interning(builder, oracle, monkeypatch):
    """Names are interned through 64-bit hashes that every line verifies against its slot's representative; a
    collision restarts the interning with another seed.  The test hook makes the first attempt collide everywhere."""
    paf = synth.generate_paf("small")
    monkeypatch.setenv("MSGPU_TEST_FORCE_COLLISION", "1")
    got = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    monkeypatch.delenv("MSGPU_TEST_FORCE_COLLISION")
    assert tables_equal(got, oracle.build(paf)) == []


def long_read_paf(n_unitigs, n_partners=3):
    """A very long read (`hub`) that shares `n_unitigs` unitigs with other reads: every unitig is also matched by one
    of `n_partners` long reads (so single edges hold hundreds of edge matches) and every third one by a short read."""
    L = util.paf_line
    rows = []
    for u in range(n_unitigs):
        a = u * 3000
        rows.append(L(f"u{u}", 5000, 0, 3000, "+", "hub", 20_000_000, a, a + 3000, 2900))
        rows.append(L(f"u{u}", 5000, 500, 3500, "+-"[u % 2], f"long{u % n_partners}", 15_000_000, a + 200, a + 3200, 2800))
        if u % 3 == 0:
            rows.append(L(f"u{u}", 5000, 1000, 4000, "+", f"p{u}", 9000, 0, 3000, 2700))
    return b"".join(rows) + L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)


@pytest.mark.parametrize("n_unitigs", [600, 1500, 4000])
def test_read_that_shares_more_unitigs_than_a_product_tile_stages(builder, oracle, n_unitigs):
    """The product tile kernel stages 512 sources; a read that shares more unitigs with partners (an ultra-long read:
    legal input, the reference has no such notion) is paired by the instantiation staged for up to 4095 sources — one
    edge then holds up to that many edge matches.  Bit-exact with the oracle."""
    paf = long_read_paf(n_unitigs)
    got = _same(builder, oracle, paf)
    assert int(np.diff(got.edge_em_begin.astype(np.int64)).max()) >= n_unitigs // 3


def hub_rows_paf(n_unitigs, per_unitig, pool=60, seed=3):
    """A read (`hub`) that shares `n_unitigs` unitigs with `per_unitig` reads each, drawn from a pool of `pool` reads:
    its row of the edge table has n_unitigs sources and n_unitigs * per_unitig candidates — more than a tile, so the row
    is split by value ranges across ALL its sources; the pool reads' rows are heavy too."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = util.paf_line
    rows = []
    for u in range(n_unitigs):
        a = u * 3000
        rows.append(L(f"u{u}", 5000, 0, 3000, "+", "hub", 20_000_000, a, a + 3000, 2900))
        for r in rng.choice(pool, min(per_unitig, pool), replace=False):
            s = int(rng.integers(0, 1500))
            rows.append(L(f"u{u}", 5000, s, s + 3000, "+-"[int(r) % 2], f"x{int(r)}", 15_000_000, a + 200, a + 3200, 2800))
    return b"".join(rows) + L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)


@pytest.mark.parametrize("n_unitigs,per_unitig", [(100, 100), (300, 30), (500, 20)])
def test_heavy_rows_with_hundreds_of_sources_vs_oracle(builder, oracle, n_unitigs, per_unitig):
    """The value-range split of a heavy row bisects over all its sources at once (split_heavy_kernel: 1, 4 or 16
    sources per lane, each with a shrinking window).  Rows with 100, 300 and 500 sources; bit-exact with the oracle."""
    paf = hub_rows_paf(n_unitigs, per_unitig, pool=max(60, per_unitig + 20))
    got = _same(builder, oracle, paf)
    assert got.counts["em"] > n_unitigs * per_unitig


def test_read_with_4096_shared_unitigs_is_the_limit(builder):
    """An edge cannot hold more edge matches than a pairing tile holds candidates (4096): a read sharing that many
    unitigs with partners (>= 10 Mb of read) is refused with MSGPU_ERR_LIMIT — not a crash, not a wrong answer."""
    from muchsalsa_b200.construct import MsGpuError
    with pytest.raises(MsGpuError) as e:
        builder.build_from_buffer(long_read_paf(4100))
    assert e.value.status == 6
