"""GPU tests of the device-side digest (msgpu_result_digest) and of what it is used for: comparing results that are
never downloaded — the slices of a multi-GPU build, the big configurations — with the CPU checkers."""
import numpy as np
import pytest
import torch

import util
from muchsalsa_b200 import multi_gpu, synth
from muchsalsa_b200.construct import GraphBuilder, MsGpuError
from oracle.pyoracle import tables_sum_digest

pytestmark = pytest.mark.gpu

SUMS = ("n_edges", "n_edge_matches", "edge_sum", "em_sum")


@pytest.mark.parametrize("name", sorted(util.edge_case_pafs()))
def test_digest_equals_the_numpy_restatement_on_the_oracle_tables(builder, oracle, name):
    paf = util.edge_case_pafs()[name]
    got = builder.build_from_buffer(paf).digest()
    want = tables_sum_digest(oracle.build(paf))
    assert got == want  # sums, key range, zero violations


@pytest.mark.parametrize("name", ["small", "ecoli", "yeast"])
def test_digest_equals_the_golden_recorded_from_the_reference(builder, name):
    """tests/golden/sum_digests.json was written by tests/golden/make_sum_digests.py from the reference's own tables."""
    g = util.load_golden("sum_digests.json")[f"{name}x1"]
    d = builder.build_from_buffer(synth.generate_paf(name)).digest()
    assert d["violations"] == 0 and all(d[k] == g[k] for k in SUMS)


def test_digest_sees_a_corrupted_table(builder):
    """The checksum is not vacuous: flipping one byte of one column on the device changes it; an unsorted edge trips
    the invariant counter."""
    from muchsalsa_b200.construct import device_view
    paf = synth.generate_paf("small")
    good = builder.build_from_buffer(paf).digest()
    t, c = builder.device_tables(), builder.counts()
    ov = device_view(t.em_ov1, c["n_edge_matches"], "i4")
    ov[7] += 1
    bad = builder.digest()
    assert bad["em_sum"] != good["em_sum"] and bad["edge_sum"] == good["edge_sum"]
    ov[7] -= 1
    ov0 = device_view(t.em_ov0, c["n_edge_matches"], "i4")
    keep = int(ov[3].item())
    ov[3] = ov0[3] + 100  # an overlap of exactly 100 is one the reference rejects (MatchMap.cpp:192)
    assert builder.digest()["violations"] == 1
    ov[3] = keep
    first = device_view(t.edge_first, c["n_edges"], "u4")
    keep = int(first[5].item())
    first[5] = first[4] - 1 if int(first[4].item()) > 0 else first[4]  # breaks the (first, second) order
    assert builder.digest()["violations"] >= 1
    first[5] = keep
    assert builder.digest() == good


@pytest.mark.parametrize("n", [2, 3])
def test_slices_of_a_sharded_build_add_up_to_the_golden(n):
    """The N-rank weak-scaling workload of bench.py (synth.rank_chunk), emulated rank after rank on cuda:0: the sums of
    the slices' digests are the digest recorded from the reference for the concatenated PAF, and the key ranges of the
    slices follow each other."""
    g = util.load_golden("sum_digests.json")[f"smallx{n}"]
    chunks = [synth.rank_chunk("small", r, n) for r in range(n)]
    ranks = [GraphBuilder(0) for _ in range(n)]
    try:
        headers, blobs = [], []
        for r, (gb, chunk) in enumerate(zip(ranks, chunks)):
            gb.shard_ingest_buffer(chunk, r == n - 1)
            h = gb.shard_header()
            blob = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
            gb.shard_export(blob.data_ptr(), blob.numel())
            headers.append(h)
            blobs.append(blob)
        torch.cuda.synchronize()
        ds = []
        for r, gb in enumerate(ranks):
            gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
            ds.append(gb.digest())
            # the contraction test needs the neighbours' rows, which live on other ranks: refused, not answered wrongly
            gb.chain_overlaps(300)
            with pytest.raises(MsGpuError) as e:
                gb.edge_filters(300)
            assert e.value.status == 6 and "slice" in e.value.message
    finally:
        for gb in ranks:
            gb.close()
    assert all(d["violations"] == 0 for d in ds)
    for k in SUMS:
        assert sum(d[k] for d in ds) % (1 << 64) == g[k], k
    live = [d for d in ds if d["n_edges"]]
    assert all(a["last_key"] < b["first_key"] for a, b in zip(live, live[1:]))


def test_scan_tickets_survive_builds_without_pairs(oracle):
    """Builds whose scans span several tiles, interleaved with builds that launch no pairing scans at all (nothing
    accepted / no candidate pairs), more often than the scans' bookkeeping has slots: every multi-tile scan must start
    its tickets at zero (a stale ticket counter would hang the look-back or drop the totals)."""
    big = synth.generate_paf("ecoli")          # 1.0e5 lines: the per-line scans have 13 tiles
    empty = util.edge_case_pafs()["nothing_accepted"]
    single = util.edge_case_pafs()["single_match"]
    want = util.load_golden("sum_digests.json")["ecolix1"]
    with GraphBuilder(0) as gb:
        for it in range(24):
            d = gb.build_from_buffer(big).digest()
            assert all(d[k] == want[k] for k in SUMS), it
            assert gb.build_from_buffer(empty).counts()["n_edges"] == 0
            if it % 3 == 0:
                assert gb.build_from_buffer(single).counts()["n_matches"] == 1
            if it % 5 == 0:
                assert gb.build_from_buffer(b"").counts()["n_lines"] == 0


def test_shortest_accepted_lines_do_not_overflow_the_id_tables(builder, oracle):
    """std::getline takes empty fields, so an accepted line can be 20-22 bytes (both names and the strand empty or one
    byte): more reads than n_bytes / 24.  The id-indexed tables are sized from the 20-byte bound."""
    names = [bytes([b]) for b in range(33, 127) if b not in (ord("\t"),)] + [bytes([b]) for b in range(161, 255)]
    rows = [b"\t\t0\t400\t\t" + nm + b"\t1\t0\t1\t400\n" for nm in names]  # 21 bytes each, distinct read names
    rows += [b"\t\t0\t400\t\t\t1\t0\t1\t400\n"] * 3                          # 20 bytes: both names empty
    paf = b"".join(rows) + util.paf_line("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)
    from oracle.pyoracle import tables_equal
    got = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    want = oracle.build(paf)
    assert len(want.read_names) == len(names) + 1
    assert tables_equal(got, want) == []


def test_names_longer_than_64_kib(builder, oracle):
    """The reference has no limit on name lengths (std::string); neither has the device parser."""
    from oracle.pyoracle import tables_equal
    L = util.paf_line
    long_r, long_u = "R" * 70000, "U" * 66000
    paf = b"".join([L(long_u, 9000, 0, 3000, "+", long_r, 20000, 100, 3100, 2900),
                    L(long_u, 9000, 1000, 4000, "-", long_r + "x", 15000, 500, 3500, 2800),
                    L("u2", 9000, 0, 3000, "+", long_r, 20000, 5000, 8000, 2900),
                    L("u2", 9000, 500, 3000, "+", "short", 20000, 5000, 7500, 2400),
                    L("zz", 900, 0, 300, "+", "zr", 1000, 0, 300, 290)])
    got = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    assert tables_equal(got, oracle.build(paf)) == []


def test_sampled_parity_equals_full_parity_on_a_small_case(builder, oracle):
    """The sampled comparison used for thaliana / human, on a case where the full comparison is also available."""
    paf = synth.generate_paf("ecoli")
    builder.build_from_buffer(paf)
    got, orient_ok = util.gpu_sample_edge_matches(builder, 16)
    want = util.checker_sample_edge_matches(oracle, paf, 16)
    assert orient_ok and got.size > 1000 and util.records_equal(got, want)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs a second GPU in the process")
def test_two_contexts_on_two_devices_in_one_process(oracle):
    """The dynamic shared-memory opt-in of the tile kernel is per device: a second context on cuda:1 must work."""
    from oracle.pyoracle import tables_equal
    paf = synth.generate_paf("small")
    want = oracle.build(paf)
    with GraphBuilder(0) as g0, GraphBuilder(1) as g1:
        a = util.gpu_to_tables(g0.build_from_buffer(paf).result())
        b = util.gpu_to_tables(g1.build_from_buffer(paf).result())
    assert tables_equal(a, want) == [] and tables_equal(b, want) == []
