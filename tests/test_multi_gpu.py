"""GPU parity tests of the multi-GPU construction path.  N ranks are emulated one after the other on cuda:0 (one
context per rank, the all-gather replaced by handing every rank the same list of blobs), so the whole data path —
chunk ingest, blob export, merge into the global id space, partitioned pairing — is checked against the oracle on the
unsplit PAF without needing N GPUs.  The collectives themselves are covered on CPU by tests/test_multi_cpu.py."""
import dataclasses

import numpy as np
import pytest
import torch

import util
from muchsalsa_b200 import multi_gpu, synth
from muchsalsa_b200.construct import GraphBuilder
from oracle.pyoracle import Orders, orders_equal, tables_digest, tables_equal

pytestmark = pytest.mark.gpu


def build_sharded(chunks, chain=False, peers=False):
    """chunks: the rank-ordered pieces of one PAF.  Returns the merged Construction.  peers: every rank has its own
    receive buffer [n][stride] and every export writes its slot of ALL of them in one kernel (the path real ranks take
    over peer-mapped memory, msgpu_shard_export_peers)."""
    n = len(chunks)
    ranks = [GraphBuilder(0) for _ in range(n)]
    try:
        headers, blobs = [], []
        for r, (gb, chunk) in enumerate(zip(ranks, chunks)):
            gb.shard_ingest_buffer(chunk, r == n - 1)
            headers.append(gb.shard_header())
        if peers:
            stride = (max(int(h.blob_bytes) for h in headers) + 255) & ~255
            recv = [torch.full((n * stride,), 0xEE, dtype=torch.uint8, device="cuda") for _ in range(n)]
            for r, gb in enumerate(ranks):
                gb.shard_export_peers([b.data_ptr() + r * stride for b in recv], stride)
            torch.cuda.synchronize()
            for b in recv[1:]:
                assert torch.equal(b, recv[0])
            # the headers every rank reads out of its own buffer are the ones the exporters described
            heads = recv[-1].view(n, stride)[:, :multi_gpu.HEADER_BYTES].contiguous().view(torch.int64)
            seen = multi_gpu.tensor_to_headers(heads.view(n, multi_gpu.HEADER_WORDS))
            assert [int(x.blob_bytes) for x in seen] == [int(h.blob_bytes) for h in headers]
        else:
            for gb, h in zip(ranks, headers):
                blob = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
                gb.shard_export(blob.data_ptr(), blob.numel())
                blobs.append(blob)
        torch.cuda.synchronize()
        parts = []
        for r, gb in enumerate(ranks):
            ptrs = [recv[r].data_ptr() + s * stride for s in range(n)] if peers else [b.data_ptr() for b in blobs]
            gb.build_from_shards(headers, ptrs, r, n)
            parts.append(gb.result())
            if chain:
                parts[-1].orders = gb.chain_overlaps(300).orders()
        return util.merge_parts(parts), parts
    finally:
        for gb in ranks:
            gb.close()


@pytest.mark.parametrize("n", [1, 2, 3, 5])
@pytest.mark.parametrize("name", ["kat8", "duplicates_ungrouped", "thresholds", "heavy_group", "long_tags"])
def test_sharded_edge_cases_vs_oracle(oracle, name, n):
    """Chunk boundaries fall inside unitig groups, duplicates of a (read, unitig) pair land in different chunks, a
    name's first appearance and its later ones are seen by different ranks."""
    paf = util.edge_case_pafs()[name]
    got, _ = build_sharded(multi_gpu.split_paf(paf, n))
    assert tables_equal(util.gpu_to_tables(got), oracle.build(paf)) == []


@pytest.mark.parametrize("n", [2, 3])
def test_export_into_every_ranks_buffer_vs_oracle(oracle, n):
    """msgpu_shard_export_peers: one kernel writes the blob into this rank's slot of every receive buffer; the result
    built from those buffers is the oracle's.  A slot that is too small receives the header only."""
    paf = util.edge_case_pafs()["duplicates_ungrouped"] if n == 2 else synth.generate_paf("small")
    got, _ = build_sharded(multi_gpu.split_paf(paf, n), peers=True)
    assert tables_equal(util.gpu_to_tables(got), oracle.build(paf)) == []
    with GraphBuilder(0) as gb:
        gb.shard_ingest_buffer(paf, True)
        h = gb.shard_header()
        small = torch.full((256,), 0xEE, dtype=torch.uint8, device="cuda")
        gb.shard_export_peers([small.data_ptr()], 128)
        torch.cuda.synchronize()
        assert int(small[:64].view(torch.int64)[7].item()) == int(h.blob_bytes) > 128
        assert bool((small[64:] == 0xEE).all())


@pytest.mark.parametrize("n", [2, 4])
def test_sharded_synthetic_vs_oracle(oracle, n):
    paf = synth.generate_paf("small")
    got, parts = build_sharded(multi_gpu.split_paf(paf, n))
    assert tables_equal(util.gpu_to_tables(got), oracle.build(paf)) == []
    assert sum(p.counts["n_edge_matches"] for p in parts) == got.counts["n_edge_matches"]


def test_sharded_ecoli_matches_reference_digest(oracle):
    """4 ranks, shuffled (ungrouped) E. coli-scale PAF: every rank sees every unitig; result = the reference's."""
    cfg = dataclasses.replace(synth.CONFIGS["ecoli"], shuffle_groups=True)
    paf = synth.generate_paf(cfg)
    got, parts = build_sharded(multi_gpu.split_paf(paf, 4))
    t = util.gpu_to_tables(got)
    g = util.load_golden("digests.json")["ecoli_shuffled"]
    assert t.counts == g["counts"] and tables_digest(t) == g["digest"]
    sizes = [p.counts["n_edge_matches"] for p in parts]
    assert min(sizes) > 0.5 * max(sizes), f"unbalanced partition {sizes}"


def test_sharded_heavy_rows(oracle):
    """Sub-tiles of one heavy row end up on different ranks: the slices still concatenate to the oracle's table."""
    paf = util.heavy_rows_paf()
    got, _ = build_sharded(multi_gpu.split_paf(paf, 3))
    assert tables_equal(util.gpu_to_tables(got), oracle.build(paf)) == []


def test_sharded_builder_world_1_equals_plain_build(builder):
    """The ShardedBuilder path with one rank (no process group) gives the plain single-GPU result."""
    paf = synth.generate_paf("small")
    plain = util.gpu_to_tables(builder.build_from_buffer(paf).result())
    with GraphBuilder(0) as gb:
        sb = multi_gpu.ShardedBuilder(gb, 0, 1, None, torch.device("cuda", 0))
        sharded = util.gpu_to_tables(sb.build_buffer(paf).result())
    assert tables_equal(plain, sharded) == []


@pytest.mark.parametrize("n", [1, 3])
def test_sharded_chain_vs_reference(reference, n):
    """Every rank chains the edges of its own slice (it holds the whole match table); the slices' orders appended in
    rank order are the reference's."""
    paf = util.edge_case_pafs()["heavy_group"] if n == 3 else synth.generate_paf("small")
    _, parts = build_sharded(multi_gpu.split_paf(paf, n), chain=True)
    keys = ("start_is_first", "left", "right", "contained", "direction", "primary", "score", "ids")
    cat = {k: np.concatenate([p.orders[k] for p in parts]) for k in keys}
    cat["edge_shadow"] = np.concatenate([p.orders["edge_shadow"] for p in parts])
    ob, ib, o0, i0 = [], [], 0, 0
    for p in parts:
        ob.append(p.orders["edge_order_begin"][:-1].astype(np.int64) + o0)
        ib.append(p.orders["ids_begin"][:-1].astype(np.int64) + i0)
        o0 += int(p.orders["edge_order_begin"][-1])
        i0 += int(p.orders["ids_begin"][-1])
    cat["edge_order_begin"] = np.concatenate(ob + [np.array([o0])]).astype(np.uint64)
    cat["ids_begin"] = np.concatenate(ib + [np.array([i0])]).astype(np.uint64)
    assert orders_equal(Orders(**cat), reference.chain(paf)) == []


def sharded_filters(chunks, wiggle=300):
    """Build + chain on n emulated ranks, exchange the filter views (here: every rank is handed the same list of
    blobs), run the sharded edge filters on every rank; returns the verdicts concatenated in rank order."""
    from oracle.pyoracle import FILTER_FIELDS
    n = len(chunks)
    ranks = [GraphBuilder(0) for _ in range(n)]
    try:
        headers, blobs = [], []
        for r, (gb, chunk) in enumerate(zip(ranks, chunks)):
            gb.shard_ingest_buffer(chunk, r == n - 1)
            headers.append(gb.shard_header())
        for gb, h in zip(ranks, headers):
            blob = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
            gb.shard_export(blob.data_ptr(), blob.numel())
            blobs.append(blob)
        torch.cuda.synchronize()
        vh, views = [], []
        for r, gb in enumerate(ranks):
            gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
            gb.chain_overlaps(wiggle)
            h = gb.filter_view_header()
            v = torch.full((int(h.blob_bytes) + 16,), 0xEE, dtype=torch.uint8, device="cuda")
            gb.filter_view_export(v.data_ptr(), v.numel())
            vh.append(h)
            views.append(v)
        torch.cuda.synchronize()
        parts = []
        for r, gb in enumerate(ranks):
            if n > 1:
                from muchsalsa_b200.construct import MsGpuError
                with pytest.raises(MsGpuError):  # the one-device entry refuses a slice
                    gb.edge_filters(wiggle)
            parts.append(gb.edge_filters_sharded(vh, [v.data_ptr() for v in views], r, n, wiggle).filters())
        return {k: np.concatenate([p[k] for p in parts]) for k in FILTER_FIELDS}, [int(h.n_edges) for h in vh]
    finally:
        for gb in ranks:
            gb.close()


@pytest.mark.parametrize("n", [1, 2, 3, 5])
@pytest.mark.parametrize("name", ["kat8", "heavy_group", "small"])
def test_sharded_edge_filters_vs_reference(reference, name, n):
    """findContractionEdges reads the edges around an order's start vertex and the third edge of every triangle — rows
    of OTHER ranks on a sharded build.  With the views exchanged, the verdicts of the slices appended in rank order are
    the reference's (src/main.cpp:416-460, 534-573)."""
    from oracle.pyoracle import FILTER_FIELDS
    paf = synth.generate_paf("small") if name == "small" else util.edge_case_pafs()[name]
    got, sizes = sharded_filters(multi_gpu.split_paf(paf, n))
    want = reference.filters(paf)
    assert [k for k in FILTER_FIELDS if got[k].shape != want[k].shape or not np.array_equal(got[k], want[k])] == []
    if name == "small" and n > 1:
        assert min(sizes) > 0 and int((got["contraction_order"] >= 0).sum()) > 0


def test_sharded_edge_filters_ecoli_equal_one_device(builder):
    """E. coli scale on 3 emulated ranks against the one-device verdicts (themselves pinned on the reference by
    tests/test_filters_gpu.py): 4.5e5 edges, thousands of contraction edges whose neighbours sit on other ranks."""
    from oracle.pyoracle import FILTER_FIELDS
    paf = synth.generate_paf("ecoli")
    want = builder.build_from_buffer(paf).chain_overlaps(300).edge_filters(300).filters()
    want = {k: np.array(want[k]) for k in FILTER_FIELDS}
    got, sizes = sharded_filters(multi_gpu.split_paf(paf, 3))
    assert [k for k in FILTER_FIELDS if not np.array_equal(got[k], want[k])] == []
    assert int((want["contraction_order"] >= 0).sum()) > 100 and min(sizes) > 0


def test_sharded_builder_world_1_edge_filters(builder):
    paf = synth.generate_paf("small")
    want = builder.build_from_buffer(paf).chain_overlaps(300).edge_filters(300).filters()
    want = {k: np.array(v) for k, v in want.items()}
    with GraphBuilder(0) as gb:
        sb = multi_gpu.ShardedBuilder(gb, 0, 1, None, torch.device("cuda", 0))
        sb.build_buffer(paf).chain_overlaps(300)
        got = sb.edge_filters(300).filters()
        assert all(np.array_equal(got[k], want[k]) for k in want)
