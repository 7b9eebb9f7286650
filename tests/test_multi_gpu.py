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


def build_sharded(chunks, chain=False):
    """chunks: the rank-ordered pieces of one PAF.  Returns the merged Construction."""
    n = len(chunks)
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
        parts = []
        for r, gb in enumerate(ranks):
            gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
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
