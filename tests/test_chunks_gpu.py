"""GPU tests of the chunked construction (msgpu_set_chunk_tiles / msgpu_chunk_next / msgpu_chunk_digest and the
overlapped download of msgpu_result_host): the edge tables produced piece by piece are the resident ones."""
import numpy as np
import pytest

import util
from muchsalsa_b200 import synth
from muchsalsa_b200.construct import GraphBuilder, device_view
from oracle.pyoracle import tables_equal

pytestmark = pytest.mark.gpu

SUMS = ("n_edges", "n_edge_matches", "edge_sum", "em_sum")


def _drain(gb):
    """All chunks of the current build as host arrays + their digests."""
    parts, digests = [], []
    while True:
        c = gb.chunk_next()
        if c is None:
            break
        E, EM = int(c.n_edges), int(c.n_edge_matches)
        g = lambda p, n, k: device_view(p, n, k).cpu().numpy().copy()
        parts.append(dict(index=int(c.index), last=int(c.last), edge_base=int(c.edge_base), em_base=int(c.em_base),
                          first=g(c.edge_first, E, "u4"), second=g(c.edge_second, E, "u4"),
                          begin=g(c.edge_em_begin, E + 1, "u8"), unitig=g(c.em_unitig, EM, "u4"),
                          ov0=g(c.em_ov0, EM, "i4"), ov1=g(c.em_ov1, EM, "i4"), score=g(c.em_score, EM, "f8"),
                          line=g(c.em_line, EM, "u4"), flags=g(c.em_flags, EM, "u1")))
        digests.append(gb.chunk_digest())
    return parts, digests


@pytest.mark.parametrize("tiles", [1, 3, 64, 100000])
@pytest.mark.parametrize("name", ["small", "heavy_rows", "ecoli"])
def test_chunks_concatenate_to_the_resident_tables(name, tiles):
    """Chunk boundaries fall between rows, inside heavy rows (between their sub-tiles) and, with one tile per chunk,
    after every tile; `100000` is one chunk holding everything."""
    if name == "ecoli" and tiles < 64:
        pytest.skip("thousands of one-tile chunks: covered by the smaller cases")
    paf = util.heavy_rows_paf() if name == "heavy_rows" else synth.generate_paf(name)
    with GraphBuilder(0) as gb:
        want = gb.build_from_buffer(paf).result()
        want_digest = gb.digest()
        gb.set_chunk_tiles(tiles)
        gb.build_from_buffer(paf)
        assert gb.counts()["n_edges"] == 0  # nothing delivered yet
        parts, digests = _drain(gb)
        c = gb.counts()
        assert c["n_edges"] == want.counts["n_edges"] and c["n_edge_matches"] == want.counts["n_edge_matches"]
        assert gb.chunk_next() is None
    assert [p["index"] for p in parts] == list(range(len(parts))) and [p["last"] for p in parts][-1] == 1
    assert all(p["last"] == 0 for p in parts[:-1])
    if tiles < 64 and name != "small":
        assert len(parts) > 3
    eb = mb = 0
    for p in parts:  # bases are running totals; em_begin continues the slice
        assert p["edge_base"] == eb and p["em_base"] == mb and int(p["begin"][0]) == mb
        eb += p["first"].size
        mb += p["unitig"].size
        assert int(p["begin"][-1]) == mb
    cat = lambda k: np.concatenate([p[k] for p in parts])
    assert np.array_equal(cat("first"), want.edge_first) and np.array_equal(cat("second"), want.edge_second)
    begin = np.concatenate([p["begin"][:-1] for p in parts] + [np.array([mb], np.int64)])
    assert np.array_equal(begin.astype(np.uint64), want.edge_em_begin)
    for k, wk in (("unitig", "unitig"), ("ov0", "ov0"), ("ov1", "ov1"), ("line", "line"), ("flags", "flags")):
        assert np.array_equal(cat(k), want.edge_matches[wk]), k
    assert np.array_equal(cat("score").view(np.uint64), want.edge_matches["score"].view(np.uint64))
    # the chunks' digests add up to the digest of the whole; no chunk breaks an invariant; key ranges follow each other
    assert all(d["violations"] == 0 for d in digests)
    for k in SUMS:
        assert sum(d[k] for d in digests) % (1 << 64) == want_digest[k], k
    live = [d for d in digests if d["n_edges"]]
    assert all(a["last_key"] < b["first_key"] for a, b in zip(live, live[1:]))


@pytest.mark.parametrize("tiles", [2, 37])
def test_host_tables_of_a_chunked_build_equal_the_resident_ones(oracle, tiles):
    """msgpu_result_host drains the chunks itself, downloading each on a second stream while the next is computed."""
    paf = synth.generate_paf("ecoli")
    with GraphBuilder(0) as gb:
        gb.set_chunk_tiles(tiles)
        got = util.gpu_to_tables(gb.build_from_buffer(paf).result())
        again = util.gpu_to_tables(gb.build_from_buffer(paf).result())  # buffers and events are reused across builds
        gb.set_chunk_tiles(0)
        plain = util.gpu_to_tables(gb.build_from_buffer(paf).result())
    assert tables_equal(got, plain) == [] and tables_equal(again, plain) == []
    assert tables_equal(got, oracle.build(paf)) == []


def test_chunked_build_of_an_empty_and_a_pairless_input():
    with GraphBuilder(0) as gb:
        gb.set_chunk_tiles(8)
        for paf in (b"", util.edge_case_pafs()["nothing_accepted"], util.edge_case_pafs()["single_match"]):
            gb.build_from_buffer(paf)
            c = gb.chunk_next()
            assert c is not None and int(c.n_edges) == 0 and int(c.last) == 1
            assert gb.chunk_next() is None
            r = gb.build_from_buffer(paf).result()
            assert r.counts["n_edges"] == 0 and r.edge_em_begin.tolist() == [0]


@pytest.mark.parametrize("tiles", [1, 5])
def test_chunks_of_a_sharded_build_add_up_to_the_golden(tiles):
    """Chunked emission on the slices of a multi-GPU build (emulated ranks): every rank plans after the merge and
    delivers ITS tiles chunk by chunk; the digests of all chunks of all ranks add up to the digest recorded from the
    reference for the whole PAF, and the host tables drained through msgpu_result_host concatenate to the same."""
    import torch
    n = 3
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
        ds, parts = [], []
        for r, gb in enumerate(ranks):
            gb.set_chunk_tiles(tiles)
            gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
            _, d = _drain(gb)
            ds += d
            gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
            parts.append(gb.result())  # drained by msgpu_result_host, downloads overlapped
    finally:
        for gb in ranks:
            gb.close()
    assert all(d["violations"] == 0 for d in ds) and len(ds) >= n
    for k in SUMS:
        assert sum(d[k] for d in ds) % (1 << 64) == g[k], k
    merged = util.merge_parts(parts)
    assert merged.counts["n_edges"] == g["n_edges"] and merged.counts["n_edge_matches"] == g["n_edge_matches"]
    util.check_invariants(merged)
