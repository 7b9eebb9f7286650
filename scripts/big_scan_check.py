#!/usr/bin/env python
"""One-GPU construction of the 8x weak-scaling yeast workload (4.0e6 matches, 230 MB of PAF: scans over more tiles than
blocks fit on the device, multi-window look-back, groups of every size class) checked against the 8-rank emulated
construction of the same input, whose per-rank ingest scans are all small: the edge tables must be identical."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = synth.scaled("yeast", n)
chunks = [synth.columns_to_paf(synth.generate_columns(cfg, (r, n), sentinel=(r == n - 1))) for r in range(n)]
COLS = ("first", "second", "begin", "unitig", "score", "line")


def feed(hs, res, em_off):
    hs["first"].update(memoryview(np.ascontiguousarray(res.edge_first)))
    hs["second"].update(memoryview(np.ascontiguousarray(res.edge_second)))
    hs["begin"].update(memoryview(res.edge_em_begin[:-1] + np.uint64(em_off)))
    hs["unitig"].update(memoryview(np.ascontiguousarray(res.edge_matches["unitig"])))
    hs["score"].update(memoryview(np.ascontiguousarray(res.edge_matches["score"])))
    hs["line"].update(memoryview(np.ascontiguousarray(res.edge_matches["line"])))
    return em_off + int(res.edge_em_begin[-1])


with GraphBuilder(0) as gb:
    gb.build_from_buffer(b"".join(chunks))
    one = gb.counts()
    h1 = {c: hashlib.sha256() for c in COLS}
    feed(h1, gb.result(), 0)
    d1 = [h1[c].hexdigest()[:12] for c in COLS]
    print("one GPU ", one, d1, flush=True)

ranks = [GraphBuilder(0) for _ in range(n)]
headers, blobs = [], []
for r, gb in enumerate(ranks):
    gb.shard_ingest_buffer(chunks[r], r == n - 1)
    h = gb.shard_header()
    blob = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
    gb.shard_export(blob.data_ptr(), blob.numel())
    headers.append(h)
    blobs.append(blob)
torch.cuda.synchronize()
h2 = {c: hashlib.sha256() for c in COLS}
tot_e = tot_em = em_off = 0
for r, gb in enumerate(ranks):
    gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
    em_off = feed(h2, gb.result(), em_off)
    tot_e += gb.counts()["n_edges"]
    tot_em += gb.counts()["n_edge_matches"]
    gb.close()
d2 = [h2[c].hexdigest()[:12] for c in COLS]
print(f"{n} ranks ", tot_e, tot_em, d2)
assert (one["n_edges"], one["n_edge_matches"]) == (tot_e, tot_em) and d1 == d2, "single-GPU and sharded results differ"
print("identical")
