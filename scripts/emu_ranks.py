#!/usr/bin/env python
"""Emulate an N-rank construction on ONE GPU (ranks one after the other, the all-gather replaced by a shared list of
blobs) on the weak-scaling workload bench.py uses at N GPUs.  For debugging under compute-sanitizer.

    python scripts/emu_ranks.py [workload] [n_ranks]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "ecoli"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = synth.scaled(workload, n)
chunks = [synth.columns_to_paf(synth.generate_columns(cfg, (r, n), sentinel=(r == n - 1))) for r in range(n)]
print([len(c) for c in chunks], flush=True)
ranks = [GraphBuilder(0) for _ in range(n)]
headers, blobs = [], []
for r, (gb, chunk) in enumerate(zip(ranks, chunks)):
    gb.shard_ingest_buffer(chunk, r == n - 1)
    h = gb.shard_header()
    print(r, {k: int(getattr(h, k)) for k, _ in h._fields_}, flush=True)
    blob = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
    gb.shard_export(blob.data_ptr(), blob.numel())
    headers.append(h)
    blobs.append(blob)
torch.cuda.synchronize()
tot = 0
for r, gb in enumerate(ranks):
    gb.build_from_shards(headers, [b.data_ptr() for b in blobs], r, n)
    c = gb.counts()
    tot += c["n_edge_matches"]
    print(r, c, flush=True)
print("total edge matches", tot)
