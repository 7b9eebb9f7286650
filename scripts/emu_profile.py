#!/usr/bin/env python
"""Per-kernel device times of rank 0's build_from_shards in an emulated N-rank construction on ONE GPU (development
aid: shows what the replicated merge + planning cost as N grows).

    python scripts/emu_profile.py [workload] [n_ranks]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "yeast"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = synth.scaled(workload, n)
ranks = [GraphBuilder(0) for _ in range(n)]
headers, blobs = [], []
for r, gb in enumerate(ranks):
    chunk = synth.columns_to_paf(synth.generate_columns(cfg, (r, n), sentinel=(r == n - 1)))
    gb.shard_ingest_buffer(chunk, r == n - 1)
    h = gb.shard_header()
    blob = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
    gb.shard_export(blob.data_ptr(), blob.numel())
    headers.append(h)
    blobs.append(blob)
torch.cuda.synchronize()
gb = ranks[0]
ptrs = [b.data_ptr() for b in blobs]
for _ in range(3):
    gb.build_from_shards(headers, ptrs, 0, n)
torch.cuda.synchronize()
import time  # noqa: E402
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    gb.build_from_shards(headers, ptrs, 0, n)
    torch.cuda.synchronize()
    ts.append(round((time.perf_counter() - t0) * 1e3, 3))
print("build_from_shards wall ms", ts)
gb.profile(True)
gb.build_from_shards(headers, ptrs, 0, n)
ks = {k: (v[0], round(v[1], 3)) for k, v in gb.kernel_stats().items()}
print(n, gb.counts(), gb.timings())
for k, v in sorted(ks.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:24s} x{v[0]:<3d} {v[1]:8.3f} ms")
