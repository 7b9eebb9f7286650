#!/usr/bin/env python
"""Small end-to-end exercise of every kernel of the library for compute-sanitizer (memcheck / racecheck):
construction from PAF (grouped and shuffled), heavy rows, chaining + edge filters, a 3-rank sharded build, the
sequence store."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dataclasses  # noqa: E402

import torch  # noqa: E402

import util  # noqa: E402
from muchsalsa_b200 import multi_gpu, synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

small = synth.generate_paf("small")
shuffled = synth.generate_paf(dataclasses.replace(synth.CONFIGS["small"], shuffle_groups=True))
heavy = util.heavy_rows_paf(k0=1500, k1=1200, shared_from=500, small_groups=10)
with GraphBuilder(0) as gb:
    for name, paf in (("small", small), ("shuffled", shuffled), ("heavy", heavy), ("kat8", util.kat8_paf())):
        print(name, gb.build_from_buffer(paf).counts(), flush=True)
        gb.chain_overlaps(300).edge_filters(300)  # the stages after construction: chain.cu, filter.cu
        f = gb.filters()
        print("  orders", int(gb.orders()["left"].size), "contraction edges", int((f["contraction_order"] >= 0).sum()),
              flush=True)
    ufa, nfa = synth.generate_sequences("small")
    gb.seq_load_buffers(ufa[:30000], nfa[:30000], False)
    uid = gb.seq_id(0, "u0")
    print(len(gb.seq_fetch_batch([(0, uid, 0, 50, True), (0, uid, 5, 90, False)])))
chunks = multi_gpu.split_paf(heavy, 3)
ranks = [GraphBuilder(0) for _ in chunks]
headers, blobs = [], []
for r, (g, c) in enumerate(zip(ranks, chunks)):
    g.shard_ingest_buffer(c, r == len(chunks) - 1)
    h = g.shard_header()
    b = torch.empty(int(h.blob_bytes) + 16, dtype=torch.uint8, device="cuda")
    g.shard_export(b.data_ptr(), b.numel())
    headers.append(h)
    blobs.append(b)
torch.cuda.synchronize()
for r, g in enumerate(ranks):
    print("rank", r, g.build_from_shards(headers, [b.data_ptr() for b in blobs], r, len(chunks)).counts(), flush=True)
    g.close()
print("done")
