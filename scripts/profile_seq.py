#!/usr/bin/env python
"""One short run of K8 for ncu: load the sequence store of a workload, fetch one batch of requests on the device.

    python scripts/profile_seq.py [workload] [n_requests]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "yeast"
n_req = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
ufa, nfa = synth.generate_sequences(workload)
g = synth.generate_columns(synth.CONFIGS[workload])["_geom"]
rng = np.random.Generator(np.random.PCG64(99))
with GraphBuilder(0) as gb:
    gb.seq_load_buffers(ufa, nfa, False)
    rid = np.array([gb.seq_id(1, f"r{int(x)}") for x in g["rname"]], np.uint32)
    idx = rng.integers(0, rid.size, n_req)
    length = g["rlen"][idx].astype(np.int64)
    want = rng.integers(50, 5001, n_req)
    left = (rng.random(n_req) * np.maximum(length - want, 1)).astype(np.int64)
    right = np.minimum(left + want - 2, length - 1)
    reqs = np.stack([np.ones(n_req, np.uint32), rid[idx], left.astype(np.uint32), right.astype(np.uint32),
                     (rng.random(n_req) < 0.5).astype(np.uint32)], axis=1).astype(np.uint32)
    d_reqs = torch.from_numpy(reqs.view(np.int32)).cuda()
    bound = int(np.minimum(want, length).sum()) + 64
    d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n_req + 1, dtype=torch.int64, device="cuda")
    d_st = torch.empty(n_req, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        gb.seq_fetch_batch_device(d_reqs.data_ptr(), n_req, d_out.data_ptr(), bound, d_off.data_ptr(), d_st.data_ptr())
    torch.cuda.synchronize()
    print("bases", int(d_off[-1].item()), "all ok", bool((d_st == 0).all().item()))
