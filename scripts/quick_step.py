#!/usr/bin/env python
"""Development helper: per-kernel device times of one construction step (yeast by default) + a table digest.

    python scripts/quick_step.py [workload]
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "yeast"
paf = synth.generate_paf(w)
dev = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
with GraphBuilder(0) as gb:
    for _ in range(3):
        gb.build_from_device(dev.data_ptr(), len(paf))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = __import__("time").perf_counter()
        gb.build_from_device(dev.data_ptr(), len(paf))
        torch.cuda.synchronize()
        ts.append((__import__("time").perf_counter() - t0) * 1e3)
    gb.profile(True)
    gb.build_from_device(dev.data_ptr(), len(paf))
    ks = {k: round(v[1], 3) for k, v in gb.kernel_stats().items()}
    c = gb.counts()
    print(w, "wall ms/step", [round(t, 3) for t in ts], "timings", gb.timings())
    print(dict(sorted(ks.items(), key=lambda kv: -kv[1])))
    print({k: c[k] for k in ("n_edges", "n_edge_matches")})
    import json  # noqa: E402
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "sum_digests.json"))).get(w + "x1")
    d = gb.digest()
    print("digest", "== golden" if gold and all(int(d[k]) == int(gold[k]) for k in ("edge_sum", "em_sum", "n_edges", "n_edge_matches")) else ("!= golden" if gold else "(no golden)"), d)
