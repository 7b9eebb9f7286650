#!/usr/bin/env python
"""Time msgpu_edge_filters (after construction + chaining) on synthetic workloads, with per-kernel device times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

for w in sys.argv[1:] or ["ecoli", "yeast"]:
    paf = synth.generate_paf(w)
    dev = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
    with GraphBuilder(0) as gb:
        for _ in range(3):
            gb.build_from_device(dev.data_ptr(), len(paf))
            gb.chain_overlaps(300)
            gb.edge_filters(300)
            ms = gb.edge_filters_ms()
        c = gb.counts()
        gb.profile(True)
        gb.edge_filters(300)
        ks = {k: round(v[1], 3) for k, v in gb.kernel_stats().items()}
        f = gb.filters()
        print(w, c["n_edges"], "filters ms", round(ms, 3), "edges/s", f"{c['n_edges'] / ms * 1e3:.3g}", ks,
              "contraction edges", int((f["contraction_order"] >= 0).sum()), "deletable", int(f["deletable"].sum()),
              flush=True)
