#!/usr/bin/env python
"""Development helper: histogram of edge matches per edge (and per direction) of a synthetic workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "yeast"
paf = synth.generate_paf(w)
with GraphBuilder(0) as gb:
    gb.build_from_buffer(paf)
    r = gb.result()
    b = r.edge_em_begin.astype(np.int64)
    n = np.diff(b)
    h = np.bincount(np.minimum(n, 40))
    print(w, "edges", len(n), "em per edge histogram (40 = 40+):")
    for i, c in enumerate(h):
        if c:
            print(f"  {i:3d}: {c:10d}  ({100 * c / len(n):.2f} %)  ems {100 * (n[np.minimum(n, 40) == i].sum()) / n.sum():.2f} %")
    print("max", n.max())
