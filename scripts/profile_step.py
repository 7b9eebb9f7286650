#!/usr/bin/env python
"""One short run of the hot path for ncu: N constructions of a synthetic workload from a device-resident PAF.

    python scripts/profile_step.py [workload] [n_steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "yeast"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
paf = synth.generate_paf(workload)
dev = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
with GraphBuilder(0) as gb:
    for _ in range(steps):
        gb.build_from_device(dev.data_ptr(), len(paf))
    print(gb.counts(), gb.timings())
