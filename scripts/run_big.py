#!/usr/bin/env python
"""Device-resident run of a big synthetic workload (thaliana: ~1e9 candidate pairs) with on-device invariant checks;
the result is too large to copy to the host or to hand to the oracle.

    python scripts/run_big.py [workload] [steps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from muchsalsa_b200 import synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402


class _Dev:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def dev_tensor(ptr, n, typestr):
    addr = C.cast(ptr, C.c_void_p).value
    return torch.as_tensor(_Dev(addr, n, typestr), device="cuda")


workload = sys.argv[1] if len(sys.argv) > 1 else "thaliana"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.time()
cols = synth.generate_columns(synth.CONFIGS[workload])
paf = synth.columns_to_paf(cols)
print(f"generated {len(paf) / 1e6:.1f} MB PAF in {time.time() - t0:.1f} s", flush=True)
dev = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
out = {"workload": workload, "paf_bytes": len(paf)}
with GraphBuilder(0) as gb:
    gb.profile(True)
    for i in range(steps):
        gb.build_from_device(dev.data_ptr(), len(paf))
        t = gb.timings()
        print(i, {k: round(v, 3) for k, v in t.items()}, flush=True)
    c = gb.counts()
    out["counts"], out["timings_ms"] = c, t
    out["kernels_ms"] = {k: round(v[1], 3) for k, v in sorted(gb.kernel_stats().items(), key=lambda kv: -kv[1][1])}
    out["matches_per_s"] = c["n_matches"] / (t["total_ms"] / 1e3)
    out["edge_matches_per_s"] = c["n_edge_matches"] / (t["total_ms"] / 1e3)
    d = gb.device_tables()
    E, EM, M = c["n_edges"], c["n_edge_matches"], c["n_matches"]
    first = dev_tensor(d.edge_first, E, "<i4").to(torch.int64)
    second = dev_tensor(d.edge_second, E, "<i4").to(torch.int64)
    key = first * (1 << 32) + second
    chk = {"first_lt_second": bool((first < second).all()), "edges_strictly_sorted": bool((key[1:] > key[:-1]).all())}
    del first, second, key
    b = dev_tensor(d.edge_em_begin, E + 1, "<i8").to(torch.int64)
    chk["csr_ok"] = bool(b[0] == 0 and b[-1] == EM and (b[1:] > b[:-1]).all())
    chk["max_edge_matches_per_edge"] = int((b[1:] - b[:-1]).max())
    del b
    ov = dev_tensor(d.em_ov1, EM, "<i4") - dev_tensor(d.em_ov0, EM, "<i4")
    chk["overlap_gt_100"] = bool((ov > 100).all())
    del ov
    sc = dev_tensor(d.em_score, EM, "<f8")
    chk["score_finite_positive"] = bool(torch.isfinite(sc).all() and (sc > 0).all())
    del sc
    mu = dev_tensor(d.m_unitig, M, "<i4").to(torch.int64)
    k = torch.bincount(mu, minlength=c["n_unitigs"])
    chk["pairs_eq_sum_k_choose_2"] = int((k * (k - 1) // 2).sum()) == c["n_pairs"]
    chk["k_max"] = int(k.max())
    out["checks"] = chk
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    free, total = torch.cuda.mem_get_info()
    out["device_mem_used_gb"] = (total - free) / 1e9
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"big_{workload}.json"), "w"), indent=1)
