#!/usr/bin/env python
"""Where does a multi-GPU step spend its time?  Host wall-clock per phase (with a device sync after each phase) on
every rank; rank 0 prints the max over ranks.  Run under torchrun.  Development aid, not a benchmark."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from muchsalsa_b200 import multi_gpu, synth  # noqa: E402
from muchsalsa_b200.construct import GraphBuilder  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev_t = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev_t)
cfg = synth.scaled(sys.argv[1] if len(sys.argv) > 1 else "yeast", world)
paf = synth.columns_to_paf(synth.generate_columns(cfg, (rank, world), sentinel=(rank == world - 1)))
stream = torch.cuda.Stream(dev_t)
torch.cuda.set_stream(stream)
dev = torch.frombuffer(bytearray(paf), dtype=torch.uint8).cuda()
gb = GraphBuilder(local)
sb = multi_gpu.ShardedBuilder(gb, rank, world, dist, dev_t, stream)
acc = {}


def phase(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
    return out


for it in range(8):
    if it == 3:
        acc.clear()
    dist.barrier()
    phase("1 shard_ingest", lambda: gb.shard_ingest_device(dev.data_ptr(), len(paf), rank == world - 1))
    with torch.cuda.stream(stream):
        h = gb.shard_header()
        heads = phase("2 gather headers (+ .cpu)", lambda: multi_gpu.tensor_to_headers(
            multi_gpu.gather_headers(multi_gpu.header_to_tensor(h, dev_t), world, dist)))
        stride = (max(int(x.blob_bytes) for x in heads) + 255) & ~255
        mine = torch.empty(stride, dtype=torch.uint8, device=dev_t)
        phase("3 export blob", lambda: gb.shard_export(mine.data_ptr(), stride))
        blobs = phase("4 all-gather blobs", lambda: multi_gpu.gather_blobs(mine, world, dist))
    ptrs = [blobs.data_ptr() + s * stride for s in range(world)]
    phase("5 build_from_shards", lambda: gb.build_from_shards(heads, ptrs, rank, world))
    t = gb.timings()
    for k in ("ingest_ms", "pair_ms", "merge_ms"):
        acc["5." + k] = acc.get("5." + k, 0.0) + t[k] / 1e3
n = 5
vals = torch.tensor([acc[k] / n * 1e3 for k in sorted(acc)], dtype=torch.float64, device=dev_t)
dist.all_reduce(vals, op=dist.ReduceOp.MAX)
if rank == 0:
    for k, v in zip(sorted(acc), vals.tolist()):
        print(f"{k:32s} {v:8.3f} ms")
dist.destroy_process_group()
