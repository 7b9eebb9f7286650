#!/usr/bin/env python
"""bench.py — graph-construction throughput (PAF matches/s, edges/s) of the B200 path, with its roofline and the
reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ecoli|yeast|small|thaliana] [--impl reference]

One "step" = one pass of the hot path (PAF text -> read/unitig/match/edge/edge-match tables) over one synthetic PAF.
  value   : kept matches / s with the PAF text already resident in HBM and the tables left in HBM (CUDA events on the
            library's stream, L2 flushed between steps)
  e2e     : the same metric through the reference-facing C ABI with HOST buffers: msgpu_build_from_paf_buffer (pinned
            host PAF -> device) + msgpu_result_host (every table -> pinned host), wall clock
  roofline: algorithmic HBM bytes of the dominant kernel / its CUDA-event duration / measured copy bandwidth
  cpu_baseline / --impl reference: the reference's own code (oracle/_ref/msref_cli, compiled from the reference
            sources) on this box's host cores on the same PAF
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PAF matches/s (graph build; edges/s alongside)"
UNIT = "matches/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

# bounded CPU sample per workload: (shard index, shard count) of the synthetic generator, None = the whole workload
CPU_SAMPLE = {"small": None, "ecoli": None, "yeast": (0, 24), "thaliana": (0, 400), "human": (0, 8000)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML during the timed region."""

    def __init__(self, index=0, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period, self.stop_flag = index, period, threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                "reasons": sorted(self.reasons)}


ON_CHIP_NOTE = ("pair_tiles keeps the candidates in shared memory: most of the bytes SURVEY.md §8(d) counts for it never "
                "cross HBM (`traffic` / `strict_frac` say what does), so `frac` is a throughput equivalent on the survey's "
                "accounting — it can exceed 1 — and the kernel is bound by instruction issue and shared memory (ncu "
                "summaries under profiles/)")


def make_workload(name, shard=None):
    from muchsalsa_b200 import synth
    cfg = synth.CONFIGS[name]
    return synth.columns_to_paf(synth.generate_columns(cfg, shard))


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic HBM bytes per launch of the kernels that can dominate (DESIGN.md "Kernels and their rooflines")
#   M kept matches, EM edge matches, E edges, P candidate pairs, B PAF bytes
# ---------------------------------------------------------------------------------------------------------------------
def kernel_bytes(name, c, paf_bytes):
    """Algorithmic bytes of one launch, two ways: (survey, strict).

    survey: SURVEY.md §8(d)'s fixed record sizes (R_m = R_e = 40 B) x the units the launch processes — the numerator
            the roofline is quoted on, so that a path that re-reads or re-sorts scores lower, not higher.  The product
            variant splits K6 + K7 over two kernels: pair_tiles does K6 (every match read once, every surviving
            candidate produced once: R_m*M + R_e*EM) and K7's pass over the candidates into edge-grouped order
            (R_e*EM); pair_emit_rows does K7's write of the ordered edge matches and the edge table (R_e*EM + 12*E).
            Together: R_m*M + 3*R_e*EM + 12*E = §8(d)'s K6 + K7.
    strict: the bytes THIS implementation has to move at least (actual record sizes, everything touched once); the
            tile kernel keeps the candidates in shared memory, so its strict figure is far below the survey's."""
    M, EM, E = c["n_matches"], c["n_edge_matches"], c["n_edges"]
    L = c["n_lines"]
    table = {
        # sort-based variant
        "radix_scatter": (2 * 12 * EM,) * 2,       # (key u64 + index u32) read once + written once per digit pass
        "radix_hist": (8 * EM,) * 2,               # keys read once
        "pair_emit": (40 * M + 40 * EM, 29 * M + 33 * EM),
        "pair_count": (12 * M,) * 2,
        "gather_sorted": (80 * EM, (12 + 25 + 25 + 1) * EM),
        "write_edges": (12 * E, 9 * EM + 16 * E),
        # product variant (pair_fused.cu)
        "pair_tiles": (40 * M + 80 * EM, 44 * M + 8 * EM),
        "pair_emit_rows": (40 * EM + 12 * E, 8 * EM + 32 * M + 25 * EM + 16 * E),
        # ingest: text read once + one packed record per line (§8(d): file bytes + R_m per record written)
        "parse_lines": (paf_bytes + 40 * L, paf_bytes + 58 * L),
        "newline_count": (paf_bytes,) * 2, "newline_write": (paf_bytes + 8 * L,) * 2,
    }
    return table.get(name)


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev_t = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev_t)
    from muchsalsa_b200 import multi_gpu, synth
    from muchsalsa_b200.construct import GraphBuilder

    # ---- workload: at N ranks the genome is N times longer and rank r holds the r-th contiguous chunk of the PAF
    # (weak scaling: every GPU adds one base-size share; reads and unitigs of a boundary region meet in the merge)
    paf = synth.rank_chunk(args.workload, rank, world)
    n_bytes = len(paf)
    # ONE explicit stream for everything: the library's kernels, torch's L2 flush and events, and the tensors NCCL
    # reads and writes.  (Not the default stream: its handle is 0, which the C ABI reads as "the context's own stream";
    # the flush and the timed work would then run unordered on two streams.)
    stream = torch.cuda.Stream(dev_t)
    torch.cuda.set_stream(stream)
    host = torch.frombuffer(bytearray(paf) if n_bytes else bytearray(1), dtype=torch.uint8)[:n_bytes].pin_memory()
    dev = host.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    gb = GraphBuilder(local)
    gb.set_stream(stream.cuda_stream)
    sb = multi_gpu.ShardedBuilder(gb, rank, world, dist, dev_t, stream)

    def barrier():
        if dist is not None:
            dist.barrier()

    def step_device():
        if world == 1:
            gb.build_from_device(dev.data_ptr(), n_bytes)
        else:
            sb.build_device(dev.data_ptr(), n_bytes)

    def step_e2e():
        if world == 1:
            gb.build_from_buffer(host)
        else:
            sb.build_buffer(host)
        gb.host_tables()

    def reduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    counts = gb.counts()
    torch.cuda.synchronize()
    barrier()
    parity = check_parity(gb, args.workload, rank, world, dist)

    # ---- timed region 1: device-resident throughput (CUDA events on the library's stream, max over ranks)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = gb.launch_count()
    ms, xch_ms = [], []
    for _ in range(args.steps):
        flush.zero_()  # evict L2 between steps (not timed)
        torch.cuda.synchronize()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step_device()
        b.record(stream)
        b.synchronize()
        ms.append(a.elapsed_time(b))
        if world > 1:
            xch_ms.append(sb.last_exchange_ms())
    torch.cuda.synchronize()
    barrier()
    launches = gb.launch_count() - launches0
    t_dev = reduce(sum(ms) / 1e3, "MAX")

    # ---- timed region 2: end to end through the C ABI with host buffers.  The edge tables are produced in chunks of
    # `--e2e-chunk-tiles` pairing tiles so that msgpu_result_host downloads one chunk while the next is paired
    gb.set_chunk_tiles(args.e2e_chunk_tiles)
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    barrier()
    t_e2e = reduce(time.perf_counter() - t0, "MAX")
    stage = gb.timings()
    gb.set_chunk_tiles(0)
    clocks = sampler.summary()

    # ---- per-kernel durations (CUDA events around every launch) for the roofline
    gb.profile(True)
    agg = {}
    for _ in range(args.steps):
        flush.zero_()
        step_device()
        for k, (n, t) in gb.kernel_stats().items():
            a = agg.setdefault(k, [0, 0.0])
            a[0] += n
            a[1] += t
    gb.profile(False)
    flush.zero_()
    step_device()  # one more plain step: its phase times (library events, no per-launch events) feed the roofline
    torch.cuda.synchronize()
    per_kernel = {k: {"launches_per_step": n / args.steps, "ms_per_step": t / args.steps, "us_per_launch": 1e3 * t / n}
                  for k, (n, t) in agg.items()}
    top = sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms_per_step"])
    peak, peak_src = peaks()
    roof = None
    for name, st in top:
        b = kernel_bytes(name, counts, n_bytes)
        if b is None:
            continue
        b_survey, b_strict = b
        achieved = b_survey / (st["us_per_launch"] * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload, {}).get(name)
        roof = {"kernel": name, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(b_survey),
                "algorithmic_bytes_definition": "SURVEY.md §8(d) record sizes (40 B per match / edge match, 12 B per "
                                                "edge) x units per launch; see DESIGN.md §4",
                "strict_bytes_per_launch": int(b_strict),
                "strict_frac": round(b_strict / (st["us_per_launch"] * 1e-6) / 1e9 / peak, 4),
                "us_per_launch": round(st["us_per_launch"], 2),
                "share_of_step": round(st["ms_per_step"] / sum(v["ms_per_step"] for v in per_kernel.values()), 3)}
        if name == "pair_tiles":
            roof["frac_note"] = ON_CHIP_NOTE
        break
    # the whole pairing + merge (K6 + K7: plan, tile sort, tile scans, emit) against the bytes it must move; device
    # time of the two phases from the library's own events
    tm = gb.timings()
    pm_ms = tm["pair_ms"] + tm["merge_ms"]
    if roof is not None and pm_ms > 0:
        M, EM, E = counts["n_matches"], counts["n_edge_matches"], counts["n_edges"]
        strict = 29 * M + 25 * EM + 16 * E                     # every match once, every edge match + edge once
        survey = 40 * M + 3 * 40 * EM + 12 * E                 # K6 + K7
        roof["pairing_and_merge"] = {
            "ms": round(pm_ms, 3), "bytes_written_once": int(strict),
            "achieved_gbs": round(strict / pm_ms / 1e6, 1), "frac": round(strict / pm_ms / 1e6 / peak, 4),
            "survey_8d_bytes": int(survey), "survey_8d_frac": round(survey / pm_ms / 1e6 / peak, 4)}
        survey = 3 * 40 * M + 3 * 40 * EM + 12 * E             # SURVEY.md §8(d) B_tot of the whole construction
        # the whole construction step against SURVEY.md §8(d)'s B_tot (the figure BASELINE.md asks for)
        step_ms = 1e3 * t_dev / args.steps
        roof["construction_step"] = {"ms": round(step_ms, 3), "survey_8d_bytes": int(survey),
                                     "survey_8d_frac": round(survey / step_ms / 1e6 / peak, 4),
                                     "bytes_written_once": int(strict + n_bytes),
                                     "frac": round((strict + n_bytes) / step_ms / 1e6 / peak, 4)}

    # ---- the stage that consumes the tables while they are in HBM (chaining + overlaps per edge);
    # reported beside the metric, not part of it.  Device time from the library's own events.
    chain_ms, filter_ms, filter_wall_ms, filter_error = [], [], [], None
    for it in range(min(args.steps, 3) + 1):
        step_device()
        gb.chain_overlaps(300)
        if world == 1:
            gb.edge_filters(300)
        elif filter_error is None:
            # the contraction test reads rows of other ranks: the views of the slices travel first (one all-gather).
            # This block is reported beside the metric; a failure here (memory for the gathered views) must not take
            # the bench line down, and every rank has to leave the collective path together.
            barrier()
            t0 = time.perf_counter()
            try:
                sb.edge_filters(300)
                torch.cuda.synchronize()
                failed = 0.0
            except Exception as e:  # noqa: BLE001
                filter_error, failed = repr(e)[:300], 1.0
            if reduce(failed, "MAX") > 0:
                filter_error = filter_error or "another rank failed"
            filter_wall = 1e3 * (time.perf_counter() - t0)
        if it:  # the first call grows the arena (cudaMalloc inside the timed span): warm-up
            chain_ms.append(gb.chain_ms())
            if filter_error is None:
                filter_ms.append(gb.edge_filters_ms())
                if world > 1:
                    filter_wall_ms.append(filter_wall)
    # the verdicts of the last pass against the reference's (golden digests; every rank takes part)
    filter_parity = check_filters_parity(gb, args.workload, rank, world, dist) if filter_error is None else None
    chain_t = reduce(sum(chain_ms) / len(chain_ms), "MAX")
    filter_t = reduce(sum(filter_ms) / len(filter_ms), "MAX") if filter_ms and filter_error is None else None

    # ---- K8: the resident sequence store + batched slice / reverse complement (rank 0 at N = 1)
    sequence = bench_sequences(gb, args.workload, stream) if world == 1 and not args.no_sequence else None

    # ---- the reference's CPU path beside it (bounded sample, rank 0 at N = 1 only)
    cpu = cpu_baseline(args.workload) if world == 1 else None

    n_edges = reduce(counts["n_edges"], "SUM")
    n_em = reduce(counts["n_edge_matches"], "SUM")
    bytes_d2h = reduce(result_bytes(counts, gb), "SUM")
    bytes_h2d = reduce(n_bytes, "SUM")
    out = {
        "metric": METRIC, "value": counts["n_matches"] * args.steps / t_dev, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32 keys + f64 scores",
        "data": "synthetic",
        "config": {"workload": workload_label(args.workload, world),
                   "paf_bytes": int(bytes_h2d), "lines": counts["n_lines"], "kept_matches": counts["n_matches"],
                   "reads": counts["n_reads"], "unitigs": counts["n_unitigs"], "pairs": counts["n_pairs"],
                   "edges": int(n_edges), "edge_matches": int(n_em),
                   "l2": "flushed between steps (256 MiB memset)", "matches_counted": "kept VertexMatch rows"},
        "edges_per_s": n_edges * args.steps / t_dev,
        "edge_matches_per_s": n_em * args.steps / t_dev,
        "e2e": {"value": counts["n_matches"] * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(bytes_h2d),
                "d2h_bytes_per_step": int(bytes_d2h), "ms_per_step": 1e3 * t_e2e / args.steps,
                "edges_per_s": n_edges * args.steps / t_e2e,
                "stages_ms_last_step": {k: round(v, 3) for k, v in stage.items()}},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "kernels": {k: {kk: round(vv, 3) for kk, vv in v.items()} for k, v in top[:12]},
        "variant": os.environ.get("MSGPU_VARIANT", "default"),
    }
    out["chain_overlaps"] = {
        "what": "per-edge max pairwise paths + overlap geometry (src/main.cpp:328-414, mpp.cpp, ol.cpp) on the "
                "device-resident edge tables; bit-exact vs the reference (tests/test_chain_gpu.py)",
        "ms": round(chain_t, 3), "edges_per_s": n_edges / chain_t * 1e3,
        "cpu_reference": (cpu or {}).get("chain")}
    out["parity"] = parity["ok"]
    out["parity_detail"] = parity
    if filter_t is not None:
        out["edge_filters"] = {
            "what": "per-edge findContractionEdges (+ sanityCheck), findDeletableEdges, computeBitweight "
                    "(src/main.cpp:416-460, 534-573, sc.cpp) on the device-resident order tables; identical to the "
                    "reference (tests/test_filters_gpu.py)",
            "ms": round(filter_t, 3), "edges_per_s": n_edges / filter_t * 1e3,
            "cpu_reference": ((cpu or {}).get("chain") or {}).get("filters"),
            "parity": (filter_parity or {}).get("ok"), "parity_detail": filter_parity}
        if world > 1 and filter_wall_ms:
            f = gb.filters()
            out["edge_filters"]["sharded"] = {
                "what": "every rank publishes a compact view of its slice (per edge: first, second, orders, shadow; per "
                        "order: offsets, flags), ONE all-gather carries the views, the passes run on the rank's own edges "
                        "against their concatenation (msgpu_edge_filters_sharded; tests/test_multi_gpu.py compares the "
                        "slices' verdicts with the reference's)",
                "ms_with_exchange": round(reduce(sum(filter_wall_ms) / len(filter_wall_ms), "MAX"), 3),
                "view_bytes_received_per_rank": int(getattr(sb, "filter_exchange_bytes", 0)),
                "contraction_edges": int(reduce(int((f["contraction_order"] >= 0).sum()), "SUM")),
                "deletable_edges": int(reduce(int(f["deletable"].sum()), "SUM"))}
    else:
        out["edge_filters"] = {"skipped": filter_error or "not run"}
    if sequence is not None:
        out["sequence"] = sequence
    if world > 1:
        x = [v for v in xch_ms if v is not None]
        x_ms = reduce(sum(x) / len(x), "MAX") if x else None
        out["exchange"] = {"what": "the ranks' match-table blobs reach every rank: " + getattr(sb, "exchange_mode", "?"),
                           "bytes_received_per_rank": sb.exchange_bytes, "ms": x_ms,
                           "gbs_per_rank": (sb.exchange_bytes / x_ms / 1e6) if x_ms else None,
                           "nvlink_peak_gbs": 770.0}
    gb.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


BIG_WORKLOADS = ("thaliana", "human")  # generated on the GPU, built chunk by chunk, compared on a unitig sample


def run_ours_big(args):
    """thaliana / human: results that do not fit the host (thaliana: 60 GB) or the GPUs (human: 10^11 edge matches).
    Same step as run_ours, except that
      * the workload is generated and formatted on the GPU (muchsalsa_b200.synth_gpu),
      * the edge tables are produced chunk by chunk (msgpu_set_chunk_tiles / msgpu_chunk_next) into two alternating
        buffer sets; a step has delivered every chunk when its clock stops,
      * parity = on-device invariants + digests of every chunk, and a comparison BY NAME with the CPU checker on a
        deterministic sample of the unitigs (oracle/sampled.py)."""
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev_t = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev_t)
    from muchsalsa_b200 import multi_gpu, synth, synth_gpu
    from muchsalsa_b200.construct import GraphBuilder
    from oracle import sampled

    cfg = synth.CONFIGS[args.workload]
    if args.genome_scale != 1.0:  # development: the same shape on a shorter genome
        import dataclasses
        cfg = dataclasses.replace(cfg, name=f"{cfg.name}*{args.genome_scale:g}", genome=int(cfg.genome * args.genome_scale))
    t_gen = time.perf_counter()
    cols = synth_gpu.generate_columns_torch(cfg, (rank, world) if world > 1 else None, sentinel=(rank == world - 1),
                                            device=dev_t)
    n_lines_local = int(cols["unitig"].numel())
    # ---- the sample: unitigs whose NAME hashes to 0 mod `sample_mod`; their lines go to the CPU checker
    uniq = torch.unique(cols["unitig"][:n_lines_local - (1 if rank == world - 1 else 0)]).cpu().numpy()
    picked = np.array([u for u in uniq if sampled.unitig_in_sample(b"u%d" % u, args.sample_mod)], np.int64)
    body = n_lines_local - (1 if rank == world - 1 else 0)  # the sentinel is the file's last line: never parsed
    in_s = torch.isin(cols["unitig"][:body], torch.from_numpy(picked).to(dev_t))
    sub_idx = torch.nonzero(in_s).squeeze(1)
    sub_cols = {k: v[:body][sub_idx] for k, v in cols.items() if torch.is_tensor(v)}
    sub_lines = synth_gpu.columns_to_paf_torch(sub_cols).cpu().numpy().tobytes().split(b"\n")[:-1] if sub_idx.numel() else []
    sub_local = sub_idx.cpu().numpy()
    dev = synth_gpu.columns_to_paf_torch(cols)
    n_bytes = int(dev.numel())
    del cols, sub_cols, in_s, sub_idx
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t_gen

    stream = torch.cuda.Stream(dev_t)
    torch.cuda.set_stream(stream)
    host = dev.cpu().pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    gb = GraphBuilder(local)
    gb.set_stream(stream.cuda_stream)
    gb.set_chunk_tiles(args.chunk_tiles)
    sb = multi_gpu.ShardedBuilder(gb, rank, world, dist, dev_t, stream)

    def barrier():
        if dist is not None:
            dist.barrier()

    def reduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    def drain(consumer=None):
        n = 0
        while True:
            c = gb.chunk_next()
            if c is None:
                return n
            n += 1
            if consumer is not None:
                consumer(c)

    def step(from_host=False):
        if world == 1:
            gb.build_from_buffer(host) if from_host else gb.build_from_device(dev.data_ptr(), n_bytes)
        else:
            sb.build_buffer(host) if from_host else sb.build_device(dev.data_ptr(), n_bytes)
        return drain()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        n_chunks = step()
    counts = gb.counts()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = gb.launch_count()
    ms = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        b.synchronize()
        ms.append(a.elapsed_time(b))
    torch.cuda.synchronize()
    barrier()
    launches = gb.launch_count() - launches0
    t_dev = reduce(sum(ms) / 1e3, "MAX")
    # end to end: the PAF travels from pinned host memory inside the timed region; what comes back per step are the
    # per-chunk row counts (the tables themselves are larger than host memory can take at these sizes)
    step(True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    torch.cuda.synchronize()
    barrier()
    t_e2e = reduce(time.perf_counter() - t0, "MAX")
    clocks = sampler.summary()

    # ---- per-kernel times of one step (events around every launch)
    gb.profile(True)
    step()
    ks = gb.kernel_stats()
    gb.profile(False)
    per_kernel = {k: {"launches_per_step": n, "ms_per_step": t, "us_per_launch": 1e3 * t / max(n, 1)} for k, (n, t) in ks.items()}
    top = sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms_per_step"])

    # ---- verification pass (not timed): digest + invariants of every chunk, rows of the sampled unitigs
    c0 = gb.counts()
    U = c0["n_unitigs"]
    t = gb.device_tables()
    un = sampled.device_name_table(t.unitig_name_off, t.unitig_names, U)
    member = torch.from_numpy(np.array([sampled.unitig_in_sample(x, args.sample_mod) for x in un], np.bool_)).cuda()
    dig = {"n_edges": 0, "n_edge_matches": 0, "edge_sum": 0, "em_sum": 0, "violations": 0}
    keys_in_order, prev_key, rows, first_key, last_key = True, -1, [], (1 << 64) - 1, 0

    def consume(c):
        nonlocal keys_in_order, prev_key, first_key, last_key
        d = gb.chunk_digest()
        for k in dig:
            dig[k] = (dig[k] + d[k]) & ((1 << 64) - 1)
        if d["n_edges"]:
            keys_in_order = keys_in_order and prev_key < d["first_key"]
            prev_key = d["last_key"]
            first_key, last_key = min(first_key, d["first_key"]), max(last_key, d["last_key"])
        rows.append(sampled.sample_rows_of_chunk(c, member))

    if world == 1:
        gb.build_from_device(dev.data_ptr(), n_bytes)
    else:
        sb.build_device(dev.data_ptr(), n_bytes)
    drain(consume)
    final = gb.counts()
    mine = sampled.concat_rows(rows)
    rn = sampled.device_name_table(t.read_name_off, t.read_names, c0["n_reads"])
    got = sampled.canon_from_ids(mine, rn, un)
    # totals over the ranks; every rank's slice must follow the previous one's
    vec = [dig[k] & 0xffffffff for k in dig] + [dig[k] >> 32 for k in dig] + \
          [first_key >> 32, first_key & 0xffffffff, last_key >> 32, last_key & 0xffffffff, n_lines_local, int(keys_in_order)]
    tv = torch.tensor(vec, dtype=torch.int64, device="cuda")
    if dist is not None:
        allv = torch.empty(world * tv.numel(), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allv, tv)
        allv = allv.view(world, -1).cpu().tolist()
    else:
        allv = [tv.cpu().tolist()]
    nk = len(dig)
    tot = {k: (sum(r[i] for r in allv) + (sum(r[nk + i] for r in allv) << 32)) & ((1 << 64) - 1) for i, k in enumerate(dig)}
    ordered, prev = all(r[2 * nk + 5] for r in allv), -1
    for r in allv:
        fk, lk = (r[2 * nk] << 32) | r[2 * nk + 1], (r[2 * nk + 2] << 32) | r[2 * nk + 3]
        if lk == 0 and fk == (1 << 64) - 1:
            continue
        ordered = ordered and prev < fk
        prev = lk
    line_base = sum(r[2 * nk + 4] for r in allv[:rank])
    payload = (got, sub_lines, (sub_local + line_base))
    if dist is not None:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
    else:
        gathered = [payload]
    sample = None
    if rank == 0:
        from oracle.pyoracle import Oracle
        g_all = np.sort(np.concatenate([g[0] for g in gathered]), order=["lo", "hi", "u"])
        lines = [ln for g in gathered for ln in g[1]]
        orig = np.concatenate([np.asarray(g[2], np.int64) for g in gathered]) if lines else np.zeros(0, np.int64)
        t0 = time.perf_counter()
        want = sampled.checker_records_for_lines(Oracle(), lines, orig)
        sample = {"unitigs_sampled": int(sum(len(set(ln[:ln.find(b"\t")] for ln in g[1])) for g in gathered)),
                  "lines": len(lines), "edge_matches_gpu": int(g_all.size), "edge_matches_checker": int(want.size),
                  "identical_by_name": bool(sampled.records_equal(g_all, want)),
                  "checker": "oracle/libms_oracle.so (pinned on the reference by tests/test_oracle.py)",
                  "checker_seconds": round(time.perf_counter() - t0, 2), "modulus": args.sample_mod}
    M_tot = counts["n_matches"]  # global (the match tables are replicated)
    parity = {"digest": tot, "slices_in_key_order": bool(ordered), "sampled": sample,
              "ok": bool(tot["violations"] == 0 and ordered and (sample is None or sample["identical_by_name"]))}
    peak, peak_src = peaks()
    EM, E = tot["n_edge_matches"], tot["n_edges"]
    roof = None
    for name, st in top:
        cnt = {"n_matches": M_tot, "n_edge_matches": EM // world, "n_edges": E // world, "n_lines": counts["n_lines"]}
        bts = kernel_bytes(name, cnt, n_bytes)
        if bts is None:
            continue
        roof = {"kernel": name, "bound": "hbm", "achieved": round(bts[0] / (st["ms_per_step"] * 1e-3) / 1e9, 1), "peak": peak,
                "unit": "GB/s", "frac": round(bts[0] / (st["ms_per_step"] * 1e-3) / 1e9 / peak, 4), "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_step_this_rank": int(bts[0]),
                "strict_frac": round(bts[1] / (st["ms_per_step"] * 1e-3) / 1e9 / peak, 4),
                "launches_per_step": st["launches_per_step"], "ms_per_step": round(st["ms_per_step"], 3),
                "share_of_step": round(st["ms_per_step"] / sum(v["ms_per_step"] for v in per_kernel.values()), 3),
                "note": "summed over the chunks of one step on rank 0 (CUDA events around every launch)"}
        if name == "pair_tiles":
            roof["frac_note"] = ON_CHIP_NOTE
        break
    step_ms = 1e3 * t_dev / args.steps
    survey = 3 * 40 * M_tot + 3 * 40 * EM + 12 * E
    if roof is not None:
        roof["construction_step"] = {"ms": round(step_ms, 3), "survey_8d_bytes": int(survey),
                                     "survey_8d_frac_of_n_gpus": round(survey / step_ms / 1e6 / (peak * world), 4)}
    cpu = cpu_baseline(args.workload) if world == 1 and args.genome_scale == 1.0 else None
    out = {"metric": METRIC, "value": M_tot * args.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "i32 keys + f64 scores", "data": "synthetic",
           "config": {"workload": f"{args.workload}-scale synthetic PAF (muchsalsa_b200.synth_gpu: G = {cfg.genome:,} bp, "
                                  f"{cfg.coverage:g}x, repeats {cfg.rep_frac:g} x {cfg.rep_kind}({cfg.rep_mean:g}) capped "
                                  f"{cfg.rep_cap}; ONE workload split over the ranks by unitig range)",
                      "paf_bytes": int(reduce(n_bytes, "SUM")), "lines": int(sum(r[2 * nk + 4] for r in allv)),
                      "kept_matches": M_tot, "reads": counts["n_reads"], "unitigs": counts["n_unitigs"],
                      "pairs": int(reduce(counts["n_pairs"], "MAX")), "edges": E, "edge_matches": EM,
                      "chunk_tiles": args.chunk_tiles, "chunks_per_step_rank0": n_chunks,
                      "l2": "flushed between steps (256 MiB memset)", "generated_in_s": round(t_gen, 1)},
           "edges_per_s": E * args.steps / t_dev, "edge_matches_per_s": EM * args.steps / t_dev,
           "e2e": {"value": M_tot * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(reduce(n_bytes, "SUM")),
                   "d2h_bytes_per_step": int(reduce(n_chunks * 80, "SUM")), "ms_per_step": 1e3 * t_e2e / args.steps,
                   "note": "PAF from pinned host memory; the tables are consumed on the device chunk by chunk, the "
                           "per-chunk row counts return to the host"},
           "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
           "kernels": {k: {kk: round(vv, 3) for kk, vv in v.items()} for k, v in top[:12]},
           "parity": parity["ok"], "parity_detail": parity,
           "device_mem_gb": round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9, 1)}
    if world > 1:
        out["exchange"] = {"what": getattr(sb, "exchange_mode", "?"), "bytes_received_per_rank": sb.exchange_bytes,
                           "ms": sb.last_exchange_ms()}
    gb.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _traffic(workload, kernel):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(tp)).get(workload, {}).get(kernel) if os.path.exists(tp) else None


def check_parity(gb, workload, rank, world, dist):
    """Is what the timed steps build the reference's result?  Every rank digests its slice of the edge tables on the
    device (msgpu_result_digest: order-independent sums + invariant violations); the sums over the ranks must equal
    the digest recorded from the CPU checkers for the same PAF (tests/golden/sum_digests.json, written by
    tests/golden/make_sum_digests.py), and the slices' key ranges must follow each other in rank order."""
    import torch
    d = gb.digest()
    keys = ("n_edges", "n_edge_matches", "edge_sum", "em_sum", "violations")
    lo_hi = [(d[k] & 0xffffffff, d[k] >> 32) for k in keys]
    v = [x for pair in lo_hi for x in pair] + [d["first_key"] >> 32, d["first_key"] & 0xffffffff,
                                               d["last_key"] >> 32, d["last_key"] & 0xffffffff]
    t = torch.tensor(v, dtype=torch.int64, device="cuda")
    if dist is not None:
        allv = torch.empty(world * t.numel(), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allv, t)
        allv = allv.view(world, -1).cpu().tolist()
    else:
        allv = [t.cpu().tolist()]
    tot = {}
    for i, k in enumerate(keys):  # 32-bit halves summed exactly, recombined mod 2^64
        lo, hi = sum(r[2 * i] for r in allv), sum(r[2 * i + 1] for r in allv)
        tot[k] = (lo + (hi << 32)) & ((1 << 64) - 1)
    ordered, prev = True, -1
    for r in allv:
        first, last = (r[10] << 32) | r[11], (r[12] << 32) | r[13]
        if (r[0] + (r[1] << 32)) == 0:
            continue  # an empty slice
        ordered = ordered and prev < first <= last
        prev = last
    gp = os.path.join(ROOT, "tests", "golden", "sum_digests.json")
    gold = json.load(open(gp)).get(f"{workload}x{world}") if os.path.exists(gp) else None
    out = {"digest": {k: tot[k] for k in keys}, "slices_in_key_order": ordered, "golden": None, "ok": None}
    if gold is not None:
        same = all(tot[k] == gold[k] for k in ("n_edges", "n_edge_matches", "edge_sum", "em_sum"))
        out["golden"] = f"tests/golden/sum_digests.json[{workload}x{world}] (recorded from the {gold['source']})"
        out["ok"] = bool(same and ordered and tot["violations"] == 0)
    else:
        out["note"] = "no golden digest for this workload: invariants only"
        out["invariants_ok"] = bool(ordered and tot["violations"] == 0)
    return out


def filters_sum_digest(f, base=0):
    """Order-independent digest of per-edge filter verdicts (the arrays of GraphBuilder.filters(), or of the
    reference's jobs): sum over the edges of a 64-bit mix of (global edge index, contraction order, deletable, kept
    orders, weight, consensus), mod 2^64.  `base` = global index of the first edge of the slice, so the digests of the
    slices of a multi-GPU build add up to the digest of the whole table."""
    import numpy as np

    def mix(x):
        x = x ^ (x >> np.uint64(30))
        x = x * np.uint64(0xbf58476d1ce4e5b9)
        x = x ^ (x >> np.uint64(27))
        x = x * np.uint64(0x94d049bb133111eb)
        return x ^ (x >> np.uint64(31))

    E = int(f["deletable"].shape[0])
    with np.errstate(over="ignore"):
        h = mix(np.arange(base, base + E, dtype=np.uint64) + np.uint64(0x9e3779b97f4a7c15))
        h = mix(h ^ f["contraction_order"].astype(np.int64).astype(np.uint64))
        small = (f["deletable"].astype(np.uint64) | (f["consensus"].astype(np.uint64) << np.uint64(8))
                 | (f["n_kept_orders"].astype(np.uint64) << np.uint64(16)))
        h = mix(h ^ small)
        h = mix(h ^ f["weight"].astype(np.uint64))
        total = int(h.sum(dtype=np.uint64)) if E else 0
    return {"n_edges": E, "sum": total & ((1 << 64) - 1),
            "contraction_edges": int((f["contraction_order"] >= 0).sum()), "deletable_edges": int(f["deletable"].sum())}


def check_filters_parity(gb, workload, rank, world, dist):
    """Are the edge-filter verdicts of the timed configuration the reference's?  Every rank digests the verdicts of its
    own edges (host copy, numpy) with the global index of its first edge as base; the sums over the ranks must equal the
    digest recorded from the REFERENCE's own jobs for the same PAF (tests/golden/sum_digests.json, key
    filters_<workload>x<N>, written by tests/golden/make_sum_digests.py --filters).  Never raises: a failure is
    reported in the line, and every rank takes part in both collectives whatever happens locally."""
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"  # (cpu: the gloo test of this function)
    E_local, err = 0, None
    try:
        E_local = int(gb.counts()["n_edges"])
    except Exception as e:  # noqa: BLE001
        err = repr(e)[:200]
    mine = torch.tensor([E_local], dtype=torch.int64, device=dev)
    if dist is not None:
        allE = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allE, mine)
        allE = allE.cpu().tolist()
    else:
        allE = [E_local]
    part = None
    if err is None:
        try:
            part = filters_sum_digest(gb.filters(), base=sum(allE[:rank]))
        except Exception as e:  # noqa: BLE001
            err = repr(e)[:200]
    vals = [0, 0, 0, 0, 0] if part is None else [1, part["sum"] & 0xffffffff, part["sum"] >> 32,
                                                  part["contraction_edges"], part["deletable_edges"]]
    t = torch.tensor(vals, dtype=torch.int64, device=dev)
    if dist is not None:
        allv = torch.empty(world * t.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allv, t)
        allv = allv.view(world, -1).cpu().tolist()
    else:
        allv = [vals]
    if not all(r[0] == 1 for r in allv):
        return {"ok": None, "error": err or "a rank could not digest its verdicts"}
    tot = {"n_edges": sum(allE), "sum": (sum(r[1] for r in allv) + (sum(r[2] for r in allv) << 32)) & ((1 << 64) - 1),
           "contraction_edges": sum(r[3] for r in allv), "deletable_edges": sum(r[4] for r in allv)}
    gp = os.path.join(ROOT, "tests", "golden", "sum_digests.json")
    gold = json.load(open(gp)).get(f"filters_{workload}x{world}") if os.path.exists(gp) else None
    out = {"digest": tot, "golden": None, "ok": None}
    if gold is not None:
        out["golden"] = f"tests/golden/sum_digests.json[filters_{workload}x{world}] (recorded from the reference's own jobs)"
        out["ok"] = bool(all(tot[k] == gold[k] for k in ("n_edges", "sum", "contraction_edges", "deletable_edges")))
    return out


def bench_sequences(gb, workload, stream, n_req=1 << 20):
    """K8 (d): load both sequence files into the resident store, then one batch of n_req
    requests shaped like assemblePath's (ap.cpp:352-579: flanks of 50-5000 bases of reads and unitigs, half of them
    reverse complemented) through msgpu_seq_fetch_batch_device (everything in HBM) and through the host-buffer entry;
    the reference's get{Illumina,Nanopore}Sequence loop on a bounded sample of the same requests beside it."""
    import numpy as np
    import torch
    from muchsalsa_b200 import synth
    if workload not in ("small", "ecoli", "yeast"):
        return None
    ufa, nfa = synth.generate_sequences(workload)
    t0 = time.perf_counter()
    gb.seq_load_buffers(ufa, nfa, False)
    load_s = time.perf_counter() - t0
    g = synth.generate_columns(synth.CONFIGS[workload])["_geom"]
    rng = np.random.Generator(np.random.PCG64(99))
    which = (rng.random(n_req) < 0.8).astype(np.uint32)  # mostly read flanks, some unitig anchors
    idx = np.where(which == 1, rng.integers(0, g["rlen"].size, n_req), rng.integers(0, g["ulen"].size, n_req))
    names_r, names_u = g["rname"], np.arange(g["ulen"].size)
    # sequence ids continue the numbering of the last build; resolve by name like the shim does
    rid = np.array([gb.seq_id(1, f"r{int(x)}") for x in names_r], np.uint32)
    uid = np.array([gb.seq_id(0, f"u{int(x)}") for x in names_u], np.uint32)
    ids = np.where(which == 1, rid[np.minimum(idx, rid.size - 1)], uid[np.minimum(idx, uid.size - 1)]).astype(np.uint32)
    length = np.where(which == 1, g["rlen"][np.minimum(idx, g["rlen"].size - 1)],
                      g["ulen"][np.minimum(idx, g["ulen"].size - 1)]).astype(np.int64)
    want = rng.integers(50, 5001, n_req)
    left = (rng.random(n_req) * np.maximum(length - want, 1)).astype(np.int64)
    right = np.minimum(left + want - 2, length - 1)  # get*Sequence returns bases left .. right + 1 inclusive
    reqs = np.stack([which, ids, left.astype(np.uint32), right.astype(np.uint32),
                     (rng.random(n_req) < 0.5).astype(np.uint32)], axis=1).astype(np.uint32)
    d_reqs = torch.from_numpy(reqs.view(np.int32)).cuda()
    bound = int(np.minimum(want, length).sum()) + 64
    d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n_req + 1, dtype=torch.int64, device="cuda")
    d_st = torch.empty(n_req, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        gb.seq_fetch_batch_device(d_reqs.data_ptr(), n_req, d_out.data_ptr(), bound, d_off.data_ptr(), d_st.data_ptr())

    for _ in range(3):
        run()
    bases = int(d_off[-1].item())
    ms = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        run()
        b.record(stream)
        b.synchronize()
        ms.append(a.elapsed_time(b))
    gb.profile(True)
    flush.zero_()
    run()
    ks = gb.kernel_stats()
    gb.profile(False)
    peak, peak_src = peaks()
    copy_ms = ks.get("slice_copy", (1, float("nan")))[1]
    dev_ms = sum(ms) / len(ms)
    # host-buffer entry (requests and output in host memory, copies inside the timed region)
    h_out, h_off, h_st = np.empty(bound, np.uint8), np.empty(n_req + 1, np.uint64), np.empty(n_req, np.uint8)
    gb.seq_fetch_batch_host(reqs, h_out, h_off, h_st)
    t0 = time.perf_counter()
    gb.seq_fetch_batch_host(reqs, h_out, h_off, h_st)
    host_s = time.perf_counter() - t0
    ok = bool((d_st == 0).all().item()) and int(h_off[-1]) == bases
    out = {"what": "msgpu_seq_load_buffers + msgpu_seq_fetch_batch(_device): strSlice / getReverseComplement / "
                   "get{Illumina,Nanopore}Sequence (SequenceUtils.cpp:27-85) over the store resident in HBM",
           "store_bytes": len(ufa) + len(nfa), "load_s": round(load_s, 3), "requests": n_req, "bases": bases,
           "request_shape": "80 % read flanks / 20 % unitig anchors, 50-5000 bases, half reverse complemented",
           "device_ms": round(dev_ms, 3), "bases_per_s": bases / dev_ms * 1e3,
           "host_buffers_ms": round(host_s * 1e3, 3), "host_buffers_bases_per_s": bases / host_s,
           "all_requests_ok": ok,
           "roofline": {"kernel": "slice_copy", "bound": "hbm", "algorithmic_bytes_per_launch": 2 * bases,
                        "traffic": _traffic(workload, "slice_copy"),
                        "traffic_note": "ncu capture of scripts/profile_seq.py (read-flank requests only: 2.48e9 bases)",
                        "us_per_launch": round(copy_ms * 1e3, 2), "achieved": round(2 * bases / copy_ms / 1e6, 1),
                        "peak": peak, "unit": "GB/s", "frac": round(2 * bases / copy_ms / 1e6 / peak, 4),
                        "peak_source": peak_src, "definition": "SURVEY.md §8(d): requested bases x (1 read + 1 write)"},
           "kernels_ms": {k: round(v[1], 4) for k, v in ks.items()}}
    # the reference numbers sequences in file order when no PAF primed its registries: same requests, its ids
    reqs_ref = reqs.copy()
    reqs_ref[:, 1] = idx.astype(np.uint32)
    out["cpu_reference"] = seq_cpu_baseline(ufa, nfa, reqs_ref)
    return out


def seq_cpu_baseline(ufa, nfa, reqs, n_sample=20000):
    """The reference's own loop (oracle/_ref/msref_cli seq: SequenceAccessor + get*Sequence, unmodified) on the first
    n_sample requests of the same batch, 1 thread and all host threads (the accessor serialises on a mutex per file)."""
    cli = os.path.join(ROOT, "oracle", "_ref", "msref_cli")
    if not os.path.exists(cli):
        return {"unavailable": "oracle/_ref/msref_cli not built"}
    import numpy as np
    td = tempfile.mkdtemp(prefix="msseq_")
    try:
        pu, pn, pr = os.path.join(td, "u.fa"), os.path.join(td, "n.fa"), os.path.join(td, "reqs.bin")
        open(pu, "wb").write(ufa)
        open(pn, "wb").write(nfa)
        sample = reqs[:n_sample].copy()
        open(pr, "wb").write(sample.astype(np.int32).tobytes())
        res = {}
        for threads in sorted({1, os.cpu_count() or 1}):
            p = subprocess.run([cli, "seq", pu, pn, pr, str(threads)], capture_output=True, text=True, timeout=900)
            if p.returncode != 0:
                res[f"threads_{threads}"] = {"error": p.returncode}
                continue
            r = json.loads(p.stdout.strip().splitlines()[-1])
            res[f"threads_{threads}"] = {"bases_per_s": r["bases"] / r["seconds"], "seconds": round(r["seconds"], 3),
                                         "requests": r["requests"], "bases": r["bases"]}
        res["sample"] = f"first {len(sample)} requests of the batch"
        res["kind"] = "reference"
        return res
    finally:
        import shutil
        shutil.rmtree(td, ignore_errors=True)


def result_bytes(c, gb):
    R, U, M, E, EM = c["n_reads"], c["n_unitigs"], c["n_matches"], c["n_edges"], c["n_edge_matches"]
    names = 0
    try:
        r = gb.result()
        names = sum(len(x) for x in r.read_names) + sum(len(x) for x in r.unitig_names)
    except Exception:
        pass
    return int(R * 8 + (R + 1) * 8 + (U + 1) * 8 + names + M * (4 * 8 + 8 + 1) + E * 8 + (E + 1) * 8 + EM * 25)


def _run_ref_cli(paf_path, threads, timeout=1800):
    cli = os.path.join(ROOT, "oracle", "_ref", "msref_cli")
    p = subprocess.run([cli, paf_path, str(threads)], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        return None, p.returncode
    return json.loads(p.stdout.strip().splitlines()[-1]), 0


def cpu_baseline(workload, steps=1):
    """The reference's own read() + calculateEdges() on this host (oracle/_ref/msref_cli), on a bounded sample."""
    cli = os.path.join(ROOT, "oracle", "_ref", "msref_cli")
    shard = CPU_SAMPLE.get(workload)
    sample = "whole workload" if shard is None else f"unitig shard {shard[0]} of {shard[1]} of the workload"
    paf = make_workload(workload, shard)
    with tempfile.NamedTemporaryFile(suffix=".paf", delete=False) as f:
        f.write(paf)
        path = f.name
    try:
        if not os.path.exists(cli):
            return cpu_baseline_port(paf, sample)
        ncpu = os.cpu_count() or 1
        r1, rc1 = _run_ref_cli(path, 1)
        if r1 is None:
            return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample,
                    "note": f"reference failed with status {rc1}"}
        res, threads, note = r1, 1, None
        if ncpu > 1:  # all host threads: the reference either exits -11 (race on the adjacency list) or crawls on its locks
            multi, rc = _run_ref_cli(path, ncpu)
            if multi is None:
                note = (f"reference with {ncpu} threads exited with status {rc} (data race on the adjacency list, "
                        "Graph.cpp:96-116 vs 212-230); timed with 1 thread, its only reliable mode")
            else:
                tn, t1 = (sum(x[k] for k in ("t_index", "t_read", "t_edges")) for x in (multi, r1))
                note = f"1 thread {t1:.2f} s, {ncpu} threads {tn:.2f} s on this sample; the faster is reported"
                if tn < t1:
                    res, threads = multi, ncpu
        t = res["t_index"] + res["t_read"] + res["t_edges"]
        chain = None
        try:  # the stage after construction on the same sample (1 thread; edges are independent jobs in the reference)
            p = subprocess.run([cli, path, "1", "chain"], capture_output=True, text=True, timeout=1800)
            if p.returncode == 0:
                c = json.loads(p.stdout.strip().splitlines()[-1])
                chain = {"edges_per_s": c["n_edges"] / c["t_chain"], "seconds": round(c["t_chain"], 3), "cores": 1,
                         "edges": c["n_edges"], "sample": sample,
                         "filters": {"edges_per_s": c["n_edges"] / max(c["t_filters"], 1e-9),
                                     "seconds": round(c["t_filters"], 3)}}
        except Exception as e:  # the baseline must not take the bench down
            chain = {"error": repr(e)}
        return {"chain": chain, "value": res["n_vm"] / t, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                "seconds": round(t, 3), "edges_per_s": res["n_edges"] / t, "edge_matches_per_s": res["n_em"] / t,
                "phases_s": {k: res[k] for k in ("t_index", "t_read", "t_edges")},
                "rows": {k: res[k] for k in ("n_reads", "n_vm", "n_edges", "n_em")}, "host_cpus": ncpu, "note": note}
    finally:
        os.unlink(path)


def cpu_baseline_port(paf, sample):
    from oracle.pyoracle import Oracle
    o = Oracle()
    t0 = time.perf_counter()
    t = o.build(paf)
    dt = time.perf_counter() - t0
    c = t.counts
    return {"value": c["vm"] / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
            "seconds": round(dt, 3), "edges_per_s": c["edges"] / dt, "note": "oracle/_ref absent: timed the C port"}


def workload_label(workload, world):
    return (f"{workload}-scale synthetic PAF x{world} (muchsalsa_b200.synth, seed fixed; at N GPUs the genome is N "
            "times longer and rank r holds the r-th chunk of the PAF)")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref/msref_cli = the unmodified
    reference sources compiled by oracle/Makefile: BlastFileAccessor + BlastFileReader::read() +
    MatchMap::calculateEdges()) on this box's host cores, on OUR arm's config.

    Every timed step builds the WHOLE workload when K steps of it fit the time budget (MSGPU_REF_BUDGET_S, default
    1500 s: yeast at N = 1 does, ~45 s per step on one core); otherwise the largest 1/d unitig shard of it that fits
    (yeast x N at N > 1, thaliana, human) and `cpu_baseline.sample` says so.  The rate is per kept match either way.
    The reference is only reliable with one worker thread (it races on the adjacency list and exits -11 with more);
    that is probed ONCE, on a small shard, not per step.  Warm-up steps run on the small shard too: a fresh CPU
    process has no state to warm beyond the page cache and the clocks."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from muchsalsa_b200 import synth
    world = max(1, args.gpus)
    cfg = synth.scaled(args.workload, world)
    budget = float(os.environ.get("MSGPU_REF_BUDGET_S", "1500"))
    ncpu = os.cpu_count() or 1

    def write(paf):
        f = tempfile.NamedTemporaryFile(suffix=".paf", delete=False)
        f.write(paf)
        f.close()
        return f.name

    probe_d = {"small": 1, "ecoli": 1}.get(args.workload, 24 * world)
    probe_path = write(synth.columns_to_paf(synth.generate_columns(cfg, (0, probe_d) if probe_d > 1 else None)))
    paths = [probe_path]
    try:
        # thread count: whichever of {1, all host threads} is faster on the probe shard — with more than one thread the
        # reference either exits -11 (unlocked adjacency-list reads race the writer, Graph.cpp:96-116 vs 212-230) or
        # crawls on its four global locks
        t0 = time.perf_counter()
        r, rc = _run_ref_cli(probe_path, 1)
        probe_s = time.perf_counter() - t0
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/msref_cli failed (status {rc})"}))
            return
        t1 = r["t_index"] + r["t_read"] + r["t_edges"]
        threads, note = 1, None
        if ncpu > 1:
            t0 = time.perf_counter()
            multi, rc = _run_ref_cli(probe_path, ncpu)
            multi_s = time.perf_counter() - t0
            if multi is None:
                note = (f"{ncpu} threads: exit status {rc} (data race on the adjacency list); 1 thread is the "
                        "reference's only reliable mode")
            else:
                tn = multi["t_index"] + multi["t_read"] + multi["t_edges"]
                note = f"probe shard: 1 thread {t1:.2f} s, {ncpu} threads {tn:.2f} s; timed with the faster"
                if tn < t1:
                    threads, probe_s = ncpu, multi_s
        for _ in range(max(0, args.warmup - 1)):
            _run_ref_cli(probe_path, threads)
        # the largest share of the workload whose K steps fit the budget
        est_full = probe_s * probe_d * 1.2
        d = 1
        while args.steps * est_full / d > budget and d < probe_d:
            d += 1
        if d == 1:
            paf = b"".join(synth.rank_chunk(args.workload, rr, world) for rr in range(world))
            sample = "whole workload"
        elif d >= probe_d:
            paf, sample = None, f"unitig shard 0 of {probe_d} of the workload"
        else:
            paf = synth.columns_to_paf(synth.generate_columns(cfg, (0, d)))
            sample = f"unitig shard 0 of {d} of the workload"
        path = probe_path if paf is None else write(paf)
        paths.append(path)
        tot, rows = 0.0, None
        for _ in range(args.steps):
            r, rc = _run_ref_cli(path, threads)
            if r is None:
                threads, note = 1, f"{ncpu}-thread run crashed mid-benchmark (status {rc}); fell back to 1 thread"
                r, rc = _run_ref_cli(path, 1)
            tot += r["t_index"] + r["t_read"] + r["t_edges"]
            rows = r
        value = rows["n_vm"] * args.steps / tot
        config = {"workload": workload_label(args.workload, world)}
        gp = os.path.join(ROOT, "tests", "golden", "sum_digests.json")
        gold = json.load(open(gp)).get(f"{args.workload}x{world}") if os.path.exists(gp) else None
        if d == 1:
            config.update(kept_matches=rows["n_vm"], reads=rows["n_reads"], edges=rows["n_edges"],
                          edge_matches=rows["n_em"])
        elif gold:
            config.update(kept_matches=gold["n_matches"], reads=gold["n_reads"], edges=gold["n_edges"],
                          edge_matches=gold["n_edge_matches"])
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32 keys + f64 scores",
               "data": "synthetic", "config": config,
               "edges_per_s": rows["n_edges"] * args.steps / tot,
               "edge_matches_per_s": rows["n_em"] * args.steps / tot,
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                                "sample": sample, "host_cpus": ncpu, "note": note,
                                "rows_per_step": {k: rows[k] for k in ("n_reads", "n_vm", "n_edges", "n_em")},
                                "warmup_on": f"unitig shard 0 of {probe_d}" if probe_d > 1 else "whole workload"},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out), flush=True)
    finally:
        for p in set(paths):
            try:
                os.unlink(p)
            except OSError:
                pass


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner there) get stderr instead."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(keep, "w")


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MSGPU_WORKLOAD", "yeast"))
    ap.add_argument("--no-sequence", action="store_true", help="skip the K8 (sequence store) block")
    ap.add_argument("--chunk-tiles", type=int, default=65536, help="thaliana / human: pairing tiles per chunk")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="thaliana / human: scale the genome (development)")
    ap.add_argument("--e2e-chunk-tiles", type=int, default=2048,
                    help="the e2e leg downloads the edge tables chunk by chunk, overlapped with the pairing; 0 = resident")
    ap.add_argument("--sample-mod", type=int, default=512, help="thaliana / human: 1 unitig in this many is compared "
                                                                  "with the CPU checker by name")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in BIG_WORKLOADS:
        run_ours_big(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
