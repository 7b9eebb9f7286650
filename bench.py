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
    if world == 1:
        paf = make_workload(args.workload)
    else:
        cfg = synth.scaled(args.workload, world)
        paf = synth.columns_to_paf(synth.generate_columns(cfg, (rank, world), sentinel=(rank == world - 1)))
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

    # ---- timed region 2: end to end through the C ABI with host buffers
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
    chain_ms, filter_ms = [], []
    for it in range(min(args.steps, 3) + 1):
        step_device()
        gb.chain_overlaps(300)
        gb.edge_filters(300)
        if it:  # the first call grows the arena (cudaMalloc inside the timed span): warm-up
            chain_ms.append(gb.chain_ms())
            filter_ms.append(gb.edge_filters_ms())
    chain_t = reduce(sum(chain_ms) / len(chain_ms), "MAX")
    filter_t = reduce(sum(filter_ms) / len(filter_ms), "MAX")

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
        "config": {"workload": f"{args.workload}-scale synthetic PAF x{world} (muchsalsa_b200.synth, seed fixed; at N "
                               "GPUs the genome is N times longer and rank r holds the r-th chunk of the PAF)",
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
    out["edge_filters"] = {
        "what": "per-edge findContractionEdges (+ sanityCheck), findDeletableEdges, computeBitweight "
                "(src/main.cpp:416-460, 534-573, sc.cpp) on the device-resident order tables; identical to the "
                "reference (tests/test_filters_gpu.py)",
        "ms": round(filter_t, 3), "edges_per_s": n_edges / filter_t * 1e3,
        "cpu_reference": ((cpu or {}).get("chain") or {}).get("filters")}
    if world > 1:
        x = [v for v in xch_ms if v is not None]
        x_ms = reduce(sum(x) / len(x), "MAX") if x else None
        out["exchange"] = {"what": "all-gather of the ranks' match-table blobs (NCCL over NVLink)",
                           "bytes_received_per_rank": sb.exchange_bytes, "ms": x_ms,
                           "gbs_per_rank": (sb.exchange_bytes / x_ms / 1e6) if x_ms else None,
                           "nvlink_peak_gbs": 770.0}
    gb.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


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
        multi, rc = _run_ref_cli(path, ncpu) if ncpu > 1 else (None, 0)
        res, threads, note = multi, ncpu, None
        if multi is None:
            note = (f"reference with {ncpu} threads exited with status {rc} (data race on the adjacency list, "
                    "Graph.cpp:96-116 vs 212-230); timed with 1 thread, its only reliable mode")
            best = None
            for _ in range(steps):
                r1, rc1 = _run_ref_cli(path, 1)
                if r1 is None:
                    return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample,
                            "note": f"reference failed with status {rc1}"}
                t = r1["t_index"] + r1["t_read"] + r1["t_edges"]
                if best is None or t < best[0]:
                    best = (t, r1)
            res, threads = best[1], 1
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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shard = CPU_SAMPLE.get(args.workload)
    paf = make_workload(args.workload, shard)
    with tempfile.NamedTemporaryFile(suffix=".paf", delete=False) as f:
        f.write(paf)
        path = f.name
    try:
        ncpu = os.cpu_count() or 1
        threads, note = ncpu, None
        probe, rc = _run_ref_cli(path, ncpu) if ncpu > 1 else (None, 0)
        if probe is None:
            threads = 1
            note = (f"{ncpu} threads: exit status {rc} (unlocked adjacency-list reads race the writer, "
                    "Graph.cpp:96-116 vs 212-230); 1 thread is the reference's only reliable mode")
        for _ in range(args.warmup):
            _run_ref_cli(path, threads)
        tot, rows = 0.0, None
        for _ in range(args.steps):
            r, rc = _run_ref_cli(path, threads)
            if r is None:
                threads, note = 1, f"{ncpu}-thread run crashed mid-benchmark (status {rc}); fell back to 1 thread"
                r, rc = _run_ref_cli(path, 1)
            tot += r["t_index"] + r["t_read"] + r["t_edges"]
            rows = r
        value = rows["n_vm"] * args.steps / tot
        sample = "whole workload" if shard is None else f"unitig shard {shard[0]} of {shard[1]} of the workload"
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32 keys + f64 scores",
               "data": "synthetic",
               "config": {"workload": f"{args.workload}-scale synthetic PAF (muchsalsa_b200.synth, seed fixed)",
                          "sample": sample, "kept_matches": rows["n_vm"], "edges": rows["n_edges"],
                          "edge_matches": rows["n_em"]},
               "edges_per_s": rows["n_edges"] * args.steps / tot,
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                                "sample": sample, "host_cpus": ncpu, "note": note},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
    finally:
        os.unlink(path)


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
