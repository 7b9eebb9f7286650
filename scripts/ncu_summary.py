#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key counters per kernel + top source lines by stall samples.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
for ri, r in enumerate(rows[2:]):
    print(f"=== [{ri}] {names[ri][:70]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {k:90s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur_file, cur_fn, agg = None, None, {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        cur_fn = r[1].split("(")[0].split("::")[-1]
        continue
    if len(r) < 8 or not r[0].strip().isdigit():
        continue
    try:
        agg.setdefault(cur_fn, []).append((cur_file, int(r[0]), r[1].strip(), int(r[4] or 0), int(r[7] or 0)))
    except ValueError:
        pass
for fn, a in agg.items():
    ts, ti = sum(x[3] for x in a) or 1, sum(x[4] for x in a) or 1
    print(f"=== source lines of {fn}: {ts} samples, {ti} warp instructions")
    for x in sorted(a, key=lambda x: -x[3])[:topn]:
        print(f"  {x[0]}:{x[1]:4d} samp {100 * x[3] / ts:5.1f}% inst {100 * x[4] / ti:5.1f}%  {x[2][:95]}")
