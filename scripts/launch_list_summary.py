#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: split it into constructions (a
construction starts with newline_count) and print, per construction, its launches, the sum of their durations and the
share of the two dominant kernels.

    python scripts/launch_list_summary.py profiles/r2_ncu_launches_bench_yeast.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("msgpu::", "").replace("void ", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    seq.append((name, v))
starts = [i for i, (n, _) in enumerate(seq) if n.startswith("newline_count")]
print(f"{len(seq)} launches, {len(starts)} constructions (durations are ncu's: serialised, cold caches)")
print(f"{'#':>3} {'launches':>8} {'sum us':>9} {'pair_tiles us':>14} {'share':>6} {'pair_emit us':>13} {'share':>6}  notes")
for bi, s in enumerate(starts):
    b = seq[s:(starts[bi + 1] if bi + 1 < len(starts) else len(seq))]
    tot = sum(v for _, v in b)
    pt = [v for n, v in b if n.startswith("pair_tiles")]
    pe = [v for n, v in b if n.startswith("pair_emit")]
    notes = []
    if len(pt) > 1:
        notes.append(f"chunked: {len(pt)} tile ranges (the e2e leg)")
    if any(n.startswith(("chain", "contraction", "edge_verdicts")) for n, _ in b):
        notes.append("+ chaining / filter stages")
    if any(n.startswith("digest") for n, _ in b):
        notes.append("+ digest (parity check)")
    if not pt:
        notes.append("no pairing (tail of the list)")
    print(f"{bi:3d} {len(b):8d} {tot:9.1f} {sum(pt):14.1f} {100 * sum(pt) / tot:5.1f}% {sum(pe):13.1f} {100 * sum(pe) / tot:5.1f}%  {'; '.join(notes)}")
