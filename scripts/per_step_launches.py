#!/usr/bin/env python
"""Per-kernel time of ONE step from an `ncu --metrics gpu__time_duration.sum --csv` launch list that covers several
identical steps (e.g. scripts/train_step_bench.py --steps 1 = 3 warm-up steps + 1 timed):  python per_step_launches.py
<csv> <steps> [--seq]"""
import collections
import csv
import sys

path, steps = sys.argv[1], int(sys.argv[2])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
per, seq = collections.OrderedDict(), []
for row in csv.DictReader(lines):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void ", "").replace("<", "", 1) if False else \
        row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
    per.setdefault(k, [0, 0.0])
    per[k][0] += 1
    per[k][1] += v
    seq.append((k, v, row["Grid Size"]))
tot = 0.0
print("| kernel | launches / step | us / step |\n|---|---|---|")
for k, (c, v) in per.items():
    print(f"| `{k}` | {c / steps:.1f} | {v / steps:.1f} |")
    tot += v / steps
print(f"| total | {len(seq) / steps:.0f} | {tot:.0f} |")
if "--seq" in sys.argv:
    n = len(seq) // steps
    for k, v, g in seq[-n:]:
        print(f"{k:34s} {v:8.1f} {g}")
