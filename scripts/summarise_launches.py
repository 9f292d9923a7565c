#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_* --csv` launch list: one eager step, per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
iN, iM, iV, iID = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = {}
for r in rows[hi + 1:]:
    if len(r) > iV:
        per.setdefault(r[iID], {"name": r[iN]})[r[iM]] = float(r[iV].replace(",", ""))
ids = sorted(per, key=int)
starts = [k for k, i in enumerate(ids) if "k_bucket_hist" in per[i]["name"]]
s, e = (starts[-2], starts[-1]) if len(starts) > 1 else (starts[-1], len(ids))
agg = collections.OrderedDict()
for i in ids[s:e]:
    d = per[i]
    n = d["name"].split("(")[0].replace("void cf::<unnamed>::", "").replace("void ", "").replace("unnamed>::", "")
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0) / 1e3
    a[2] += d.get("dram__bytes_read.sum", 0) / 1e6
    a[3] += d.get("dram__bytes_write.sum", 0) / 1e6
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | us | share | DRAM read MB | DRAM written MB |\n|---|---|---|---|---|---|")
for n, a in agg.items():
    print(f"| `{n}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f} % | {a[2]:.0f} | {a[3]:.0f} |")
print(f"| total | {sum(a[0] for a in agg.values())} | {tot:.0f} | | | |")
