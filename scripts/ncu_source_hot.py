#!/usr/bin/env python
"""Per-source-line hot spots of a kernel in an .ncu-rep (needs -lineinfo and --import-source on):
aggregates the SASS rows of `ncu --page source --csv --print-source cuda,sass` by CUDA source line.
Usage: python scripts/ncu_source_hot.py file.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # find header rows (several files / functions may follow each other)
    agg = defaultdict(lambda: [0, 0, ""])
    hdr = None
    cur_file = ""
    tot_s = tot_i = 0
    sass = []
    last_line = ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            iL, iS = 0, 1
            iSamp = hdr.index("# Samples")
            iInst = hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            s, n = int(r[iSamp] or 0), int(r[iInst] or 0)
        except ValueError:
            continue
        if not r[iL].strip():   # a SASS row under the preceding source line: kept for the SASS top list only
            sass.append((s, n, last_line, r[3].strip()))
            continue
        last_line = f"{cur_file}:{r[iL]}"
        key = (cur_file, r[iL])
        agg[key][0] += s
        agg[key][1] += n
        agg[key][2] = r[iS]
        tot_s += s
        tot_i += n
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for key, (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * s / max(tot_s, 1):5.1f}% samples {100.0 * n / max(tot_i, 1):5.1f}% inst  {key[0]}:{key[1]:>5}  {src.strip()[:110]}")
    print("-- SASS instructions with the most stall samples")
    for s, n, line, ins in sorted(sass, key=lambda t: -t[0])[:top]:
        print(f"{100.0 * s / max(tot_s, 1):5.1f}% samples  {line:>24}  {ins[:90]}")


if __name__ == "__main__":
    main()

