#!/usr/bin/env python
"""Developer micro-benchmark: cf_knn_query on the finest scale of a bench workload, timed alone with CUDA events
(median of --reps, L2 flushed between calls).  Not part of the bench contract."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcf_b200 as dcf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg1")
    ap.add_argument("--reps", type=int, default=15)
    a = ap.parse_args()
    wl = dcf.synthetic.make_workload(a.workload, seed=100)
    dev = torch.device("cuda")
    ops = dcf.ops
    to = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    points, counts = to(wl["points"]), to(wl["num_points"])
    grid = ops.BucketGrid(*dcf.geometry.bucket_grid(wl["config"], None))
    start, srt, _ = ops.bucket_points(points, counts, grid)
    sc = wl["scales"][0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(a.reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        knn = ops.knn_query(start, srt, grid, sc["H"], sc["W"], sc["geom"], wl["radius"], wl["k"])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[3:])
    print(json.dumps({"workload": a.workload, "H": sc["H"], "W": sc["W"], "k": wl["k"], "frames": int(points.shape[0]),
                      "knn_ms": ts[len(ts) // 2], "live_fraction": float((knn[..., 0] >= 0).float().mean()),
                      "checksum": int(knn.long().sum())}))


if __name__ == "__main__":
    main()
