#!/usr/bin/env python
"""Developer micro-benchmark: cf_fusion_fwd per scale of a bench workload, out of place and in place,
each call timed alone with CUDA events (median of --reps).  Not part of the bench contract."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcf_b200 as dcf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg1")
    ap.add_argument("--mode", default=None)
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--groups", default="")
    ap.add_argument("--empty", action="store_true", help="no LiDAR points: every cell is a pure bev -> out copy")
    a = ap.parse_args()
    wl = dcf.synthetic.make_workload(a.workload, seed=100)
    mode = a.mode or wl["workload"]["mode"]
    dev = torch.device("cuda")
    ops = dcf.ops
    to = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    if a.empty:
        wl["num_points"][:] = 0
    points, counts, img = to(wl["points"]), to(wl["num_points"]), to(wl["img_feat"])
    grid = ops.BucketGrid(*dcf.geometry.bucket_grid(wl["config"], None))
    size = (float(wl["config"]["image_width"]), float(wl["config"]["image_height"]))
    start, srt, _ = ops.bucket_points(points, counts, grid)
    feat, _ = ops.point_gather(img, points, counts, calib=wl["calib"], img_size=size)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fine = None
    want = set(int(g) for g in a.groups.split(",") if g)
    for sc in wl["scales"]:
        w = [to(x) for x in sc["weights"]]
        T = ops.point_mlp1(feat, points, counts, w[0], w[1], mode=mode)
        if fine is None:
            knn = ops.knn_query(start, srt, grid, sc["H"], sc["W"], sc["geom"], wl["radius"], wl["k"])
            fine, fs = knn, sc["stride"]
        else:
            knn = ops.knn_subsample(fine, sc["stride"] // fs, sc["H"], sc["W"])
        if want and sc["group"] not in want:
            continue
        bev = to(sc["bev"])
        pk = ops.PackedWeights()
        packed = pk.w23(w[2], w[4], mode)
        res = {}
        for name, inplace in (("out_of_place", False), ("in_place", True)):
            out = bev.clone() if inplace else torch.empty_like(bev)
            src = out if inplace else bev
            ts = []
            for _ in range(a.reps + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.fusion_fwd(src, T, knn, sc["geom"], w[0], w[2], w[3], w[4], w[5], mode=mode, packed=packed, out=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[name] = float(np.median(ts[3:]))
        live = float((knn[..., 0] >= 0).float().mean())
        print(f"g{sc['group']} C={sc['C']} {sc['H']}x{sc['W']} live={live:.3f}  out_of_place {res['out_of_place']*1e3:.1f} us  "
              f"in_place {res['in_place']*1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
