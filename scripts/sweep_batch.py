#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput sweep over the batch size -- the fusion hot path of configs[1] (K=5, five scales of a
700x800 BEV, fp32 mode) plus the greedy SAT NMS on 2 000 boxes per frame.  One GPU, or `torchrun --nproc-per-node N`: every
rank runs its own batch of that size (frames are independent, no collective on the path), the step time is the maximum over
the ranks and the rate the aggregate.  Frames of the 4-frame synthetic workload are repeated to fill larger batches.
Prints one JSON line per batch size on rank 0; not part of the bench contract."""
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
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = dcf.dist_util.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dcf.dist_util.init("nccl", dev)
    wl = dcf.synthetic.make_workload("cfg1", seed=dcf.dist_util.rank_seed(100, rank))
    mode = wl["workload"]["mode"]
    to = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    base = {"points": to(wl["points"]), "counts": to(wl["num_points"]), "img": to(wl["img_feat"]),
            "bev": [to(sc["bev"]) for sc in wl["scales"]]}
    grid = dcf.ops.BucketGrid(*dcf.geometry.bucket_grid(wl["config"], None))
    size = (float(wl["config"]["image_width"]), float(wl["config"]["image_height"]))
    layers = []
    for sc in wl["scales"]:
        layer = dcf.ContinuousFusion(base["img"].shape[1], sc["C"], k=wl["k"], radius=wl["radius"], geom=sc["geom"], mode=mode).to(dev)
        with torch.no_grad():
            for prm, w in zip((layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias, layer.fc3.weight,
                               layer.fc3.bias), sc["weights"]):
                prm.copy_(to(w))
        layers.append(layer.eval())
    boxes4 = np.stack([dcf.synthetic.nms_boxes(500 + f, 2000) for f in range(4)])
    rep = lambda t, B: t.repeat((B + 3) // 4, *([1] * (t.dim() - 1)))[:B].contiguous()
    for B in [int(x) for x in a.batches.split(",")]:
        pts, cnt, img = rep(base["points"], B), rep(base["counts"], B), rep(base["img"], B)
        bevs = [rep(t, B) for t in base["bev"]]
        boxes = torch.zeros((B, 2048, 7), device=dev)
        boxes[:, :2000] = rep(to(boxes4), B)
        bcnt = torch.full((B,), 2000, dtype=torch.int32, device=dev)

        def step():
            with torch.no_grad():
                frames = dcf.FrameContext(pts, cnt, grid)
                frames.gather(img, calib=wl["calib"], img_size=size)
                outs = dcf.fuse_scales(frames, layers, bevs)
                keep, kc = dcf.ops.nms_sat(boxes, bcnt)
            return outs, kc

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()      # as in bench.py: the ~25 short launches of a step are replayed as one graph
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        dcf.dist_util.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            graph.replay()
        e1.record()
        dcf.dist_util.barrier()
        torch.cuda.synchronize()
        ms = dcf.dist_util.max_over_ranks([e0.elapsed_time(e1) / a.steps], device=dev)[0]
        if rank == 0:
            print(json.dumps({"batch_per_gpu": B, "n_gpus": world, "ms_per_step": round(ms, 4), "frames_per_sec": round(B * world / ms * 1e3, 1),
                              "includes": "fusion (5 scales, fp32 mode, out of place) + SAT NMS on 2000 boxes/frame",
                              "launch": "cuda_graph_replay"}), flush=True)
        del graph, pts, cnt, img, bevs, boxes
        torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
