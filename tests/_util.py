"""Shared helpers of the GPU parity tests: run the product path (CUDA, through the C ABI) and the CPU oracle
on the same seeded inputs."""
import numpy as np
import torch


def rel_err(out, ref):
    """Appendix A13: ||out - ref||_inf / max(||ref||_inf, 1e-6), per tensor."""
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(out - ref).max() / max(np.abs(ref).max(), 1e-6)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def cuda_knn(dcf, wl, scale, cell=None, radius=None, k=None):
    """Product path: bucket + knn for one scale of a synthetic workload -> (B,H,W,K) int32 numpy."""
    grid = dcf.ops.BucketGrid(*dcf.geometry.bucket_grid(wl["config"], cell))
    pts = dev(wl["points"])
    cnt = dev(wl["num_points"])
    start, srt, _ = dcf.ops.bucket_points(pts, cnt, grid)
    knn = dcf.ops.knn_query(start, srt, grid, scale["H"], scale["W"], scale["geom"], radius or wl["radius"],
                            k or wl["k"])
    torch.cuda.synchronize()
    return knn.cpu().numpy(), (start.cpu().numpy(), srt.cpu().numpy())


def oracle_knn(oracle, wl, scale, radius=None, k=None):
    r = np.float32(radius or wl["radius"])
    x0, y0, dx, dy = scale["geom"]
    return np.stack([oracle.knn_bruteforce(wl["points"][b], int(wl["num_points"][b]), scale["H"], scale["W"], x0, y0,
                                           dx, dy, r * r, k or wl["k"]) for b in range(wl["points"].shape[0])])


def cuda_fusion(dcf, wl, mode, use_uv=False, channels_last=False, inplace=False):
    """Product path through the nn.Module: returns [out per scale], [knn per scale].  inplace: out aliases bev."""
    pts, cnt = dev(wl["points"]), dev(wl["num_points"])
    img = dev(wl["img_feat"])
    if channels_last:
        img = img.contiguous(memory_format=torch.channels_last)
    cfg = wl["config"]
    frames = dcf.prepare_frames(pts, cnt, img, config=cfg, calib=None if use_uv else wl["calib"],
                                uv=dev(wl["uv"]) if use_uv else None)
    outs, knns = [], []
    for sc in wl["scales"]:
        layer = dcf.ContinuousFusion(img.shape[1], sc["C"], k=wl["k"], radius=wl["radius"], geom=sc["geom"], mode=mode).cuda()
        w1, b1, w2, b2, w3, b3 = sc["weights"]
        with torch.no_grad():
            layer.fc1.weight.copy_(dev(w1)); layer.fc1.bias.copy_(dev(b1))
            layer.fc2.weight.copy_(dev(w2)); layer.fc2.bias.copy_(dev(b2))
            layer.fc3.weight.copy_(dev(w3)); layer.fc3.bias.copy_(dev(b3))
            bev = dev(sc["bev"])
            out, knn = layer(bev, frames=frames, return_knn=True, out=bev if inplace else None)
            assert (out.data_ptr() == bev.data_ptr()) == bool(inplace)
        outs.append(out.cpu().numpy())
        knns.append(knn.cpu().numpy())
    torch.cuda.synchronize()
    return outs, knns


def oracle_fusion(oracle, wl, use_uv=False):
    outs, knns = [], []
    B = wl["points"].shape[0]
    for sc in wl["scales"]:
        o = np.empty_like(sc["bev"])
        kk = []
        for b in range(B):
            n = int(wl["num_points"][b])
            res, knn = oracle.fusion_forward(sc["bev"][b], wl["img_feat"][b], wl["points"][b], n, sc["geom"],
                                             wl["radius"], wl["k"], sc["weights"],
                                             calib=None if use_uv else wl["calib"],
                                             uv=wl["uv"][b] if use_uv else None, return_knn=True)
            o[b] = res
            kk.append(knn)
        outs.append(o)
        knns.append(np.stack(kk))
    return outs, knns
