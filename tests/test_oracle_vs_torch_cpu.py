"""CPU: the fusion oracle (parity unpinned -- the reference has no fusion layer) cross-checked against an
INDEPENDENT PyTorch restatement of SURVEY Appendix A built from library ops: stable sort for the KNN order,
F.grid_sample(align_corners=False, padding_mode='zeros') for the bilinear gather, nn.functional.linear for the MLP."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def torch_fusion(bev, img, pts, n, geom, radius, K, weights, calib, img_w=640.0, img_h=480.0):
    bev, img, pts = torch.from_numpy(bev), torch.from_numpy(img), torch.from_numpy(pts)
    C, H, W = bev.shape
    x0, y0, dx, dy = [torch.tensor(np.float32(g)) for g in geom]
    cx = x0 + torch.arange(H, dtype=torch.float32) * dx
    cy = y0 + torch.arange(W, dtype=torch.float32) * dy
    p = pts[:n]
    ddx = p[None, None, :, 0] - cx[:, None, None]
    ddy = p[None, None, :, 1] - cy[None, :, None]
    d2 = ddx * ddx + ddy * ddy                                             # (H,W,n), separately rounded fp32
    r2 = torch.tensor(np.float32(radius) * np.float32(radius))
    d2m = torch.where(d2 <= r2, d2, torch.tensor(float("inf")))
    order = torch.sort(d2m, dim=-1, stable=True)                            # ascending (d2, idx)  (A5)
    idx = order.indices[..., :K].clone()
    idx[order.values[..., :K] == float("inf")] = -1
    if idx.shape[-1] < K:
        idx = torch.cat([idx, torch.full((H, W, K - idx.shape[-1]), -1, dtype=idx.dtype)], dim=-1)
    q = torch.cat([p, torch.ones(n, 1)], dim=1) @ torch.from_numpy(calib)  # (A6)
    u, v = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
    grid = torch.stack([2 * (u + 0.5) / img_w - 1, 2 * (v + 0.5) / img_h - 1], dim=-1).view(1, 1, n, 2)
    feat = F.grid_sample(img[None], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0, :, 0].T  # (n,Ci)
    w1, b1, w2, b2, w3, b3 = [torch.from_numpy(w) for w in weights]
    out = bev.clone()
    safe = idx.clamp(min=0)
    for k in range(K):
        j = safe[..., k]
        off = torch.stack([p[j, 0] - cx[:, None], p[j, 1] - cy[None, :], p[j, 2]], dim=-1)        # (A8)
        x = torch.cat([feat[j], off], dim=-1)
        y = F.linear(F.relu(F.linear(F.relu(F.linear(x, w1, b1)), w2, b2)), w3, b3)                 # (A9)
        out += (y * (idx[..., k] >= 0)[..., None]).permute(2, 0, 1)                                  # (A10)
    return out.numpy(), idx.numpy().astype(np.int32)


@pytest.mark.parametrize("seed,K,radius", [(1, 3, 2.0), (2, 5, 0.7), (3, 1, 3.0)])
def test_oracle_matches_torch_restatement(oracle, dcf, seed, K, radius):
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("tiny"), batch=1, scales=(1, 2), max_num_pc=512,
                                          n_az=60), seed=seed, c_img=16, img_hw=(24, 32))
    n = int(wl["num_points"][0])
    assert 50 < n <= 512
    for sc in wl["scales"]:
        out_o, knn_o = oracle.fusion_forward(sc["bev"][0], wl["img_feat"][0], wl["points"][0], n, sc["geom"], radius, K,
                                             sc["weights"], calib=wl["calib"], return_knn=True)
        out_t, knn_t = torch_fusion(sc["bev"][0], wl["img_feat"][0], wl["points"][0], n, sc["geom"], radius, K,
                                    sc["weights"], wl["calib"])
        assert np.array_equal(knn_o, knn_t)
        assert (knn_o >= 0).mean() > 0.05
        err = np.abs(out_o - out_t).max() / np.abs(out_t).max()
        assert err < 2e-6, err


def test_gather_matches_grid_sample_including_borders(oracle):
    rng = np.random.default_rng(0)
    img = rng.standard_normal((8, 12, 16), dtype=np.float32)
    # uv spanning well outside the 640x480 image on every side, plus exact pixel centres and corners
    u = np.concatenate([rng.uniform(-80, 720, 400), [0.0, 639.0, 640.0, -0.5, 19.5, 20.0]]).astype(np.float32)
    v = np.concatenate([rng.uniform(-80, 560, 400), [0.0, 479.0, 480.0, -0.5, 19.5, 20.0]]).astype(np.float32)
    uv = np.stack([u, v], axis=1)
    got = oracle.gather_points(img, uv)
    grid = torch.from_numpy(np.stack([2 * (u + 0.5) / 640 - 1, 2 * (v + 0.5) / 480 - 1], axis=-1)).view(1, 1, -1, 2)
    ref = F.grid_sample(torch.from_numpy(img)[None], grid, mode="bilinear", padding_mode="zeros",
                        align_corners=False)[0, :, 0].T.numpy()
    assert np.abs(got - ref).max() < 2e-5


def test_knn_tie_break_and_padding_rows(oracle):
    """Equal distances resolve to the smaller index; rows >= n_valid are never candidates even though the zero
    padding row (0,0,0) would be the nearest point (A1, A5)."""
    pts = np.zeros((8, 3), np.float32)
    pts[:4, :2] = [[1, 0], [-1, 0], [0, 1], [0, -1]]
    knn = oracle.knn_bruteforce(pts, 4, 1, 1, 0.0, 0.0, 1.0, 1.0, np.float32(4.0), 3)
    assert knn.reshape(-1).tolist() == [0, 1, 2]
    knn = oracle.knn_bruteforce(pts, 8, 1, 1, 0.0, 0.0, 1.0, 1.0, np.float32(4.0), 3)
    assert knn.reshape(-1).tolist() == [4, 5, 6]
    knn = oracle.knn_bruteforce(pts, 2, 1, 1, 0.0, 0.0, 1.0, 1.0, np.float32(4.0), 3)
    assert knn.reshape(-1).tolist() == [0, 1, -1]
    knn = oracle.knn_bruteforce(pts, 4, 1, 1, 0.0, 0.0, 1.0, 1.0, np.float32(0.5), 3)
    assert knn.reshape(-1).tolist() == [-1, -1, -1]
