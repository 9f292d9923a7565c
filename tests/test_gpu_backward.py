"""GPU: K-4b.  Gradients of the fused layer (w.r.t. the BEV map, the camera feature map and all six MLP
parameters) against PyTorch autograd on a float64 restatement of the same layer (grid_sample gather, gathered rows,
three Linear layers, masked sum-pool).  Mode "simt": every GEMM of the backward on fp32 CUDA cores, tolerance 2e-3 relative L2
(2e-2 max norm) per gradient.  Mode "fp32": the GEMMs on tcgen05 with split-bf16 operands (~1.5e-5 relative per product; the
GEMM kernels themselves are held to 1e-4 in test_gpu_bwd_gemm.py), tolerance 5e-3 / 5e-2.  The tolerances are this wide
because of the ReLU kinks, not of the arithmetic: a gradient is either within ~1e-6 (simt) / ~1e-5 (fp32) of the float64
reference, or a pre-activation that is zero to rounding switches side and ONE full-size term of the sum changes (~1e-3
of the tensor's norm at these sizes); how often that happens is proportional to the rounding error of the recomputed
pre-activations, so the tensor-core mode sees a few more of them."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _util import dev

pytestmark = pytest.mark.gpu


def torch_layer(bev, img, pts, n_valid, knn, geom, calib, w, size=(640.0, 480.0)):
    """float64 reference of one scale, differentiable w.r.t. bev, img and the weights."""
    B, C, H, W = bev.shape
    x0, y0, dx, dy = [float(g) for g in geom]
    cx = x0 + torch.arange(H, device=bev.device, dtype=torch.float64) * dx
    cy = y0 + torch.arange(W, device=bev.device, dtype=torch.float64) * dy
    w1, b1, w2, b2, w3, b3 = w
    out = []
    for b in range(B):
        n = int(n_valid[b])
        p = pts[b, :n]
        q = torch.cat([p, torch.ones(n, 1, device=p.device, dtype=p.dtype)], 1) @ calib
        u, v = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
        grid = torch.stack([2 * (u + 0.5) / size[0] - 1, 2 * (v + 0.5) / size[1] - 1], -1).view(1, 1, n, 2)
        feat = F.grid_sample(img[b:b + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0, :, 0].T
        idx = knn[b].long()                                  # (H,W,K)
        valid = (idx >= 0)
        j = idx.clamp(min=0)
        off = torch.stack([p[j, 0] - cx[:, None, None], p[j, 1] - cy[None, :, None], p[j, 2]], -1)
        x = torch.cat([feat[j], off], -1)
        h = F.relu(F.linear(F.relu(F.linear(x, w1, b1)), w2, b2))
        pooled = (h * valid[..., None]).sum(2)
        y = F.linear(pooled, w3) + valid.sum(2, keepdim=True) * b3
        out.append(bev[b] + y.permute(2, 0, 1))
    return torch.stack(out)


@pytest.mark.parametrize("mode", ["fp32", "simt"])
def test_backward_matches_autograd(dcf, mode):
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("tiny"), scales=(1, 2)), seed=21, c_img=32, img_hw=(24, 32))
    cfg = wl["config"]
    img = dev(wl["img_feat"]).requires_grad_(True)
    pts, cnt = dev(wl["points"]), dev(wl["num_points"])
    frames = dcf.FrameContext(pts, cnt, dcf.ops.BucketGrid(*dcf.geometry.bucket_grid(cfg)))
    frames.gather(img, calib=wl["calib"])
    layers, bevs, outs = [], [], []
    for sc in wl["scales"]:
        layer = dcf.ContinuousFusion(32, sc["C"], k=wl["k"], radius=wl["radius"], geom=sc["geom"], mode=mode).cuda()
        with torch.no_grad():
            for p_, w_ in zip((layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias, layer.fc3.weight,
                               layer.fc3.bias), sc["weights"]):
                p_.copy_(dev(w_))
        bev = dev(sc["bev"]).requires_grad_(True)
        out, knn = layer(bev, frames=frames, return_knn=True)
        layers.append(layer); bevs.append(bev); outs.append((out, knn))
    gen = torch.Generator(device="cuda").manual_seed(5)
    seeds = [torch.randn(o.shape, device="cuda", generator=gen) for o, _ in outs]
    loss = sum((o * s).sum() for (o, _), s in zip(outs, seeds))
    loss.backward()

    # float64 autograd reference on the same indices
    img64 = dev(wl["img_feat"]).double().requires_grad_(True)
    calib64 = dev(wl["calib"]).double()
    ref_loss, ref_params, ref_bevs = 0.0, [], []
    for sc, layer, (o, knn), s in zip(wl["scales"], layers, outs, seeds):
        w64 = [p_.detach().double().requires_grad_(True) for p_ in (layer.fc1.weight, layer.fc1.bias, layer.fc2.weight,
                                                                    layer.fc2.bias, layer.fc3.weight, layer.fc3.bias)]
        b64 = dev(sc["bev"]).double().requires_grad_(True)
        ro = torch_layer(b64, img64, pts.double(), wl["num_points"], knn, sc["geom"], calib64, w64)
        assert (ro.float() - o.detach()).abs().max() / ro.abs().max() < 1e-4
        ref_loss = ref_loss + (ro * s.double()).sum()
        ref_params.append(w64); ref_bevs.append(b64)
    ref_loss.backward()

    report, bad = [], []

    base_tol = 2e-3 if mode == "simt" else 5e-3

    def close(a, b, name, tol=None):
        tol = base_tol if tol is None else tol
        # relative L2 error, plus a looser max-norm bound: a ReLU whose pre-activation is within fp32 rounding of zero
        # can switch between the fp32 kernels and the fp64 reference, which perturbs a few entries, not the bulk
        l2 = (a.double() - b).norm().item() / max(b.norm().item(), 1e-12)
        mx = (a.double() - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
        report.append(f"{name}: rel L2 err {l2:.3e}, rel max err {mx:.3e}")
        if not (l2 < tol and mx < 10 * tol):
            bad.append(report[-1])

    close(img.grad, img64.grad, "d img_feat")
    for g, (layer, w64, bev, b64) in enumerate(zip(layers, ref_params, bevs, ref_bevs)):
        close(bev.grad, b64.grad, f"scale {g} d bev", 1e-6)
        ci = layer.c_img
        close(layer.fc1.weight.grad[:, :ci], w64[0].grad[:, :ci], f"scale {g} d W1[:, image]")
        close(layer.fc1.weight.grad[:, ci:], w64[0].grad[:, ci:], f"scale {g} d W1[:, offset]")
        for name, p_, r_ in zip(["b1", "W2", "b2", "W3", "b3"], (layer.fc1.bias, layer.fc2.weight, layer.fc2.bias,
                                                                  layer.fc3.weight, layer.fc3.bias), w64[1:]):
            close(p_.grad, r_.grad, f"scale {g} d {name}")
    print("\n".join(report))
    assert not bad, "\n".join(["gradients outside the tolerance:"] + bad + ["all:"] + report)


def _gate_kinks(knn, pre_min, eps):
    """Drop the (cell, k) slots whose smallest |pre-activation| is below eps and close the gaps (valid slots stay a
    prefix, order kept): the ReLU of such a slot may fall on either side in fp32, which changes one full-size term of
    every gradient sum; without them 1e-4 can be demanded of the rest."""
    valid = knn >= 0
    keep = valid & (pre_min >= eps)
    order = torch.argsort((~keep).to(torch.int8), dim=-1, stable=True)
    out = torch.where(torch.gather(keep, -1, order), torch.gather(knn, -1, order), torch.full_like(knn, -1))
    return out.contiguous(), int(valid.sum()), int(keep.sum())


def _pre_min(bev_shape, img, pts, n_valid, knn, geom, calib, w, size=(640.0, 480.0)):
    """min over channels of |z1|, |z2| (the two pre-activations) per (b, i, j, k), float64; inf where the slot is empty."""
    B, C, H, W = bev_shape
    x0, y0, dx, dy = [float(g) for g in geom]
    cx = x0 + torch.arange(H, device=img.device, dtype=torch.float64) * dx
    cy = y0 + torch.arange(W, device=img.device, dtype=torch.float64) * dy
    w1, b1, w2, b2 = w[:4]
    out = []
    for b in range(B):
        n = int(n_valid[b])
        p = pts[b, :n]
        q = torch.cat([p, torch.ones(n, 1, device=p.device, dtype=p.dtype)], 1) @ calib
        u, v = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
        grid = torch.stack([2 * (u + 0.5) / size[0] - 1, 2 * (v + 0.5) / size[1] - 1], -1).view(1, 1, n, 2)
        feat = F.grid_sample(img[b:b + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0, :, 0].T
        idx = knn[b].long()
        j = idx.clamp(min=0)
        off = torch.stack([p[j, 0] - cx[:, None, None], p[j, 1] - cy[None, :, None], p[j, 2]], -1)
        z1 = F.linear(torch.cat([feat[j], off], -1), w1, b1)
        z2 = F.linear(F.relu(z1), w2, b2)
        m = torch.minimum(z1.abs().amin(-1), z2.abs().amin(-1))
        out.append(torch.where(idx >= 0, m, torch.full_like(m, float("inf"))))
    return torch.stack(out)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("channels", [(128, 192), (256, 96), (32, 64)])
def test_backward_matches_autograd_wide_layers(dcf, mode, channels):
    """K-4b through cf_fusion_bwd at every width the backbone uses (C = 32 ... 256), against float64 autograd.
    fp32 mode: slots with a pre-activation within EPS of zero are removed from the neighbour table on both sides
    (_gate_kinks), every gradient must then agree to 1e-4 relative L2.  bf16 mode: the backward reuses the forward's
    per-point table, which is bf16-accurate there, so the recomputed activation pattern differs from float64 on ~1e-2
    of the units by design and no gating is possible; gradients within 6e-2 relative L2 (measured 2e-2 ... 3.4e-2; A12's 1e-2 is a forward bound)."""
    EPS = 2e-4
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("tiny"), scales=(1, 2)), seed=24, c_img=32, img_hw=(24, 32),
                                     channels=list(channels) + [32, 32, 32])
    cfg = wl["config"]
    img = dev(wl["img_feat"]).requires_grad_(True)
    img64 = dev(wl["img_feat"]).double().requires_grad_(True)
    calib64 = dev(wl["calib"]).double()
    pts, cnt = dev(wl["points"]), dev(wl["num_points"])
    frames = dcf.FrameContext(pts, cnt, dcf.ops.BucketGrid(*dcf.geometry.bucket_grid(cfg)))
    frames.gather(img, calib=wl["calib"])
    gen = torch.Generator(device="cuda").manual_seed(6)
    loss, ref_loss, sets = 0.0, 0.0, []
    for sc in wl["scales"]:
        assert sc["C"] in channels
        layer = dcf.ContinuousFusion(32, sc["C"], k=wl["k"], radius=wl["radius"], geom=sc["geom"], mode=mode).cuda()
        with torch.no_grad():
            for p_, w_ in zip((layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias, layer.fc3.weight,
                               layer.fc3.bias), sc["weights"]):
                p_.copy_(dev(w_))
        w64 = [p_.detach().double().requires_grad_(True) for p_ in (layer.fc1.weight, layer.fc1.bias, layer.fc2.weight,
                                                                    layer.fc2.bias, layer.fc3.weight, layer.fc3.bias)]
        H, W = sc["H"], sc["W"]
        knn = frames.knn(H, W, sc["geom"], wl["radius"], wl["k"])
        if mode == "fp32":
            with torch.no_grad():
                pm = _pre_min(sc["bev"].shape, img64, pts.double(), wl["num_points"], knn, sc["geom"], calib64, w64)
            knn, n_all, n_kept = _gate_kinks(knn, pm, EPS)
            assert n_kept > 0.5 * n_all, (n_all, n_kept)
            key = [k for k in frames._knn_cache if k[0] == H and k[1] == W][0]
            frames._knn_cache[key] = knn
        bev = dev(sc["bev"]).requires_grad_(True)
        b64 = dev(sc["bev"]).double().requires_grad_(True)
        out, knn_used = layer(bev, frames=frames, return_knn=True)
        assert torch.equal(knn_used, knn)
        seed = torch.randn(out.shape, device="cuda", generator=gen)
        loss = loss + (out * seed).sum()
        ro = torch_layer(b64, img64, pts.double(), wl["num_points"], knn, sc["geom"], calib64, w64)
        fwd_tol = 1e-4 if mode == "fp32" else 1e-2
        assert (ro.float() - out.detach()).abs().max() / ro.abs().max() < fwd_tol
        ref_loss = ref_loss + (ro * seed.double()).sum()
        sets.append((sc, layer, bev, b64, w64))
    loss.backward()
    ref_loss.backward()
    tol = 1e-4 if mode == "fp32" else 6e-2
    report, bad = [], []

    def close(a, b, name, t=tol):
        l2 = (a.double() - b).norm().item() / max(b.norm().item(), 1e-12)
        report.append(f"{name}: rel L2 err {l2:.3e}")
        if not l2 < t:
            bad.append(report[-1])

    close(img.grad, img64.grad, "d img_feat")
    for sc, layer, bev, b64, w64 in sets:
        g = f"C={sc['C']}"
        close(bev.grad, b64.grad, f"{g} d bev", 1e-6)
        ci = layer.c_img
        close(layer.fc1.weight.grad[:, :ci], w64[0].grad[:, :ci], f"{g} d W1[:, image]")
        close(layer.fc1.weight.grad[:, ci:], w64[0].grad[:, ci:], f"{g} d W1[:, offset]")
        for name, p_, r_ in zip(["b1", "W2", "b2", "W3", "b3"], (layer.fc1.bias, layer.fc2.weight, layer.fc2.bias,
                                                                  layer.fc3.weight, layer.fc3.bias), w64[1:]):
            close(p_.grad, r_.grad, f"{g} d {name}")
    print("\n".join(report))
    assert not bad, "\n".join(["gradients outside the tolerance:"] + bad + ["all:"] + report)


def test_backward_run_to_run_spread_is_rounding(dcf):
    """The reductions of K-4b use float atomics: two runs on the same inputs may differ in the last bits, never by more
    than accumulated rounding (1e-5 relative L2 per gradient)."""
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("tiny"), scales=(1, 2)), seed=25, c_img=32, img_hw=(24, 32),
                                     channels=[128, 64, 32, 32, 32])
    pts, cnt = dev(wl["points"]), dev(wl["num_points"])
    grads = []
    for _ in range(2):
        img = dev(wl["img_feat"]).requires_grad_(True)
        frames = dcf.FrameContext(pts, cnt, dcf.ops.BucketGrid(*dcf.geometry.bucket_grid(wl["config"])))
        frames.gather(img, calib=wl["calib"])
        loss, params = 0.0, []
        for sc in wl["scales"]:
            layer = dcf.ContinuousFusion(32, sc["C"], k=wl["k"], radius=wl["radius"], geom=sc["geom"], mode="fp32").cuda()
            with torch.no_grad():
                for p_, w_ in zip((layer.fc1.weight, layer.fc1.bias, layer.fc2.weight, layer.fc2.bias, layer.fc3.weight,
                                   layer.fc3.bias), sc["weights"]):
                    p_.copy_(dev(w_))
            out = layer(dev(sc["bev"]), frames=frames)
            loss = loss + (out * torch.linspace(-1, 1, out.numel(), device="cuda").view_as(out)).sum()
            params += list(layer.parameters())
        loss.backward()
        grads.append([img.grad.clone()] + [p_.grad.clone() for p_ in params])
    for a, b in zip(*grads):
        assert (a - b).norm().item() <= 1e-5 * max(b.norm().item(), 1e-12)


def test_training_step_through_the_dropin_model(dcf):
    """Gradient step of ObjectDetection_DCF with fusion on: every fusion / camera parameter receives a finite gradient,
    the fusion and camera weights get non-zero ones, and a small step along -grad lowers the loss (first-order check).
    BatchNorm runs in eval mode, as it does in the reference (Test.__init__ calls net.eval(), train.py:76, test.py:37)."""
    cfg = dcf.geometry.carla_config(fusion_scales=(1, 2, 3), fusion_k=3, max_num_pc=4096)
    torch.manual_seed(0)
    model = dcf.ObjectDetection_DCF(cfg).cuda().eval()
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("yaml"), batch=2, max_num_pc=4096, n_az=300), seed=22)
    x_lidar = torch.rand(2, 32, 384, 256, device="cuda")
    x_image = torch.randint(0, 255, (2, 3, 480, 640), device="cuda", dtype=torch.uint8)
    target = torch.randn(2, 32, 96, 64, device="cuda")
    extra = dict(pointcloud_raw=dev(wl["points"]), num_points_raw=torch.from_numpy(wl["num_points"]),
                 projected_loc_uv=dev(wl["uv"]))

    def loss_fn():
        pred = model(x_lidar, x_image, **extra)
        assert pred.shape == (2, 32, 96, 64)
        return F.mse_loss(pred[:, :18].double(), target[:, :18].double())

    loss0 = loss_fn()
    loss0.backward()
    gsq = 0.0
    for name, p_ in model.named_parameters():
        if name.startswith(("fusion.", "image_backbone.")):
            assert p_.grad is not None and torch.isfinite(p_.grad).all(), name
            gsq += float(p_.grad.double().pow(2).sum())
    assert model.fusion["group1"].fc2.weight.grad.abs().max() > 0
    assert model.fusion["group3"].fc1.weight.grad.abs().max() > 0
    assert model.image_backbone.out.weight.grad.abs().max() > 0
    # step only the fusion + camera parameters: the decrease must match  -lr * |g|^2  to first order
    lr = 1e-5 * loss0.item() / gsq   # tiny: a random-init deep net is strongly curved (the pure-PyTorch groups too)
    with torch.no_grad():
        for name, p_ in model.named_parameters():
            if name.startswith(("fusion.", "image_backbone.")):
                p_ -= lr * p_.grad
        loss1 = loss_fn()
    predicted = lr * gsq
    assert loss1.item() < loss0.item()
    assert abs((loss0.item() - loss1.item()) - predicted) < 0.3 * predicted


def test_dropin_model_inference_equals_training_forward(dcf):
    """Under torch.no_grad() the drop-in model fuses each group's map IN PLACE with tables / KNN precomputed on side
    streams (one layer-1 launch for all scales); with autograd on it runs the out-of-place path layer by layer.  Both
    forwards must give the same prediction tensor, bit for bit, and calling the model without the extra tensors must
    still behave as the LiDAR-only reference."""
    cfg = dcf.geometry.carla_config(fusion_scales=(1, 2, 3, 4, 5), fusion_k=3, max_num_pc=4096)
    torch.manual_seed(1)
    model = dcf.ObjectDetection_DCF(cfg).cuda().eval()
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("yaml"), batch=2, max_num_pc=4096, n_az=300), seed=23)
    x_lidar = torch.rand(2, 32, 384, 256, device="cuda")
    x_image = torch.randint(0, 255, (2, 3, 480, 640), device="cuda", dtype=torch.uint8)
    extra = dict(pointcloud_raw=dev(wl["points"]), num_points_raw=torch.from_numpy(wl["num_points"]))
    with torch.no_grad():
        a = model(x_lidar, x_image, **extra)
        a2 = model(x_lidar, x_image, **extra)
        lidar_only = model(x_lidar, x_image)
    b = model(x_lidar, x_image, **extra)
    torch.cuda.synchronize()
    assert b.requires_grad and not a.requires_grad
    assert torch.equal(a, a2)                       # deterministic, inputs untouched by the in-place fusion
    assert torch.equal(a, b.detach())
    assert (a - lidar_only).abs().max() > 1e-4      # the fusion actually contributes
