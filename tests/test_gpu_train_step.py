"""BASELINE.json configs[3] on the GPU: a full train step of the drop-in model (LiDAR backbone + camera trunk + continuous
fusion at every residual group) closed with the drop-in LossTotal (device-side target assignment), forward + backward + Adam."""
import os
import sys

import numpy as np
import pytest
import torch

import dcf_b200 as dcf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_step_runs_and_every_part_gets_gradients():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from train_step_bench import synthetic_labels
    dev = torch.device("cuda")
    torch.manual_seed(0)
    cfg = dcf.geometry.carla_config(fusion_scales=(1, 2, 3, 4, 5), fusion_k=3)
    model = dcf.ObjectDetection_DCF(cfg).to(dev).eval()          # BatchNorm in eval mode, as in the reference (test.py:37)
    crit = dcf.LossTotal(cfg, batch_reduction="sum", generator=torch.Generator(device=dev).manual_seed(1)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.999))
    B = 2
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("yaml"), batch=B), seed=7)
    to = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    x_lidar = torch.rand(B, 32, 384, 256, device=dev)
    x_image = torch.randint(0, 255, (B, 3, 480, 640), device=dev, dtype=torch.uint8)
    ref, num = synthetic_labels(B, 11, dev)
    before = {n: p.detach().clone() for n, p in model.named_parameters() if n.startswith("fusion.") or n.startswith("image_backbone.stem")}
    assert before, "the drop-in model exposes fusion.* and image_backbone.* parameters"
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        pred = model(x_lidar, x_image, pointcloud_raw=to(wl["points"]), num_points_raw=to(wl["num_points"]), projected_loc_uv=to(wl["uv"]))
        pred_cls, pred_reg, _ = torch.split(pred, [4, 14, 14], dim=1)     # train.py:32
        loss = crit(ref, num, pred_cls, pred_reg).sum()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)), losses
    got_grad = {n: p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().sum()) > 0
                for n, p in model.named_parameters() if n in before}
    assert all(got_grad.values()), [n for n, ok in got_grad.items() if not ok]
    moved = [n for n, p in model.named_parameters() if n in before and not torch.equal(p.detach(), before[n])]
    assert len(moved) == len(before), "Adam moved every fusion / camera-stem parameter"
    # the reference's reduction (last frame only, loss.py:71) is the default and returns a (1,) tensor
    v = dcf.LossTotal(cfg).to(dev)(ref, num, pred_cls.detach(), pred_reg.detach())
    assert tuple(v.shape) == (1,) and bool(torch.isfinite(v).all())
