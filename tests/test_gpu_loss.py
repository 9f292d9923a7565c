"""SURVEY 8(f-4) on the GPU: cf_loss_targets and the drop-in LossTotal against the reference-generated fixtures
(tests/golden/loss_total_rt*.npz, made by the unmodified loss.py with the same random draws) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import dcf_b200 as dcf
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(rt):
    fx = np.load(os.path.join(GOLD, f"loss_total_rt{rt}.npz"))
    cfg = dcf.geometry.carla_config(regress_type=rt)
    dev = torch.device("cuda")
    t = {k: torch.from_numpy(fx[k]).to(dev) for k in ("ref", "num", "pred_cls", "pred_reg", "keys", "cand")}
    return fx, cfg, t


@pytest.mark.parametrize("rt", [0, 1])
def test_targets_bit_exact_against_reference_lists(rt):
    fx, cfg, t = _load(rt)
    loss = dcf.LossTotal(cfg).cuda()
    pos, npos, neg, nneg, reg = loss.targets(t["ref"], t["num"], 96, 64, draws=(t["keys"], t["cand"]))
    np.testing.assert_array_equal(npos.cpu().numpy(), fx["npos"])
    np.testing.assert_array_equal(nneg.cpu().numpy(), fx["nneg"])
    np.testing.assert_array_equal(pos.cpu().numpy(), fx["pos"])           # integer work: bit-exact, in the reference's order
    np.testing.assert_array_equal(neg.cpu().numpy(), fx["neg"])
    ref_reg = O.loss_targets(fx["ref"], fx["num"], 96, 64, dcf.geometry.voxel_scales(cfg), 4, cfg["positive_range"], rt,
                             cfg["pos_sample_threshold"], cfg["neg_sample_threshold"], fx["keys"], fx["cand"])[4]
    np.testing.assert_array_equal(reg.cpu().numpy(), ref_reg)


@pytest.mark.parametrize("rt", [0, 1])
def test_loss_value_matches_the_reference(rt):
    fx, cfg, t = _load(rt)
    loss = dcf.LossTotal(cfg).cuda()
    per = loss.per_frame(t["ref"], t["num"], t["pred_cls"], t["pred_reg"], draws=(t["keys"], t["cand"]))
    np.testing.assert_allclose(per.cpu().numpy(), fx["losses"], rtol=1e-5)   # fp32 sums in a different order
    # the reference returns the LAST frame's loss only (loss.py:71), shape (1,)
    v = loss(t["ref"], t["num"], t["pred_cls"], t["pred_reg"], draws=(t["keys"], t["cand"]))
    assert tuple(v.shape) == (1,)
    np.testing.assert_allclose(float(v), fx["losses"][-1], rtol=1e-5)
    s = dcf.LossTotal(cfg, batch_reduction="sum").cuda()(t["ref"], t["num"], t["pred_cls"], t["pred_reg"], draws=(t["keys"], t["cand"]))
    np.testing.assert_allclose(float(s), fx["losses"].sum(), rtol=1e-5)


def test_loss_gradient_is_the_directional_derivative():
    fx, cfg, t = _load(0)
    loss = dcf.LossTotal(cfg, batch_reduction="sum").cuda()
    draws = (t["keys"], t["cand"])
    pc = t["pred_cls"].double().requires_grad_(True)
    pr = t["pred_reg"].double().requires_grad_(True)
    loss.double()
    v = loss(t["ref"], t["num"], pc, pr, draws=draws)
    v.backward()
    g = torch.Generator(device="cuda").manual_seed(3)
    for x, gx in ((pc, pc.grad), (pr, pr.grad)):
        d = torch.randn(x.shape, device="cuda", generator=g, dtype=torch.float64)
        eps = 1e-5
        with torch.no_grad():
            args = lambda a, b: (t["ref"], t["num"], a, b)
            if x is pc:
                up, dn = loss(*args(pc + eps * d, pr), draws=draws), loss(*args(pc - eps * d, pr), draws=draws)
            else:
                up, dn = loss(*args(pc, pr + eps * d), draws=draws), loss(*args(pc, pr - eps * d), draws=draws)
        num = float(up - dn) / (2 * eps)
        assert abs(num - float((gx * d).sum())) <= 1e-6 * max(1.0, abs(num))


def test_device_draws_and_edge_cases():
    cfg = dcf.geometry.carla_config()
    dev = torch.device("cuda")
    loss = dcf.LossTotal(cfg, generator=torch.Generator(device="cuda").manual_seed(7)).cuda()
    B, M = 3, 20
    ref = torch.zeros((B, M, 8), device=dev)
    ref[:, :, 3:6] = 1.0
    ref[0, 0, :2] = torch.tensor([35.0, 0.0])
    ref[1, 0, :2] = torch.tensor([500.0, 0.0])        # outside the map: no positives, no regression cells (loss.py:88-89)
    num = torch.tensor([1, 1, 0], device=dev)          # a frame without ground truth
    pred_cls = torch.randn(B, 4, 96, 64, device=dev, requires_grad=True)
    pred_reg = torch.randn(B, 14, 96, 64, device=dev, requires_grad=True)
    pos, npos, neg, nneg, reg = loss.targets(ref, num, 96, 64)
    assert npos.tolist() == [25, 0, 0] and nneg.tolist() == [129, 129, 129]
    assert int((reg[1:] >= 0).sum()) == 0 and int((reg[0, 0] >= 0).sum()) == 25
    taken = set(pos[0, :25].tolist())
    assert not (taken & set(neg[0].tolist()))          # negatives never hit a positive cell
    per = loss.per_frame(ref, num, pred_cls, pred_reg)
    assert torch.isfinite(per).all()
    per.sum().backward()
    assert torch.isfinite(pred_cls.grad).all() and torch.isfinite(pred_reg.grad).all()
    # same generator seed -> same draws -> same targets (the draw is the only source of randomness)
    l2 = dcf.LossTotal(cfg, generator=torch.Generator(device="cuda").manual_seed(7)).cuda()
    p2 = l2.targets(ref, num, 96, 64)[0]
    assert torch.equal(p2, pos)


def test_loss_step_is_capturable():
    """No host synchronisation anywhere: targets + both terms + backward replay as one CUDA graph."""
    cfg = dcf.geometry.carla_config()
    fx, _, t = _load(0)
    loss = dcf.LossTotal(cfg, batch_reduction="sum").cuda()
    pc = t["pred_cls"].clone().requires_grad_(True)
    pr = t["pred_reg"].clone().requires_grad_(True)
    draws = (t["keys"], t["cand"])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            v = loss(t["ref"], t["num"], pc, pr, draws=draws)
            v.backward()
    torch.cuda.current_stream().wait_stream(s)
    pc.grad = None
    pr.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        v = loss(t["ref"], t["num"], pc, pr, draws=draws)
        v.backward()
    g.replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(float(v), fx["losses"].sum(), rtol=1e-5)
