"""GPU: cf_voxelize_project (dataset-side voxelisation + projection) against the reference's own tensors
(golden fixtures) and against the numpy restatement on random sweeps.  Voxel grid, points and counts: bit-exact;
uv: bit-exact vs the restatement (same summation order), 2e-3 px vs the reference (BLAS order)."""
import numpy as np
import pytest
import torch

from _util import dev

pytestmark = pytest.mark.gpu


def _run(dcf, raws, cfg, crt, nraw=None):
    B = len(raws)
    nraw = nraw or max(r.shape[0] for r in raws)
    buf = np.full((B, nraw, 3), 7.5, dtype=np.float32)            # padding rows would pass the range filter if read
    for b, r in enumerate(raws):
        buf[b, :r.shape[0]] = r
    cnt = torch.tensor([r.shape[0] for r in raws], dtype=torch.int64)
    out = dcf.ops.voxelize_project(dev(buf), cnt.cuda(), cfg, crt)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def test_voxelize_matches_reference_fixtures(dcf, golden):
    cfg = dcf.geometry.carla_config()
    gs = [golden("voxelize_a.npz"), golden("voxelize_b.npz")]
    vox, pts, uv, num = _run(dcf, [g["raw"] for g in gs], cfg, gs[0]["crt"])
    for b, g in enumerate(gs):
        ref = np.zeros(int(np.prod(g["vox_shape"])), np.float32)
        ref[g["vox_idx"]] = g["vox_val"]
        n = int(g["num_points_raw"])
        assert int(num[b]) == n
        assert np.array_equal(vox[b].ravel(), ref)
        assert np.array_equal(pts[b, :n + 8], g["pointcloud_raw"])
        assert np.abs(uv[b, :n + 8] - g["projected_loc_uv"]).max() < 2e-3
        assert not pts[b, n:].any() and not uv[b, n:].any()


@pytest.mark.parametrize("seed", [21, 22])
def test_voxelize_matches_restatement_random(dcf, oracle, seed):
    """Dense random sweeps (many points per voxel: the last-write-wins rule matters), ragged batch, an empty frame."""
    rng = np.random.default_rng(seed)
    cfg = dcf.geometry.carla_config()
    crt = dcf.geometry.calibration_crt()
    raws = [np.stack([rng.uniform(-5, 75, n), rng.uniform(-35, 35, n), rng.uniform(-3, 1.2, n)], 1).astype(np.float32)
            for n in (30000, 0, 4097)]
    raws[2][::7, 1] = 0.0     # exact zeros in y: the nonzero()/3 bookkeeping drops trailing points
    vox, pts, uv, num = _run(dcf, raws, cfg, crt)
    for b, r in enumerate(raws):
        rv, rp, ru, rn = oracle.voxelize_project(r, cfg, crt)
        assert int(num[b]) == rn
        assert np.array_equal(vox[b], rv)
        assert np.array_equal(pts[b], rp)
        assert np.array_equal(uv[b], ru)


def test_voxelize_output_feeds_the_fusion_inputs(dcf):
    """The device-side tensors have the layouts the model consumes: (B,32,384,256) voxels, zero-padded (B,20000,3)
    points whose uv agree with cf_point_gather's own projection to 1e-3 px."""
    cfg = dcf.geometry.carla_config()
    raw = dcf.synthetic.lidar_sweep(np.random.default_rng(23), 32, 500)
    crt = dcf.geometry.calibration_crt()
    vox, pts, uv, num = _run(dcf, [raw], cfg, crt)
    assert vox.shape == (1, 32, 384, 256) and pts.shape == (1, 20000, 3) and uv.shape == (1, 20000, 2)
    n = int(num[0])
    assert 1000 < n <= 20000 and abs(float(vox.sum()) - float((vox > 0).sum() > 0) * float(vox.sum())) < 1e-3
    p = pts[0, :n]
    q = np.concatenate([p, np.ones((n, 1), np.float32)], 1) @ crt
    assert np.abs(q[:, :2] / q[:, 2:3] - uv[0, :n]).max() < 1e-3
