"""GPU: K-1/K-2 (bucketing + grid-hashed bounded-radius KNN) through the C ABI, BIT-EXACT against the
brute-force oracle (Appendix A1-A5)."""
import numpy as np
import pytest
import torch

from _util import cuda_knn, dev, oracle_knn

pytestmark = pytest.mark.gpu


def _check_bucketing(wl, start, srt, grid_dims):
    """bucket_start is a valid CSR over the sorted points and sorted holds a permutation of the valid rows."""
    B, N, _ = wl["points"].shape
    nbx, nby = grid_dims
    for b in range(B):
        n = int(wl["num_points"][b])
        s = start[b]
        assert s[0] == 0 and s[-1] == n and np.all(np.diff(s) >= 0)
        idx = srt[b, :n, 3].view(np.int32)
        assert np.array_equal(np.sort(idx), np.arange(n, dtype=np.int32))
        assert np.array_equal(srt[b, :n, :3], wl["points"][b][idx])


@pytest.mark.parametrize("name,seed", [("tiny", 1), ("yaml", 2)])
def test_knn_matches_oracle_all_scales(dcf, oracle, name, seed):
    wl = dcf.synthetic.make_workload(name, seed=seed)
    for sc in wl["scales"]:
        got, (start, srt) = cuda_knn(dcf, wl, sc)
        _, _, _, nbx, nby = dcf.geometry.bucket_grid(wl["config"])
        _check_bucketing(wl, start, srt, (nbx, nby))
        ref = oracle_knn(oracle, wl, sc)
        assert got.dtype == np.int32 and got.shape == ref.shape
        assert np.array_equal(got, ref), f"scale {sc['group']}: {(got != ref).sum()} index mismatches"


@pytest.mark.parametrize("K", [1, 2, 4, 7, 10, 16])
def test_knn_every_k(dcf, oracle, K):
    wl = dcf.synthetic.make_workload("tiny", seed=3)
    sc = wl["scales"][0]
    got, _ = cuda_knn(dcf, wl, sc, k=K)
    assert np.array_equal(got, oracle_knn(oracle, wl, sc, k=K))


@pytest.mark.parametrize("K", [None, 9])   # the workload's K (row-major gather) and K >= 8 (centre-out gather of the patch kernel)
@pytest.mark.parametrize("radius,cell", [(0.3, 0.5), (2.0, 0.25), (2.0, 1.7), (5.0, 0.5), (40.0, 2.0)])
def test_knn_radius_and_bucket_pitch(dcf, oracle, radius, cell, K):
    """Result must not depend on the bucket pitch; radius from 'mostly empty' to 'everything is a candidate'."""
    wl = dcf.synthetic.make_workload("tiny", seed=4)
    kw = {} if K is None else {"k": K}
    for sc in wl["scales"][:2]:   # scale 0: patch kernel, scale 1 onwards: whichever the patch-size rule picks
        got, _ = cuda_knn(dcf, wl, sc, cell=cell, radius=radius, **kw)
        assert np.array_equal(got, oracle_knn(oracle, wl, sc, radius=radius, **kw))


def test_knn_ties_duplicates_and_lattice(dcf, oracle):
    """Adversarial: points on the cell-centre lattice (exact d2 ties in every direction), exact duplicates,
    points on bucket boundaries, points outside the bucket grid (clamped), and a point at the origin."""
    wl = dcf.synthetic.make_workload("tiny", seed=5)
    sc = wl["scales"][0]
    x0, y0, dx, dy = [float(v) for v in sc["geom"]]
    rng = np.random.default_rng(5)
    ii, jj = np.meshgrid(np.arange(0, sc["H"], 4), np.arange(0, sc["W"], 4), indexing="ij")
    lattice = np.stack([np.float32(x0) + ii.ravel().astype(np.float32) * np.float32(dx),
                        np.float32(y0) + jj.ravel().astype(np.float32) * np.float32(dy),
                        np.zeros(ii.size, np.float32)], axis=1)
    half = lattice + np.array([dx / 2, dy / 2, 0], np.float32)             # equidistant from 4 centres
    edges = np.stack([np.arange(0, 8, 0.5, dtype=np.float32), np.full(16, -30.0, np.float32) + np.arange(16) * 0.5,
                      np.zeros(16, np.float32)], axis=1)                     # on bucket boundaries
    outside = np.array([[-3.0, -31.0, 0], [80.0, 40.0, 0], [0.0, 0.0, 0.0], [-0.25, -29.9, 0]], np.float32)
    pts = np.concatenate([lattice, half, lattice[::2], edges, outside, lattice[::5]])  # duplicates included
    pts = pts[rng.permutation(pts.shape[0])]
    N = wl["points"].shape[1]
    assert pts.shape[0] <= N
    wl["points"][:] = 0
    wl["points"][0, :pts.shape[0]] = pts
    wl["points"][1, :pts.shape[0]] = pts[::-1]
    wl["num_points"][:] = pts.shape[0]
    for K in (3, 8):
        got, _ = cuda_knn(dcf, wl, sc, k=K, radius=1.0)
        assert np.array_equal(got, oracle_knn(oracle, wl, sc, k=K, radius=1.0))


def test_knn_empty_and_ragged_frames(dcf, oracle):
    """num_points = 0, 1, K-1 and N (zero padding rows must never be candidates: A1)."""
    wl = dcf.synthetic.make_workload(dict(dcf.synthetic.workload("tiny"), batch=4), seed=6)
    N = wl["points"].shape[1]
    full = wl["points"][0].copy()
    n_full = int(wl["num_points"][0])
    reps = int(np.ceil(N / max(n_full, 1)))
    wl["points"][3] = np.tile(full[:n_full], (reps, 1))[:N] + np.float32(0.001) * np.arange(N, dtype=np.float32)[:, None]
    wl["num_points"][:] = [0, 1, 2, N]
    sc = wl["scales"][0]
    for K in (3, 12):
        got, (start, _) = cuda_knn(dcf, wl, sc, k=K)
        ref = oracle_knn(oracle, wl, sc, k=K)
        assert np.array_equal(got, ref)
        assert (got[0] == -1).all() and start[0].max() == 0
        assert (got[1] <= 0).all() and (got[2] <= 1).all()


def test_knn_rejects_bad_arguments(dcf):
    wl = dcf.synthetic.make_workload("tiny", seed=7)
    grid = dcf.ops.BucketGrid(*dcf.geometry.bucket_grid(wl["config"]))
    pts, cnt = dev(wl["points"]), dev(wl["num_points"])
    start, srt, _ = dcf.ops.bucket_points(pts, cnt, grid)
    sc = wl["scales"][0]
    with pytest.raises(RuntimeError, match="K=17"):
        dcf.ops.knn_query(start, srt, grid, sc["H"], sc["W"], sc["geom"], 2.0, 17)
    with pytest.raises(RuntimeError, match="radius"):
        dcf.ops.knn_query(start, srt, grid, sc["H"], sc["W"], sc["geom"], -1.0, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dcf.ops.bucket_points(pts.cpu(), cnt, grid)
    with pytest.raises(TypeError):
        dcf.ops.bucket_points(pts.double(), cnt, grid)


def test_knn_full_size_properties(dcf, oracle):
    """BASELINE config 0 size (700x800 cells, ~20k points, K=3): size-independent properties on every cell
    (in-radius, ascending (d2, idx), -1 only as a suffix, no duplicates) + bit-exact oracle check on a
    seeded sample of 4096 rows of cells."""
    wl = dcf.synthetic.make_workload("cfg0", seed=8)
    sc = wl["scales"][0]
    got, _ = cuda_knn(dcf, wl, sc)
    H, W, K = sc["H"], sc["W"], wl["k"]
    x0, y0, dx, dy = sc["geom"]
    pts = wl["points"][0]
    ii, jj = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    cx, cy = x0 + ii * dx, y0 + jj * dy
    idx = got[0]
    valid = idx >= 0
    assert np.all(valid[..., :-1] >= valid[..., 1:])            # -1 only as a suffix
    safe = np.where(valid, idx, 0)
    ddx = pts[safe, 0] - cx[..., None]
    ddy = pts[safe, 1] - cy[..., None]
    d2 = ddx * ddx + ddy * ddy
    r2 = np.float32(wl["radius"]) ** 2
    assert np.all(d2[valid] <= r2)
    both = valid[..., :-1] & valid[..., 1:]
    asc = (d2[..., :-1] < d2[..., 1:]) | ((d2[..., :-1] == d2[..., 1:]) & (idx[..., :-1] < idx[..., 1:]))
    assert np.all(asc[both])
    assert valid.mean() > 0.2
    rows = np.sort(np.random.default_rng(8).choice(H, 6, replace=False))
    for r in rows:
        ref = oracle.knn_bruteforce(pts, int(wl["num_points"][0]), H, W, x0, y0, dx, dy, r2, K,
                                    cell_range=(int(r) * W, (int(r) + 1) * W))
        assert np.array_equal(idx[r], ref)
