"""Seeded synthetic inputs with the reference's shapes and layouts (there is no dataset offline).

LiDAR frames are ring structured (elevation beams hitting a ground plane, plus car-sized clusters), pushed
through a restatement of the reference's range / image filter and zero padding
(data_import_carla.py:196-229, 261-267), so `pointcloud_raw`, `projected_loc_uv` and `num_points_raw`
have the reference-made layout.  tests/test_oracle_cpu.py checks that restatement against the reference's
own Voxelization_Projection when /root/reference is present.
"""
from __future__ import annotations

import math

import numpy as np

from . import geometry as G

SCALE_STRIDES = (1, 2, 4, 8, 16)  # residual groups 1..5 of ResnetCustomed (model.py:64-79)


def lidar_sweep(rng: np.random.Generator, n_beams: int, n_az: int, sensor_h: float = 2.2, n_cars: int = 10,
                pts_per_car: int = 160) -> np.ndarray:
    """Raw (M,3) float32 returns in the LiDAR frame (x forward, y left, z up), forward half-plane."""
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_beams))
    az = (rng.random((n_beams, n_az)) - 0.5) * math.pi
    el = elev[:, None] + rng.normal(0.0, 2e-4, (n_beams, n_az))
    down = el < -1e-3
    rng_m = np.where(down, sensor_h / np.tan(-np.where(down, el, -1.0)), np.inf)
    rng_m = rng_m * (1.0 + rng.normal(0.0, 2e-3, rng_m.shape))
    ok = np.isfinite(rng_m) & (rng_m < 90.0)
    x = rng_m * np.cos(az)
    y = rng_m * np.sin(az)
    z = -sensor_h + rng.normal(0.0, 0.01, rng_m.shape)
    ground = np.stack([x[ok], y[ok], z[ok]], axis=1)
    cars = []
    for _ in range(n_cars):
        c = np.array([rng.uniform(6.0, 60.0), rng.uniform(-20.0, 20.0), -sensor_h + 0.75])
        size = np.array([4.0, 2.0, 1.5]) * np.exp(0.1 * rng.normal(size=3))
        yaw = rng.uniform(-math.pi, math.pi)
        u = rng.random((pts_per_car, 3)) - 0.5
        face = rng.integers(0, 3, pts_per_car)
        sign = np.where(rng.random(pts_per_car) < 0.5, -0.5, 0.5)
        u[np.arange(pts_per_car), face] = sign  # snap to a face
        local = u * size
        rot = np.array([[math.cos(yaw), -math.sin(yaw), 0.0], [math.sin(yaw), math.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
        cars.append(local @ rot.T + c)
    pts = np.concatenate([ground] + cars, axis=0) if cars else ground
    return np.ascontiguousarray(pts.astype(np.float32))


def filter_and_project(points: np.ndarray, config: dict, crt: np.ndarray | None = None):
    """Range filter + camera projection + image filter, reference semantics.

    Range (data_import_carla.py:215-226): strict > min and < max - delta on x, y, z.
    Projection (:196-201): [x y z 1] @ CRT, divide by the third component.
    Image filter (:202-205): 0 < u < image_height and 0 < v < image_width (the reference's swapped bounds).
    Returns (points_kept (n,3) f32, uv (n,2) f32).
    """
    crt = G.calibration_crt() if crt is None else crt
    d = np.float32(config["delta"])
    p = points.astype(np.float32)
    keep = ((p[:, 0] > np.float32(config["lidar_x_min"])) & (p[:, 0] < np.float32(config["lidar_x_max"]) - d) &
            (p[:, 1] > np.float32(config["lidar_y_min"])) & (p[:, 1] < np.float32(config["lidar_y_max"]) - d) &
            (p[:, 2] > np.float32(config["lidar_z_min"])) & (p[:, 2] < np.float32(config["lidar_z_max"]) - d))
    p = p[keep]
    q = np.concatenate([p, np.ones((p.shape[0], 1), np.float32)], axis=1) @ crt.astype(np.float32)
    uv = (q[:, :2] / q[:, 2:3]).astype(np.float32)
    inside = (uv[:, 0] > 0) & (uv[:, 0] < config["image_height"]) & (uv[:, 1] > 0) & (uv[:, 1] < config["image_width"])
    return np.ascontiguousarray(p[inside]), np.ascontiguousarray(uv[inside])


def make_points(seed: int, config: dict, n_beams: int = 32, n_az: int = 1900, batch: int = 1, uniform: bool = False):
    """Zero-padded (B,N,3) points, (B,N,2) uv, (B,) int64 counts, N = config["max_num_pc"]."""
    N = int(config["max_num_pc"])
    pts = np.zeros((batch, N, 3), np.float32)
    uvs = np.zeros((batch, N, 2), np.float32)
    cnt = np.zeros((batch,), np.int64)
    for b in range(batch):
        rng = np.random.default_rng(seed * 1000 + b)
        if uniform:
            m = int(N * 2.6)
            raw = np.stack([rng.uniform(config["lidar_x_min"], config["lidar_x_max"], m),
                            rng.uniform(config["lidar_y_min"], config["lidar_y_max"], m),
                            rng.uniform(config["lidar_z_min"], config["lidar_z_max"], m)], axis=1).astype(np.float32)
        else:
            raw = lidar_sweep(rng, n_beams, n_az)
        p, uv = filter_and_project(raw, config)
        if p.shape[0] > N:  # the reference raises on overflow (:263-264); the generator subsamples instead
            sel = np.sort(rng.choice(p.shape[0], N, replace=False))
            p, uv = p[sel], uv[sel]
        n = p.shape[0]
        pts[b, :n], uvs[b, :n], cnt[b] = p, uv, n
    return pts, uvs, cnt


def linear_init(rng: np.random.Generator, out_f: int, in_f: int):
    """nn.Linear's default init: weight, bias ~ U(-1/sqrt(in), 1/sqrt(in))."""
    bound = 1.0 / math.sqrt(in_f)
    w = rng.uniform(-bound, bound, (out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, (out_f,)).astype(np.float32)
    return w, b


def mlp_weights(seed: int, c_img: int, c: int):
    rng = np.random.default_rng(77000 + seed * 131 + c)
    w1, b1 = linear_init(rng, c, c_img + 3)
    w2, b2 = linear_init(rng, c, c)
    w3, b3 = linear_init(rng, c, c)
    return w1, b1, w2, b2, w3, b3


def scale_shapes(config: dict, channels=None):
    """[(C_s, H_s, W_s)] of the five residual groups for the config's voxel grid (model.py:64-79)."""
    lm = config["lidar_module"]
    ch = channels or [lm[f"out_feature{i}"] for i in range(1, 6)]
    H, W = int(config["voxel_length"]), int(config["voxel_width"])
    out = []
    for s, c in zip(SCALE_STRIDES, ch):
        out.append((c, H, W))
        H, W = (H + 1) // 2, (W + 1) // 2  # 3x3 stride-2 pad-1 conv
    return out


def nms_boxes(seed: int, n: int) -> np.ndarray:
    """(n,7) float32 [x,y,z,l,w,h,yaw]: the distribution of the SURVEY 3.2 probe (car-sized, uniform centres)."""
    rng = np.random.default_rng(4242 + seed)
    c = np.stack([rng.uniform(0, 70, n), rng.uniform(-30, 30, n), rng.uniform(-2.0, -1.0, n)], axis=1)
    s = np.array([4.0, 2.0, 1.5]) * np.exp(0.1 * rng.normal(size=(n, 3)))
    y = rng.uniform(-math.pi, math.pi, (n, 1))
    return np.concatenate([c, s, y], axis=1).astype(np.float32)


# workloads of BASELINE.json `configs`
def workload(name: str) -> dict:
    if name == "cfg0":  # single CARLA-shaped frame, 700x800 BEV, K=3, one scale, batch 1
        return dict(name=name, batch=1, bev=(700, 800), scales=(1,), k=3, max_num_pc=20000, n_beams=32, n_az=1500,
                    mode="fp32")
    if name == "cfg1":  # batch 4, K=5, every residual group, fp32
        return dict(name=name, batch=4, bev=(700, 800), scales=(1, 2, 3, 4, 5), k=5, max_num_pc=20000, n_beams=32,
                    n_az=1500, mode="fp32")
    if name == "cfg2":  # 64-beam density, K=10, bf16 MLP, batch 8
        return dict(name=name, batch=8, bev=(700, 800), scales=(1, 2, 3, 4, 5), k=10, max_num_pc=131072, n_beams=64,
                    n_az=4800, mode="bf16")
    if name == "yaml":  # the reference YAML's own 384x256 grid
        return dict(name=name, batch=2, bev=(384, 256), scales=(1, 2, 3, 4, 5), k=3, max_num_pc=20000, n_beams=32,
                    n_az=1500, mode="fp32")
    if name == "tiny":
        return dict(name=name, batch=2, bev=(96, 80), scales=(1, 2), k=3, max_num_pc=2048, n_beams=24, n_az=200,
                    mode="fp32", lidar_range=dict(lidar_x_max=24.0, lidar_y_min=-10.0, lidar_y_max=10.0))
    raise KeyError(name)


def make_workload(name_or_dict, seed: int = 0, c_img: int = 128, img_hw=(120, 160), channels=None):
    """All host-side numpy inputs of one fusion step: points, uv, counts, image map, BEV maps, weights."""
    wl = workload(name_or_dict) if isinstance(name_or_dict, str) else dict(name_or_dict)
    cfg = G.carla_config(voxel_length=wl["bev"][0], voxel_width=wl["bev"][1], max_num_pc=wl["max_num_pc"],
                         **wl.get("lidar_range", {}))
    B = wl["batch"]
    pts, uv, cnt = make_points(seed, cfg, wl["n_beams"], wl["n_az"], B, uniform=wl.get("uniform", False))
    rng = np.random.default_rng(9000 + seed)
    img = rng.standard_normal((B, c_img, img_hw[0], img_hw[1]), dtype=np.float32)
    shapes = scale_shapes(cfg, channels)
    scales = []
    for g in wl["scales"]:
        C, H, W = shapes[g - 1]
        scales.append(dict(group=g, stride=SCALE_STRIDES[g - 1], C=C, H=H, W=W,
                           geom=G.scale_geometry(cfg, SCALE_STRIDES[g - 1]),
                           bev=rng.standard_normal((B, C, H, W), dtype=np.float32),
                           weights=mlp_weights(seed, c_img, C)))
    return dict(workload=wl, config=cfg, points=pts, uv=uv, num_points=cnt, img_feat=img, scales=scales,
                calib=G.calibration_crt(), radius=float(G.fusion_option(cfg, "fusion_radius")), k=wl["k"])
