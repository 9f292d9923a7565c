"""ctypes front-end of the CPU oracle (oracle/cf_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; the product package never does (its ops raise if the CUDA library is missing).

All functions take / return numpy arrays.  Semantics and reference citations are in cf_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcf_oracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "cf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libcf_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.cfo_set_threads.argtypes = [C.c_int]
        L.cfo_set_threads.restype = C.c_int
        L.cfo_knn_bruteforce.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                         C.c_float, C.c_float, C.c_float, C.c_int32, _i32p, C.c_int64, C.c_int64]
        L.cfo_project_points.argtypes = [_f32p, C.c_int32, _f32p, _f32p]
        L.cfo_gather_points.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_int32, _f32p, C.c_int32, C.c_float,
                                        C.c_float, _f32p]
        L.cfo_fusion_mlp.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_int32, _f32p, C.c_int32, _f32p, _i32p,
                                     C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, _f32p, _f32p, _f32p,
                                     _f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_int64]
        L.cfo_get_vertice_rect.argtypes = [_f32p, _f32p]
        L.cfo_sat_axes.argtypes = [_f32p, _f32p]
        L.cfo_set_sq_mode.argtypes = [C.c_int]
        L.cfo_sat_overlap.argtypes = [_f32p, _f32p]
        L.cfo_sat_overlap.restype = C.c_int
        L.cfo_sat_polygons.argtypes = [_f32p, C.c_int32, _f32p, C.c_int32]
        L.cfo_sat_polygons.restype = C.c_int
        L.cfo_nms_sat.argtypes = [_f32p, C.c_int32, _i32p]
        L.cfo_nms_sat.restype = C.c_int32
        L.cfo_sat_matrix.argtypes = [_f32p, C.c_int32, _u8p]
        L.cfo_get_3d_box.argtypes = [_f32p, _f64p]
        L.cfo_box3d_iou.argtypes = [_f32p, _f32p, C.c_float, _f64p, _f64p]
        L.cfo_box3d_iou_corners.argtypes = [_f64p, _f64p, _f64p, _f64p]
        L.cfo_box3d_iou_matrix.argtypes = [_f32p, C.c_int32, _f32p, C.c_int32, C.c_float, _f64p, _f64p]
        L.cfo_nms_iou.argtypes = [_f32p, C.c_int32, C.c_float, _i32p]
        L.cfo_nms_iou.restype = C.c_int32
        L.cfo_get_bboxes.argtypes = [_f32p, _f32p, C.c_int32, C.c_int32, C.c_float, _f32p]
        L.cfo_get_bboxes.restype = C.c_int32
        _lib = L
    return _lib


def set_threads(n: int) -> int:
    """Use n OpenMP threads (returns the count in effect)."""
    return int(lib().cfo_set_threads(int(n)))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------------ fusion
def knn_bruteforce(pts, n_valid, H, W, x0, y0, dx, dy, r2, K, cell_range=None):
    """(H*W or range, K) int32 indices, -1 = empty slot.  One frame."""
    pts = _f32(pts)
    b, e = (0, H * W) if cell_range is None else cell_range
    out = np.empty((e - b, K), dtype=np.int32)
    lib().cfo_knn_bruteforce(_p(pts, _f32p), int(n_valid), H, W, np.float32(x0), np.float32(y0), np.float32(dx),
                             np.float32(dy), np.float32(r2), K, _p(out, _i32p), b, e)
    return out if cell_range is not None else out.reshape(H, W, K)


def project_points(pts, calib):
    pts = _f32(pts)
    calib = _f32(calib)
    assert calib.shape == (4, 3)
    uv = np.empty((pts.shape[0], 2), dtype=np.float32)
    lib().cfo_project_points(_p(pts, _f32p), pts.shape[0], _p(calib, _f32p), _p(uv, _f32p))
    return uv


def gather_points(img, uv, img_w=640.0, img_h=480.0):
    """img (Ci,Hf,Wf), uv (n,2) -> (n,Ci)."""
    img = _f32(img)
    uv = _f32(uv)
    Ci, Hf, Wf = img.shape
    out = np.empty((uv.shape[0], Ci), dtype=np.float32)
    lib().cfo_gather_points(_p(img, _f32p), Ci, Hf, Wf, _p(uv, _f32p), uv.shape[0], img_w, img_h, _p(out, _f32p))
    return out


def fusion_mlp(bev, feat, pts, knn, geom, weights, cell_range=None):
    """bev (C,H,W); feat (n,Ci); pts (N,3); knn (H,W,K); geom=(x0,y0,dx,dy); weights=(W1,b1,W2,b2,W3,b3)."""
    bev = _f32(bev)
    Cc, H, W = bev.shape
    feat = _f32(feat)
    pts = _f32(pts)
    knn = np.ascontiguousarray(knn, dtype=np.int32)
    K = knn.shape[-1]
    W1, b1, W2, b2, W3, b3 = [_f32(w) for w in weights]
    Ci = feat.shape[1]
    assert W1.shape == (Cc, Ci + 3) and W2.shape == (Cc, Cc) and W3.shape == (Cc, Cc)
    x0, y0, dx, dy = [np.float32(g) for g in geom]
    if cell_range is None:
        out = np.empty_like(bev)
        lib().cfo_fusion_mlp(_p(bev, _f32p), Cc, H, W, _p(feat, _f32p), Ci, _p(pts, _f32p), _p(knn, _i32p), K, x0, y0,
                             dx, dy, _p(W1, _f32p), _p(b1, _f32p), _p(W2, _f32p), _p(b2, _f32p), _p(W3, _f32p),
                             _p(b3, _f32p), _p(out, _f32p), 0, H * W)
        return out
    # bounded sample (bench cpu_baseline): knn is (range,K); output written into a full-size buffer
    b, e = cell_range
    out = np.zeros_like(bev)
    lib().cfo_fusion_mlp(_p(bev, _f32p), Cc, H, W, _p(feat, _f32p), Ci, _p(pts, _f32p), _p(knn, _i32p), K, x0, y0, dx,
                         dy, _p(W1, _f32p), _p(b1, _f32p), _p(W2, _f32p), _p(b2, _f32p), _p(W3, _f32p), _p(b3, _f32p),
                         _p(out, _f32p), b, e)
    return out


def fusion_forward(bev, img, pts, n_valid, geom, radius, K, weights, calib=None, uv=None, img_w=640.0, img_h=480.0,
                   return_knn=False):
    """One frame, one scale, the whole layer: KNN -> projection -> gather -> MLP -> pool -> add."""
    bev = _f32(bev)
    _, H, W = bev.shape
    pts = _f32(pts)
    x0, y0, dx, dy = geom
    r2 = np.float32(radius) * np.float32(radius)
    knn = knn_bruteforce(pts, n_valid, H, W, x0, y0, dx, dy, r2, K)
    if uv is None:
        uv = project_points(pts[:n_valid], calib)
    feat = gather_points(img, _f32(uv)[:n_valid], img_w, img_h)
    out = fusion_mlp(bev, feat, pts, knn, geom, weights)
    return (out, knn) if return_knn else out


# ------------------------------------------------------------------------------------ post-process
def get_vertice_rect(box7):
    b = _f32(box7)
    o = np.empty(8, dtype=np.float32)
    lib().cfo_get_vertice_rect(_p(b, _f32p), _p(o, _f32p))
    return o.reshape(4, 2)


def sat_axes(box7):
    b = _f32(box7)
    o = np.empty(8, dtype=np.float32)
    lib().cfo_sat_axes(_p(b, _f32p), _p(o, _f32p))
    return o.reshape(4, 2)


def set_sq_mode(m: int):
    lib().cfo_set_sq_mode(int(m))


def sat_overlap(box_a, box_b) -> bool:
    a, b = _f32(box_a), _f32(box_b)
    return bool(lib().cfo_sat_overlap(_p(a, _f32p), _p(b, _f32p)))


def sat_polygons(va, vb) -> bool:
    """separating_axis_theorem on two convex polygons given as (n,2) vertex arrays."""
    a, b = _f32(va).reshape(-1, 2), _f32(vb).reshape(-1, 2)
    return bool(lib().cfo_sat_polygons(_p(a, _f32p), a.shape[0], _p(b, _f32p), b.shape[0]))


def nms_sat(boxes):
    """(n,7) -> ascending kept input indices (int32)."""
    boxes = _f32(boxes).reshape(-1, 7)
    keep = np.empty(max(boxes.shape[0], 1), dtype=np.int32)
    n = lib().cfo_nms_sat(_p(boxes, _f32p), boxes.shape[0], _p(keep, _i32p))
    return keep[:n].copy()


def sat_matrix(boxes):
    boxes = _f32(boxes).reshape(-1, 7)
    n = boxes.shape[0]
    m = np.empty((n, n), dtype=np.uint8)
    lib().cfo_sat_matrix(_p(boxes, _f32p), n, _p(m, _u8p))
    return m


def get_3d_box(box7):
    b = _f32(box7)
    o = np.empty(24, dtype=np.float64)
    lib().cfo_get_3d_box(_p(b, _f32p), _p(o, _f64p))
    return o.reshape(8, 3)


def box3d_iou(box_a, box_b, nudge_b=0.0):
    a, b = _f32(box_a), _f32(box_b)
    i3, i2 = C.c_double(), C.c_double()
    lib().cfo_box3d_iou(_p(a, _f32p), _p(b, _f32p), np.float32(nudge_b), C.byref(i3), C.byref(i2))
    return i3.value, i2.value


def box3d_iou_corners(c1, c2):
    """IOU.box3d_iou on two (8,3) float64 corner arrays."""
    c1 = np.ascontiguousarray(c1, dtype=np.float64)
    c2 = np.ascontiguousarray(c2, dtype=np.float64)
    i3, i2 = C.c_double(), C.c_double()
    lib().cfo_box3d_iou_corners(_p(c1, _f64p), _p(c2, _f64p), C.byref(i3), C.byref(i2))
    return i3.value, i2.value


def box3d_iou_matrix(boxes_a, boxes_b, nudge_b=0.0):
    a = _f32(boxes_a).reshape(-1, 7)
    b = _f32(boxes_b).reshape(-1, 7)
    i3 = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    i2 = np.empty_like(i3)
    lib().cfo_box3d_iou_matrix(_p(a, _f32p), a.shape[0], _p(b, _f32p), b.shape[0], np.float32(nudge_b),
                               _p(i3, _f64p), _p(i2, _f64p))
    return i3, i2


def nms_iou(boxes, thr=0.01):
    boxes = _f32(boxes).reshape(-1, 7)
    keep = np.empty(max(boxes.shape[0], 1), dtype=np.int32)
    n = lib().cfo_nms_iou(_p(boxes, _f32p), boxes.shape[0], np.float32(thr), _p(keep, _i32p))
    return keep[:n].copy()


def get_bboxes(cls4, box14, thr=0.8):
    """One frame: cls (4,H,W), decoded boxes (14,H,W) -> (n,7) in anchor-major / row-major order."""
    cls4 = _f32(cls4)
    box14 = _f32(box14)
    _, H, W = cls4.shape
    out = np.empty((2 * H * W, 7), dtype=np.float32)
    n = lib().cfo_get_bboxes(_p(cls4, _f32p), _p(box14, _f32p), H, W, np.float32(thr), _p(out, _f32p))
    return out[:n].copy()


# ---------------------------------------------------------------------------------------------------------------
# Dataset-side voxelisation + projection: numpy restatement of CarlaDataset.Voxelization_Projection
# and .Projection, data_import_carla.py:196-267, statement by statement.  Pinned against the reference's own code
# run on synthetic sweeps (tests/golden/voxelize.npz, oracle/gen_golden.py).
# ---------------------------------------------------------------------------------------------------------------
def voxel_matrix(config):
    """pc_to_voxel_indice, data_import_carla.py:35-43: integer scales / offsets computed with Python floats."""
    xs = int(config["voxel_length"] / (config["lidar_x_max"] - config["lidar_x_min"]))
    ys = int(config["voxel_width"] / (config["lidar_y_max"] - config["lidar_y_min"]))
    zs = int(config["voxel_channel"] / (config["lidar_z_max"] - config["lidar_z_min"]))
    return (xs, ys, zs, int(-config["lidar_x_min"] * xs), int(-config["lidar_y_min"] * ys), int(-config["lidar_z_min"] * zs))


def voxelize_project(raw, config, crt):
    """raw (n,3) float32 -> (lidar_voxel (Z,X,Y) f32, pointcloud_raw (max_num_pc,3), uv (max_num_pc,2), num).

    Follows the reference line by line, including its quirks:
      * the range filter zeroes rejected points and keeps the first count_nonzero/3 columns of nonzero() (:214-229);
      * the trilinear splat uses `voxel[idx] += w` with advanced indexing, i.e. for points that fall into the same
        voxel within one of the 8 statements only the LAST one counts (:236-258);
      * the image filter compares u with image_height and v with image_width (:202-205) and has the same
        nonzero()/2 rule (:206-207)."""
    f32 = np.float32
    raw = np.ascontiguousarray(raw, dtype=f32)
    x, y, z = raw[:, 0], raw[:, 1], raw[:, 2]
    d = config["delta"]
    # the thresholds are Python floats (max - delta in double), promoted to float32 by the tensor comparison
    keep = ((x > f32(config["lidar_x_min"])) & (x < f32(config["lidar_x_max"] - d)) & (y > f32(config["lidar_y_min"])) &
            (y < f32(config["lidar_y_max"] - d)) & (z > f32(config["lidar_z_min"])) & (z < f32(config["lidar_z_max"] - d)))
    surv = np.nonzero(keep)[0]
    nz = int((raw[surv] != 0).sum())
    rows = np.nonzero((raw[surv] != 0).T)          # nonzero() of the (3, N) tensor: row-major order
    cols = rows[1][: nz // 3]                        # first third of the entries = column indices
    pts = raw[surv][cols]
    xs, ys, zs, xo, yo, zo = voxel_matrix(config)
    fx = pts[:, 0] * f32(xs) + f32(xo)
    fy = pts[:, 1] * f32(ys) + f32(yo)
    fz = pts[:, 2] * f32(zs) + f32(zo)
    xl, yl, zl = fx.astype(np.int64), fy.astype(np.int64), fz.astype(np.int64)
    dx, dy, dz = fx - xl.astype(f32), fy - yl.astype(f32), fz - zl.astype(f32)
    one = f32(1)
    vox = np.zeros((config["voxel_channel"], config["voxel_length"], config["voxel_width"]), dtype=f32)
    for zi, xi, yi, w in ((zl, xl, yl, (one - dx) * (one - dy) * (one - dz)), (zl + 1, xl, yl, (one - dx) * (one - dy) * dz),
                          (zl, xl + 1, yl, dx * (one - dy) * (one - dz)), (zl + 1, xl + 1, yl, dx * (one - dy) * dz),
                          (zl, xl, yl + 1, (one - dx) * dy * (one - dz)), (zl + 1, xl, yl + 1, (one - dx) * dy * dz),
                          (zl, xl + 1, yl + 1, dx * dy * (one - dz)), (zl + 1, xl + 1, yl + 1, dx * dy * dz)):
        vox[zi, xi, yi] += w                         # numpy fancy `+=`: last write wins, like torch index_put_
    # Projection
    crt = np.asarray(crt, dtype=f32)
    q = (pts[:, 0:1] * crt[0] + pts[:, 1:2] * crt[1]) + (pts[:, 2:3] * crt[2] + crt[3])
    u, v = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
    ok = (u > 0) & (u < f32(config["image_height"])) & (v > 0) & (v < f32(config["image_width"]))
    s2 = np.nonzero(ok)[0]
    uv2 = np.stack([u[s2], v[s2]], axis=0)
    r2 = np.nonzero(uv2 != 0)
    c2 = r2[1][: int((uv2 != 0).sum()) // 2]
    num = c2.shape[0]
    N = int(config["max_num_pc"])
    pc = np.zeros((N, 3), dtype=f32)
    uvp = np.zeros((N, 2), dtype=f32)
    pc[:num] = pts[s2][c2]
    uvp[:num] = uv2.T[c2]
    return vox, pc, uvp, num


# ---- target assignment and value of LossTotal (loss.py:33-189), restated with explicit random draws -----
def loss_targets(ref_boxes, num_ref, H, W, scales, reduced_scale, positive_range, regress_type, pos_thr, neg_thr, keys, cand):
    """getPositionOfPositive (loss.py:74-110) + getPositionOfNegative (loss.py:112-127) for every frame, plain loops.
    RNG contract (include/cf_b200.h, cf_loss_targets): np.random.shuffle(list) == list[stable argsort(keys[:len(list)])],
    the rejection loop's (x, y) draws are cand[b] in order.  Returns pos (B,pos_thr), npos (B), neg (B,neg_thr+1), nneg (B),
    reg (B,M,R*R): int32 linear cells x*W+y, -1 padded."""
    ref_boxes = np.asarray(ref_boxes, dtype=np.float32)
    B, M, _ = ref_boxes.shape
    R = int(positive_range)
    xs, ys, xo, yo = (np.float32(v) for v in scales)
    rs = np.float32(reduced_scale)
    pos = -np.ones((B, pos_thr), np.int32)
    neg = -np.ones((B, neg_thr + 1), np.int32)
    reg = -np.ones((B, M, R * R), np.int32)
    npos, nneg = np.zeros(B, np.int32), np.zeros(B, np.int32)
    for b in range(B):
        plist = []
        for i in range(int(num_ref[b])):
            # loss.py:86-87: float32 tensor arithmetic, int() truncates toward zero
            px = int(np.float32(np.float32(ref_boxes[b, i, 0] * xs) + xo) / rs)
            py = int(np.float32(np.float32(ref_boxes[b, i, 1] * ys) + yo) / rs)
            if px < 0 or px > H - 1 or py < 0 or py > W - 1:
                continue
            for xi in range(R):
                x = px - int(R / 2) + xi
                for yi in range(R):
                    y = py - int(R / 2) + yi
                    if x < 0 or x > H - 1 or y < 0 or y > W - 1:
                        continue
                    plist.append(x * W + y)
                    if regress_type == 0 or (x == px and y == py):
                        reg[b, i, xi * R + yi] = x * W + y
        order = np.argsort(np.asarray(keys[b][:len(plist)], dtype=np.float32), kind="stable")
        plist = [plist[j] for j in order][:pos_thr]
        npos[b] = len(plist)
        pos[b, :len(plist)] = plist
        taken = set(plist)
        k = 0
        for x, y in np.asarray(cand[b]):
            if int(x) * W + int(y) in taken:
                continue
            neg[b, k] = int(x) * W + int(y)
            k += 1
            if k > neg_thr:
                break
        nneg[b] = k
    return pos, npos, neg, nneg, reg


def loss_per_frame(ref_boxes, num_ref, pred_cls, pred_reg, anchors, targets, gain):
    """Value of LossTotal for every frame (loss.py:52-71, 129-189) from the assigned targets, frame by frame and box by
    box with torch CPU ops as the reference does (CrossEntropyLoss mean, SmoothL1Loss none); float32."""
    import torch
    import torch.nn.functional as F
    pos, npos, neg, nneg, reg = targets
    pred_cls, pred_reg = torch.as_tensor(pred_cls), torch.as_tensor(pred_reg)
    anchors = torch.as_tensor(anchors)                     # (14, H, W)
    ref_boxes = torch.as_tensor(np.asarray(ref_boxes, dtype=np.float32))
    B, _, H, W = pred_cls.shape
    out = []
    for b in range(B):
        pc, ng = torch.as_tensor(pos[b, :npos[b]]).long(), torch.as_tensor(neg[b, :nneg[b]]).long()
        total = torch.zeros(())
        for a in range(2):                                 # loss.py:62-63: channels [:2] and [2:4]
            logits = pred_cls[b, 2 * a:2 * a + 2].reshape(2, H * W)
            total = total + F.cross_entropy(logits[:, ng].T, torch.zeros(len(ng), dtype=torch.long))
            if len(pc):
                total = total + F.cross_entropy(logits[:, pc].T, torch.ones(len(pc), dtype=torch.long))
        regl = torch.zeros(())
        for i in range(int(num_ref[b])):
            cells = torch.as_tensor(reg[b, i][reg[b, i] >= 0]).long()
            if len(cells) == 0:
                continue
            p = pred_reg[b].reshape(14, H * W)[:, cells].T.reshape(-1, 2, 7)
            an = anchors.reshape(2, 7, H * W)[:, :, cells].permute(2, 0, 1)
            r = ref_boxes[b, i][None, None, :].expand(len(cells), 2, -1)
            xy = (r[..., :2] - an[..., :2]) / torch.sqrt(an[..., 3:4] ** 2 + an[..., 4:5] ** 2)
            z = (r[..., 2:3] - an[..., 2:3]) / an[..., 5:6]
            whd = torch.log(r[..., 3:6] / an[..., 3:6])
            d = r[..., 6] - an[..., 6]
            tgt = torch.cat((xy, z, whd, torch.atan2(torch.sin(d), torch.cos(d))[..., None]), dim=-1)
            regl = regl + F.smooth_l1_loss(p, tgt, reduction="none").sum() / (len(cells) * 14)
        out.append(float(total + gain * regl))
    return np.asarray(out, dtype=np.float64)
