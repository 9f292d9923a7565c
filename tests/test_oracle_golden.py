"""CPU: the oracle (oracle/cf_oracle.c) against the golden vectors produced by the reference itself
(oracle/gen_golden.py) and against the reference's own known answers."""
import numpy as np
import pytest


def test_sat_known_answers(oracle):
    # separation_axis_theorem.py:98-105 prints True True True for these three polygon pairs
    a = [(0, 0), (70, 70), (70, 0), (0, 70)]
    b = [(70, 70), (150, 70), (150, 150), (70, 150)]
    c = [(30, 30), (150, 70), (70, 150)]
    assert oracle.sat_polygons(a, b) is True
    assert oracle.sat_polygons(a, c) is True
    assert oracle.sat_polygons(b, c) is True
    # a clearly separated pair, and touching-counts-as-overlap (closed intervals, :53)
    sq = lambda x: [(x, 0), (x + 1, 0), (x + 1, 1), (x, 1)]
    assert oracle.sat_polygons(sq(0), sq(3)) is False
    assert oracle.sat_polygons(sq(0), sq(1)) is True
    assert oracle.sat_polygons(sq(0), sq(1.0001)) is False


def test_iou_known_answer(oracle, golden):
    g = golden("box_iou_known.npz")
    i3, i2 = oracle.box3d_iou_corners(g["corners_pred"], g["corners_gt"])
    # IOU.py:161-168 -> (0.6879605829952761, 0.7003692488090808)
    assert abs(g["iou"][0] - 0.6879605829952761) < 1e-15 and abs(g["iou"][1] - 0.7003692488090808) < 1e-15
    assert abs(i3 - g["iou"][0]) < 1e-12 and abs(i2 - g["iou"][1]) < 1e-12


def test_iou_edge_cases(oracle):
    box = np.array([0, 0, 0, 4, 2, 1.5, 0], np.float32)
    assert oracle.box3d_iou(box, box) == pytest.approx((1.0, 1.0), abs=1e-12)           # identical
    inner = np.array([0, 0, 0, 2, 1, 1, 0], np.float32)
    assert oracle.box3d_iou(inner, box) == pytest.approx((1 / 6, 1 / 4), abs=1e-12)     # contained
    assert oracle.box3d_iou(box, np.array([4, 0, 0, 4, 2, 1.5, 0], np.float32)) == (0.0, 0.0)  # edge touching


@pytest.mark.parametrize("name", ["uniform_0", "uniform_1", "uniform_2", "uniform_50", "uniform_200", "uniform_600",
                                  "uniform_2000", "dense_300", "edge_cases"])
def test_nms_sat_keep_lists(oracle, golden, name):
    g = golden("nms_sat.npz")
    keep = oracle.nms_sat(g[f"{name}__boxes"])
    assert np.array_equal(keep, g[f"{name}__keep"])  # bit-exact index lists (Appendix A14)


def test_sat_pairs_and_vertices(oracle, golden):
    g = golden("sat_pairs.npz")
    boxes = g["boxes"]
    for i in range(boxes.shape[0]):
        assert np.array_equal(oracle.get_vertice_rect(boxes[i]), g["vertices"][i])  # fp32 bit-exact corners
    assert np.array_equal(oracle.sat_matrix(boxes), g["overlap"])
    # the reference evaluates v**2 through libm powf; the oracle's powf mode reproduces that too
    oracle.set_sq_mode(1)
    try:
        assert np.array_equal(oracle.sat_matrix(boxes), g["overlap"])
    finally:
        oracle.set_sq_mode(0)


def test_box_iou_matrix(oracle, golden):
    g = golden("box_iou.npz")
    i3, i2 = oracle.box3d_iou_matrix(g["boxes_a"], g["boxes_b"])
    ok = (g["qhull_error"] == 0) & np.isfinite(g["iou3d"]) & np.isfinite(g["iou2d"])
    # float32 cos/sin of the heading differ by <= 1 ulp between numpy's SIMD kernels and libm -> 1e-6
    assert np.abs(i3 - g["iou3d"])[ok].max() < 1e-6
    assert np.abs(i2 - g["iou2d"])[ok].max() < 1e-6
    assert (g["iou2d"][ok] > 0).sum() > 100


def test_nms_iou_keep_list(oracle, golden):
    g = golden("nms_iou.npz")
    assert np.array_equal(oracle.nms_iou(g["boxes"], float(g["thr"])), g["keep"])


def test_get_bboxes(oracle, golden):
    g = golden("get_bboxes.npz")
    outs = [oracle.get_bboxes(g["cls"][b], g["box"][b], float(g["thr"])) for b in range(g["cls"].shape[0])]
    assert [o.shape[0] for o in outs] == g["counts"].tolist()
    assert np.array_equal(np.concatenate(outs), g["boxes"])


def test_projection_and_filter(oracle, golden, dcf):
    """synthetic.filter_and_project restates data_import_carla.py:196-229,261-267; CRT restates :31-34,180-194."""
    g = golden("projection.npz")
    assert np.allclose(dcf.geometry.calibration_crt(), g["crt"], rtol=0, atol=1e-4)
    cfg = dcf.geometry.carla_config()
    p, uv = dcf.synthetic.filter_and_project(g["raw"], cfg, g["crt"])
    n = int(g["num_points_raw"])
    assert p.shape[0] == n
    assert np.array_equal(p, g["pointcloud_raw"][:n])
    assert np.allclose(uv, g["projected_loc_uv"][:n], rtol=0, atol=2e-3)   # BLAS vs explicit summation order
    assert not g["pointcloud_raw"][n:].any() and not g["projected_loc_uv"][n:].any()  # zero padding rows
    # oracle projection (Appendix A6) agrees with the dataset's precomputed uv to 1e-3 px
    uv_o = oracle.project_points(p, g["crt"])
    assert np.abs(uv_o - g["projected_loc_uv"][:n]).max() < 2e-3


@pytest.mark.parametrize("tag", ["a", "b"])
def test_voxelize_project_restatement_matches_reference(oracle, golden, dcf, tag):
    """oracle.voxelize_project restates CarlaDataset.Voxelization_Projection (data_import_carla.py:212-267); the
    fixture holds the reference's own tensors (sequential index_put_ semantics, see oracle/gen_golden.py)."""
    g = golden(f"voxelize_{tag}.npz")
    cfg = dcf.geometry.carla_config()
    vox, pc, uv, num = oracle.voxelize_project(g["raw"], cfg, g["crt"])
    ref = np.zeros(int(np.prod(g["vox_shape"])), np.float32)
    ref[g["vox_idx"]] = g["vox_val"]
    n = int(g["num_points_raw"])
    assert num == n
    assert np.array_equal(vox.ravel(), ref)                                   # voxel grid: bit-exact
    assert np.array_equal(pc[:n + 8], g["pointcloud_raw"])                    # points: exact, zero padded
    assert np.abs(uv[:n + 8] - g["projected_loc_uv"]).max() < 2e-3            # BLAS vs explicit summation order
    # host helpers reproduce pc_to_voxel_indice and the float32 thresholds
    assert dcf.geometry.voxel_matrix(cfg) == oracle.voxel_matrix(cfg) == (5, 4, 10, 0, 120, 24)


def test_postprocess_chain(oracle, golden):
    """The evaluator's chain (test.py:79-85: get_bboxes -> NMS_SAT) as run by the reference on one prediction map."""
    g = golden("postprocess_chain.npz")
    for b in range(3):
        boxes = oracle.get_bboxes(g["cls"][b], g["box"][b], float(g["thr"]))
        assert boxes.shape[0] == int(g["counts"][b]) and np.array_equal(boxes.reshape(-1, 7), g[f"boxes_{b}"])
        keep = oracle.nms_sat(boxes) if boxes.shape[0] else np.zeros((0,), np.int32)
        assert np.array_equal(keep, g[f"keep_{b}"])
