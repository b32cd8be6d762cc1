"""The CUDA path against fixtures produced by the REFERENCE'S OWN code (oracle/_ref: the reference's LidarFeatureAssociate.cpp, CostFunction.h,
Equirectangular.{h,cpp} compiled where they lie with the stand-in container types of oracle/shim; tests/make_golden.py: golden_ref_assoc /
golden_ref_path) - no oracle in between.  The pair used here is the one whose pose estimate is the identity (ASSOC_CASES[4]): exactly representable as
an angle-axis block, so the float32 world clouds are the same bits on both sides."""
import os

import numpy as np
import pytest

from test_reference_pinning import ASSOC_CASES, assoc_case

G = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.gpu
CI = 4


def _fixture():
    g = np.load(os.path.join(G, "ref_assoc.npz"))
    seed, n_az, perturb, tol, thr_p, thr_l = ASSOC_CASES[CI]
    A, B, RB, tB = assoc_case(None, seed, n_az, perturb)
    assert np.array_equal(RB, np.eye(3)) and np.all(tB == 0) and np.array_equal(A["R_wl"], np.eye(3)) and np.all(A["t_wl"] == 0)
    return {k[len(f"c{CI}_"):]: g[k] for k in g.files if k.startswith(f"c{CI}_")}, A, B, tol, thr_p, thr_l


def _line_frame(f, R, t):
    from panovlm_b200 import LineFrame
    return LineFrame(f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["end_points"], R, t)


def test_point2plane_kernel_equals_the_reference_association(gpu_ctx):
    """k_associate (A1) == AssociatePoint2Plane of the reference: same accepted queries in the same order, query points bit-identical, planes to 1e-9."""
    exp, A, B, tol, thr_p, _ = _fixture()
    gpu_ctx.frames_set([A["surfLessFlat"], B["surfLessFlat"]], [A["surfFlat"], B["surfFlat"]])
    e, q, pt, pl = gpu_ctx.frames_associate_point2plane(np.zeros((2, 6)), [0], [1], tol, thr_p, 10)
    assert len(pt) == len(exp["p2plane_point"]) > 200 and np.all(np.diff(q) > 0)
    assert np.abs(pt - exp["p2plane_point"]).max() < 1e-12
    assert np.abs(pl - exp["p2plane_plane"]).max() < 1e-9


def test_line_associations_equal_the_reference_association(gpu_ctx):
    """A2 / A3: AssociateLine2Line, AssociateLine2LineKNN, AssociatePoint2Line, AssociatePoint2LineSegmentKNN, AssociatePoint2LineSegment."""
    exp, A, B, _, _, thr = _fixture()
    I, z = np.eye(3), np.zeros(3)
    fa, fb = _line_frame(A, I, z), _line_frame(B, I, z)
    for name, got in (("l2l", gpu_ctx.line2line_associate(fa, fb, thr)), ("l2lknn", gpu_ctx.line2line_knn_associate(fa, fb, thr))):
        nl, rl, a, b = got
        assert np.array_equal(nl, exp[name + "_nei"]) and np.array_equal(rl, exp[name + "_ref"]) and len(nl) >= 10, name
        assert np.abs(a - exp[name + "_a"]).max() < 1e-12 and np.abs(b - exp[name + "_b"]).max() < 1e-12, name
    for name, got in (("p2lsk", gpu_ctx.point2line_segment_knn_associate(fa, fb, thr)), ("p2ls", gpu_ctx.point2line_segment_associate(fa, fb, thr))):
        _, _, pt, a, b = got
        assert len(pt) == len(exp[name + "_point"]) > 100 and np.abs(pt - exp[name + "_point"]).max() < 1e-12, name
        assert np.abs(a - exp[name + "_a"]).max() < 1e-12 and np.abs(b - exp[name + "_b"]).max() < 1e-12, name
    gpu_ctx.frames_set_corners([A["cornerLessSharp"], B["cornerLessSharp"]])
    _, _, pt, a, b = gpu_ctx.frames_associate_point2line(np.zeros((2, 6)), np.array([0], np.int32), np.array([1], np.int32), thr)
    ea, eb = exp["p2l_a"], exp["p2l_b"]
    assert len(pt) == len(exp["p2l_point"]) > 100 and np.abs(pt - exp["p2l_point"]).max() < 1e-12
    same = np.maximum(np.abs(a - ea).max(1), np.abs(b - eb).max(1))
    flip = np.maximum(np.abs(a - eb).max(1), np.abs(b - ea).max(1))                     # the PCA direction's sign is the eigen solver's choice
    assert np.minimum(same, flip).max() < 1e-9


def test_projection_kernel_equals_the_reference_projection(gpu_ctx):
    """k_project (P / E1 / E2): CamToImage of the reference in float32 (FastAtan2), bit for bit, on the fixture's 4008 camera-frame points."""
    g = np.load(os.path.join(G, "ref_geometry.npz"))
    cam = np.concatenate([g["cam_f"], np.zeros((len(g["cam_f"]), 1), np.float32)], axis=1)
    out = gpu_ctx.project_equirect(cam, np.eye(4), int(g["rows"]), int(g["cols"]))
    assert np.array_equal(out[:, :2], g["px_f"])


def test_camera_lidar_kernel_equals_the_reference_association(gpu_ctx):
    """k_angle_votes + tail (A4) == the reference's AssociateByAngle + Filter(false, true) + UniqueLinePair: pair lists, float32 scores, masks."""
    from test_reference_pinning import CAMLIDAR_VARIANTS, _camlidar_masks, camlidar_case
    g = np.load(os.path.join(G, "ref_camlidar.npz"))
    A, rows, cols, T, lines = camlidar_case()
    fr = _line_frame(A, A["R_wl"], A["t_wl"])
    n_seg = len(A["segment_coeffs"])
    for name, multi, masked in CAMLIDAR_VARIANTS:
        im, lm = _camlidar_masks(g, len(lines), n_seg) if masked else (None, None)
        il, ll, s, e, ang = gpu_ctx.camera_lidar_associate(rows, cols, lines, fr, T, True, multi, im, lm)
        assert np.array_equal(il, g[name + "_image"]) and np.array_equal(ll, g[name + "_lidar"]) and np.array_equal(ang, g[name + "_score"]), name
        assert np.abs(s - g[name + "_start"]).max() < 1e-12 and np.abs(e - g[name + "_end"]).max() < 1e-12, name


def test_depth_splat_kernel_equals_the_reference_image(gpu_ctx):
    """k_project + k_splat_finalize (P) == ProjectLidar2PanoramaDepth of the reference: identical uint16 images (window clipping, last writer wins)."""
    from test_reference_pinning import camlidar_case
    g = np.load(os.path.join(G, "ref_camlidar.npz"))
    A, _, _, T, _ = camlidar_case()
    for key, rows, cols, size in (("depth_720", 720, 1440, 3), ("depth_360", 360, 720, 4)):
        assert np.array_equal(gpu_ctx.project_depth_image(A["cloud"], T, rows, cols, size), g[key]), key


def test_generate_line_tracks_equals_the_reference_tracks(gpu_ctx):
    """pvb_generate_line_tracks (device vote matrices per pair + host union-find) == LidarLineMatch::GenerateTracks of the reference."""
    from panovlm_b200 import Context
    from test_reference_pinning import TRACK_CASES, track_case
    g = np.load(os.path.join(G, "ref_assoc.npz"))
    for ci, (nf, n_az, k, min_len, no_pose) in enumerate(TRACK_CASES):
        frames = track_case(nf, n_az)
        exp = [g[f"tr{ci}_feat"][g[f"tr{ci}_off"][t]:g[f"tr{ci}_off"][t + 1]] for t in range(len(g[f"tr{ci}_off"]) - 1)]
        pv = np.ones(nf, np.uint8)
        if no_pose is not None:
            pv[no_pose] = 0
        nbrs = Context.find_neighbors(np.array([f["t_wl"] for f in frames]), pv, None, k)
        got = gpu_ctx.generate_line_tracks([_line_frame(f, f["R_wl"], f["t_wl"]) for f in frames], nbrs, pv, 0.3, min_len)
        assert len(got) == len(exp) and all(np.array_equal(a, b) for a, b in zip(got, exp)), ci


def test_transform_and_undistort_kernels_equal_the_reference_clouds(gpu_ctx):
    """k_transform_simple (T1) bit for bit against the cloud left by the reference's Transform2LidarWorld; k_undistort against its UndistortCloud, with the
    tolerance of tests/test_undistort.py (device sin vs glibc: at most one float32 rounding flip on <= 1e-4 of the values)."""
    from test_reference_pinning import velodyne_case
    g = np.load(os.path.join(G, "ref_velodyne.npz"))
    cloud, R_wl, t_wl, sweeps = velodyne_case()
    assert np.array_equal(gpu_ctx.transform_cloud(cloud, R_wl, t_wl), g["world"])
    T_wl = np.eye(4); T_wl[:3, :3] = R_wl; T_wl[:3, 3] = t_wl
    n = len(cloud)
    off = (np.arange(len(sweeps) + 1) * n).astype(np.int32)
    T_we = np.stack([np.block([[R, t[:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) for R, t in sweeps])
    out = gpu_ctx.undistort_clouds(np.concatenate([cloud] * len(sweeps)), off, np.stack([T_wl] * len(sweeps)), T_we)
    bad = 0
    for k in range(len(sweeps)):
        got, ref = out[off[k]:off[k + 1]], g[f"undistorted{k}"]
        assert np.array_equal(got[:, 3], ref[:, 3])
        d = np.abs(got[:, :3].view(np.int32).astype(np.int64) - ref[:, :3].view(np.int32).astype(np.int64))
        assert d.max() <= 1, k
        bad += int((d > 0).sum())
    assert bad <= 1e-4 * 3 * off[-1] + 2


def test_pixel_knn_kernel_equals_the_reference_first_stage(gpu_ctx):
    """k_pixel_knn3 (A5, first stage) + pvb_pixel_line_candidates == the candidate lists the reference's own pixel-space Associate() hands to its RANSAC fit."""
    from panovlm_b200 import Context
    from test_reference_pinning import camlidar_case
    g = np.load(os.path.join(G, "ref_camlidar.npz"))
    A, rows, cols, T, lines = camlidar_case()
    cloud = A["cloud"][::4]
    line3, _, _ = gpu_ctx.pixel_line_neighbors(rows, cols, lines, cloud, T)
    off, idx = Context.pixel_line_candidates(len(lines), line3, 6)
    cam = gpu_ctx.transform_cloud(cloud, T[:3, :3], T[:3, 3])[:, :3]
    got = [cam[idx[off[l]:off[l + 1]]] for l in range(len(lines)) if off[l + 1] > off[l]]
    exp = [g["px_xyz"][g["px_off"][k]:g["px_off"][k + 1]] for k in range(len(g["px_off"]) - 1)]
    assert len(got) == len(exp) >= 10 and all(np.array_equal(a, b) for a, b in zip(got, exp))


def test_dense_kernel_equals_the_systems_built_from_the_reference_pieces(gpu_ctx):
    """k_associate<10, reduce> in dense mode (the bench / smoke path, configs[4]) against per-frame normal equations rebuilt from the reference's OWN Transform2LidarWorld,
    AssociatePoint2Plane and Point2Plane_Meter (tests/golden/ref_dense.npz): identical association counts, systems / gradients / costs to 1e-8 (smoke()'s gate)."""
    import panovlm_b200
    from panovlm_b200 import synth
    from test_reference_pinning import DENSE_CASE as c
    g = np.load(os.path.join(G, "ref_dense.npz"))
    d = synth.make_dense_sweep(n_target=c["n_target"], n_frames=c["n_frames"], pts_per_frame=c["pts_per_frame"], seed=c["seed"])
    gpu_ctx.dense_set_target(d["target"])
    gpu_ctx.dense_set_sources(d["src_local"], d["src_off"])
    prm = gpu_ctx.dense_params(c["plane_tol"], c["dist_thr"], 10, panovlm_b200.P2PLANE_METER, 1, c["huber"], c["weight"])
    s_gpu = gpu_ctx.dense_evaluate(d["poses_lw_init"], prm)
    exp = g["systems"]
    assert np.array_equal(s_gpu[:, 28], exp[:, 28])
    assert np.abs(s_gpu - exp).max() < 1e-8 * np.abs(exp).max()


def test_ceres_bridge_serves_the_device_rows_through_the_ceres_surface(gpu_ctx, oracle):
    """(b) boundary: include/panovlm_b200_ceres_adapter.hpp driven the way Ceres drives it (tests/adapter_harness.cpp): AddBlocks registers one SizedCostFunction per
    block on the callers' pose lists with loss == nullptr, PrepareForEvaluation launches ONE device evaluation, every CostFunction::Evaluate then returns the row the
    kernel wrote (loss-corrected), null Jacobian blocks are skipped.  Compared with the oracle's one-functor-at-a-time evaluation."""
    import ctypes as C
    import cases
    from conftest import build_adapter_harness
    L = C.CDLL(build_adapter_harness())
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)  # noqa: E731
    c = cases.random_blocks(5, 3000)
    n = len(c["type"])
    blk = oracle.Blocks(c["type"], c["ref"], c["nei"], c["consts"], c["huber"], c["normalize"])
    r_o, J_o, _ = blk.evaluate(c["poses"], apply_loss=True)
    keep = [np.ascontiguousarray(c[k], np.int32) for k in ("type", "ref", "nei", "normalize")] + [np.ascontiguousarray(c[k], np.float64) for k in ("huber", "consts", "poses")]
    for null_block in (0, 1):
        r, J = np.zeros(n), np.zeros((n, 12))
        rc = L.adapter_run(0, C.c_long(n), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]), p(keep[4]), p(keep[5]), C.c_int(c["nb"]), p(keep[6]), C.c_int(null_block), p(r), p(J))
        assert rc == 0, rc
        assert (np.abs(r - r_o) / np.maximum(1e-9, np.abs(r_o))).max() < 1e-5
        J_exp = J_o.copy()
        if null_block:
            for i in range(n):
                J_exp[i, 3 * (i % 4):3 * (i % 4) + 3] = 0
        assert (np.abs(J - J_exp).max(1) / np.maximum(1e-9, np.abs(J_o).max(1))).max() < 1e-6
