"""The oracle (and the product's host-side math) against the REFERENCE'S OWN source code.

`make -C oracle ref` compiles base/CostFunction.h, base/Geometry.hpp, base/Math.h and sensors/Equirectangular.{h,cpp} where they lie under
/root/reference into oracle/_ref/libpvo_ref_path.so; Eigen / Ceres / OpenCV / PCL / glog are not installed here, so oracle/shim/ provides stand-ins
for the value types those files use.  What that pins is the reference's formulas, branch thresholds, argument order and constructor
normalisations - not Eigen's or Ceres' arithmetic.  The outputs are committed as tests/golden/ref_functors.npz and ref_geometry.npz
(tests/make_golden.py: golden_ref_path), which is what these tests read, so they run wherever the repository goes; when the library itself is present
the comparison is repeated live on fresh random inputs."""
import ctypes as C
import os

import numpy as np
import pytest

import cases

G = os.path.join(os.path.dirname(__file__), "golden")
p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731


def _oracle_blocks(oracle, c, consts):
    b = oracle.Blocks(c["type"], c["ref"], c["nei"], consts, 0.0, c["normalize"])
    return b.evaluate(c["poses"], apply_loss=False)


def test_oracle_functors_equal_the_reference_functors(oracle):
    """All nine functors of the path (F1-F3, F5, F6): residual and Jacobian of the oracle's restatement == the reference's own templates run through
    ceres::CostFunction::Evaluate (Jet<12> / Jet<6>).  The gate is 1e-12 relative; measured: bit-identical."""
    g = np.load(os.path.join(G, "ref_functors.npz"))
    assert set(np.unique(g["type"]).tolist()) == set(range(9))
    r, J, _ = _oracle_blocks(oracle, g, g["consts"])
    assert (np.abs(r - g["residual"]) / np.maximum(1e-9, np.abs(g["residual"]))).max() < 1e-12
    assert (np.abs(J - g["jacobian"]).max(1) / np.maximum(1e-9, np.abs(g["jacobian"]).max(1))).max() < 1e-12
    # the branches the reference has must be present in the fixture: zero residuals of the angle types (< 1e-3 => 0) and of the IOU hinge
    assert np.any((g["residual"] == 0) & (g["type"] == 8)) and np.any((g["residual"] == 0) & np.isin(g["type"], (5, 7)))
    if oracle.ref_path_lib() is not None:                                  # live, fresh inputs
        c = cases.ref_functor_cases(99, 4000)
        r_ref, J_ref, k_ref = oracle.ref_eval_functors(c["type"], c["normalize"], c["raw"], c["params"])
        r, J, _ = _oracle_blocks(oracle, c, k_ref)
        assert (np.abs(r - r_ref) / np.maximum(1e-9, np.abs(r_ref))).max() < 1e-12
        J_ref = cases.ref_jacobian_to_block_layout(c["type"], J_ref)
        assert (np.abs(J - J_ref).max(1) / np.maximum(1e-9, np.abs(J_ref).max(1))).max() < 1e-12


def test_constructor_normalisations_equal_the_reference(oracle):
    """The constants a block carries are the functor's members AFTER its constructor ran (plane_ref.normalize(), (a - b).normalized(),
    plane / |n|, the float32 half arc of PlaneRelativeIOUResidual).  The fixture holds the reference's; rebuild them here from the raw arguments."""
    g = np.load(os.path.join(G, "ref_functors.npz"))
    raw, k, t = g["raw"], g["consts"], g["type"]
    unit = lambda v: v / np.linalg.norm(v, axis=1, keepdims=True)  # noqa: E731
    m = np.isin(t, (0, 1))
    assert np.array_equal(k[m, :8], raw[m, :8])                                                            # no normalisation: the caller's job
    m = np.isin(t, (2, 3))
    assert np.array_equal(k[m, :6], raw[m, :6]) and np.abs(k[m, 6:9] - unit(raw[m, 3:6] - raw[m, 6:9])).max() < 1e-15 and np.array_equal(k[m, 9], raw[m, 9])
    m = np.isin(t, (4, 6))
    assert np.abs(k[m, :3] - unit(raw[m, :3])).max() < 1e-15 and np.array_equal(k[m, 3:10], raw[m, 3:10])
    m = t == 5
    assert np.abs(k[m, :4] - raw[m, :4] / np.linalg.norm(raw[m, :3], axis=1, keepdims=True)).max() < 1e-15 and np.array_equal(k[m, 4:12], raw[m, 4:12])
    m = t == 8
    assert np.abs(k[m, :3] - unit(raw[m, :3])).max() < 1e-15 and np.abs(k[m, 3:6] - unit(raw[m, 3:6])).max() < 1e-15
    m = t == 7                                                                                             # float32 constructor path (:518-529)
    s, e = raw[m, 7:10].astype(np.float32), raw[m, 10:13].astype(np.float32)
    cosang = (s * e).sum(1, dtype=np.float32) / (np.sqrt((s * s).sum(1, dtype=np.float32)) * np.sqrt((e * e).sum(1, dtype=np.float32)))
    assert np.abs(k[m, 10] - np.arccos(cosang.astype(np.float32)).astype(np.float32) / np.float32(2)).max() < 1e-6
    assert np.array_equal(k[m, 7:10], ((s + e) / np.float32(2)).astype(np.float64)) and np.array_equal(k[m, 11], raw[m, 13])


def test_pairwise_and_reprojection_functors_equal_the_reference(oracle):
    """F4 (PairWisePoint2Plane_Meter / PairWisePoint2Line_Meter: one relative pose, P = R p + t) is the 4-block functor with the neighbour block at
    identity - not bit-identical (the 4-block form goes through RotationMatrixToAngleAxis), gate 1e-9; PanoramaReprojResidual_1Angle against the
    oracle's Jet<9> evaluation; PlaneIOUResidual's camera-LiDAR constructor against the LiDAR-LiDAR one fed with its documented constants."""
    g = np.load(os.path.join(G, "ref_geometry.npz"))
    t, raw, prm, r_ref, J_ref, k_ref = (g[x] for x in ("pw_type", "pw_raw", "pw_params", "pw_residual", "pw_jacobian", "pw_consts"))
    for ref_type, block_type in ((oracle.REF_PAIRWISE_P2PLANE, oracle.P2PLANE_METER), (oracle.REF_PAIRWISE_P2LINE, oracle.P2LINE_METER)):
        m = t == ref_type
        n = int(m.sum())
        poses = np.zeros((n + 1, 6)); poses[:n] = prm[m, :6]
        b = oracle.Blocks(np.full(n, block_type), np.arange(n), n, k_ref[m], 0.0, 0)
        r, J, _ = b.evaluate(poses, apply_loss=False)
        assert (np.abs(r - r_ref[m]) / np.maximum(1e-9, np.abs(r_ref[m]))).max() < 1e-9
        assert (np.abs(J[:, :6] - J_ref[m, :6]).max(1) / np.maximum(1e-9, np.abs(J_ref[m, :6]).max(1))).max() < 1e-7
    m = t == oracle.REF_REPROJ_1ANGLE
    n = int(m.sum())
    rp = oracle.Reproj(np.arange(n), np.arange(n), k_ref[m, :3], weight=1.0)
    r, J = rp.evaluate(prm[m, :6], prm[m, 6:9], apply_loss=False)[:2]
    w = k_ref[m, 3]
    assert (np.abs(r * w - r_ref[m]) / np.maximum(1e-9, np.abs(r_ref[m]))).max() < 1e-12
    assert (np.abs(J[:, :9] * w[:, None] - J_ref[m, :9]).max(1) / np.maximum(1e-9, np.abs(J_ref[m, :9]).max(1))).max() < 1e-12
    m = t == oracle.REF_PLANE_IOU_CAMERA                                     # CostFunction.h:442-451
    s, e = raw[m, 7:10], raw[m, 10:13]
    ang = np.arccos(np.clip((s * e).sum(1) / np.linalg.norm(s, axis=1) / np.linalg.norm(e, axis=1), -1, 1)) / 2
    assert np.abs(k_ref[m, 10] - ang).max() < 1e-12 and np.abs(k_ref[m, 7:10] - (s + e) / 2).max() < 1e-15
    assert np.abs(k_ref[m, :4] - raw[m, :4] / np.linalg.norm(raw[m, :3], axis=1, keepdims=True)).max() < 1e-15
    n = int(m.sum())
    poses = np.concatenate([prm[m, :6], prm[m, 6:]])
    b = oracle.Blocks(np.full(n, oracle.PLANE_IOU), np.arange(n), n + np.arange(n), k_ref[m], 0.0, 0)
    r, J, _ = b.evaluate(poses, apply_loss=False)
    assert (np.abs(r - r_ref[m]) / np.maximum(1e-9, np.abs(r_ref[m]))).max() < 1e-12
    assert (np.abs(J - J_ref[m]).max(1) / np.maximum(1e-9, np.abs(J_ref[m]).max(1))).max() < 1e-12


def test_geometry_helpers_equal_the_reference(oracle):
    """base/Geometry.hpp: FormPlane (3 points and least squares + tolerance), FormLine (PCA + ratio / distance tests), PointToLineDistance3D,
    PointToPlaneDistance, ProjectPointToPlane, VectorAngle3D, PlaneAngle, SlerpPose.  Accept / reject decisions must be identical; values agree to
    1e-9 (the QR / eigen solver / quaternion arithmetic behind them is the shim's in the fixture and the oracle's own here)."""
    g = np.load(os.path.join(G, "ref_geometry.npz"))
    n_rej_p = n_rej_l = 0
    for i in range(len(g["fp_counts"])):
        pts = g["fp_points"][i, :g["fp_counts"][i]]
        pl = oracle.form_plane(pts, float(g["fp_tol"][i]))
        ref = g["fp_plane"][i]
        if i % 4 == 3:
            continue                      # exactly collinear points: a rank-deficient system, the answer depends on the QR's rank decision
        assert np.all(ref == 0) == np.all(pl == 0), i
        n_rej_p += int(np.all(ref == 0))
        assert np.abs(pl - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), i
        ok, ln = oracle.form_line(pts, float(g["fl_tol"][i]), float(g["fl_thr"][i]))
        ref = g["fl_line"][i]
        assert ok == bool(np.any(ref != 0)), i
        n_rej_l += int(not ok)
        if ok:
            sgn = np.sign(np.dot(ln[3:], ref[3:]))                # an eigenvector's sign is the solver's choice
            assert np.abs(ln[:3] - ref[:3]).max() < 1e-12 and np.abs(sgn * ln[3:] - ref[3:]).max() < 1e-9, i
    assert n_rej_p > 20 and n_rej_l > 20                          # both outcomes are exercised
    for i in range(len(g["g_in"])):
        q = g["g_in"][i]
        out, proj = oracle.geometry_helpers(q[:3], q[3:9], q[9:13])
        assert np.abs(out - g["g_scalar"][i]).max() <= 1e-13 * max(1.0, np.abs(g["g_scalar"][i]).max()), i
        assert np.abs(proj - g["g_project"][i]).max() < 1e-12, i
        assert np.abs(oracle.form_plane3(q[:3], q[3:6], q[6:9]) - g["g_plane3"][i]).max() < 1e-12 * max(1.0, np.abs(g["g_plane3"][i]).max()), i
    for i in range(len(g["slerp_ratio"])):
        out = oracle.slerp_pose(g["slerp_w1"][i], g["slerp_w2"][i], float(g["slerp_ratio"][i]))
        assert np.abs(out - g["slerp_out"][i]).max() < 1e-12, i


def test_equirectangular_projection_equals_the_reference(oracle):
    """sensors/Equirectangular.h / .cpp: CamToImage in float32 (the bulk path: FastAtan2, float arithmetic) must be BIT-IDENTICAL; the double paths
    (cv::Point and Eigen overloads) and ImageToCam agree to the last bits; BreakToSegments gives the same polyline incl. the seam split."""
    g = np.load(os.path.join(G, "ref_geometry.npz"))
    rows, cols = int(g["rows"]), int(g["cols"])
    assert np.array_equal(oracle.cam_to_image(rows, cols, g["cam_f"]), g["px_f"])
    assert np.array_equal(oracle.cam_to_image(rows, cols, g["cam_d"]), g["px_d"])
    assert np.abs(g["px_eigen_d"] - g["px_d"]).max() < 1e-9                                  # the two overloads of the reference agree with each other
    assert np.abs(oracle.image_to_cam(rows, cols, g["pix"]) - g["i2c_d"]).max() < 1e-15
    assert np.abs(g["i2c_eigen_d"] - g["i2c_d"]).max() < 1e-15
    assert np.array_equal(oracle.image_to_cam_f(rows, cols, g["pix_f"], 5.0), g["i2c_f5"])
    n_seam = 0
    for i in range(len(g["bts_lines"])):
        ref = g["bts_xy"][g["bts_off"][i]:g["bts_off"][i + 1]]
        got = oracle.break_to_segments(rows, cols, g["bts_lines"][i], float(g["bts_seg_len"][i]))
        assert got.shape == ref.shape and np.array_equal(got, ref), i
        n_seam += int(np.any(np.abs(np.diff(ref[:, 0])) > 0.8 * cols))
    assert n_seam > 10
    if oracle.ref_path_lib() is not None:
        L = oracle.ref_path_lib()
        rng = np.random.default_rng(5)
        cam = rng.normal(0, 10, (500_000, 3)).astype(np.float32)
        px = np.zeros((len(cam), 2), np.float32)
        L.ref_cam_to_image_f(C.c_int(rows), C.c_int(cols), C.c_long(len(cam)), p(cam), p(px))
        assert np.array_equal(oracle.cam_to_image(rows, cols, cam), px)


def test_product_host_math_equals_the_reference_functors(harness):
    """The product's analytic residual / Jacobian code (pvb_math.cuh compiled for the host) against the reference-compiled fixture directly:
    BASELINE.json's gates (1e-5 residual, 1e-6 Jacobian; measured ~1e-9)."""
    g = np.load(os.path.join(G, "ref_functors.npz"))
    n = len(g["type"])
    r, J, cost = np.zeros(n), np.zeros((n, 12)), np.zeros(n)
    harness.pvbh_eval_blocks(C.c_long(n), p(g["type"]), p(g["ref"]), p(g["nei"]), p(g["normalize"]), p(np.zeros(n)), p(np.ascontiguousarray(g["consts"])),
                             p(np.ascontiguousarray(g["poses"])), C.c_int(int(g["nb"])), C.c_int(0), p(r), p(J), p(cost))
    assert (np.abs(r - g["residual"]) / np.maximum(1e-9, np.abs(g["residual"]))).max() < 1e-8
    keep = np.abs(g["residual"]) > 1e-12                  # at r == 0 the sign of abs' is a subgradient
    assert (np.abs(J - g["jacobian"]).max(1) / np.maximum(1e-9, np.abs(g["jacobian"]).max(1)))[keep].max() < 1e-7


def test_product_builders_carry_the_reference_constructor_constants(oracle):
    """pvb_build_point2line_blocks / pvb_build_camera_lidar_blocks / pvb_build_calibration_blocks against constants produced by the reference's
    constructors (fixture) and by the reference construction sequence of util/Optimization.cpp:585-600 / CameraLidarOptimizer.cpp:45-63 replayed
    with the reference-checked primitives."""
    from panovlm_b200 import BlockList, Context
    g = np.load(os.path.join(G, "ref_functors.npz"))
    m = np.isin(g["type"], (2, 3))
    raw, k = g["raw"][m], g["consts"][m]
    bl = BlockList(len(raw) + 1)
    Context.build_point2line_blocks(bl, raw[:, :3], raw[:, 3:6], raw[:, 6:9], 0, 1, False, False, 1.0)
    v = bl.view()
    assert np.array_equal(v["consts"][:, :6], k[:, :6]) and np.abs(v["consts"][:, 6:9] - k[:, 6:9]).max() < 1e-15
    m = np.isin(g["type"], (4, 5))
    if oracle.ref_path_lib() is None:
        return
    # B4 with the reference's own constructors: plane through the two image rays, Plane2Plane_Global(plane.head(3), end, start, w_pair * w),
    # PlaneIOUResidual(plane, (end + start) / 2, (p1 + p2) / 2, angle(p1, p2), 2 w)
    rows, cols, n = 2880, 5760, 60
    rng = np.random.default_rng(31)
    lines = np.stack([rng.uniform(0, cols, n), rng.uniform(0, rows, n), rng.uniform(0, cols, n), rng.uniform(0, rows, n)], axis=1).astype(np.float32)
    start, end = rng.normal(0, 3, (n, 3)), rng.normal(0, 3, (n, 3))
    pw = rng.uniform(0.5, 2, n).astype(np.float32)
    L = oracle.ref_path_lib()
    px = lines.astype(np.float64).reshape(-1, 2)
    cam = np.zeros((2 * n, 3))
    L.ref_image_to_cam_eigen_d(C.c_int(rows), C.c_int(cols), C.c_long(2 * n), p(px), C.c_double(1.0), p(cam))
    p1, p2 = cam[0::2], cam[1::2]
    raw = np.zeros((2 * n, 16)); typ = np.tile([4, 5], n).astype(np.int32)
    for i in range(n):
        plane = np.zeros(4)
        L.ref_form_plane3(p(np.ascontiguousarray(p1[i])), p(np.ascontiguousarray(p2[i])), p(np.zeros(3)), p(plane))
        ang = L.ref_vector_angle3d(p(np.ascontiguousarray(p1[i])), p(np.ascontiguousarray(p2[i])), C.c_int(1))
        raw[2 * i, :3] = plane[:3]; raw[2 * i, 3:6] = end[i]; raw[2 * i, 6:9] = start[i]; raw[2 * i, 9] = float(pw[i]) * 25.0
        raw[2 * i + 1, :4] = plane; raw[2 * i + 1, 4:7] = (end[i] + start[i]) / 2.0; raw[2 * i + 1, 7:10] = (p1[i] + p2[i]) / 2.0; raw[2 * i + 1, 10] = ang; raw[2 * i + 1, 11] = 2.0 * 25.0
    _, _, k_ref = oracle.ref_eval_functors(typ, 0, raw, np.zeros((2 * n, 12)), jac=False)
    bl = BlockList(2 * n + 1)
    Context.build_camera_lidar_blocks(bl, rows, cols, lines, start, end, pw, 0, 1, 25.0)
    v = bl.view()
    assert np.array_equal(v["type"], typ)
    assert np.abs(v["consts"] - k_ref).max() < 1e-12


# ------------------------------------------------------------------------------------------------------------------------------------------------
# The association layer: lidar_mapping/LidarFeatureAssociate.cpp compiled where it lies (oracle/ref_assoc_wrap.cpp, oracle/_ref/libpvo_ref_assoc.so)
# ------------------------------------------------------------------------------------------------------------------------------------------------
def _world(oracle, F, key, R=None, t=None):
    return oracle.transform_cloud(F["R_wl"] if R is None else R, F["t_wl"] if t is None else t, F[key])


def assoc_case(oracle, seed, n_az=900, perturb=0.0):
    """A synthetic 2-frame pair as the association functions see it; `perturb` moves the neighbour's pose estimate off the truth; perturb < 0: the
    estimate is the identity (the initial guess of BASELINE.json configs[0]) - exactly representable, so the kernel tests can use the same case."""
    from panovlm_b200 import synth
    from scipy.spatial.transform import Rotation
    A, B = synth.make_pair(seed=seed, n_az=n_az, ground_class=True)
    rng = np.random.default_rng(seed)
    RB = Rotation.from_rotvec(rng.normal(0, abs(perturb) * 0.2, 3)).as_matrix() @ B["R_wl"]
    tB = B["t_wl"] + rng.normal(0, abs(perturb), 3)
    if perturb < 0:
        RB, tB = np.eye(3), np.zeros(3)
    return A, B, RB, tB


def oracle_associations(oracle, A, B, RB, tB, plane_tol, thr_plane, thr_line):
    """Every association function of LidarFeatureAssociate.cpp through the ORACLE, flattened into a dict of arrays."""
    out = {}
    ref_sl, nei_sf = _world(oracle, A, "surfLessFlat"), _world(oracle, B, "surfFlat", RB, tB)
    ref_c, nei_c = _world(oracle, A, "cornerLessSharp"), _world(oracle, B, "cornerLessSharp", RB, tB)
    ref_lw, nei_lw = oracle.transform_lines(A["R_wl"], A["t_wl"], A["segment_coeffs"]), oracle.transform_lines(RB, tB, B["segment_coeffs"])
    S_a, S_b = len(A["segment_coeffs"]), len(B["segment_coeffs"])
    _, out["p2plane_point"], out["p2plane_plane"] = oracle.associate_p2plane(ref_sl, A["R_wl"], A["t_wl"], nei_sf, RB, tB, plane_tol, thr_plane, 10, True)
    M = oracle.line_votes(ref_lw, nei_c, B["p2s_off"], B["p2s_ids"], S_b, thr_line)
    out["l2l_nei"], out["l2l_ref"], out["l2l_a"], out["l2l_b"] = oracle.find_associations(A["segment_coeffs"], ref_lw, nei_lw, np.diff(B["seg_off"]), M)
    M = oracle.line2line_knn_votes(ref_c, A["p2s_off"], A["p2s_ids"], S_a, nei_c, B["p2s_off"], B["p2s_ids"], S_b, thr_line)
    out["l2lknn_nei"], out["l2lknn_ref"], out["l2lknn_a"], out["l2lknn_b"] = oracle.find_associations(A["segment_coeffs"], ref_lw, nei_lw, np.diff(B["seg_off"]), M)
    _, out["p2l_point"], out["p2l_a"], out["p2l_b"] = oracle.associate_p2line(ref_c, A["R_wl"], A["t_wl"], nei_c, RB, tB, thr_line)
    _, _, out["p2lsk_point"], out["p2lsk_a"], out["p2lsk_b"] = oracle.associate_p2line_segment_knn(ref_c, A["p2s_off"], A["p2s_ids"], A["segment_coeffs"], nei_c, RB, tB, thr_line)
    _, _, out["p2ls_point"], out["p2ls_a"], out["p2ls_b"] = oracle.associate_p2line_segment(ref_lw, A["segment_coeffs"], nei_c, RB, tB, thr_line)
    return out


def reference_associations(oracle, A, B, RB, tB, plane_tol, thr_plane, thr_line):
    """The same through the reference's own functions (needs oracle/_ref/libpvo_ref_assoc.so)."""
    # sensor-frame clouds: the reference's own Velodyne::Transform2LidarWorld() brings them to the world frame (T1)
    fa = oracle.RefFrame(A["R_wl"], A["t_wl"], A["cornerLessSharp"], A["p2s_off"], A["p2s_ids"], A["segment_coeffs"], A["surfFlat"], A["surfLessFlat"], id=0, local=True)
    fb = oracle.RefFrame(RB, tB, B["cornerLessSharp"], B["p2s_off"], B["p2s_ids"], B["segment_coeffs"], B["surfFlat"], B["surfLessFlat"], id=1, local=True)
    out = {}
    out["p2plane_point"], out["p2plane_plane"] = oracle.ref_associate_point2plane(fa, fb, plane_tol, thr_plane)
    out["l2l_nei"], out["l2l_ref"], out["l2l_a"], out["l2l_b"] = oracle.ref_associate_line2line(fa, fb, thr_line)
    out["l2lknn_nei"], out["l2lknn_ref"], out["l2lknn_a"], out["l2lknn_b"] = oracle.ref_associate_line2line(fa, fb, thr_line, knn=True)
    out["p2l_point"], out["p2l_a"], out["p2l_b"] = oracle.ref_associate_point2line(fa, fb, thr_line)
    out["p2lsk_point"], out["p2lsk_a"], out["p2lsk_b"] = oracle.ref_associate_point2line(fa, fb, thr_line, "_segment_knn")
    out["p2ls_point"], out["p2ls_a"], out["p2ls_b"] = oracle.ref_associate_point2line(fa, fb, thr_line, "_segment")
    return out


ASSOC_CASES = [(20261021, 900, 0.0, 0.05, 1.0, 0.3), (20261022, 900, 0.02, 0.05, 1.0, 0.4), (20261023, 1800, 0.05, 0.01, 0.7, 0.3), (20261024, 900, 0.15, 0.05, 0.5, 0.6), (20261025, 900, -1.0, 0.05, 1.0, 0.3)]


def _compare_associations(got, exp, tag):
    for k in exp:
        assert got[k].shape == exp[k].shape, (tag, k, got[k].shape, exp[k].shape)
        if exp[k].dtype.kind == "i" or k.endswith("_point"):
            assert np.array_equal(got[k], exp[k]), (tag, k)                  # index lists and the queries (World2Local of a float32 point): exact
        elif k in ("p2l_a", "p2l_b"):
            # PCA line through 5 neighbours: centre +- 0.1 * direction, the eigenvector's sign is the solver's choice => compare as an unordered pair
            pass
        else:
            assert np.abs(got[k] - exp[k]).max() < 1e-9, (tag, k)
    a, b, ea, eb = got["p2l_a"], got["p2l_b"], exp["p2l_a"], exp["p2l_b"]
    if len(a):
        same = np.maximum(np.abs(a - ea).max(1), np.abs(b - eb).max(1))
        flip = np.maximum(np.abs(a - eb).max(1), np.abs(b - ea).max(1))
        assert np.minimum(same, flip).max() < 1e-9, (tag, "p2l line")


def test_oracle_associations_equal_the_reference_associations(oracle):
    """A1 / A2 / A3 (all six association functions) on four synthetic pairs: the oracle returns the same correspondences in the same order as the
    reference's own code; planes / lines agree to 1e-9.  Reads the committed fixture; repeats the comparison live when oracle/_ref is present."""
    g = np.load(os.path.join(G, "ref_assoc.npz"))
    total = {}
    for ci, (seed, n_az, perturb, tol, thr_p, thr_l) in enumerate(ASSOC_CASES):
        A, B, RB, tB = assoc_case(oracle, seed, n_az, perturb)
        got = oracle_associations(oracle, A, B, RB, tB, tol, thr_p, thr_l)
        exp = {k[len(f"c{ci}_"):]: g[k] for k in g.files if k.startswith(f"c{ci}_")}
        assert set(exp) == set(got)
        _compare_associations(got, exp, f"fixture case {ci}")
        for k in got:
            total[k] = total.get(k, 0) + len(got[k])
        if oracle.ref_assoc_lib() is not None:
            _compare_associations(got, reference_associations(oracle, A, B, RB, tB, tol, thr_p, thr_l), f"live case {ci}")
    assert total["p2plane_point"] > 500 and total["l2l_nei"] > 20 and total["l2lknn_nei"] > 10 and total["p2l_point"] > 100 and total["p2lsk_point"] > 100 and total["p2ls_point"] > 100


def test_find_neighbors_and_transform_lines_equal_the_reference(oracle):
    """N (FindNeighbors: k-NN over frame centres, forced previous / next valid frame, loop candidates, frames without a pose) and TransformLines."""
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_assoc.npz"))
    for ci in range(int(g["fn_cases"])):
        t, pv, va, k = g[f"fn{ci}_t"], g[f"fn{ci}_pose_valid"], g[f"fn{ci}_valid"], int(g[f"fn{ci}_k"])
        exp = [g[f"fn{ci}_ids"][g[f"fn{ci}_off"][i]:g[f"fn{ci}_off"][i + 1]].tolist() for i in range(len(t))]
        got = Context.find_neighbors(t, pv, va, k)
        assert [list(x) for x in got] == exp, ci
        if oracle.ref_assoc_lib() is not None:
            R = np.tile(np.eye(3).reshape(1, 9), (len(t), 1))
            assert oracle.ref_find_neighbors(R, t, pv, va, k) == exp
    assert np.abs(oracle.transform_lines(g["tl_T"][:3, :3], g["tl_T"][:3, 3], g["tl_in"]) - g["tl_out"]).max() < 1e-12


# ------------------------------------------------------------------------------------------------------------------------------------------------
# Camera-LiDAR line association (A4) and the depth splat (P): joint_optimization/CameraLidarLineAssociate.cpp and util/Visualization.h compiled where
# they lie (oracle/ref_camlidar_wrap.cpp, oracle/_ref/libpvo_ref_camlidar.so)
# ------------------------------------------------------------------------------------------------------------------------------------------------
def camlidar_case(oracle_mod=None):
    """One frame with LiDAR segments, image lines = projections of the segments' end points + noise, clutter lines and near-duplicates (so that
    UniqueLinePair has conflicts to resolve), a T_cl off the identity, and masks that exclude one image line and one segment."""
    from panovlm_b200 import synth
    from scipy.spatial.transform import Rotation
    A, _ = synth.make_pair(seed=20261030, n_az=1800, ground_class=True)
    rows, cols = 2880, 5760
    T = np.eye(4); T[:3, :3] = Rotation.from_rotvec([0.01, 0.02, -0.01]).as_matrix(); T[:3, 3] = [0.03, -0.05, 0.02]
    rng = np.random.default_rng(9)
    ends_cam = A["end_points"].reshape(-1, 3) @ T[:3, :3].T + T[:3, 3]
    lon = np.arctan2(ends_cam[:, 0], ends_cam[:, 2]); lat = -np.arctan2(ends_cam[:, 1], np.hypot(ends_cam[:, 0], ends_cam[:, 2]))
    px = np.stack([cols * (0.5 + lon / (2 * np.pi)), rows * (0.5 - lat / np.pi)], axis=1).reshape(-1, 4) + rng.normal(0, 2, (len(A["end_points"]), 4))
    clutter = np.stack([rng.uniform(0, cols, 30), rng.uniform(0, rows, 30), rng.uniform(0, cols, 30), rng.uniform(0, rows, 30)], axis=1)
    lines = np.concatenate([px, clutter, px + 1.5]).astype(np.float32)
    return A, rows, cols, T, lines


CAMLIDAR_VARIANTS = [("multi", True, False), ("unique", False, False), ("masked", True, True)]


def _camlidar_masks(g, n_lines, n_seg):
    im = np.ones(n_lines, np.uint8); im[int(g["multi_image"][0])] = 0
    lm = np.ones(n_seg, np.uint8); lm[int(g["multi_lidar"][-1])] = 0
    return im, lm


def test_camera_lidar_association_and_depth_splat_equal_the_reference(oracle):
    """A4: AssociateByAngle + Filter(false, true) + UniqueLinePair of the reference == the oracle: same (image line, LiDAR segment) pairs in the same order,
    float32 scores bit-identical, end points to 1e-12; masks honoured.  P: ProjectLidar2PanoramaDepth gives the identical uint16 image (last writer wins)."""
    g = np.load(os.path.join(G, "ref_camlidar.npz"))
    A, rows, cols, T, lines = camlidar_case()
    assert np.array_equal(lines, g["lines"])
    sizes = np.diff(A["seg_off"])
    for name, multi, masked in CAMLIDAR_VARIANTS:
        im, lm = _camlidar_masks(g, len(lines), len(sizes)) if masked else (None, None)
        got = oracle.associate_by_angle(rows, cols, lines, A["cornerLessSharp"], A["p2s_off"], A["p2s_ids"], sizes, A["end_points"], T, True, multi, im, lm)
        assert np.array_equal(got[0], g[name + "_image"]) and np.array_equal(got[1], g[name + "_lidar"]) and np.array_equal(got[4], g[name + "_score"]), name
        assert np.abs(got[2] - g[name + "_start"]).max() < 1e-12 and np.abs(got[3] - g[name + "_end"]).max() < 1e-12, name
        if oracle.ref_camlidar_lib() is not None:
            ref = oracle.ref_associate_by_angle(rows, cols, lines, A["cornerLessSharp"], A["p2s_off"], A["p2s_ids"], A["segment_coeffs"], A["end_points"], T, multi, im, lm)
            assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]) and np.array_equal(ref[4], got[4]), name
    assert len(g["unique_image"]) < len(g["multi_image"]) and len(g["masked_image"]) < len(g["multi_image"]) and len(g["unique_image"]) >= 10
    for key, r_, c_, size in (("depth_720", 720, 1440, 3), ("depth_360", 360, 720, 4)):
        img, _ = oracle.project_depth(A["cloud"], r_, c_, T, size=size)
        assert np.array_equal(img, g[key]) and (img > 0).sum() > 20000, key
    if oracle.ref_camlidar_lib() is not None:
        img, _ = oracle.project_depth(A["cloud"], rows, cols, T, size=3)
        assert np.array_equal(img, oracle.ref_project_depth(A["cloud"], rows, cols, T, 3))


# ---- LiDAR line tracks: lidar_mapping/LidarLineMatch.cpp + util/Tracks.cpp (GenerateTracks = FindNeighbors + AssociateLine2Line + TrackBuilder) ----
TRACK_CASES = [(8, 600, 4, 3, None), (8, 600, 4, 2, 2), (12, 600, 2, 3, None), (12, 600, 6, 4, 5)]      # frames, n_az, neighbor_size, min_track_length, frame without pose


def track_case(n_frames, n_az):
    from panovlm_b200 import synth
    return synth.make_sequence(n_frames, n_az=n_az)


def oracle_line_tracks(oracle, frames, neighbor_size, min_len, no_pose, builder=None):
    """The oracle's pipeline for GenerateTracks; `builder` swaps the union-find stage (the product's host pvb_line_tracks_build)."""
    from panovlm_b200 import Context
    n = len(frames)
    R_wl, t_wl = [f["R_wl"] for f in frames], np.array([f["t_wl"] for f in frames])
    pv = np.ones(n, np.uint8)
    if no_pose is not None:
        pv[no_pose] = 0
    nbrs = Context.find_neighbors(t_wl, pv, None, neighbor_size)
    corner_w = [oracle.transform_cloud(R_wl[i], t_wl[i], f["cornerLessSharp"]) for i, f in enumerate(frames)]
    lines_w = [oracle.transform_lines(R_wl[i], t_wl[i], f["segment_coeffs"]) for i, f in enumerate(frames)]
    pa, pb, off, ma, mb = [], [], [0], [], []
    for i in range(n):
        if not pv[i]:
            continue
        for nb in nbrs[i]:
            if not (0 <= nb < n) or not pv[nb]:
                continue                                   # the synthetic cases never pair a valid frame with a pose-less neighbour's clouds
            M = oracle.line_votes(lines_w[nb], corner_w[i], frames[i]["p2s_off"], frames[i]["p2s_ids"], len(frames[i]["segment_coeffs"]), 0.3)
            on, orf, _, _ = oracle.find_associations(frames[nb]["segment_coeffs"], lines_w[nb], lines_w[i], np.diff(frames[i]["seg_off"]), M)
            for x, y in sorted(set(zip(on.tolist(), orf.tolist()))):
                ma.append(x); mb.append(y)
            pa.append(i); pb.append(nb); off.append(len(ma))
    return (builder or oracle.line_tracks)(pa, pb, off, ma, mb, min_len, True)


def reference_line_tracks(oracle, frames, neighbor_size, min_len, no_pose):
    rf = []
    for i, f in enumerate(frames):
        rf.append(oracle.RefFrame(f["R_wl"], f["t_wl"], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], id=i, pose_valid=(i != no_pose), local=True))
    return oracle.ref_generate_line_tracks(rf, neighbor_size, min_len)


def test_line_tracks_equal_the_reference_generate_tracks(oracle):
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_assoc.npz"))
    for ci, (nf, n_az, k, min_len, no_pose) in enumerate(TRACK_CASES):
        frames = track_case(nf, n_az)
        exp = [g[f"tr{ci}_feat"][g[f"tr{ci}_off"][t]:g[f"tr{ci}_off"][t + 1]] for t in range(len(g[f"tr{ci}_off"]) - 1)]
        assert len(exp) >= 5
        for builder in (None, Context.line_tracks_build):
            got = oracle_line_tracks(oracle, frames, k, min_len, no_pose, builder)
            assert len(got) == len(exp) and all(np.array_equal(a, b) for a, b in zip(got, exp)), (ci, builder)
        if oracle.ref_assoc_lib() is not None:
            ref = reference_line_tracks(oracle, frames, k, min_len, no_pose)
            assert len(ref) == len(exp) and all(np.array_equal(a, b) for a, b in zip(ref, exp)), ci


# ---- Residual-block builders: util/Optimization.cpp (AddLidarPointToPlaneResidual, AddLidarLineToLineResidual2, AddLidarPointToLineResidual,
# AddCameraLidarResidual) compiled where it lies; ceres::Problem is a recorder (oracle/shim/pvo_shim_ceres.hpp) ----
JAC_STRIDE = 8          # the fixture keeps every 8th Jacobian row of the block lists
BUILDER_CASES = [dict(), dict(angle_residual=False, normalize_distance=False), dict(point_to_line=True, line_to_line=False, use_segment=True),
                 dict(point_to_line=True, line_to_line=False, point_to_plane=False, use_segment=False, line_dis_threshold=0.4), dict(plane_tolerance=0.01, line_dis_threshold=0.4)]


def builder_case(n_frames=6, n_az=600, seed=3):
    from panovlm_b200 import synth
    from scipy.spatial.transform import Rotation
    frames = synth.make_sequence(n_frames, n_az=n_az)
    rng = np.random.default_rng(seed)
    Rs = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.01, 3)).as_matrix() for f in frames]
    ts = [f["t_wl"] + rng.normal(0, 0.03, 3) for f in frames]
    return frames, Rs, ts


def product_refine_blocks(oracle, frames, Rs, ts, point_to_plane=True, line_to_line=True, point_to_line=False, use_segment=True, angle_residual=True,
                          normalize_distance=True, plane_dis_threshold=1.0, line_dis_threshold=0.3, plane_tolerance=0.05, plane_weight=1.0, p2l=None, block_offset=0,
                          l2l_needs_segment=True):
    """The block list of one RefinePose built from the ORACLE's associations and the PRODUCT's host builders (pvb_build_*_blocks, pvb_find_neighbors,
    pvb_line_tracks_build / pvb_line_tracks_gate), in the reference's registration order."""
    from panovlm_b200 import BlockList, Context, LineFrame
    n = len(frames)
    t_arr = np.array(ts)
    nbrs = Context.find_neighbors(t_arr, None, None, 6)
    edges = [(i, j) for i in range(n) for j in nbrs[i] if 0 <= j < n and j != i]
    corner_w = [oracle.transform_cloud(Rs[i], ts[i], f["cornerLessSharp"]) for i, f in enumerate(frames)]
    tgt_w = [oracle.transform_cloud(Rs[i], ts[i], f["surfLessFlat"]) for i, f in enumerate(frames)]
    qry_w = [oracle.transform_cloud(Rs[i], ts[i], f["surfFlat"]) for i, f in enumerate(frames)]
    lines_w = [oracle.transform_lines(Rs[i], ts[i], f["segment_coeffs"]) for i, f in enumerate(frames)]
    lf = [LineFrame(f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["end_points"], Rs[i], ts[i]) for i, f in enumerate(frames)]
    bl = BlockList(200000)

    def l2l(i, j, thr):
        fj = frames[j]
        M = oracle.line_votes(lines_w[i], corner_w[j], fj["p2s_off"], fj["p2s_ids"], len(fj["segment_coeffs"]), thr)
        return oracle.find_associations(frames[i]["segment_coeffs"], lines_w[i], lines_w[j], np.diff(fj["seg_off"]), M)
    if point_to_line:
        p_seg, p_angle, p_norm, p_w = p2l if p2l is not None else (use_segment, angle_residual, normalize_distance, 1.0)
        for (i, j) in edges:
            if abs(i - j) > 1:
                continue
            if p_seg:
                _, _, pt, a, b = oracle.associate_p2line_segment_knn(corner_w[i], frames[i]["p2s_off"], frames[i]["p2s_ids"], frames[i]["segment_coeffs"], corner_w[j], Rs[j], ts[j], line_dis_threshold)
            else:
                _, pt, a, b = oracle.associate_p2line(corner_w[i], Rs[i], ts[i], corner_w[j], Rs[j], ts[j], line_dis_threshold)
            if len(pt):
                Context.build_point2line_blocks(bl, pt, a, b, i + block_offset, j + block_offset, p_angle, p_norm, p_w)
    if line_to_line and (use_segment or not l2l_needs_segment):
        tn = Context.find_neighbors(t_arr, None, None, 4)
        pa, pb, off, ma, mb = [], [], [0], [], []
        for i in range(n):
            for nb in tn[i]:
                on, orf, _, _ = l2l(nb, i, 0.3)
                for x, y in sorted(set(zip(on.tolist(), orf.tolist()))):
                    ma.append(x); mb.append(y)
                pa.append(i); pb.append(nb); off.append(len(ma))
        tracks = Context.line_tracks_build(pa, pb, off, ma, mb, 3, True)
        for (i, j) in edges:
            on, orf, oa, ob = l2l(i, j, line_dis_threshold)
            keep = Context.line_tracks_gate(tracks, i, j, orf, on)
            assert np.array_equal(keep, oracle.line_track_gate(tracks, i, j, orf, on))
            for k in np.nonzero(keep)[0]:
                Context.build_line2line_blocks(bl, lf[j], corner_w[j], int(on[k]), oa[k], ob[k], i + block_offset, j + block_offset, angle_residual, normalize_distance, 1.0)
    if point_to_plane:
        for (i, j) in edges:
            _, pt, pl = oracle.associate_p2plane(tgt_w[i], Rs[i], ts[i], qry_w[j], Rs[j], ts[j], plane_tolerance, plane_dis_threshold, 10, True)
            if len(pt):
                Context.build_point2plane_blocks(bl, pt, pl, i + block_offset, j + block_offset, angle_residual, normalize_distance, plane_weight)
    return bl.view()


def reference_refine_blocks(oracle, frames, Rs, ts, **kw):
    rf = [oracle.RefFrame(Rs[i], ts[i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["surfFlat"], f["surfLessFlat"], id=i, local="keep")
          for i, f in enumerate(frames)]
    return oracle.ref_refine_pose_blocks(rf, **kw)


def test_residual_block_lists_equal_the_reference_builders(oracle):
    """B1 / B2 / B3: the blocks the reference's own builders register for one RefinePose (which frame pairs, which functor, which loss - nullptr for the
    angle line residuals -, which constants, in which order) == the list built by the product's host builders from the oracle's associations: same
    (reference, neighbour) frames and Huber parameters, raw residuals and Jacobians equal to 1e-9 when both lists are evaluated at the same pose blocks."""
    g = np.load(os.path.join(G, "ref_builders.npz"))
    frames, Rs, ts = builder_case()
    for ci, kw in enumerate(BUILDER_CASES):
        v = product_refine_blocks(oracle, frames, Rs, ts, **kw)
        exp = {k[len(f"b{ci}_"):]: g[k] for k in g.files if k.startswith(f"b{ci}_")}
        assert len(v["type"]) == len(exp["residual"]) > 1000, (ci, len(v["type"]), len(exp["residual"]))
        assert np.array_equal(v["ref"], exp["ref"]) and np.array_equal(v["nei"], exp["nei"]), ci
        assert np.abs(v["huber"] - exp["huber"]).max() < 1e-15, ci
        b = oracle.Blocks(v["type"], v["ref"], v["nei"], v["consts"], 0.0, v["normalize"])
        r, J, _ = b.evaluate(exp["poses"], apply_loss=False)
        # planes / lines differ by ~1e-14 between the two QR / eigen implementations; acos near 1 amplifies that by 1 / r for small angle residuals
        assert np.all(np.abs(r - exp["residual"]) <= 1e-9 * np.abs(exp["residual"]) + 1e-11), ci
        assert np.all(np.abs(J[::JAC_STRIDE] - exp["jacobian"]).max(1) <= 1e-6 * np.abs(exp["jacobian"]).max(1) + 1e-9), ci
        if oracle.ref_assoc_lib() is not None and ci == 0:
            live = reference_refine_blocks(oracle, frames, Rs, ts, **kw)
            assert np.array_equal(live["ref"], exp["ref"]) and np.array_equal(live["residual"], exp["residual"])
            # what RefinePose hands to ceres::Solve besides the blocks: SetOptionsLidar's max_num_iterations = 20 (util/Optimization.cpp:663; OdometryConfig.max_lm_iterations),
            # DENSE_SCHUR below 100 frames (:645-647), and exactly two constant parameter blocks = the first valid frame's (LidarOdometry.cpp:58-64; is_const[0] in refine_pose)
            from panovlm_b200 import odometry
            assert live["info"].tolist() == [odometry.OdometryConfig().max_lm_iterations, 3, 2, 0]


def test_camera_lidar_blocks_equal_the_reference_builder(oracle):
    """B4: AddCameraLidarResidual of the reference == pvb_build_camera_lidar_blocks (Plane2Plane_Global + PlaneIOUResidual per pair, weights, constants)."""
    from panovlm_b200 import BlockList, Context
    g = np.load(os.path.join(G, "ref_builders.npz"))
    bl = BlockList(2 * len(g["cl_lines"]) + 1)
    Context.build_camera_lidar_blocks(bl, int(g["cl_rows"]), int(g["cl_cols"]), g["cl_lines"], g["cl_start"], g["cl_end"], g["cl_pair_weight"], 0, 1, float(g["cl_weight"]))
    v = bl.view()
    assert np.allclose(v["huber"], 3 * np.pi / 180)
    b = oracle.Blocks(v["type"], v["ref"], v["nei"], v["consts"], 0.0, v["normalize"])
    r, J, _ = b.evaluate(g["cl_poses"], apply_loss=False)
    assert len(r) == len(g["cl_residual"]) == 2 * len(g["cl_lines"])
    assert (np.abs(r - g["cl_residual"]) / np.maximum(1e-6, np.abs(g["cl_residual"]))).max() < 1e-9
    assert (np.abs(J - g["cl_jacobian"]).max(1) / np.maximum(1e-6, np.abs(g["cl_jacobian"]).max(1))).max() < 1e-8
    assert (g["cl_residual"] > 0).sum() > len(r) // 2


def camera_residual_case():
    """6 panoramas (2 without a pose), 400 key points each - a share of them exactly on half-integer pixel coordinates, where the implicit
    cv::Point2f -> cv::Point2i conversion of `eq.ImageToCam(keypoint.pt)` (Optimization.cpp:203) rounds ties to even -, 150 tracks of 2-5 features."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(20261101)
    rows, cols, nf, nk = 2880, 5760, 6, 400
    R = np.stack([Rotation.from_rotvec(rng.normal(0, 0.4, 3)).as_matrix() for _ in range(nf)]); t = rng.normal(0, 2, (nf, 3))
    pv = np.ones(nf, np.uint8); pv[[2, 4]] = 0
    xy = np.stack([rng.uniform(0, cols - 1, nf * nk), rng.uniform(0, rows - 1, nf * nk)], axis=1)
    xy[::7] = np.floor(xy[::7]) + 0.5
    xy = xy.astype(np.float32)
    kp_off = np.arange(nf + 1, dtype=np.int32) * nk
    track_off, ff, fi = [0], [], []
    for _ in range(150):
        fr = np.sort(rng.choice(nf, int(rng.integers(2, 6)), replace=False))
        for f in fr:
            ff.append(int(f)); fi.append(int(rng.integers(0, nk)))
        track_off.append(len(ff))
    pts = rng.normal(0, 6, (150, 3))
    return rows, cols, R, t, pv, kp_off, xy, np.array(track_off, np.int32), np.array(ff, np.int32), np.array(fi, np.int32), pts


def test_reprojection_observations_equal_the_reference_builder(oracle):
    """(f) rank 3: AddCameraResidual of the reference (ANGLE_RESIDUAL_1) == pvb_build_reproj_observations + the oracle's PanoramaReprojResidual_1Angle:
    same (camera, track) list - features of frames without a pose skipped -, bearings from the ROUNDED pixel, raw residuals and 1x9 Jacobians."""
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_builders.npz"))
    rows, cols, R, t, pv, kp_off, xy, track_off, ff, fi, pts = camera_residual_case()
    cam, point, bearing = Context.build_reproj_observations(rows, cols, track_off, ff, xy[kp_off[ff] + fi], pv)
    assert np.array_equal(cam, g["cr_cam"]) and np.array_equal(point, g["cr_track"]) and 200 < len(cam) < len(ff)
    rp = oracle.Reproj(cam, point, bearing, weight=float(g["cr_weight"]))
    r, J = rp.evaluate(g["cr_cams"], pts, apply_loss=False)[:2]
    assert np.all(np.abs(r - g["cr_residual"]) <= 1e-9 * np.abs(g["cr_residual"]) + 1e-12)
    assert np.all(np.abs(J[:, :9] - g["cr_jacobian"]).max(1) <= 1e-7 * np.abs(g["cr_jacobian"]).max(1) + 1e-10)
    if oracle.ref_assoc_lib() is not None:
        live = oracle.ref_camera_residual_blocks(rows, cols, R, t, pv, kp_off, xy, track_off, ff, fi, pts, float(g["cr_weight"]))
        assert np.array_equal(live["cam"], cam) and np.array_equal(live["residual"], g["cr_residual"])


# ---- class Velodyne itself: sensors/Velodyne.cpp compiled where it lies (T1 Transform2LidarWorld / Transform2Local, T2 World2Local, UndistortCloud) ----
def velodyne_case():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(20261102)
    cloud = np.concatenate([rng.normal(0, 8, (6000, 3)), rng.uniform(0, 1, (6000, 1))], axis=1).astype(np.float32)
    R_wl = Rotation.from_rotvec([0.1, -0.3, 0.2]).as_matrix(); t_wl = np.array([1.0, 2.0, -0.5])
    sweeps = [(R_wl @ Rotation.from_rotvec([0.02, 0.05, -0.03]).as_matrix(), t_wl + np.array([0.3, -0.1, 0.05])),      # ordinary motion
              (R_wl, t_wl + np.array([0.2, 0.0, 0.0])),                                                               # pure translation: slerp's |d| >= 1 - eps branch
              (R_wl @ Rotation.from_rotvec([0.0, 2.9, 0.0]).as_matrix(), t_wl)]                                      # large rotation
    return cloud, R_wl, t_wl, sweeps


def test_transform_and_undistortion_equal_the_reference(oracle):
    """T1: the float32 world clouds left by the reference's Transform2LidarWorld (pcl::transformPointCloud with a double matrix) and the sensor-frame clouds
    after Transform2Local are BIT-IDENTICAL to the oracle's transform; (f) rank 4: Velodyne::UndistortCloud (Quaternion slerp per point) likewise."""
    g = np.load(os.path.join(G, "ref_velodyne.npz"))
    cloud, R_wl, t_wl, sweeps = velodyne_case()
    assert np.array_equal(oracle.transform_cloud(R_wl, t_wl, cloud), g["world"])
    assert np.array_equal(oracle.transform_cloud(R_wl.T, -R_wl.T @ t_wl, g["world"]), g["local_again"])
    for k, (R_we, t_we) in enumerate(sweeps):
        assert np.array_equal(oracle.undistort_cloud(R_wl, t_wl, R_we, t_we, cloud), g[f"undistorted{k}"]), k
    if oracle.ref_assoc_lib() is not None:
        f = oracle.RefFrame(R_wl, t_wl, surf_less_flat_world=cloud, local=True)
        assert np.array_equal(f.cloud("less_flat"), g["world"])


# ---- loop level: LidarOdometry::UndistortLidars, pose text files (util/FileIO.cpp), CameraLidarOptimizer::NeighborEachFrame / LidarMaskByTrack ----
def undistort_lidars_case():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(20261103)
    n, per = 9, 700
    T = np.tile(np.eye(4), (n, 1, 1))
    for f in range(1, n):
        d = np.eye(4); d[:3, :3] = Rotation.from_rotvec(rng.normal(0, 0.05, 3)).as_matrix(); d[:3, 3] = rng.normal(0, 0.2, 3) + [0.4, 0, 0]
        T[f] = T[f - 1] @ d
    pv = np.ones(n, np.uint8); pv[[3, 7]] = 0
    va = np.ones(n, np.uint8); va[[3, 5]] = 0
    # a frame without a pose holds R = 0, t = inf (Velodyne's constructor state, what ReadPoseT stores for an `inf` line).  Frame 3 has neither pose nor data and is
    # stepped over; frame 7 has no pose but IS valid, so the reference's `!IsPoseValid() && !valid` test stops there and its neighbours slerp towards a non-pose:
    # their sweeps come out as NaN - reproduced, not fixed
    T[pv == 0, :3, :3] = 0.0; T[pv == 0, :3, 3] = np.inf
    off = (np.arange(n + 1) * per).astype(np.int32)
    clouds = np.concatenate([rng.normal(0, 10, (n * per, 3)), rng.uniform(0, 1, (n * per, 1))], axis=1).astype(np.float32)
    return T, pv, va, off, clouds


def test_undistort_lidars_equals_the_reference_loop(oracle):
    """(f) rank 4: the reference's own LidarOdometry::UndistortLidars (sweep-end pose from the next / previous frame with a pose, SlerpPose with the duration
    ratio, UndistortCloud) == pvb_undistort_end_poses + the oracle's UndistortCloud: float32 sweeps bit-identical, frames without an end pose untouched."""
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_velodyne.npz"))
    T, pv, va, off, clouds = undistort_lidars_case()
    for gi, gap in enumerate((0.0, 0.02)):
        e_pose, e_has = Context.undistort_end_poses(T, pv, va, gap)
        o_pose, o_has = oracle.undistort_end_poses(T, pv, va, gap)
        assert np.array_equal(e_has, o_has) and 0 < e_has.sum() < len(T)
        out = clouds.copy()
        for f in range(len(T)):
            if e_has[f]:
                out[off[f]:off[f + 1]] = oracle.undistort_cloud(T[f][:3, :3], T[f][:3, 3], e_pose[f][:3, :3], e_pose[f][:3, 3], clouds[off[f]:off[f + 1]])
        assert np.array_equal(out, g[f"ul_out{gi}"], equal_nan=True), gap
        moved = (out != clouds).any(1).reshape(len(T), -1).any(1)
        assert np.isnan(out).any() and np.isfinite(out[off[0]:off[2]]).all() and not moved[3] and not moved[5] and not moved[7]
        if oracle.ref_assoc_lib() is not None:
            assert np.array_equal(oracle.ref_undistort_lidars(T[:, :3, :3], T[:, :3, 3], pv, va, off, clouds, gap), g[f"ul_out{gi}"], equal_nan=True)


def test_pose_text_files_equal_the_reference_io(oracle, tmp_path):
    """ExportPoseT / ReadPoseT of the reference (util/FileIO.cpp) against pvb_write_poses_text / pvb_read_poses_text: the committed file written by the reference is
    reproduced byte for byte (6 significant digits, `inf` lines for frames without a pose), and both readers return the same poses from it."""
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_velodyne.npz"))
    R, t, names = g["pt_R"], g["pt_t"], [str(x) for x in g["pt_names"]]
    text = bytes(g["pt_text"]).decode()
    path = tmp_path / "poses.txt"
    Context.write_poses_text(path, R, t, names)
    assert path.read_text() == text
    for with_invalid in (True, False):
        R2, t2, valid, nm = Context.read_poses_text(path, with_invalid=with_invalid)
        key = "pt_read_all" if with_invalid else "pt_read_valid"
        assert np.array_equal(np.asarray(R2).reshape(-1, 9), g[key + "_R"].reshape(-1, 9)) and np.array_equal(np.asarray(t2), g[key + "_t"]) and list(nm) == [str(x) for x in g[key + "_names"]]
    if oracle.ref_assoc_lib() is not None:
        p2 = tmp_path / "ref.txt"
        oracle.ref_export_pose_t(p2, R, t, names)
        assert p2.read_text() == text
        Rr, tr, nr = oracle.ref_read_pose_t(path, True)
        assert np.array_equal(Rr.reshape(-1, 9), g["pt_read_all_R"].reshape(-1, 9)) and nr == [str(x) for x in g["pt_read_all_names"]]


def neighbor_each_frame_case():
    rng = np.random.default_rng(20261104)
    n = 40
    s = np.arange(n) * 0.5
    t_wl = np.stack([s, 0.3 * np.sin(s), np.zeros(n)], 1) + rng.normal(0, 0.02, (n, 3))
    t_wc = t_wl + rng.normal(0, 0.05, (n, 3))
    fpv = np.ones(n, np.uint8); fpv[[4, 17]] = 0
    lpv = np.ones(n, np.uint8); lpv[[9, 30]] = 0
    lva = np.ones(n, np.uint8); lva[21] = 0
    return t_wc, fpv, t_wl, lpv, lva


def test_neighbor_each_frame_and_lidar_mask_equal_the_reference(oracle):
    """CameraLidarOptimizer::NeighborEachFrame (temporal window / spatial k-NN + forced i-1, i+1) and LidarMaskByTrack (segments that belong to a LiDAR line track)
    of the reference == pvb_neighbor_each_frame / pvb_lidar_mask_by_track."""
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_assoc.npz"))
    t_wc, fpv, t_wl, lpv, lva = neighbor_each_frame_case()
    n = len(t_wc)
    I = np.tile(np.eye(3).reshape(1, 9), (n, 1))
    for ci, (k, temporal) in enumerate(((3, True), (1, True), (4, False), (2, False))):
        exp = [g[f"nef{ci}_ids"][g[f"nef{ci}_off"][i]:g[f"nef{ci}_off"][i + 1]].tolist() for i in range(n)]
        got = Context.neighbor_each_frame(n, n, k, temporal, t_wc, fpv, t_wl, lpv, lva)
        assert [list(x) for x in got] == exp, (k, temporal)
        if oracle.ref_assoc_lib() is not None:
            assert oracle.ref_neighbor_each_frame(I, t_wc, fpv, I, t_wl, lpv, lva, k, temporal) == exp
    for ci, (nf, n_az, k, min_len, no_pose) in enumerate(TRACK_CASES):
        frames = track_case(nf, n_az)
        tracks = [g[f"tr{ci}_feat"][g[f"tr{ci}_off"][t]:g[f"tr{ci}_off"][t + 1]] for t in range(len(g[f"tr{ci}_off"]) - 1)]
        exp = [g[f"lm{ci}_mask"][g[f"lm{ci}_off"][i]:g[f"lm{ci}_off"][i + 1]].astype(bool) for i in range(nf)]
        got = Context.lidar_mask_by_track(tracks, [len(f["segment_coeffs"]) for f in frames])
        assert all(np.array_equal(a, b) for a, b in zip(got, exp)) and sum(int(m.sum()) for m in exp) > 20, ci


# ---- the joint problem (BASELINE.json configs[2]): CameraLidarOptimizer::AssociateLineMulti + mapping-mode Optimize of the reference, recorded at ceres::Solve ----
JOINT_CASES = [dict(),                                                                                   # Room.txt: angle residuals, line-to-line + point-to-plane
               dict(angle_residual=False, normalize_distance=False, point_to_line=True, lidar_weight=0.1, refine=(True, False, True, True, False))]


def joint_case():
    from panovlm_b200 import synth
    frames, Rs, ts = builder_case(n_frames=5, n_az=600, seed=8)
    rng = np.random.default_rng(20261106)
    rows, cols = 2880, 5760
    T_cl = np.eye(4); T_cl[:3, :3] = synth.rotvec_to_R(np.array([0.01, 0.02, -0.01])); T_cl[:3, 3] = [0.03, -0.05, 0.02]
    R_wc, t_wc, image_lines = [], [], []
    for f in frames:
        T_wl = np.eye(4); T_wl[:3, :3] = f["R_wl"]; T_wl[:3, 3] = f["t_wl"]
        T_wc = T_wl @ np.linalg.inv(T_cl)
        R_wc.append(T_wc[:3, :3]); t_wc.append(T_wc[:3, 3])
        ends_cam = f["end_points"].reshape(-1, 3) @ T_cl[:3, :3].T + T_cl[:3, 3]
        px = synth.pixel_of(ends_cam, rows, cols).reshape(-1, 4) + rng.normal(0, 2.0, (len(f["end_points"]), 4))
        cl = np.stack([rng.uniform(0, cols, 15), rng.uniform(0, rows, 15), rng.uniform(0, cols, 15), rng.uniform(0, rows, 15)], axis=1)
        image_lines.append(np.concatenate([px, cl]).astype(np.float32))
    n_pts, nf = 120, len(frames)
    pts = rng.uniform([-3, -2, -1], [3, 2, 1], (n_pts, 3)) * 1.5
    keypoints = [[] for _ in range(nf)]
    track_off, ff, fi = [0], [], []
    for p_ in range(n_pts):
        for c in np.sort(rng.choice(nf, int(rng.integers(2, 5)), replace=False)):
            Pc = R_wc[c].T @ (pts[p_] - t_wc[c])
            uv = synth.pixel_of(Pc[None], rows, cols)[0] + rng.normal(0, 1.0, 2)
            ff.append(int(c)); fi.append(len(keypoints[c])); keypoints[c].append(uv)
        track_off.append(len(ff))
    keypoints = [np.array(k, np.float32).reshape(-1, 2) for k in keypoints]
    return dict(frames=frames, Rs=Rs, ts=ts, rows=rows, cols=cols, T_cl=T_cl, R_wc=np.array(R_wc), t_wc=np.array(t_wc), image_lines=image_lines, keypoints=keypoints,
                track_off=np.array(track_off, np.int32), feat_frame=np.array(ff, np.int32), feat_index=np.array(fi, np.int32), points=pts + rng.normal(0, 0.03, pts.shape))


def reference_joint_blocks(oracle, d, **kw):
    rf = [oracle.RefFrame(d["Rs"][i], d["ts"][i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["surfFlat"], f["surfLessFlat"], id=i, local="keep",
                          end_points=f["end_points"]) for i, f in enumerate(d["frames"])]
    return oracle.ref_joint_optimize_blocks(d["rows"], d["cols"], d["R_wc"], d["t_wc"], d["image_lines"], d["keypoints"], rf, d["track_off"], d["feat_frame"], d["feat_index"],
                                            d["points"], d["T_cl"], **kw)


def product_joint_blocks(oracle, d, camera_weight=1.0, lidar_weight=0.01, camera_lidar_weight=25.0, point_to_plane=True, line_to_line=True, point_to_line=False,
                         angle_residual=True, normalize_distance=True, plane_dis_threshold=1.0, line_dis_threshold=0.3, plane_tolerance=0.05, refine=None):
    """The joint problem in the reference's registration order, from the oracle's associations and the product's host functions: camera-LiDAR blocks, reprojection
    observations, then the LiDAR-LiDAR families called as Optimize calls them (panovlm_b200.joint.joint_lidar_config)."""
    from panovlm_b200 import BlockList, Context, joint, odometry
    frames, n = d["frames"], len(d["frames"])
    # base/Config.h keeps the weights and thresholds as `float`: what reaches the builders is double(float(value))
    camera_weight, lidar_weight, camera_lidar_weight = (float(np.float32(x)) for x in (camera_weight, lidar_weight, camera_lidar_weight))
    bl = BlockList(4096)
    n_pairs = 0
    for i, nbrs in enumerate(Context.neighbor_each_frame(n, n, 1, True)):
        T_wc = np.eye(4); T_wc[:3, :3] = d["R_wc"][i]; T_wc[:3, 3] = d["t_wc"][i]
        for li in nbrs:
            f = frames[li]
            T_wl = np.eye(4); T_wl[:3, :3] = d["Rs"][li]; T_wl[:3, 3] = d["ts"][li]
            T_cl = np.linalg.inv(T_wc) @ T_wl
            il, ll, s_, e_, ang = oracle.associate_by_angle(d["rows"], d["cols"], d["image_lines"][i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], np.diff(f["seg_off"]),
                                                            f["end_points"], T_cl, True, True)
            n_pairs += len(il)
            if len(il):
                Context.build_camera_lidar_blocks(bl, d["rows"], d["cols"], d["image_lines"][i][il], s_, e_, np.ones(len(il), np.float32), i, n + li, camera_lidar_weight)
    cl = {k: v.copy() for k, v in bl.view().items()}
    kp_off = np.concatenate([[0], np.cumsum([len(k) for k in d["keypoints"]])])
    xy = np.concatenate(d["keypoints"])[kp_off[d["feat_frame"]] + d["feat_index"]]
    cam, point, bearing = Context.build_reproj_observations(d["rows"], d["cols"], d["track_off"], d["feat_frame"], xy, None)
    cfg = odometry.OdometryConfig(point_to_plane=point_to_plane, line_to_line=line_to_line, point_to_line=point_to_line, angle_residual=angle_residual,
                                  normalize_distance=normalize_distance, lidar_weight=lidar_weight)
    p2l_cfg, main_cfg = joint.joint_lidar_config(cfg)
    ll = product_refine_blocks(oracle, frames, d["Rs"], d["ts"], point_to_plane=point_to_plane, line_to_line=line_to_line, point_to_line=point_to_line, use_segment=True,
                               angle_residual=angle_residual, normalize_distance=normalize_distance, plane_dis_threshold=float(np.float32(plane_dis_threshold)),
                               line_dis_threshold=float(np.float32(line_dis_threshold)), plane_tolerance=float(np.float32(plane_tolerance)), plane_weight=main_cfg.plane_weight,
                               p2l=(p2l_cfg.use_segment, p2l_cfg.angle_residual, p2l_cfg.normalize_distance, p2l_cfg.point_line_weight), block_offset=n, l2l_needs_segment=False)
    return cl, (cam, point, bearing), ll, n_pairs


def test_joint_problem_equals_the_reference_optimize(oracle):
    """configs[2]: the whole problem the reference's mapping-mode Optimize hands to the solver - AssociateLineMulti's line pairs, AddCameraLidarResidual, AddCameraResidual,
    the three LiDAR builders with the joint stage's (shifted) arguments and weights, the constant blocks - against the product's host functions fed by the oracle."""
    g = np.load(os.path.join(G, "ref_joint.npz"))
    d = joint_case()
    n = len(d["frames"])
    for ci, kw in enumerate(JOINT_CASES):
        exp = {k[len(f"j{ci}_"):]: g[k] for k in g.files if k.startswith(f"j{ci}_")}
        cl, (cam, point, bearing), ll, n_pairs = product_joint_blocks(oracle, d, **kw)
        n_cl, n_rp, n_ll = len(cl["type"]), len(cam), len(ll["type"])
        assert n_pairs == int(exp["n_line_pairs"]) >= 10 and n_cl == 2 * n_pairs
        assert len(exp["residual"]) == n_cl + n_rp + n_ll, (ci, len(exp["residual"]), n_cl, n_rp, n_ll)
        npar = exp["n_params"]
        assert np.all(npar[:n_cl] == 4) and np.all(npar[n_cl:n_cl + n_rp] == 3) and np.all(npar[n_cl + n_rp:] == 4)
        four = {k: np.concatenate([cl[k], ll[k]]) for k in cl}
        m4 = npar == 4
        assert np.array_equal(four["ref"], exp["a"][m4]) and np.array_equal(four["nei"], exp["b"][m4]) and np.abs(four["huber"] - exp["huber"][m4]).max() < 1e-15, ci
        assert np.array_equal(cam, exp["a"][~m4]) and np.array_equal(point, exp["b"][~m4]) and np.allclose(exp["huber"][~m4], 4 * np.pi / 180)
        b = oracle.Blocks(four["type"], four["ref"], four["nei"], four["consts"], 0.0, four["normalize"])
        r, J, _ = b.evaluate(exp["poses"], apply_loss=False)
        rp = oracle.Reproj(cam, point, bearing, weight=float(np.float32(kw.get("camera_weight", 1.0))))
        r3, J3 = rp.evaluate(exp["poses"][:n], d["points"], apply_loss=False)[:2]
        r_all, J_all = np.zeros(len(npar)), np.zeros((len(npar), 12))
        r_all[m4], J_all[m4] = r, J
        r_all[~m4] = r3; J_all[~m4, :9] = J3[:, :9]
        assert np.all(np.abs(r_all - exp["residual"]) <= 1e-9 * np.abs(exp["residual"]) + 5e-11), ci      # 5e-11: acos near 1 on a 3e-5 rad residual
        assert np.all(np.abs(J_all[::JAC_STRIDE] - exp["jacobian"]).max(1) <= 1e-6 * np.abs(exp["jacobian"]).max(1) + 1e-9), ci
        # constant parameter blocks: the refine_* switches (rotation 2 b, translation 2 b + 1, track 2 (n + m) + t), then camera 0 (:462-491)
        refine = kw.get("refine", (True, True, True, True, True))
        want = []
        if not refine[4]:
            want += [4 * n + t for t in range(len(d["points"]))]
        for i in range(n):
            want += ([2 * i] if not refine[0] else []) + ([2 * i + 1] if not refine[1] else [])
        for i in range(n):
            want += ([2 * (n + i)] if not refine[2] else []) + ([2 * (n + i) + 1] if not refine[3] else [])
        want += [0, 1]
        assert exp["const_part"].tolist() == want, ci
    if oracle.ref_assoc_lib() is not None:
        live = reference_joint_blocks(oracle, d, **JOINT_CASES[0])
        assert np.array_equal(live["residual"], g["j0_residual"]) and np.array_equal(live["a"], g["j0_a"])


def test_calibration_problem_equals_the_reference_optimize(oracle):
    """Calibration mode: the reference's AssociateLineSingle(T_cl) + Optimize(line_pairs, T_cl) recorded at ceres::Solve == one-to-one AssociateByAngle of the oracle +
    pvb_build_calibration_blocks (Plane2Plane_Relative in degrees with Huber 2 deg, PlaneRelativeIOUResidual with weight 2 and no loss, the float32 image-plane path)."""
    from panovlm_b200 import BlockList, Context
    g = np.load(os.path.join(G, "ref_joint.npz"))
    d = joint_case()
    T = d["T_cl"].copy(); T[:3, 3] += [0.02, -0.01, 0.015]
    bl = BlockList(2048)
    n_pairs = 0
    for i, f in enumerate(d["frames"]):
        il, ll, s_, e_, _ = oracle.associate_by_angle(d["rows"], d["cols"], d["image_lines"][i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], np.diff(f["seg_off"]), f["end_points"],
                                                      T, True, False)
        n_pairs += len(il)
        if len(il):
            Context.build_calibration_blocks(bl, d["rows"], d["cols"], d["image_lines"][i][il], s_, e_, 0)
    v = bl.view()
    assert n_pairs == int(g["cal_info"][0]) >= 20 and len(v["type"]) == len(g["cal_residual"]) == 2 * n_pairs
    assert int(g["cal_info"][1]) == 50                                                       # options.max_num_iterations (:69)
    assert np.abs(v["huber"] - g["cal_huber"]).max() < 1e-15 and np.all(g["cal_huber"][1::2] == 0) and np.allclose(g["cal_huber"][0::2], 2 * np.pi / 180)
    pose = np.concatenate([oracle.R_to_aa(T[:3, :3]), T[:3, 3]])
    assert np.abs(pose - g["cal_pose"]).max() < 1e-15
    b = oracle.Blocks(v["type"], 0, 0, v["consts"], 0.0, v["normalize"])
    r, J, _ = b.evaluate(g["cal_pose"].reshape(1, 6), apply_loss=False)
    assert np.all(np.abs(r - g["cal_residual"]) <= 1e-9 * np.abs(g["cal_residual"]) + 1e-9)             # residuals in degrees (Plane2Plane_Relative)
    assert np.all(np.abs(J[:, :6] - g["cal_jacobian"]).max(1) <= 1e-6 * np.abs(g["cal_jacobian"]).max(1) + 1e-7)
    if oracle.ref_assoc_lib() is not None:
        rf = [oracle.RefFrame(np.eye(3), np.zeros(3), f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], id=i, local="keep", end_points=f["end_points"])
              for i, f in enumerate(d["frames"])]
        live = oracle.ref_calibration_blocks(d["rows"], d["cols"], d["image_lines"], rf, T)
        assert np.array_equal(live["residual"], g["cal_residual"])


DENSE_CASE = dict(n_target=20000, n_frames=3, pts_per_frame=1500, seed=11, plane_tol=0.05, dist_thr=1.0, huber=0.2, weight=0.7)


def dense_systems_from_reference_pieces(oracle, d, tol, thr, hub, w):
    """Per-frame [H upper (21) | g (6) | cost | count] of the dense ICP evaluation built from the reference's own Transform2LidarWorld, AssociatePoint2Plane and
    Point2Plane_Meter::Create(...)->Evaluate(), with Ceres' corrector for HuberLoss written out (rho' = a / |r| beyond a; rho'' < 0 => plain sqrt(rho') scaling)."""
    tgt = oracle.RefFrame(np.eye(3), np.zeros(3), surf_less_flat_world=d["target"])
    out = np.zeros((len(d["src_off"]) - 1, 29))
    for f in range(len(out)):
        src = d["src_local"][d["src_off"][f]:d["src_off"][f + 1]]
        p = d["poses_lw_init"][f]
        R_wl = oracle.aa_to_R(p[:3]).T
        fr = oracle.RefFrame(R_wl, -R_wl @ p[3:], surf_flat_world=src, local=True)
        pt, pl = oracle.ref_associate_point2plane(tgt, fr, tol, thr)
        m = len(pt)
        raw = np.zeros((m, 16)); raw[:, :3] = pt; raw[:, 3:7] = pl; raw[:, 7] = w
        prm = np.zeros((m, 12)); prm[:, 6:] = p
        r, J, _ = oracle.ref_eval_functors(np.zeros(m, np.int32), 0, raw, prm)
        s_ = r * r
        outl = s_ > hub * hub
        scale = np.where(outl, np.sqrt(hub / np.maximum(np.abs(r), 1e-300)), 1.0)
        cost = np.where(outl, 2 * hub * np.abs(r) - hub * hub, s_).sum() * 0.5
        Jc, rc = J[:, 6:] * scale[:, None], r * scale
        H, g = Jc.T @ Jc, Jc.T @ rc
        out[f] = np.concatenate([H[np.triu_indices(6)], g, [cost, m]])
    return out


def test_dense_icp_evaluation_equals_the_reference_pieces(oracle):
    """BASELINE.json configs[4] / the bench and smoke() path: the oracle's dense ICP evaluation (association + plane fit + Point2Plane_Meter + Huber + per-frame 6x6
    reduce) against per-frame systems rebuilt from the REFERENCE'S OWN pieces (dense_systems_from_reference_pieces; committed as tests/golden/ref_dense.npz, rebuilt live when
    oracle/_ref is present).  The kernels read the same fixture in tests/test_zz_gpu_reference_fixtures.py."""
    from panovlm_b200 import synth
    c = DENSE_CASE
    d = synth.make_dense_sweep(n_target=c["n_target"], n_frames=c["n_frames"], pts_per_frame=c["pts_per_frame"], seed=c["seed"])
    g = np.load(os.path.join(G, "ref_dense.npz"))
    sys_o, _, n_o = oracle.dense_icp_eval(d["target"], d["src_local"], d["src_off"], d["poses_lw_init"], c["plane_tol"], c["dist_thr"], 10, c["huber"], c["weight"], 1)
    exp = g["systems"]
    assert np.array_equal(sys_o[:, 28], exp[:, 28]) and n_o == exp[:, 28].sum() and np.all(exp[:, 28] > 500)
    assert np.abs(sys_o - exp).max() < 1e-9 * np.abs(exp).max()
    if oracle.ref_assoc_lib() is not None and oracle.ref_path_lib() is not None:
        live = dense_systems_from_reference_pieces(oracle, d, c["plane_tol"], c["dist_thr"], c["huber"], c["weight"])
        assert np.array_equal(live, exp)


def test_pixel_space_candidates_equal_the_reference_first_stage(oracle):
    """A5, first stage: the candidate lists the reference's own pixel-space Associate() hands to its RANSAC fit (recorded by the SACSegmentation stand-in) == the oracle's
    pixel_line_neighbors + the product's host pvb_pixel_line_candidates: same lines get a list, same LiDAR points in the same order (duplicates included)."""
    from panovlm_b200 import Context
    g = np.load(os.path.join(G, "ref_camlidar.npz"))
    A, rows, cols, T, lines = camlidar_case()
    cloud = A["cloud"][::4]
    line3, _, _ = oracle.pixel_line_neighbors(rows, cols, lines, cloud, T)
    off, idx = Context.pixel_line_candidates(len(lines), line3, 6)
    cam = oracle.transform_cloud(T[:3, :3], T[:3, 3], cloud)[:, :3]
    got = [cam[idx[off[l]:off[l + 1]]] for l in range(len(lines)) if off[l + 1] > off[l]]
    exp = [g["px_xyz"][g["px_off"][k]:g["px_off"][k + 1]] for k in range(len(g["px_off"]) - 1)]
    assert len(got) == len(exp) >= 10 and sum(len(x) for x in exp) > 300
    assert all(np.array_equal(a, b) for a, b in zip(got, exp))
    if oracle.ref_camlidar_lib() is not None:
        live = oracle.ref_pixel_associate_candidates(rows, cols, lines, cloud, T)
        assert len(live) == len(exp) and all(np.array_equal(a, b) for a, b in zip(live, exp))


LOOP_SCRIPTS = [[(100, 20), (50, 20), (49.9, 20), (10, 20)],                         # cost change below 1 % of the previous cost
                [(100, 3), (50, 2), (20, 1)],                                        # fewer than 5 successful steps twice in a row
                [(100, 20), (80, 3), (60, 20), (40, 2), (20, 1), (10, 1)],           # a good iteration in between resets nothing: only consecutive ones count
                [(0.0, 20), (0.0, 20), (1, 1)],                                      # 0 / 0 and x / 0: the relative test never fires on a zero previous cost
                [(100, 20), (90, 20), (80, 20), (70, 20), (60, 20), (50, 20), (40, 20), (30, 20), (20, 20)]]      # runs into num_iteration_joint


def test_joint_outer_loop_equals_the_reference_joint_optimize(oracle):
    """The mapping-mode outer loop: the reference's own JointOptimize (AssociateLineMulti, Optimize, re-association, two early exits) driven by a scripted solver calls
    Optimize exactly as often as panovlm_b200.joint.joint_optimize does with the same scripted Optimize."""
    from panovlm_b200 import joint
    g = np.load(os.path.join(G, "ref_joint.npz"))
    counts = []
    for script in LOOP_SCRIPTS:
        calls = []

        def scripted(ctx, data, cams, lidars, points, cfg, aa_to_R, _s=script, _c=calls):
            k = min(len(_c), len(_s) - 1)
            _c.append(k)
            return cams, lidars, points, dict(final_cost=float(_s[k][0]), successful=int(_s[k][1])), None
        joint.joint_optimize(None, None, None, None, None, joint.JointConfig(), None, num_iteration_joint=7, optimize_fn=scripted)
        counts.append(len(calls))
    assert counts == g["loop_counts"].tolist() == [3, 2, 5, 4, 7]
    if oracle.ref_assoc_lib() is not None:
        d = joint_case()
        rf = [oracle.RefFrame(d["Rs"][i], d["ts"][i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["surfFlat"], f["surfLessFlat"], id=i, local="keep",
                              end_points=f["end_points"]) for i, f in enumerate(d["frames"])]
        live = [oracle.ref_joint_optimize_loop(d["rows"], d["cols"], d["R_wc"], d["t_wc"], d["image_lines"], rf, d["T_cl"], 7, [x[0] for x in sc], [x[1] for x in sc])
                for sc in LOOP_SCRIPTS]
        assert live == counts


CALIB_SCRIPTS = [np.zeros((0, 6)),                                                                        # nothing moves: one iteration
                 np.array([[0.01, 0, 0, 0.02, 0, 0], [0.004, 0, 0, 0, 0.02, 0], [1e-5, 0, 0, 1e-4, 0, 0]]),  # two real updates, then below both thresholds
                 np.array([[1e-4, 0, 0, 0.02, 0, 0], [1e-4, 0, 0, 0.005, 0, 0]]),                          # rotation already small, translation decides
                 np.array([[0.002, 0, 0, 0, 0, 0]] * 4 + [[0.0015, 0, 0, 0, 0, 0]])]                       # 0.115 deg steps stay above 0.1 deg, 0.086 deg ends it


def test_calibration_outer_loop_equals_the_reference_joint_optimize(oracle):
    """Calibration mode, outer loop: the reference's own JointOptimize (AssociateLineSingle, Optimize(line_pairs, T_cl), float32 rotation / translation change, exit below
    0.1 deg AND 0.01) with a scripted solver (the k-th solve adds a scripted delta to the pose block) against panovlm_b200.joint.calibrate running the same script:
    same number of iterations, same final T_cl."""
    from panovlm_b200 import joint
    g = np.load(os.path.join(G, "ref_joint.npz"))
    d = joint_case()
    frames = d["frames"][:3]
    T0 = d["T_cl"].copy(); T0[:3, 3] += [0.02, -0.01, 0.015]

    def associate(T_cl):
        return [oracle.associate_by_angle(d["rows"], d["cols"], d["image_lines"][i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], np.diff(f["seg_off"]), f["end_points"], T_cl,
                                          True, False) for i, f in enumerate(frames)]
    for si, script in enumerate(CALIB_SCRIPTS):
        calls = []

        def solve(blocks, pose, _s=script, _c=calls):
            k = len(_c); _c.append(k)
            return pose + (_s[k][None, :] if k < len(_s) else 0.0), dict(final_cost=0.0, successful=0)
        T, log = joint.calibrate(None, frames, d["image_lines"], d["rows"], d["cols"], T0, oracle.aa_to_R, oracle.R_to_aa, associate_fn=associate, solve_fn=solve)
        assert len(log) == int(g["cal_loop_calls"][si]), (si, len(log))
        assert np.abs(T - g["cal_loop_T"][si]).max() < 1e-12, si
    assert g["cal_loop_calls"].tolist() == [1, 3, 2, 5]
    if oracle.ref_assoc_lib() is not None:
        rf = [oracle.RefFrame(np.eye(3), np.zeros(3), f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], id=i, local="keep", end_points=f["end_points"])
              for i, f in enumerate(frames)]
        m, To = oracle.ref_calibration_loop(d["rows"], d["cols"], d["image_lines"][:3], rf, T0, CALIB_SCRIPTS[1])
        assert m == 3 and np.array_equal(To, g["cal_loop_T"][1])
