"""Context.pixel_associate = CameraLidarLineAssociate::Associate(lines, cloud, T_cl) (CameraLidarLineAssociate.cpp:22-188) end to end.  The first stage is pinned
against the reference's own code (tests/test_zz_gpu_reference_fixtures.py::test_pixel_knn_kernel_equals_the_reference_first_stage); the line fit's RANSAC restates
PCL and is parity-unpinned (tests/test_pixel_fit_line.py).  Here: the composition (CPU, the oracle standing in for the two device calls) and the same call on the
device (named to run after every other GPU test: first run on a B200 = the round-end run)."""
import numpy as np
import pytest

from panovlm_b200 import Context
from test_reference_pinning import camlidar_case


class _OracleBackedCtx(Context):
    def __init__(self, pvo):                                                         # no device: the two device calls go to the oracle
        self._pvo = pvo

    def pixel_line_neighbors(self, rows, cols, lines, cloud, T):
        return self._pvo.pixel_line_neighbors(rows, cols, lines, cloud, T)

    def transform_cloud(self, c, R, t):
        return self._pvo.transform_cloud(R, t, c)

    def __del__(self):
        pass


def _case():
    A, rows, cols, T, lines = camlidar_case()
    return rows, cols, lines, A["cloud"][::4], T


def _check(out, rows, cols, lines, cloud, T):
    ids, s, e, ang = out
    assert len(ids) >= 3 and np.all(np.diff(ids) > 0) and s.shape == e.shape == (len(ids), 3) and ang.dtype == np.float32
    assert np.all(ang < 5.0)                                                          # Filter: great-circle planes within 5 degrees (:652-655)
    # both ends lie on a line through the LiDAR cloud: some cloud point within the RANSAC threshold (0.1 m) + slack of each end's neighbourhood
    d_s = np.linalg.norm(cloud[None, :, :3] - s[:, None, :], axis=2).min(1)
    d_e = np.linalg.norm(cloud[None, :, :3] - e[:, None, :], axis=2).min(1)
    assert d_s.max() < 1.0 and d_e.max() < 1.0
    assert np.all(np.linalg.norm(s - e, axis=1) > 0.1)


def test_pixel_associate_composition_on_the_oracle(oracle):
    rows, cols, lines, cloud, T = _case()
    ctx = _OracleBackedCtx(oracle)
    out = ctx.pixel_associate(rows, cols, lines, cloud, T)
    _check(out, rows, cols, lines, cloud, T)
    again = ctx.pixel_associate(rows, cols, lines, cloud, T)
    assert all(np.array_equal(a, b) for a, b in zip(out, again))


@pytest.mark.gpu
def test_pixel_associate_on_the_device_equals_the_oracle_backed_composition(gpu_ctx, oracle):
    rows, cols, lines, cloud, T = _case()
    got = gpu_ctx.pixel_associate(rows, cols, lines, cloud, T)
    exp = _OracleBackedCtx(oracle).pixel_associate(rows, cols, lines, cloud, T)
    assert all(np.array_equal(a, b) for a, b in zip(got, exp))                        # first stage and transform are bit-exact, the rest is the same host code
    _check(got, rows, cols, lines, cloud, T)


def test_associate_lines_takes_the_pixel_path_for_frames_without_segments(oracle):
    """AssociateLineMulti (CameraLidarOptimizer.cpp:360-367): a LiDAR frame whose edge_segmented is empty goes through Associate(lines, cornerLessSharp, T_cl)."""
    from panovlm_b200 import joint
    A, rows, cols, T, lines = camlidar_case()
    frame = dict(cornerLessSharp=A["cloud"][::4], p2s_off=np.zeros(1, np.int32), p2s_ids=np.zeros(0, np.int32), segment_coeffs=np.zeros((0, 6)), end_points=np.zeros((0, 6)))
    cam = np.concatenate([oracle.R_to_aa(T[:3, :3]), T[:3, 3]])[None, :]
    ctx = _OracleBackedCtx(oracle)
    pairs = joint.associate_lines(ctx, [frame], [lines], cam, np.zeros((1, 6)), rows, cols, oracle.aa_to_R)
    il, ll, s, e, ang = pairs[(0, 0)]
    exp = ctx.pixel_associate(rows, cols, lines, frame["cornerLessSharp"], T)
    assert np.array_equal(il, exp[0]) and np.all(ll == -1) and len(il) >= 3
    assert np.abs(s - exp[1]).max() < 1e-9 and np.abs(e - exp[2]).max() < 1e-9          # T_cl rebuilt from the angle-axis block


def test_everything_after_the_ransac_equals_the_reference_own_code(oracle):
    """tests/golden/ref_pixel_fit.npz = the reference's own Associate() (oracle/_ref) with the RANSAC's inliers scripted to the product's (tests/make_golden.py:
    golden_ref_pixel_fit).  Pins what follows the RANSAC against the reference's code: the < 3 test, the refit, the farthest pair with its position-as-index quirk
    (:136-137), ProjectPoint2Line3D, Filter(true, true), the transform back to the LiDAR frame."""
    import os
    from make_golden import pixel_fit_script
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pixel_fit.npz"))
    rows, cols, lines, cloud, T = _case()
    script = pixel_fit_script(oracle, rows, cols, lines, cloud, T)
    assert np.array_equal(np.concatenate(script), g["inl_idx"]) and np.array_equal(np.cumsum([len(x) for x in script]), g["inl_off"][1:])     # the script is in sync
    ids, s, e, ang = _OracleBackedCtx(oracle).pixel_associate(rows, cols, lines, cloud, T)
    assert len(ids) == len(g["angle"]) >= 10 and np.array_equal(lines[ids], g["image_line"])
    assert np.array_equal(ang, g["angle"])
    assert np.abs(s - g["start"]).max() < 1e-9 and np.abs(e - g["end"]).max() < 1e-9
    if oracle.ref_camlidar_lib() is not None:                                                # live, where oracle/_ref is built
        il, s2, e2, a2 = oracle.ref_pixel_associate_scripted(rows, cols, lines, cloud, T, script)
        assert np.array_equal(il, g["image_line"]) and np.array_equal(s2, g["start"]) and np.array_equal(e2, g["end"]) and np.array_equal(a2, g["angle"])


def test_calibration_loop_takes_the_pixel_path_for_frames_without_segments(oracle):
    """AssociateLineSingle (CameraLidarOptimizer.cpp:309-314): with edge_segmented empty the pairs of the calibration problem come from Associate(lines, cornerLessSharp, T_cl);
    two residual blocks per pair on the one pose block (Optimize(line_pairs, T_cl) :32-64).  The solver is scripted (no device)."""
    from panovlm_b200 import joint
    A, rows, cols, T, lines = camlidar_case()
    frame = dict(cornerLessSharp=A["cloud"][::4], p2s_off=np.zeros(1, np.int32), p2s_ids=np.zeros(0, np.int32), segment_coeffs=np.zeros((0, 6)), end_points=np.zeros((0, 6)))
    ctx = _OracleBackedCtx(oracle)
    seen = []

    def solve(v, pose):
        seen.append(v)
        return pose.copy(), dict(final_cost=1.0, successful=1)

    T_out, log = joint.calibrate(ctx, [frame], [lines], rows, cols, T, oracle.aa_to_R, oracle.R_to_aa, max_iterations=3, solve_fn=solve)
    n_pairs = len(ctx.pixel_associate(rows, cols, lines, frame["cornerLessSharp"], T)[0])
    assert len(log) == 1 and log[0]["n_pairs"] == n_pairs >= 10                       # pose unchanged by the scripted solver: the loop stops after one iteration
    assert len(seen[0]["type"]) == 2 * n_pairs and np.all(seen[0]["ref"] == 0)
    assert np.abs(T_out - T).max() < 1e-12


def test_calibration_problem_over_the_pixel_path_equals_the_reference_optimize(oracle):
    """The reference's own AssociateLineSingle + Optimize(line_pairs, T_cl) on a frame WITHOUT LiDAR segments (RANSAC answers scripted, recorded at ceres::Solve;
    tests/golden/ref_pixel_fit.npz cal_*) == Context.pixel_associate + pvb_build_calibration_blocks: same pairs, residuals (degrees) and Jacobians of the one pose block."""
    import os
    from panovlm_b200 import BlockList
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pixel_fit.npz"))
    rows, cols, lines, cloud, T = _case()
    il, s, e, _ = _OracleBackedCtx(oracle).pixel_associate(rows, cols, lines, cloud, T)
    bl = BlockList(2 * len(il) + 4)
    Context.build_calibration_blocks(bl, rows, cols, lines[il], s, e, 0)
    v = bl.view()
    assert len(il) == int(g["cal_info"][0]) and len(v["type"]) == len(g["cal_residual"]) == 2 * len(il)
    assert np.abs(v["huber"] - g["cal_huber"]).max() < 1e-15
    pose = np.concatenate([oracle.R_to_aa(T[:3, :3]), T[:3, 3]])
    assert np.abs(pose - g["cal_pose"]).max() < 1e-15
    r, J, _ = oracle.Blocks(v["type"], 0, 0, v["consts"], 0.0, v["normalize"]).evaluate(g["cal_pose"].reshape(1, 6), apply_loss=False)
    assert np.all(np.abs(r - g["cal_residual"]) <= 1e-9 * np.abs(g["cal_residual"]) + 1e-9)
    assert np.all(np.abs(J[:, :6] - g["cal_jacobian"]).max(1) <= 1e-6 * np.abs(g["cal_jacobian"]).max(1) + 1e-7)
    if oracle.ref_assoc_lib() is not None:                                               # live, where oracle/_ref is built
        from make_golden import pixel_calibration_from_reference, pixel_fit_script
        live = pixel_calibration_from_reference(oracle, rows, cols, lines, cloud, T, pixel_fit_script(oracle, rows, cols, lines, cloud, T))
        assert np.array_equal(live["residual"], g["cal_residual"]) and np.array_equal(live["jacobian"], g["cal_jacobian"])


def _segments():
    from make_golden import segments_of
    A, rows, cols, T, lines = camlidar_case()
    return rows, cols, lines, segments_of(A), T


def test_segmented_associate_equals_the_reference_own_code(oracle):
    """Context.pixel_associate_segmented = the SEGMENTED overload Associate(lines, segmented_cloud, T_cl) (CameraLidarLineAssociate.cpp:191-338; the reference's own call
    sites of it are commented out).  tests/golden/ref_pixel_fit.npz seg_* = that function of the reference (oracle/_ref) with the RANSAC's inliers scripted to the product's:
    the script is in sync (the same image lines reach the fit: 6-point test, 70 % single-segment test, first-maximum rule), and everything after the RANSAC is the
    reference's code - including its quirk of fitting the segment in the LiDAR frame and applying T_lc afterwards anyway."""
    import os
    from make_golden import pixel_fit_script_segmented
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pixel_fit.npz"))
    rows, cols, lines, segs, T = _segments()
    script = pixel_fit_script_segmented(oracle, rows, cols, lines, segs, T)
    assert len(script) == len(g["seg_inl_off"]) - 1 >= 10
    assert np.array_equal(np.concatenate(script), g["seg_inl_idx"]) and np.array_equal(np.cumsum([len(x) for x in script]), g["seg_inl_off"][1:])
    ids, s, e, ang = _OracleBackedCtx(oracle).pixel_associate_segmented(rows, cols, lines, segs, T)
    assert len(ids) == len(g["seg_angle"]) >= 3 and np.array_equal(lines[ids], g["seg_image_line"])
    assert np.array_equal(ang, g["seg_angle"])
    assert np.abs(s - g["seg_start"]).max() < 1e-9 and np.abs(e - g["seg_end"]).max() < 1e-9
    if oracle.ref_camlidar_lib() is not None:                                                # live, where oracle/_ref is built
        live = oracle.ref_pixel_associate_segmented_scripted(rows, cols, lines, segs, T, script)
        assert live is not None and np.array_equal(live[0], g["seg_image_line"]) and np.array_equal(live[1], g["seg_start"]) and np.array_equal(live[3], g["seg_angle"])
        # a script of the wrong length is noticed: the number of fits is part of what is pinned
        assert oracle.ref_pixel_associate_segmented_scripted(rows, cols, lines, segs, T, script[:-1]) is None
    # no segments / empty segments
    out = _OracleBackedCtx(oracle).pixel_associate_segmented(rows, cols, lines, [], T)
    assert len(out[0]) == 0 and out[1].shape == (0, 3)


@pytest.mark.gpu
def test_segmented_associate_on_the_device_equals_the_oracle_backed_composition(gpu_ctx, oracle):
    rows, cols, lines, segs, T = _segments()
    got = gpu_ctx.pixel_associate_segmented(rows, cols, lines, segs, T)
    exp = _OracleBackedCtx(oracle).pixel_associate_segmented(rows, cols, lines, segs, T)
    assert len(got[0]) >= 3 and all(np.array_equal(a, b) for a, b in zip(got, exp))
