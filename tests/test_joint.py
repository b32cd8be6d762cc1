"""The joint camera-LiDAR refinement end to end (BASELINE.json configs[2]; joint_optimization/CameraLidarOptimizer.cpp:330-548): line association by
angle on the device, camera-LiDAR / LiDAR-LiDAR residual blocks from the host builders, reprojection observations, ONE trust-region problem over
[cameras | LiDARs | points].  The block lists the product builds are handed to the oracle's dense LM; both must take the same steps."""
import numpy as np
import pytest

from panovlm_b200 import joint, synth


@pytest.mark.gpu
@pytest.mark.parametrize("refine_structure", [True, False])
def test_joint_optimize_matches_oracle(gpu_ctx, oracle, refine_structure):
    d = synth.make_joint_problem(n_frames=6, n_points=200, n_az=600)
    n = 6
    cfg = joint.JointConfig(refine_structure=refine_structure, max_lm_iterations=12)
    cams, lidars, points, summ, (v, const, pt_const) = joint.optimize(gpu_ctx, d, d["cams"], d["lidars"], d["points"], cfg, oracle.aa_to_R, device_blocks=False)
    # the same call with the point-to-plane blocks built on the device: same problem (rows in another order), same answer up to rounding
    cams2, lidars2, points2, summ2, _ = joint.optimize(gpu_ctx, d, d["cams"], d["lidars"], d["points"], cfg, oracle.aa_to_R, device_blocks=True)
    for k in ("iterations", "successful", "unsuccessful", "termination", "n_camera_lidar_blocks", "n_lidar_blocks"):
        assert summ[k] == summ2[k], (k, summ, summ2)
    assert abs(summ["final_cost"] - summ2["final_cost"]) < 1e-9 * summ["final_cost"]
    assert np.abs(np.concatenate([cams2, lidars2]) - np.concatenate([cams, lidars])).max() < 1e-8 and np.abs(points2 - points).max() < 1e-8
    assert summ["n_line_pairs"] >= 10 and summ["n_camera_lidar_blocks"] == 2 * summ["n_line_pairs"] and summ["n_lidar_blocks"] > 1000
    # the same problem through the oracle's dense LM
    blk = oracle.Blocks(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"])
    rep = oracle.Reproj(d["cam"], d["point"], d["bearing"], weight=cfg.camera_weight, huber=4.0 * np.pi / 180.0)
    mask = np.concatenate([const.ravel(), np.repeat(np.zeros(len(points), np.uint8) if pt_const is None else pt_const, 3)])
    start = np.concatenate([d["cams"], d["lidars"]])
    e_x, e_p, e_s = oracle.joint_solve_lm(blk, rep, start, d["points"], mask, max_iter=12)
    for k in ("iterations", "successful", "unsuccessful", "termination"):
        assert e_s[k] == summ[k], (k, e_s, summ)
    assert abs(e_s["final_cost"] - summ["final_cost"]) < 1e-5 * e_s["final_cost"]
    got = np.concatenate([cams, lidars])
    assert np.abs(got - e_x).max() < 1e-4 * np.abs(e_x - start).max()                    # pose deltas: 1e-4 relative (BASELINE.json)
    if refine_structure:
        assert np.abs(points - e_p).max() < 1e-4 * np.abs(e_p - d["points"]).max()
    else:
        assert np.array_equal(points, d["points"])
    assert np.array_equal(cams[0], d["cams"][0])
    assert summ["final_cost"] < 0.5 * summ["initial_cost"]


@pytest.mark.gpu
def test_calibration_loop_matches_oracle_loop(gpu_ctx, oracle):
    """Calibration mode (CameraLidarOptimizer.cpp:195-233): AssociateLineSingle + Optimize(line_pairs, T_cl) iterated from a perturbed extrinsic; the same
    loop driven by oracle primitives must take the same iterations and end at the same T_cl, closer to the true one than the start."""
    from scipy.spatial.transform import Rotation
    d = synth.make_joint_problem(n_frames=5, n_points=10, n_az=900, clutter=15, pixel_noise=1.0)
    frames, lines, rows, cols = d["frames"], d["image_lines"], d["rows"], d["cols"]
    T0 = d["T_cl"].copy()
    T0[:3, :3] = T0[:3, :3] @ Rotation.from_rotvec([0.006, -0.004, 0.005]).as_matrix()
    T0[:3, 3] += [0.02, -0.015, 0.01]
    T_gpu, log = joint.calibrate(gpu_ctx, frames, lines, rows, cols, T0, oracle.aa_to_R, oracle.R_to_aa, max_iterations=6)
    # the oracle's loop
    T = T0.copy()
    sizes = [np.diff(f["seg_off"]) for f in frames]

    def associate(T_cl):
        return [oracle.associate_by_angle(rows, cols, lines[i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], sizes[i], f["end_points"], T_cl, True, False)
                for i, f in enumerate(frames)]

    pairs = associate(T)
    n_iter = 0
    for it in range(6):
        typ, hub, consts = [], [], []
        for i, (il, ll, s, e, ang) in enumerate(pairs):
            if len(il):
                t_, h_, c_ = oracle.build_calibration_blocks(rows, cols, lines[i][il], s, e)
                typ.append(t_); hub.append(h_); consts.append(c_)
        typ, hub, consts = np.concatenate(typ), np.concatenate(hub), np.concatenate(consts)
        assert len(typ) == 2 * log[it]["n_pairs"]
        blk = oracle.Blocks(typ, np.zeros(len(typ), np.int32), np.zeros(len(typ), np.int32), consts, hub, 1)
        pose = np.concatenate([oracle.R_to_aa(T[:3, :3]), T[:3, 3]])[None, :]
        new_pose, summ = blk.solve_lm(pose, None, 50)
        assert abs(summ["iterations"] - log[it]["iterations"]) <= 1 and abs(summ["final_cost"] - log[it]["final_cost"]) < 1e-6 * max(summ["final_cost"], 1e-12)
        T_new = np.eye(4); T_new[:3, :3] = oracle.aa_to_R(new_pose[0, :3]); T_new[:3, 3] = new_pose[0, 3:]
        rot = np.float32(np.arccos(np.clip((np.trace(T[:3, :3].T @ T_new[:3, :3]) - 1) / 2.0, -1.0, 1.0))) * np.float32(180.0 / np.pi)
        tr = np.float32(np.linalg.norm(T[:3, 3] - T_new[:3, 3]))
        T = T_new
        pairs = associate(T)
        n_iter += 1
        if rot < 0.1 and tr < 0.01:
            break
    assert n_iter == len(log)
    assert np.abs(T_gpu - T).max() < 1e-4 * np.abs(T - T0).max()                       # pose deltas: 1e-4 relative (BASELINE.json)
    err0 = np.abs(T0 - d["T_cl"]).max()
    assert np.abs(T_gpu - d["T_cl"]).max() < err0
    assert log[0]["n_pairs"] >= 10
