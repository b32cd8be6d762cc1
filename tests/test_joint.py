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
