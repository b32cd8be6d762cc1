"""Control flow of the joint refinement's outer loop (CameraLidarOptimizer::JointOptimize, mapping mode, CameraLidarOptimizer.cpp:256-287) with a
scripted Optimize: the early exits must fire exactly like the reference's."""
import numpy as np

from panovlm_b200 import joint


def _scripted(costs, steps):
    calls = {"n": 0}

    def fn(ctx, data, cams, lidars, points, cfg, aa_to_R):
        i = calls["n"]
        calls["n"] += 1
        return cams + 1.0, lidars, points, {"final_cost": costs[i], "successful": steps[i]}, None

    return fn, calls


def test_outer_loop_early_exits():
    z = np.zeros((2, 6))
    # runs all iterations while the cost keeps falling by more than 1 % and the LM keeps taking >= 5 steps
    fn, calls = _scripted([100.0, 50.0, 25.0, 12.0, 6.0], [9, 9, 9, 9, 9])
    cams, _, _, log = joint.joint_optimize(None, None, z, z, None, None, None, 5, fn)
    assert calls["n"] == 5 and len(log) == 5 and np.all(cams == 5.0)
    # cost change below 1 % of the previous cost stops after that iteration (:271); the first iteration never stops (division by last_cost = 0)
    fn, calls = _scripted([100.0, 99.5, 1.0], [9, 9, 9])
    joint.joint_optimize(None, None, z, z, None, None, None, 5, fn)
    assert calls["n"] == 2
    fn, calls = _scripted([0.0, 0.0, 0.0], [9, 9, 9])            # 0 / 0 is NaN: the comparison is false, like in the reference
    joint.joint_optimize(None, None, z, z, None, None, None, 3, fn)
    assert calls["n"] == 3
    # fewer than 5 successful steps twice in a row (:276); one short iteration alone does not stop
    fn, calls = _scripted([100.0, 50.0, 25.0, 12.0], [9, 3, 9, 9])
    joint.joint_optimize(None, None, z, z, None, None, None, 4, fn)
    assert calls["n"] == 4
    fn, calls = _scripted([100.0, 50.0, 25.0, 12.0], [9, 3, 4, 9])
    joint.joint_optimize(None, None, z, z, None, None, None, 4, fn)
    assert calls["n"] == 3
    fn, calls = _scripted([100.0, 50.0], [3, 9])                 # last_step starts at INT32_MAX: the first short iteration does not stop
    joint.joint_optimize(None, None, z, z, None, None, None, 2, fn)
    assert calls["n"] == 2


def test_estimate_pose_early_exits():
    """LidarOdometry::EstimatePose (lidar_mapping/LidarOdometry.cpp:166-183) with a scripted RefinePose."""
    from panovlm_b200 import odometry

    def scripted(costs, steps):
        calls = {"n": 0}

        def fn(ctx, frames, poses, cfg, aa_to_R):
            i = calls["n"]
            calls["n"] += 1
            return poses + 1.0, {"final_cost": costs[i], "successful": steps[i]}

        return fn, calls

    z = np.zeros((2, 6))
    fn, calls = scripted([100.0, 60.0, 30.0, 10.0, 5.0, 2.0, 1.0], [9] * 7)
    poses, log = odometry.estimate_pose(None, None, z, None, None, 7, fn)
    assert calls["n"] == 7 and len(log) == 7 and np.all(poses == 7.0)
    fn, calls = scripted([100.0, 99.2, 50.0], [9, 9, 9])                 # |99.2 - 100| / 100 < 1 %: stop after the second iteration
    odometry.estimate_pose(None, None, z, None, None, 7, fn)
    assert calls["n"] == 2
    fn, calls = scripted([100.0, 98.9, 50.0, 49.9], [9, 9, 9, 9])        # 1.1 % of the previous cost is not enough; 0.2 % is
    odometry.estimate_pose(None, None, z, None, None, 7, fn)
    assert calls["n"] == 4
    fn, calls = scripted([100.0, 50.0, 25.0, 12.0], [4, 9, 3, 2])
    odometry.estimate_pose(None, None, z, None, None, 7, fn)
    assert calls["n"] == 4
    fn, calls = scripted([100.0, 50.0, 25.0], [4, 4, 9])
    odometry.estimate_pose(None, None, z, None, None, 7, fn)
    assert calls["n"] == 2
