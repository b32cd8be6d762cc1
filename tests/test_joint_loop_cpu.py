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


def test_joint_stage_calls_the_lidar_builders_like_the_reference(monkeypatch):
    """CameraLidarOptimizer::Optimize passes the LiDAR builders other arguments than RefinePose does (CameraLidarOptimizer.cpp:436-458; pinned against the
    reference's own Optimize in test_reference_pinning.py): joint_lidar_config derives the two configurations, _lidar_blocks keeps the reference's registration order
    (point-to-line first) and the return shape of odometry.build_problem for both values of host_point2plane."""
    import numpy as np
    from panovlm_b200 import BlockList, joint, odometry
    cfg = odometry.OdometryConfig(point_to_line=True, angle_residual=False, normalize_distance=True, lidar_weight=0.1)
    p2l, main = joint.joint_lidar_config(cfg)
    assert (p2l.use_segment, p2l.angle_residual, p2l.normalize_distance, p2l.point_line_weight) == (False, True, True, 1.0)       # every argument one slot early
    assert not p2l.point_to_plane and not p2l.line_to_line and p2l.point_to_line
    assert main.plane_weight == 0.1 and not main.point_to_line and main.line_to_line and main.point_to_plane
    assert cfg.plane_weight == 1.0 and cfg.point_to_line                                                                          # the caller's config is untouched
    zero = joint.joint_lidar_config(odometry.OdometryConfig(point_to_line=True, lidar_weight=0.0))[0]
    assert zero.normalize_distance is False                                                                                       # bool(lidar_weight)
    calls = []

    def fake_build(ctx, frames, poses, c, aa_to_R, frame_range=None, host_point2plane=True):
        calls.append((c.point_to_line, c.point_to_plane, host_point2plane))
        bl = BlockList(8)
        k = 2 if c.point_to_line else 3
        bl.extend(dict(type=np.full(k, 3 if c.point_to_line else 1, np.int32), ref=np.zeros(k, np.int32), nei=np.ones(k, np.int32), normalize=np.ones(k, np.int32),
                       huber=np.zeros(k), consts=np.zeros((k, 12))))
        return (bl, [(0, 1)]) if host_point2plane else (bl, [(0, 1)], [(0, 1)])
    monkeypatch.setattr(odometry, "build_problem", fake_build)
    (bl, edges), m = joint._lidar_blocks(None, None, None, cfg, None)
    assert bl.view()["type"].tolist() == [3, 3, 1, 1, 1] and edges == [(0, 1)] and m.plane_weight == 0.1
    (bl, edges, mine), _ = joint._lidar_blocks(None, None, None, cfg, None, host_point2plane=False)
    assert bl.n == 5 and mine == [(0, 1)]
    (bl, edges), _ = joint._lidar_blocks(None, None, None, odometry.OdometryConfig(), None)
    assert bl.n == 3 and calls[-1] == (False, True, True)
