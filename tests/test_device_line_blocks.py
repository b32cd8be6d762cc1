"""pvb_frames_line2line_blocks_device (csrc/pvb_lines.cuh): the FindAssociations tail, the line-track gate and the Point2Line blocks of
AddLidarLineToLineResidual2 (util/Optimization.cpp:329-441) produced on the device.  The host-tail path (pvb_frames_line2line_blocks, pinned against the reference's
own RefinePose recording in test_reference_pinning / test_zz_gpu_reference_fixtures) is the checker: same block count, same normal equations, same RefinePose."""
import numpy as np
import pytest


def perturbed_sequence(oracle, n, seed):
    from panovlm_b200 import odometry, synth
    from scipy.spatial.transform import Rotation
    frames = synth.make_sequence(n, n_az=600, tilt=0.3)
    rng = np.random.default_rng(seed)
    R0 = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.005, 3) * (i > 0)).as_matrix() for i, f in enumerate(frames)]
    t0 = [f["t_wl"] + rng.normal(0, 0.02, 3) * (i > 0) for i, f in enumerate(frames)]
    return frames, odometry.pose_blocks_from_world(R0, t0, oracle.R_to_aa)


def system_of(ctx, frames, poses, cfg, aa_to_R, device_line_blocks):
    from panovlm_b200 import odometry
    bl, _, mine = odometry.build_problem(ctx, frames, poses, cfg, aa_to_R, host_point2plane=False, device_line_blocks=device_line_blocks)
    ref, nei = np.array([e[0] for e in mine], np.int32), np.array([e[1] for e in mine], np.int32)
    n = ctx.frames_point2plane_blocks(poses, ref, nei, cfg.plane_tolerance, cfg.plane_dis_threshold, cfg.angle_residual, cfg.normalize_distance, cfg.plane_weight, len(frames),
                                      extra=bl.view())
    ctx.blocks_evaluate(poses, want_rows=False, want_system=True)
    H, g, c = ctx.blocks_dense_system()
    return bl.n, n, H, g, c


@pytest.mark.gpu
@pytest.mark.parametrize("tracks", [True, False], ids=["track-gate", "no-gate"])
@pytest.mark.parametrize("angle", [True, False], ids=["angle", "metre"])
def test_device_line_blocks_give_the_normal_equations_of_the_host_line_blocks(gpu_ctx, oracle, tracks, angle):
    from panovlm_b200 import odometry
    frames, poses = perturbed_sequence(oracle, 8, 23)
    cfg = odometry.OdometryConfig(line_tracks=tracks, angle_residual=angle)
    n_line_host, n_host, H0, g0, c0 = system_of(gpu_ctx, frames, poses, cfg, oracle.aa_to_R, False)
    n_line_dev, n_dev, H1, g1, c1 = system_of(gpu_ctx, frames, poses, cfg, oracle.aa_to_R, True)
    assert n_line_host > 0 and n_line_dev == 0                      # the device path hands no line block to the host
    assert n_dev == n_host                                          # same number of blocks in all
    # the same blocks, summed in another order (one reduction edge per pose-graph edge instead of the host list's grouping)
    assert abs(c1 - c0) <= 1e-12 * abs(c0)
    assert np.abs(H1 - H0).max() <= 1e-11 * np.abs(H0).max() and np.abs(g1 - g0).max() <= 1e-11 * np.abs(g0).max()
    # and the point-to-plane family alone gives something else: the line blocks are really in the system
    cfg_p = odometry.OdometryConfig(line_to_line=False, angle_residual=angle)
    _, n_plane, Hp, _, _ = system_of(gpu_ctx, frames, poses, cfg_p, oracle.aa_to_R, True)
    assert n_plane == n_host - n_line_host and np.abs(Hp - H0).max() > 1e-6 * np.abs(H0).max()


@pytest.mark.gpu
def test_refine_pose_with_device_line_blocks(gpu_ctx, oracle):
    from panovlm_b200 import odometry
    frames, poses = perturbed_sequence(oracle, 8, 29)
    cfg = odometry.OdometryConfig()
    exp_poses, exp = odometry.refine_pose(gpu_ctx, frames, poses, cfg, oracle.aa_to_R, device_line_blocks=False)
    got_poses, got = odometry.refine_pose(gpu_ctx, frames, poses, cfg, oracle.aa_to_R, device_line_blocks=True)
    assert got["n_blocks"] == exp["n_blocks"] and got["iterations"] == exp["iterations"] and got["termination"] == exp["termination"]
    assert abs(got["final_cost"] - exp["final_cost"]) <= 1e-9 * exp["final_cost"]
    assert np.abs(got_poses - exp_poses).max() < 1e-9
    # a pending list is dropped by the next vote pass: a host-tail build after a device-tail call does not pick the old blocks up
    lf_edges = odometry.pose_graph_edges(poses, cfg, oracle.aa_to_R)
    bl, _, mine = odometry.build_problem(gpu_ctx, frames, poses, cfg, oracle.aa_to_R, host_point2plane=False, device_line_blocks=True)
    bl2, _, _ = odometry.build_problem(gpu_ctx, frames, poses, cfg, oracle.aa_to_R, host_point2plane=False, device_line_blocks=False)
    ref, nei = np.array([e[0] for e in mine], np.int32), np.array([e[1] for e in mine], np.int32)
    n = gpu_ctx.frames_point2plane_blocks(poses, ref, nei, cfg.plane_tolerance, cfg.plane_dis_threshold, cfg.angle_residual, cfg.normalize_distance, cfg.plane_weight, len(frames),
                                          extra=bl2.view())
    assert n == exp["n_blocks"] and len(lf_edges) == exp["n_edges"]
