"""Sweep undistortion and pose interpolation: Velodyne::UndistortCloud (sensors/Velodyne.cpp:1642-1674), SlerpPose
(base/Geometry.hpp:572-583) and the sweep-end pose selection of LidarOdometry::UndistortLidars (lidar_mapping/LidarOdometry.cpp:203-243).
CPU: the oracle is pinned against scipy.spatial.transform; the product's per-point math (compiled for the host) and its host functions are
compared with the oracle.  GPU: the batched kernel through the C ABI vs the oracle, plus exact properties at 10 M points."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

import panovlm_b200

p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731


def _pose(rng, rot=0.3, trans=1.0):
    T = np.eye(4)
    T[:3, :3] = Rot.from_rotvec(rng.normal(size=3) * rot).as_matrix()
    T[:3, 3] = rng.normal(size=3) * trans
    return T


def _trajectory(rng, n):
    out, T = [], np.eye(4)
    for _ in range(n):
        T = T @ _pose(rng, 0.05, 0.3)
        out.append(T.copy())
    return np.stack(out)


def test_oracle_slerp_pose_matches_scipy(oracle):
    rng = np.random.default_rng(11)
    for _ in range(20):
        P1, P2 = _pose(rng), _pose(rng)
        for ratio in (0.0, 0.25, 0.7, 1.0, 1.3):
            T21 = np.linalg.inv(P2) @ P1
            Ts = np.eye(4)
            Ts[:3, :3] = Rot.from_rotvec(Rot.from_matrix(T21[:3, :3]).as_rotvec() * ratio).as_matrix()
            Ts[:3, 3] = T21[:3, 3] * ratio
            assert np.abs(oracle.slerp_pose(P1, P2, ratio) - P1 @ np.linalg.inv(Ts)).max() < 1e-13
    # end points: ratio 0 keeps pose 1, ratio 1 gives pose 2
    assert np.abs(oracle.slerp_pose(P1, P2, 0.0) - P1).max() < 1e-14
    assert np.abs(oracle.slerp_pose(P1, P2, 1.0) - P2).max() < 1e-13


def test_oracle_undistort_matches_scipy(oracle):
    rng = np.random.default_rng(12)
    for n in (1, 7, 2000):
        A, E = _pose(rng), _pose(rng)
        cloud = (rng.normal(size=(n, 4)) * 10).astype(np.float32)
        out = oracle.undistort_cloud(A[:3, :3], A[:3, 3], E[:3, :3], E[:3, 3], cloud)
        R_se, t_se = A[:3, :3].T @ E[:3, :3], A[:3, :3].T @ (E[:3, 3] - A[:3, 3])
        rv = Rot.from_matrix(R_se).as_rotvec()
        ratio = (np.arange(n, dtype=np.float32) / np.float32(n)).astype(np.float64)
        exp = np.stack([Rot.from_rotvec(rv * r).apply(cloud[i, :3].astype(np.float64)) + r * t_se for i, r in enumerate(ratio)])
        assert np.abs(out[:, :3] - exp).max() < 4e-6                 # float32 store of values ~ 30
        assert np.array_equal(out[:, 3], cloud[:, 3])
        assert np.array_equal(out[0, :3], cloud[0, :3])             # ratio 0: the first point of the sweep does not move


def test_device_math_on_host_is_bit_identical_to_oracle(oracle, harness):
    rng = np.random.default_rng(13)
    for n, rot in ((5000, 0.3), (3000, 0.0), (3000, 3.0)):           # rot = 0: Eigen's linear branch of slerp; 3.0: trace <= 0 branch of the quaternion
        A, E = _pose(rng, rot), _pose(rng, rot)
        if rot == 0.0:
            E[:3, :3] = A[:3, :3]
        cloud = (rng.normal(size=(n, 4)) * 20).astype(np.float32)
        ref = oracle.undistort_cloud(A[:3, :3], A[:3, 3], E[:3, :3], E[:3, 3], cloud)
        out = np.empty_like(cloud)
        harness.pvbh_undistort_cloud(p(np.ascontiguousarray(A[:3, :3])), p(np.ascontiguousarray(A[:3, 3])), p(np.ascontiguousarray(E[:3, :3])),
                                     p(np.ascontiguousarray(E[:3, 3])), p(cloud), C.c_long(n), p(out))
        assert np.array_equal(out, ref)


def test_host_pose_functions_match_oracle(oracle):
    rng = np.random.default_rng(14)
    for _ in range(10):
        P1, P2 = _pose(rng), _pose(rng)
        for ratio in (0.0, 0.1, 0.5, 1.0, 1.2):
            assert np.abs(panovlm_b200.Context.slerp_pose(P1, P2, ratio) - oracle.slerp_pose(P1, P2, ratio)).max() < 1e-13
    for n in (0, 1, 2, 9):
        poses = _trajectory(rng, n) if n else np.zeros((0, 4, 4))
        for trial in range(6):
            pv = (rng.random(n) > 0.3).astype(np.uint8) if trial else np.ones(n, np.uint8)
            fv = (rng.random(n) > 0.3).astype(np.uint8) if trial else np.ones(n, np.uint8)
            e_pose, e_has = oracle.undistort_end_poses(poses, pv, fv, 0.02)
            g_pose, g_has = panovlm_b200.Context.undistort_end_poses(poses, pv, fv, 0.02)
            assert np.array_equal(e_has, g_has)
            assert np.abs(e_pose - g_pose).max(initial=0.0) < 1e-12
    # all frames valid: every frame but a 1-frame sequence gets an end pose; the first one lies on the way to frame 1
    poses = _trajectory(rng, 5)
    g_pose, g_has = panovlm_b200.Context.undistort_end_poses(poses, np.ones(5, np.uint8), np.ones(5, np.uint8), 0.0)
    assert g_has.all()
    assert np.abs(g_pose[0] - poses[1]).max() < 1e-12              # gap 0: the sweep ends exactly at the next frame's pose


@pytest.mark.gpu
def test_undistort_clouds_kernel_matches_oracle(gpu_ctx, oracle):
    rng = np.random.default_rng(15)
    sizes = [28800, 0, 1, 255, 256, 257, 12345]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    cloud = (rng.normal(size=(off[-1], 4)) * 15).astype(np.float32)
    T_wl = np.stack([_pose(rng) for _ in sizes])
    T_we = np.stack([T_wl[f] @ _pose(rng, 0.05, 0.2) for f in range(len(sizes))])
    T_we[3] = T_wl[3]                                               # identity motion
    T_we[4][:3, :3] = T_wl[4][:3, :3]                               # pure translation (linear slerp branch)
    has = np.array([1, 1, 1, 1, 1, 0, 1], np.uint8)
    out = gpu_ctx.undistort_clouds(cloud, off, T_wl, T_we, has)
    bad = 0
    for f, n in enumerate(sizes):
        a = cloud[off[f]:off[f + 1]]
        ref = oracle.undistort_cloud(T_wl[f][:3, :3], T_wl[f][:3, 3], T_we[f][:3, :3], T_we[f][:3, 3], a) if has[f] else a
        got = out[off[f]:off[f + 1]]
        assert np.array_equal(got[:, 3], a[:, 3])
        d = np.abs(got[:, :3].view(np.int32).astype(np.int64) - ref[:, :3].view(np.int32).astype(np.int64))
        assert d.max(initial=0) <= 1                                # device sin() vs glibc: at most a float32 rounding flip
        bad += int((d > 0).sum())
    assert bad <= 1e-4 * 3 * off[-1]
    assert np.array_equal(out[off[3]:off[4]], cloud[off[3]:off[4]])  # identity motion leaves the sweep untouched
    assert np.array_equal(out[off[5]:off[6]], cloud[off[5]:off[6]])  # no end pose: passes through
    # has_end = NULL means every frame is undistorted
    out2 = gpu_ctx.undistort_clouds(cloud[:off[1]], off[:2], T_wl[:1], T_we[:1])
    assert np.array_equal(out2, out[:off[1]])


@pytest.mark.gpu
def test_undistort_exact_properties_at_full_size(gpu_ctx):
    """10 M points in 64 sweeps (BASELINE.json configs[4] shape): a pure translation has a closed form in float64 -> float32 that the kernel
    must hit bit for bit; the identity motion must return the input."""
    rng = np.random.default_rng(16)
    n_frames, per = 64, 156250
    off = (np.arange(n_frames + 1) * per).astype(np.int32)
    cloud = (rng.normal(size=(n_frames * per, 4)) * 30).astype(np.float32)
    # rotations that are exact in floating point (signed axis permutations), so that R_wl^T R_we is the identity bit for bit
    perms = [np.eye(3)[list(pm)] * np.array(sg)[:, None] for pm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)) for sg in ((1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1))]
    T_wl = np.stack([_pose(rng) for _ in range(n_frames)])
    for f in range(n_frames):
        T_wl[f][:3, :3] = perms[f % len(perms)]
    T_we = T_wl.copy()
    shift = rng.normal(size=(n_frames, 3))
    T_we[:, :3, 3] += shift
    out = gpu_ctx.undistort_clouds(cloud, off, T_wl, T_we)
    ratio = (np.arange(per, dtype=np.float32) / np.float32(per)).astype(np.float64)
    for f in (0, 17, 63):
        t_se = T_wl[f][:3, :3].T @ (T_we[f][:3, 3] - T_wl[f][:3, 3])
        a = cloud[off[f]:off[f + 1], :3].astype(np.float64)
        assert np.array_equal(out[off[f]:off[f + 1], :3], (a + ratio[:, None] * t_se[None, :]).astype(np.float32))
    assert np.array_equal(gpu_ctx.undistort_clouds(cloud, off, T_wl, T_wl), cloud)


def test_pose_text_round_trip_and_format(tmp_path):
    """ReadPoseT / ExportPoseT (util/FileIO.cpp:11-73, 168-191): 12 numbers per line with 6 significant digits, optional name, inf / nan lines."""
    rng = np.random.default_rng(17)
    n = 7
    R = np.stack([Rot.from_rotvec(rng.normal(size=3)).as_matrix() for _ in range(n)])
    t = rng.normal(size=(n, 3)) * 100
    t[3] = np.inf                                                      # a frame without pose (ReadPoseT's marker)
    names = [f"frame_{i:04d}.pcd" for i in range(n)]
    path = tmp_path / "poses.txt"
    panovlm_b200.Context.write_poses_text(path, R, t, names)
    lines = path.read_text().splitlines()
    assert len(lines) == n
    for i, ln in enumerate(lines):                                     # exactly what `out << double` prints: "%g"
        vals = [R[i, 0, 0], R[i, 0, 1], R[i, 0, 2], t[i, 0], R[i, 1, 0], R[i, 1, 1], R[i, 1, 2], t[i, 1], R[i, 2, 0], R[i, 2, 1], R[i, 2, 2], t[i, 2]]
        assert ln == names[i] + " " + " ".join("%g" % v for v in vals)
    R2, t2, valid, nm = panovlm_b200.Context.read_poses_text(path, with_invalid=True)
    assert nm == names and valid.tolist() == [True, True, True, False, True, True, True]
    ok = valid
    assert np.abs(R2[ok] - R[ok]).max() < 5e-6 and np.abs(t2[ok] - t[ok]).max() < 5e-4     # the 6-digit precision of the format
    assert np.all(R2[3] == 0) and np.all(np.isinf(t2[3]))
    R3, t3, valid3, nm3 = panovlm_b200.Context.read_poses_text(path, with_invalid=False)
    assert len(R3) == n - 1 and valid3.all() and nm3 == names[:3] + names[4:]
    # files without names
    panovlm_b200.Context.write_poses_text(path, R[:2], t[:2])
    R4, t4, v4, nm4 = panovlm_b200.Context.read_poses_text(path)
    assert len(R4) == 2 and nm4 == ["", ""] and np.abs(R4 - R[:2]).max() < 5e-6
