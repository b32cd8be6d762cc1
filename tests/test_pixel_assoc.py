"""Pixel-space camera-LiDAR association, first stage (joint_optimization/CameraLidarLineAssociate.cpp:22-102): image lines ->
sub-line mid points, LiDAR points -> pixels -> 3 nearest mid points within 60 px -> per-line candidate lists.  CPU: the oracle against an
independent scipy cKDTree + numpy restatement, the product's host functions against the oracle.  GPU: the kernel through the C ABI, bit-exact."""
import numpy as np
import pytest
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

import panovlm_b200
from panovlm_b200 import synth

ROWS, COLS = 2880, 5760


def _scene(seed=31, n_az=900):
    fr = synth.make_pair(seed=20260925 + seed, n_az=n_az)[0]
    rng = np.random.default_rng(seed)
    T = np.eye(4)
    T[:3, :3] = Rotation.from_rotvec([0.01, 0.02, -0.01]).as_matrix()
    T[:3, 3] = [0.03, -0.05, 0.02]
    return fr, T, rng


def _lines(oracle, fr, T, rng, clutter=40):
    ends_cam = fr["end_points"].reshape(-1, 3) @ T[:3, :3].T + T[:3, 3]
    px = oracle.cam_to_image(ROWS, COLS, ends_cam).reshape(-1, 4) + rng.normal(0, 2, (len(fr["end_points"]), 4))
    c = np.stack([rng.uniform(0, COLS, clutter), rng.uniform(0, ROWS, clutter), rng.uniform(0, COLS, clutter), rng.uniform(0, ROWS, clutter)], axis=1)
    seam = np.array([[COLS - 40.0, 900.0, 60.0, 1100.0], [30.0, 200.0, COLS - 20.0, 260.0]])          # lines across the +-pi seam
    return np.concatenate([px, c, seam]).astype(np.float32)


def test_oracle_pixel_neighbors_match_kdtree_restatement(oracle):
    fr, T, rng = _scene()
    lines = _lines(oracle, fr, T, rng)
    cloud = fr["cloud"][::7]
    line3, d2, px = oracle.pixel_line_neighbors(ROWS, COLS, lines, cloud, T)
    mid, s2l = oracle.pixel_sub_lines(ROWS, COLS, lines)
    assert len(mid) > len(lines) and s2l.max() == len(lines) - 1
    # independent: pcl transform (oracle, pinned elsewhere) + CamToImage (pinned by tests/golden/fast_atan2.npz) + an exact kd-tree
    cam = oracle.transform_cloud(T[:3, :3], T[:3, 3], cloud)
    assert np.array_equal(px, oracle.cam_to_image(ROWS, COLS, np.ascontiguousarray(cam[:, :3])))
    dist, idx = cKDTree(mid.astype(np.float64)).query(px.astype(np.float64), k=3)
    exp = np.where(dist ** 2 > 3600.0, -1, s2l[idx])
    # float32 distances vs the float64 tree: identical except where two mid points are closer than float rounding / at the 60 px gate
    assert (exp != line3).mean() < 1e-3
    assert np.abs(np.sqrt(d2.astype(np.float64)) - dist).max() < 1e-2
    assert np.all(np.diff(d2, axis=1) >= 0)
    assert (line3 >= 0).any() and (line3 < 0).any()


def test_host_sub_lines_and_candidate_lists_match_oracle(oracle):
    fr, T, rng = _scene(32)
    lines = _lines(oracle, fr, T, rng)
    mid, s2l = panovlm_b200.Context.pixel_sub_lines(ROWS, COLS, lines)
    mid_o, s2l_o = oracle.pixel_sub_lines(ROWS, COLS, lines)
    assert np.array_equal(mid, mid_o) and np.array_equal(s2l, s2l_o)
    line3, _, _ = oracle.pixel_line_neighbors(ROWS, COLS, lines, fr["cloud"][::5], T)
    off, idx = panovlm_b200.Context.pixel_line_candidates(len(lines), line3, 6)
    # restatement of :62-97 with Python containers
    line_lidar = {}
    for i, row in enumerate(line3):
        for l in row:
            if l >= 0:
                line_lidar.setdefault(int(l), []).append(i)
    for l in range(len(lines)):
        got = idx[off[l]:off[l + 1]].tolist()
        exp = line_lidar.get(l, [])
        assert got == (exp if len(exp) >= 6 else [])
    assert off[-1] == len(idx) and len(idx) > 0
    # empty inputs
    off0, idx0 = panovlm_b200.Context.pixel_line_candidates(3, np.zeros((0, 3), np.int32), 6)
    assert off0.tolist() == [0, 0, 0, 0] and len(idx0) == 0
    m0, s0 = panovlm_b200.Context.pixel_sub_lines(ROWS, COLS, np.zeros((0, 4), np.float32))
    assert len(m0) == 0 and len(s0) == 0


@pytest.mark.gpu
def test_pixel_line_neighbors_kernel_is_bit_exact(gpu_ctx, oracle):
    fr, T, rng = _scene(33, n_az=1800)
    for clutter, cloud in ((40, fr["cloud"]), (2500, fr["cloud"][::3]), (0, fr["cloud"][:1])):       # 2500 clutter lines: several shared-memory tiles of mid points
        lines = _lines(oracle, fr, T, rng, clutter)
        line3, d2, px = gpu_ctx.pixel_line_neighbors(ROWS, COLS, lines, cloud, T)
        o3, od2, opx = oracle.pixel_line_neighbors(ROWS, COLS, lines, cloud, T)
        assert np.array_equal(px, opx)
        assert np.array_equal(d2, od2)
        assert np.array_equal(line3, o3)
    # no image lines at all: every neighbour is -1
    l3, _, _ = gpu_ctx.pixel_line_neighbors(ROWS, COLS, np.zeros((0, 4), np.float32), fr["cloud"][:100], T)
    assert np.all(l3 == -1)
    # fewer than three mid points
    l3, dd, _ = gpu_ctx.pixel_line_neighbors(ROWS, COLS, np.array([[100.0, 100.0, 120.0, 110.0]], np.float32), fr["cloud"][:100], T)
    o3, od, _ = oracle.pixel_line_neighbors(ROWS, COLS, np.array([[100.0, 100.0, 120.0, 110.0]], np.float32), fr["cloud"][:100], T)
    assert np.array_equal(l3, o3) and np.all(l3[:, 1:] == -1) and np.array_equal(dd, od)


def test_filter_line_pairs_matches_oracle_and_geometry(oracle):
    """CameraLidarLineAssociate::Filter (:628-715): both branches against the oracle, and the angle branch against its geometric meaning."""
    fr, T, rng = _scene(34)
    ends_cam = fr["end_points"].reshape(-1, 2, 3) @ T[:3, :3].T + T[:3, 3]
    S = len(ends_cam)
    px = oracle.cam_to_image(ROWS, COLS, ends_cam.reshape(-1, 3)).reshape(-1, 4)
    # every LiDAR segment paired with (a) its own projection, (b) the projection stretched beyond both ends, (c) a tilted line, (d) a far away line
    own = px + rng.normal(0, 1.0, px.shape)
    centre = (px[:, :2] + px[:, 2:]) / 2
    stretched = np.concatenate([centre + 1.6 * (px[:, :2] - centre), centre + 1.6 * (px[:, 2:] - centre)], axis=1)
    tilted = stretched + np.concatenate([np.full((S, 1), 400.0), np.zeros((S, 3))], axis=1)
    far = stretched + 700.0
    lines = np.concatenate([own, stretched, tilted, far]).astype(np.float32)
    start = np.tile(ends_cam[:, 0], (4, 1)); end = np.tile(ends_cam[:, 1], (4, 1))
    for by_angle, by_length in ((True, True), (True, False), (False, True), (False, False)):
        k_o, a_o = oracle.filter_line_pairs(ROWS, COLS, lines, start, end, by_angle, by_length)
        k_p, a_p = panovlm_b200.Context.filter_line_pairs(ROWS, COLS, lines, start, end, by_angle, by_length)
        assert np.array_equal(k_o, k_p) and np.array_equal(a_o, a_p)
        if not by_angle and not by_length:
            assert k_p.all()
    k, ang = panovlm_b200.Context.filter_line_pairs(ROWS, COLS, lines, start, end, True, False)
    assert k[S:2 * S].mean() > 0.8                               # a longer image line that contains the LiDAR line passes
    assert not k[3 * S:].any()                                   # a line hundreds of pixels away fails
    assert np.all(ang[k] <= 5.0) and k[2 * S:3 * S].mean() < 0.5           # one end moved by 400 px: the great-circle planes differ by more than 5 deg
    # length branch: a segment projected shorter than 100 px is dropped
    short = np.array([[0.0, 0.0, 5.0]]), np.array([[0.02, 0.0, 5.0]])
    kk, _ = panovlm_b200.Context.filter_line_pairs(ROWS, COLS, np.array([[2880.0, 1440.0, 2900.0, 1440.0]], np.float32), short[0], short[1], False, True)
    assert not kk[0]
    k0, a0 = panovlm_b200.Context.filter_line_pairs(ROWS, COLS, np.zeros((0, 4), np.float32), np.zeros((0, 3)), np.zeros((0, 3)), True, True)
    assert len(k0) == 0
