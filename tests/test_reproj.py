"""Camera-camera reprojection residuals and their bundle adjustment: PanoramaReprojResidual_1Angle
(base/CostFunction.h:218-247), AddCameraResidual (util/Optimization.cpp:172-222), SfMGlobalBA (util/Optimization.cpp:10-82).
CPU: the oracle (Jet<9> autodiff, dense normal equations, dense LM) is pinned against a torch float64 autograd twin, finite differences and
a numpy Schur-complement solve; the host-side observation builder against a numpy restatement.  GPU: rows, reduced blocks and the
Schur-complement LM through the C ABI vs the oracle."""
import numpy as np
import pytest
import torch

import panovlm_b200
from panovlm_b200 import synth

HUBER = 4.0 * np.pi / 180.0


def _twin(cams, pts, cam, point, bearing, weight):
    """independent twin: torch float64 reverse-mode autograd over the closed-form Rodrigues rotation"""
    r, J = np.zeros(len(cam)), np.zeros((len(cam), 9))
    for i in range(len(cam)):
        aa = torch.tensor(cams[cam[i], :3], requires_grad=True)
        t = torch.tensor(cams[cam[i], 3:], requires_grad=True)
        X = torch.tensor(pts[point[i]], requires_grad=True)
        th = torch.linalg.norm(aa)
        k = aa / th
        K = torch.zeros(3, 3, dtype=torch.float64)
        K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -k[2], k[1], k[2], -k[0], -k[1], k[0]
        R = torch.eye(3, dtype=torch.float64) + torch.sin(th) * K + (1 - torch.cos(th)) * (K @ K)
        P = R @ X + t
        s = torch.tensor(bearing[i] / np.linalg.norm(bearing[i]))
        res = weight * torch.acos((P @ s) / torch.linalg.norm(P))
        res.backward()
        r[i] = res.item()
        J[i] = np.concatenate([aa.grad.numpy(), t.grad.numpy(), X.grad.numpy()])
    return r, J


def test_oracle_functor_matches_autograd_twin_and_finite_differences(oracle):
    d = synth.make_ba_problem(n_cams=5, n_points=40, seed=3)
    bearing = d["bearing"] * 1.7                        # the functor normalises the bearing itself
    R = oracle.Reproj(d["cam"], d["point"], bearing, weight=1.3)
    r, J, cost = R.evaluate(d["cams"], d["points"], apply_loss=False)
    r2, J2 = _twin(d["cams"], d["points"], d["cam"], d["point"], bearing, 1.3)
    assert np.abs(r - r2).max() < 1e-12
    assert (np.abs(J - J2).max(1) / np.abs(J2).max(1)).max() < 1e-8
    assert np.allclose(cost, 0.5 * r * r)
    # central finite differences on the packed parameter vector
    h = 1e-6
    for k in range(9):
        num = np.zeros(len(r))
        for i in range(len(r)):
            cp2, pp2, cm2, pm2 = d["cams"].copy(), d["points"].copy(), d["cams"].copy(), d["points"].copy()
            if k < 6:
                cp2[d["cam"][i], k] += h; cm2[d["cam"][i], k] -= h
            else:
                pp2[d["point"][i], k - 6] += h; pm2[d["point"][i], k - 6] -= h
            one = oracle.Reproj(d["cam"][i:i + 1], d["point"][i:i + 1], bearing[i:i + 1], weight=1.3)
            num[i] = (one.evaluate(cp2, pp2, False, False)[0][0] - one.evaluate(cm2, pm2, False, False)[0][0]) / (2 * h)
        assert (np.abs(num - J[:, k]) / np.maximum(1e-3, np.abs(J).max(1))).max() < 1e-5
    # closed form: the residual is the angle between the bearing and the direction of the point in the camera frame
    Pc = np.stack([synth.rotvec_to_R(d["cams"][c, :3]) @ d["points"][p] + d["cams"][c, 3:] for c, p in zip(d["cam"], d["point"])])
    ang = np.arccos(np.sum(Pc * d["bearing"], 1) / np.linalg.norm(Pc, axis=1))
    assert np.abs(r - 1.3 * ang).max() < 1e-12
    # Huber corrector: cost = rho(s)/2 and the row is scaled by sqrt(rho')
    rl, Jl, cl = oracle.Reproj(d["cam"], d["point"], bearing, weight=40.0, huber=HUBER).evaluate(d["cams"], d["points"], apply_loss=True)
    r0 = 40.0 / 1.3 * r
    out = np.abs(r0) > HUBER
    assert out.any() and (~out).any()
    assert np.allclose(cl[out], 0.5 * (2 * HUBER * np.abs(r0[out]) - HUBER ** 2)) and np.allclose(cl[~out], 0.5 * r0[~out] ** 2)
    assert np.allclose(rl[out], r0[out] * np.sqrt(HUBER / np.abs(r0[out])))


def test_oracle_lm_equals_a_schur_complement_solve_and_converges(oracle):
    d = synth.make_ba_problem(n_cams=6, n_points=60, seed=5)
    R = oracle.Reproj(d["cam"], d["point"], d["bearing"], huber=HUBER)
    nc, npt = 6, 60
    H, g, cost = R.normal_equations(d["cams"], d["points"])
    # the damped step of the full system equals eliminate-the-points-first (what DENSE_SCHUR does)
    free = np.arange(6, 6 * nc + 3 * npt)
    Hf, gf = H[np.ix_(free, free)], g[free]
    Hd = Hf + np.diag(np.clip(np.diag(Hf), 1e-6, 1e32)) / 1e4
    full = np.linalg.solve(Hd, -gf)
    ncf = 6 * (nc - 1)
    B, E, Cc = Hd[:ncf, :ncf], Hd[:ncf, ncf:], Hd[ncf:, ncf:]
    Ci = np.linalg.inv(Cc)
    yc = np.linalg.solve(B - E @ Ci @ E.T, -gf[:ncf] + E @ Ci @ gf[ncf:])
    yp = -Ci @ (gf[ncf:] + E.T @ yc)
    assert np.abs(np.concatenate([yc, yp]) - full).max() < 1e-7 * np.abs(full).max()
    mask = np.zeros(6 * nc + 3 * npt, np.uint8)
    mask[:6] = 1
    cams, pts, summ = R.solve_lm(d["cams"], d["points"], mask, max_iter=30)
    assert summ["final_cost"] < 0.2 * summ["initial_cost"] and summ["successful"] >= 3
    assert np.array_equal(cams[0], d["cams"][0])


def test_observation_builder_matches_numpy_restatement():
    d = synth.make_ba_problem(n_cams=7, n_points=50, seed=6)
    # tracks as CSR (observations are point-major already)
    track_off = np.concatenate([[0], np.cumsum(np.bincount(d["point"], minlength=50))]).astype(np.int32)
    pose_valid = np.ones(7, np.uint8)
    pose_valid[3] = 0
    cam, point, bearing = panovlm_b200.Context.build_reproj_observations(d["rows"], d["cols"], track_off, d["cam"], d["pixels"], pose_valid)
    keep = d["cam"] != 3
    assert np.array_equal(cam, d["cam"][keep]) and np.array_equal(point, d["point"][keep])
    # float32 restatement of ImageToSphere / SphereToCam on the rounded pixel
    px = np.rint(d["pixels"][keep]).astype(np.float32)
    lon = ((np.float32(2) * px[:, 0] / np.float32(d["cols"]) - np.float32(1)).astype(np.float64) * np.pi).astype(np.float32)
    lat = ((0.5 - (px[:, 1] / np.float32(d["rows"])).astype(np.float64)) * np.pi).astype(np.float32)
    cy = np.cos(lat)
    exp = np.stack([cy * np.sin(lon), -np.sin(lat), cy * np.cos(lon)], 1)
    assert exp.dtype == np.float32
    assert np.abs(bearing - exp.astype(np.float64)).max() < 2e-7          # libm float vs numpy float: 1 ulp
    assert np.abs(np.sum(bearing * d["bearing"][keep], 1) - 1).max() < 1e-5  # within the rounding to a pixel of the true bearing
    # empty input
    c0, p0, b0 = panovlm_b200.Context.build_reproj_observations(10, 20, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros((0, 2), np.float32))
    assert len(c0) == 0 and len(p0) == 0 and b0.shape == (0, 3)


def _blocks_from_dense(H, g, d, nc, npt):
    B = np.stack([H[6 * c:6 * c + 6, 6 * c:6 * c + 6][np.triu_indices(6)] for c in range(nc)])
    gc = g[:6 * nc].reshape(nc, 6)
    o = 6 * nc
    Cp = np.stack([H[o + 3 * p:o + 3 * p + 3, o + 3 * p:o + 3 * p + 3][np.triu_indices(3)] for p in range(npt)])
    gp = g[o:].reshape(npt, 3)
    return B, gc, Cp, gp


@pytest.mark.gpu
def test_reproj_rows_and_blocks_match_oracle(gpu_ctx, oracle):
    d = synth.make_ba_problem(n_cams=9, n_points=400, seed=7)
    nc, npt = 9, 400
    for weight, huber in ((1.0, HUBER), (35.0, HUBER), (2.0, 0.0)):
        R = oracle.Reproj(d["cam"], d["point"], d["bearing"] * 0.8, weight=weight, huber=huber)
        r, J, cost = R.evaluate(d["cams"], d["points"], apply_loss=True)
        gpu_ctx.reproj_set(d["cam"], d["point"], d["bearing"] * 0.8, nc, npt, weight=weight, huber=huber)
        gpu_ctx.reproj_evaluate(d["cams"], d["points"], True, True)
        r2, J2 = gpu_ctx.reproj_rows()
        assert (np.abs(r2 - r) / np.maximum(1e-9, np.abs(r))).max() < 1e-7             # gate: 1e-5 relative (acos amplifies rounding at small angles)
        assert (np.abs(J2 - J).max(1) / np.maximum(1e-12, np.abs(J).max(1))).max() < 1e-7   # gate: 1e-6 relative
        assert abs(gpu_ctx.reproj_cost() - cost.sum()) < 1e-10 * max(1.0, cost.sum())
        H, g, _ = R.normal_equations(d["cams"], d["points"])
        B, gc, Cp, gp = _blocks_from_dense(H, g, d, nc, npt)
        B2, gc2, Cp2, gp2, E2 = gpu_ctx.reproj_blocks()
        for a, b in ((B, B2), (gc, gc2), (Cp, Cp2), (gp, gp2)):
            assert np.abs(a - b).max() < 1e-9 * max(1.0, np.abs(a).max())
        # coupling blocks: E_o = Jc^T Jp; their sum over the observations of (camera, point) is the off-diagonal block of H
        assert np.abs(E2 - J[:, :6, None] * J[:, None, 6:]).max() < 1e-9 * max(1.0, np.abs(E2).max())
    # the rows come back in the caller's order even when it is not point-major
    perm = np.random.default_rng(0).permutation(len(d["cam"]))
    gpu_ctx.reproj_set(d["cam"][perm], d["point"][perm], d["bearing"][perm], nc, npt, weight=2.0, huber=0.0)
    gpu_ctx.reproj_evaluate(d["cams"], d["points"], True, False)
    r3, J3 = gpu_ctx.reproj_rows()
    R = oracle.Reproj(d["cam"], d["point"], d["bearing"], weight=2.0, huber=0.0)
    r, J, _ = R.evaluate(d["cams"], d["points"])
    assert (np.abs(r3 - r[perm]) / np.maximum(1e-9, np.abs(r[perm]))).max() < 1e-7
    assert np.abs(J3 - J[perm]).max() < 1e-7 * np.abs(J).max()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["all", "fixed_structure", "fixed_rotations", "some_points_fixed"])
def test_reproj_schur_lm_matches_dense_oracle_lm(gpu_ctx, oracle, mode):
    """SfMGlobalBA's solve: the Schur-complement LM on the device follows the oracle's dense LM over [cameras | points] step for step."""
    d = synth.make_ba_problem(n_cams=8, n_points=120, seed=8)
    nc, npt = 8, 120
    cam_const = np.zeros((nc, 6), np.uint8)
    cam_const[0] = 1                                              # first valid camera constant (Optimization.cpp:50-57)
    pt_const = np.zeros(npt, np.uint8)
    if mode == "fixed_structure":
        pt_const[:] = 1                                           # refine_structure = false (:45-47)
    if mode == "fixed_rotations":
        cam_const[:, :3] = 1                                      # refine_rotation = false (:40-41)
    if mode == "some_points_fixed":
        pt_const[::3] = 1
    mask = np.concatenate([cam_const.ravel(), np.repeat(pt_const, 3)])
    R = oracle.Reproj(d["cam"], d["point"], d["bearing"], huber=HUBER)
    e_c, e_p, e_s = R.solve_lm(d["cams"], d["points"], mask, max_iter=25)
    gpu_ctx.reproj_set(d["cam"], d["point"], d["bearing"], nc, npt, huber=HUBER)
    g_c, g_p, g_s = gpu_ctx.reproj_solve_lm(d["cams"], d["points"], cam_const, pt_const, max_iterations=25)
    for k in ("iterations", "successful", "unsuccessful", "termination"):
        assert e_s[k] == g_s[k], (k, e_s, g_s)
    assert abs(e_s["initial_cost"] - g_s["initial_cost"]) < 1e-10 * e_s["initial_cost"]
    # eliminating the points first and solving the dense system directly round differently on the ill-conditioned damped systems
    # (radius 1e4 and beyond): the two trajectories drift apart at the 1e-7 level over 25 iterations
    assert abs(e_s["final_cost"] - g_s["final_cost"]) < 1e-5 * e_s["final_cost"]
    dc, dp = e_c - d["cams"], e_p - d["points"]
    assert np.abs(g_c - e_c).max() < 1e-4 * max(np.abs(dc).max(), 1e-12)      # pose deltas: 1e-4 relative (BASELINE.json)
    assert np.abs(g_p - e_p).max() < 1e-4 * max(np.abs(dp).max(), 1e-12)
    assert np.array_equal(g_c[cam_const.astype(bool)], d["cams"][cam_const.astype(bool)])
    assert np.array_equal(g_p[pt_const.astype(bool)], d["points"][pt_const.astype(bool)])
    assert g_s["final_cost"] < g_s["initial_cost"]


@pytest.mark.gpu
def test_reproj_room_scale_bundle_adjustment(gpu_ctx):
    """Room-shaped joint problem (BASELINE.json configs[2]): 454 panoramas, 40 k points, ~200 k observations.  No dense oracle at this size:
    size-independent properties — the cost falls, the gradient at the solution is tiny against the start, the constant blocks stay put and
    two runs are bit-identical (every reduction has a fixed order)."""
    d = synth.make_ba_problem(n_cams=454, n_points=40000, track_len=(3, 8), seed=9)
    nc, npt = 454, 40000
    cam_const = np.zeros((nc, 6), np.uint8)
    cam_const[0] = 1
    gpu_ctx.reproj_set(d["cam"], d["point"], d["bearing"], nc, npt, huber=HUBER)
    c1, p1, s1 = gpu_ctx.reproj_solve_lm(d["cams"], d["points"], cam_const, None, max_iterations=20)
    c2, p2, s2 = gpu_ctx.reproj_solve_lm(d["cams"], d["points"], cam_const, None, max_iterations=20)
    assert s1 == s2 and np.array_equal(c1, c2) and np.array_equal(p1, p2)
    assert s1["final_cost"] < 0.05 * s1["initial_cost"]
    assert np.array_equal(c1[0], d["cams"][0])
    # closer to the ground truth than the perturbed start (the gauge is pinned by camera 0 only up to scale: compare rotations)
    assert np.abs(c1[:, :3] - d["cams_gt"][:, :3]).mean() < 0.3 * np.abs(d["cams"][:, :3] - d["cams_gt"][:, :3]).mean()


def _joint_problem(nc=6, npt=150, n_blocks=1500, seed=21):
    """Pose blocks [cameras 0..nc) | LiDARs nc..2nc): a BA problem on the cameras plus plane / line residual blocks between random pairs of all pose
    blocks (the LiDAR-LiDAR and camera-LiDAR families of the joint problem), consistent with one ground truth."""
    from scipy.spatial.transform import Rotation
    d = synth.make_ba_problem(n_cams=nc, n_points=npt, seed=seed)
    rng = np.random.default_rng(seed + 1)
    nb = 2 * nc
    truth = np.concatenate([d["cams_gt"], np.concatenate([rng.normal(0, 0.3, (nc, 3)), rng.normal(0, 1.0, (nc, 3))], 1)])
    typ = rng.integers(0, 4, n_blocks).astype(np.int32)
    ref = rng.integers(0, nb, n_blocks).astype(np.int32)
    nei = ((ref + rng.integers(1, nb, n_blocks)) % nb).astype(np.int32)
    consts = np.zeros((n_blocks, 12))
    for i in range(n_blocks):
        pw = rng.normal(0, 4, 3)
        Rr, tr = Rotation.from_rotvec(truth[ref[i], :3]).as_matrix(), truth[ref[i], 3:]
        Rn, tn = Rotation.from_rotvec(truth[nei[i], :3]).as_matrix(), truth[nei[i], 3:]
        p_ref, p_nei = Rr @ pw + tr, Rn @ pw + tn
        consts[i, :3] = p_nei + rng.normal(0, 0.01, 3)
        if typ[i] < 2:
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            dd = -nrm @ p_ref
            if dd < 0:
                nrm, dd = -nrm, -dd
            consts[i, 3:6] = nrm; consts[i, 6] = dd; consts[i, 7] = 0.05
        else:
            dr = rng.normal(size=3); dr /= np.linalg.norm(dr)
            consts[i, 3:6] = p_ref + 0.3 * dr; consts[i, 6:9] = dr; consts[i, 9] = 0.05
    hub = np.where(typ % 2 == 1, 2 * np.pi / 180, 0.2)
    start = truth.copy()
    start[:nc] = d["cams"]
    start[nc:] += np.concatenate([rng.normal(0, 0.01, (nc, 3)), rng.normal(0, 0.03, (nc, 3))], axis=1)
    return d, nb, (typ, ref, nei, consts, hub), start


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["all", "fixed_camera_rotations", "fixed_structure"])
def test_joint_camera_lidar_lm_matches_dense_oracle_lm(gpu_ctx, oracle, mode):
    """CameraLidarOptimizer::Optimize's solve: reprojection blocks + pose-graph blocks in one trust-region problem, points eliminated on the device."""
    nc, npt = 6, 150
    d, nb, (typ, ref, nei, consts, hub), start = _joint_problem(nc, npt)
    pose_const = np.zeros((nb, 6), np.uint8)
    pose_const[0] = 1                                             # camera 0 constant (CameraLidarOptimizer.cpp:490-491)
    pt_const = np.zeros(npt, np.uint8)
    if mode == "fixed_camera_rotations":
        pose_const[:nc, :3] = 1                                   # refine_camera_rotation = false (:469-474)
    if mode == "fixed_structure":
        pt_const[:] = 1                                           # refine_structure = false (:462-465)
    blk = oracle.Blocks(typ, ref, nei, consts, hub, 1)
    rep = oracle.Reproj(d["cam"], d["point"], d["bearing"], huber=HUBER)
    mask = np.concatenate([pose_const.ravel(), np.repeat(pt_const, 3)])
    e_x, e_p, e_s = oracle.joint_solve_lm(blk, rep, start, d["points"], mask, max_iter=15)
    gpu_ctx.blocks_set(typ, ref, nei, consts, hub, 1, nb)
    gpu_ctx.reproj_set(d["cam"], d["point"], d["bearing"], nc, npt, huber=HUBER)
    g_x, g_p, g_s = gpu_ctx.joint_solve_lm(start, d["points"], pose_const, pt_const, max_iterations=15)
    for k in ("iterations", "successful", "unsuccessful", "termination"):
        assert e_s[k] == g_s[k], (k, e_s, g_s)
    assert abs(e_s["initial_cost"] - g_s["initial_cost"]) < 1e-10 * e_s["initial_cost"]
    assert abs(e_s["final_cost"] - g_s["final_cost"]) < 1e-5 * e_s["final_cost"]
    assert e_s["final_cost"] < 0.5 * e_s["initial_cost"]
    assert np.abs(g_x - e_x).max() < 1e-4 * np.abs(e_x - start).max()          # pose deltas: 1e-4 relative (BASELINE.json)
    if mode != "fixed_structure":
        assert np.abs(g_p - e_p).max() < 1e-4 * np.abs(e_p - d["points"]).max()
    assert np.array_equal(g_x[pose_const.astype(bool)], start[pose_const.astype(bool)])
    assert np.array_equal(g_p[pt_const.astype(bool)], d["points"][pt_const.astype(bool)])


GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "reproj.npz")


def test_oracle_against_committed_golden_vectors(oracle):
    """tests/golden/reproj.npz (tests/make_golden.py: torch float64 autograd twin, scipy rotation-vector interpolation) pins the oracle."""
    g = np.load(GOLDEN)
    r, J, _ = oracle.Reproj(g["cam"], g["point"], g["bearing"], weight=float(g["weight"]), huber=0.0).evaluate(g["cams"], g["points"], apply_loss=False)
    assert np.abs(r - g["residual"]).max() < 1e-10                       # acos at small angles amplifies the rounding of two different evaluation orders
    assert (np.abs(J - g["jacobian"]).max(1) / np.abs(g["jacobian"]).max(1)).max() < 1e-8
    und = oracle.undistort_cloud(g["T_wl"][:3, :3], g["T_wl"][:3, 3], g["T_we"][:3, :3], g["T_we"][:3, 3], g["sweep"])
    exp = g["undistorted"].astype(np.float32)
    ulp = np.abs(und[:, :3].view(np.int32).astype(np.int64) - exp.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp > 0).mean() < 0.01                    # quaternion slerp vs rotation-vector scaling: equal up to a float32 rounding flip
    assert np.array_equal(und[:, 3], g["sweep"][:, 3])


@pytest.mark.gpu
def test_kernels_against_committed_golden_vectors(gpu_ctx):
    g = np.load(GOLDEN)
    gpu_ctx.reproj_set(g["cam"], g["point"], g["bearing"], len(g["cams"]), len(g["points"]), weight=float(g["weight"]), huber=0.0)
    gpu_ctx.reproj_evaluate(g["cams"], g["points"], True, False)
    r, J = gpu_ctx.reproj_rows()
    assert (np.abs(r - g["residual"]) / np.maximum(1e-9, np.abs(g["residual"]))).max() < 1e-7        # gate: 1e-5 relative
    assert (np.abs(J - g["jacobian"]).max(1) / np.abs(g["jacobian"]).max(1)).max() < 1e-7            # gate: 1e-6 relative
    und = gpu_ctx.undistort_clouds(g["sweep"], np.array([0, len(g["sweep"])], np.int32), g["T_wl"][None], g["T_we"][None])
    exp = g["undistorted"].astype(np.float32)
    ulp = np.abs(und[:, :3].view(np.int32).astype(np.int64) - exp.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp > 0).mean() < 0.01
