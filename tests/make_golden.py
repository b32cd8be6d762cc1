"""Generates tests/golden/*.npz — the fixtures that pin the oracle (and through it the CUDA path).

The reference ships no golden vectors and cannot be built here, so the expected values come from
INDEPENDENT implementations, never from the oracle itself:
  functors.npz   residuals + 1x12 Jacobians of all six cost functors from tests/twin.py (torch float64 reverse-mode
                 autograd over the closed form P = R_r R_n^T (p - t_n) + t_r)
  functors_f6.npz  same for the calibration-mode functors (Plane2Plane_Relative, PlaneRelativeIOUResidual, Line2Line_Angle)
  rotations.npz  scipy.spatial.transform.Rotation matrices / rotation vectors
  assoc_pair.npz a small 2-frame pair (feature clouds) with the expected point-to-plane associations computed with
                 scipy cKDTree (float64 on the float32 world points) + numpy lstsq / eigh
  fast_atan2.npz dense grid of atan2 values (the polynomial must stay within 1.7e-4 rad of them)
  ref_fast_atan2.npz  outputs of the REFERENCE'S OWN FastAtan2 (base/Math.h compiled where it lies, oracle/_ref, `make -C oracle ref`) in float32 and
                 float64 on a dense angle sweep, random points and the axis / signed-zero / extreme-ratio cases: the one fixture produced by reference code
  ref_functors.npz  residuals + Jacobians + post-constructor constants of all nine functors of the path evaluated by the REFERENCE'S OWN
                 base/CostFunction.h through `Functor::Create(...)->Evaluate()` (compiled where it lies with the stand-in Eigen / Ceres / OpenCV types of
                 oracle/shim: oracle/_ref/libpvo_ref_path.so), same keys as functors.npz so the kernel tests read it the same way
  ref_geometry.npz  FormPlane / FormLine / SlerpPose / PointToLineDistance3D ... of the reference's base/Geometry.hpp and the projection functions of
                 sensors/Equirectangular.{h,cpp} (CamToImage float / double, ImageToCam, BreakToSegments incl. seam crossings), same build
  ref_assoc.npz  correspondences returned by the reference's own lidar_mapping/LidarFeatureAssociate.cpp (all six association functions, FindNeighbors,
                 TransformLines; compiled where it lies with the PCL / Eigen stand-ins of oracle/shim: oracle/_ref/libpvo_ref_assoc.so) on four synthetic pairs
  ref_camlidar.npz  (image line, LiDAR segment) pairs of the reference's own CameraLidarLineAssociate::AssociateByAngle (+ Filter + UniqueLinePair, masks) and uint16
                 depth images of its ProjectLidar2PanoramaDepth (oracle/_ref/libpvo_ref_camlidar.so)
  ref_builders.npz  the residual blocks (frame pairs, loss, raw residual, raw Jacobian, in registration order) that the reference's own util/Optimization.cpp builders
                 register for one RefinePose in five configurations, and AddCameraLidarResidual for one frame pair (ceres::Problem = a recorder, oracle/shim)
  ref_velodyne.npz  float32 clouds after the reference's own Transform2LidarWorld / Transform2Local and UndistortCloud (sensors/Velodyne.cpp compiled where it lies)
  ref_joint.npz  the problem of the joint stage (configs[2]) as the reference's own AssociateLineMulti + Optimize assemble it, recorded at ceres::Solve
  ref_pixel_fit.npz  the pairs of the reference's own pixel-space Associate() run to its end with the RANSAC's inliers scripted to what pvb_pixel_fit_line reports
                 (PCL is not available: oracle/shim's SACSegmentation replays the script; everything after the RANSAC is the reference's code)
  ref_dense.npz  per-frame 6x6 systems of a small dense ICP evaluation (the bench / smoke path) rebuilt from the reference's own association + functor code
  reproj.npz     residuals + 1x9 Jacobians of PanoramaReprojResidual_1Angle from a torch float64 autograd twin (Rodrigues closed form), and the
                 undistortion of a small sweep with scipy.spatial.transform (rotation vector scaling instead of quaternion slerp)
Run from the repo root:  python tests/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import twin  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_functors():
    c = cases.random_blocks(20260925, 240)
    r = np.zeros(len(c["type"]))
    J = np.zeros((len(c["type"]), 12))
    for i in range(len(r)):
        r[i], J[i] = twin.residual_and_jacobian(int(c["type"][i]), c["consts"][i], bool(c["normalize"][i]), c["poses"][c["ref"][i]], c["poses"][c["nei"][i]])
    np.savez_compressed(os.path.join(OUT, "functors.npz"), residual=r, jacobian=J, **c)


def golden_functors_f6():
    c = cases.random_blocks_f6(20260926, 120)
    r = np.zeros(len(c["type"]))
    J = np.zeros((len(c["type"]), 12))
    for i in range(len(r)):
        r[i], J[i] = twin.residual_and_jacobian(int(c["type"][i]), c["consts"][i], bool(c["normalize"][i]), c["poses"][c["ref"][i]], c["poses"][c["nei"][i]])
    np.savez_compressed(os.path.join(OUT, "functors_f6.npz"), residual=r, jacobian=J, **c)


def golden_rotations():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(7)
    aa = rng.normal(size=(64, 3))
    aa = aa / np.linalg.norm(aa, axis=1, keepdims=True) * np.concatenate([np.geomspace(1e-7, 3.1, 60), [3.14, 3.1415, 1e-9, 0.5]])[:, None]
    R = Rotation.from_rotvec(aa).as_matrix()
    np.savez_compressed(os.path.join(OUT, "rotations.npz"), aa=aa, R=R)


def golden_assoc():
    from scipy.spatial import cKDTree
    A, B = synth.make_pair(seed=20260925, n_az=900, ground_class=True)
    R_ref, t_ref = A["R_wl"], A["t_wl"]
    R_nei, t_nei = np.eye(3), np.zeros(3)               # initial guess of frame B = identity
    # float32 world clouds exactly as pcl::transformPointCloud(…, Matrix4d): double math, float32 store
    def to_world(R, t, cloud):
        w = cloud.copy()
        w[:, :3] = (cloud[:, :3].astype(np.float64) @ R.T + t).astype(np.float32)
        return w
    refw, neiw = to_world(R_ref, t_ref, A["surfLessFlat"]), to_world(R_nei, t_nei, B["surfFlat"])
    tree = cKDTree(refw[:, :3].astype(np.float64))
    k, thr, tol = 10, 1.0, 0.05
    dd, ii = tree.query(neiw[:, :3].astype(np.float64), k=k)
    q_idx, planes, points = [], [], []
    for qi in range(len(neiw)):
        d2 = ((refw[ii[qi], :3] - neiw[qi, :3]) ** 2).astype(np.float32)
        d2 = (d2[:, 0] + d2[:, 1]) + d2[:, 2]
        if d2.max() > np.float32(thr) * np.float32(thr):
            continue
        if not np.all(refw[ii[qi], 3] == neiw[qi, 3]):
            continue
        pl = (refw[ii[qi], :3].astype(np.float64) - t_ref) @ R_ref        # R^T (p - t)
        x = np.linalg.lstsq(pl, -np.ones(k), rcond=None)[0]
        n = x / np.linalg.norm(x); d = 1.0 / np.linalg.norm(x)
        if np.any(np.abs(pl @ n + d) > tol):
            continue
        c = pl - pl.mean(0)
        w = np.linalg.eigvalsh(c.T @ c)
        if w[2] > 3.0 * w[1]:
            continue
        q_idx.append(qi); planes.append(np.concatenate([n, [d]])); points.append((neiw[qi, :3].astype(np.float64) - t_nei) @ R_nei)
    np.savez_compressed(os.path.join(OUT, "assoc_pair.npz"), ref_local=A["surfLessFlat"], nei_local=B["surfFlat"], R_ref=R_ref, t_ref=t_ref,
                        R_nei=R_nei, t_nei=t_nei, ref_world=refw, nei_world=neiw, knn_idx=ii.astype(np.int32),
                        query=np.array(q_idx, np.int32), plane=np.array(planes), point=np.array(points), k=k, thr=thr, tol=tol)


def golden_atan2():
    rng = np.random.default_rng(3)
    ang = np.linspace(-np.pi, np.pi, 4001)[1:]
    r = rng.uniform(0.1, 50, size=ang.shape)
    y, x = r * np.sin(ang), r * np.cos(ang)
    np.savez_compressed(os.path.join(OUT, "fast_atan2.npz"), y=y, x=x, atan2=np.arctan2(y, x))


def golden_reproj():
    import torch
    from scipy.spatial.transform import Rotation
    d = synth.make_ba_problem(n_cams=6, n_points=80, seed=20261002)
    weight = 1.3
    n = len(d["cam"])
    r, J = np.zeros(n), np.zeros((n, 9))
    for i in range(n):
        aa = torch.tensor(d["cams"][d["cam"][i], :3], requires_grad=True)
        t = torch.tensor(d["cams"][d["cam"][i], 3:], requires_grad=True)
        X = torch.tensor(d["points"][d["point"][i]], requires_grad=True)
        th = torch.linalg.norm(aa)
        k = aa / th
        K = torch.zeros(3, 3, dtype=torch.float64)
        K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -k[2], k[1], k[2], -k[0], -k[1], k[0]
        R = torch.eye(3, dtype=torch.float64) + torch.sin(th) * K + (1 - torch.cos(th)) * (K @ K)
        P = R @ X + t
        s = torch.tensor(d["bearing"][i])
        res = weight * torch.acos((P @ s) / torch.linalg.norm(P))
        res.backward()
        r[i] = res.item()
        J[i] = np.concatenate([aa.grad.numpy(), t.grad.numpy(), X.grad.numpy()])
    # a small sweep undistorted with scipy (float64), stored as the float32 the reference writes back
    rng = np.random.default_rng(20261003)
    m = 500
    cloud = (rng.normal(size=(m, 4)) * 10).astype(np.float32)
    A, E = np.eye(4), np.eye(4)
    A[:3, :3] = Rotation.from_rotvec([0.2, -0.1, 0.3]).as_matrix(); A[:3, 3] = [1.0, 2.0, -0.5]
    E[:3, :3] = A[:3, :3] @ Rotation.from_rotvec([0.02, 0.05, -0.03]).as_matrix(); E[:3, 3] = A[:3, 3] + [0.1, -0.05, 0.02]
    R_se, t_se = A[:3, :3].T @ E[:3, :3], A[:3, :3].T @ (E[:3, 3] - A[:3, 3])
    rv = Rotation.from_matrix(R_se).as_rotvec()
    ratio = (np.arange(m, dtype=np.float32) / np.float32(m)).astype(np.float64)
    und = np.stack([Rotation.from_rotvec(rv * q).apply(cloud[i, :3].astype(np.float64)) + q * t_se for i, q in enumerate(ratio)])
    np.savez_compressed(os.path.join(OUT, "reproj.npz"), cam=d["cam"], point=d["point"], bearing=d["bearing"], cams=d["cams"], points=d["points"], weight=weight,
                        residual=r, jacobian=J, sweep=cloud, T_wl=A, T_we=E, undistorted=und)


def golden_ref_math():
    from oracle import pvo
    if pvo.ref_lib() is None:
        print("oracle/_ref not built (no /root/reference here): ref_fast_atan2.npz left as committed")
        return
    rng = np.random.default_rng(20261004)
    ang = np.linspace(-np.pi, np.pi, 6001)
    r = rng.uniform(0.05, 80, size=ang.shape)
    y = np.concatenate([r * np.sin(ang), rng.normal(0, 10, 6000), [0, 0, 1, -1, 0, -0.0, 0.0, 1e-30, 1e30, -1e-30, 5, 5, -5, -5]])
    x = np.concatenate([r * np.cos(ang), rng.normal(0, 10, 6000), [0, 1, 0, 0, -1, -1, -0.0, 1e30, 1e-30, -1e30, 5, -5, 5, -5]])
    yf, xf = y.astype(np.float32), x.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "ref_fast_atan2.npz"), y=y, x=x, out_f64=pvo.ref_fast_atan2(y, x), yf=yf, xf=xf, out_f32=pvo.ref_fast_atan2(yf, xf))


def golden_ref_path():
    from oracle import pvo
    L = pvo.ref_path_lib()
    if L is None:
        print("oracle/_ref/libpvo_ref_path.so not built (no /root/reference here): ref_functors.npz / ref_geometry.npz left as committed")
        return
    import ctypes as C
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    c = cases.ref_functor_cases(20261017, 600)
    r, J, consts = pvo.ref_eval_functors(c["type"], c["normalize"], c["raw"], c["params"])
    assert np.all(np.isfinite(r)) and np.all(np.isfinite(J))
    c["consts"] = consts
    np.savez_compressed(os.path.join(OUT, "ref_functors.npz"), residual=r, jacobian=cases.ref_jacobian_to_block_layout(c["type"], J), jacobian_call_order=J, **c)
    # pairwise functors and the reprojection functor (2 and 3 parameter blocks)
    rng = np.random.default_rng(20261018)
    n = 200
    typ = np.repeat([pvo.REF_PAIRWISE_P2PLANE, pvo.REF_PAIRWISE_P2LINE, pvo.REF_REPROJ_1ANGLE, pvo.REF_PLANE_IOU_CAMERA], n).astype(np.int32)
    raw, params = np.zeros((4 * n, 16)), np.zeros((4 * n, 12))
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    raw[:n, :3] = rng.normal(0, 3, (n, 3)); raw[:n, 3:6] = nrm; raw[:n, 6] = rng.normal(0, 2, n); raw[:n, 7] = rng.uniform(0.5, 2, n)
    raw[n:2 * n, :9] = rng.normal(0, 3, (n, 9)); raw[n:2 * n, 9] = rng.uniform(0.5, 2, n)
    raw[2 * n:3 * n, :3] = rng.normal(0, 1, (n, 3)); raw[2 * n:3 * n, 3] = rng.uniform(0.5, 2, n)
    raw[3 * n:, :4] = rng.normal(0, 1, (n, 4)); raw[3 * n:, 4:7] = rng.normal(0, 3, (n, 3)); raw[3 * n:, 7:13] = rng.normal(0, 1, (n, 6)); raw[3 * n:, 13] = rng.uniform(0.5, 2, n)
    params[:, :3] = rng.normal(0, 0.5, (4 * n, 3)); params[:, 3:6] = rng.normal(0, 1, (4 * n, 3))
    params[2 * n:3 * n, 6:9] = rng.normal(0, 4, (n, 3))
    params[3 * n:, 6:9] = rng.normal(0, 0.5, (n, 3)); params[3 * n:, 9:] = rng.normal(0, 1, (n, 3))
    params[::9, :3] = 0.0
    r2, J2, c2 = pvo.ref_eval_functors(typ, 0, raw, params)
    assert np.all(np.isfinite(r2)) and np.all(np.isfinite(J2))
    # geometry helpers
    out = dict(pw_type=typ, pw_raw=raw, pw_params=params, pw_residual=r2, pw_jacobian=J2, pw_consts=c2)
    m = 300
    planes, lines, line_ok = np.zeros((m, 4)), np.zeros((m, 6)), np.zeros(m)
    pts_all = np.zeros((m, 10, 3)); counts = np.zeros(m, np.int32)
    tol_p, tol_l, thr_l = np.zeros(m), np.zeros(m), np.zeros(m)
    for i in range(m):
        k = int(rng.integers(3, 11)); counts[i] = k
        kind = i % 4
        base = rng.normal(0, 5, 3); u = rng.normal(size=3); u /= np.linalg.norm(u); v = np.cross(u, rng.normal(size=3)); v /= np.linalg.norm(v)
        if kind == 0:   pts = base + rng.normal(0, 1, (k, 1)) * u + rng.normal(0, 1, (k, 1)) * v + rng.normal(0, 0.005, (k, 3))     # plane
        elif kind == 1: pts = base + rng.normal(0, 1, (k, 1)) * u + rng.normal(0, 0.01, (k, 3))                                       # line
        elif kind == 2: pts = base + rng.normal(0, 1, (k, 3))                                                                         # blob
        else:           pts = base + rng.normal(0, 1, (k, 1)) * u                                                                     # exactly collinear: rank deficient
        pts_all[i, :k] = pts
        tol_p[i] = [0.0, 0.02, 0.05][i % 3]; tol_l[i] = [3.0, 10.0][i % 2]; thr_l[i] = [0.0, 0.05][(i // 2) % 2]
        L.ref_form_plane(C.c_int(k), p(np.ascontiguousarray(pts)), C.c_double(tol_p[i]), p(planes[i:i + 1]))
        L.ref_form_line(C.c_int(k), p(np.ascontiguousarray(pts)), C.c_double(tol_l[i]), C.c_double(thr_l[i]), p(lines[i:i + 1]))
    out.update(fp_points=pts_all, fp_counts=counts, fp_tol=tol_p, fp_plane=planes, fl_tol=tol_l, fl_thr=thr_l, fl_line=lines)
    q = rng.normal(0, 3, (m, 16))
    s = np.zeros((m, 5)); proj = np.zeros((m, 2, 3)); plane3 = np.zeros((m, 4))
    for i in range(m):
        pt, ln, pl = q[i, :3].copy(), q[i, 3:9].copy(), q[i, 9:13].copy()
        s[i, 0] = L.ref_point_to_line_distance3d(p(pt), p(ln))
        for nz in (0, 1):
            pln = pl.copy()
            if nz: pln[:3] /= np.linalg.norm(pln[:3])
            s[i, 1 + nz] = L.ref_point_to_plane_distance(p(pln), p(pt), C.c_int(nz))
            L.ref_project_point_to_plane(p(pt), p(pln), p(proj[i, nz:nz + 1]), C.c_int(nz))
        s[i, 3] = L.ref_vector_angle3d(p(pt), p(ln[:3].copy()), C.c_int(0))
        s[i, 4] = L.ref_plane_angle(p(pt), p(ln[:3].copy()), C.c_int(0))
        L.ref_form_plane3(p(pt), p(ln[:3].copy()), p(ln[3:].copy()), p(plane3[i:i + 1]))
    out.update(g_in=q, g_scalar=s, g_project=proj, g_plane3=plane3)
    from scipy.spatial.transform import Rotation
    n_s = 80
    w1, w2, ratio, so = np.zeros((n_s, 4, 4)), np.zeros((n_s, 4, 4)), rng.uniform(0, 1, n_s), np.zeros((n_s, 4, 4))
    for i in range(n_s):
        for w in (w1, w2):
            w[i] = np.eye(4); w[i, :3, :3] = Rotation.from_rotvec(rng.normal(0, [0.05, 1.0, 2.5][i % 3], 3)).as_matrix(); w[i, :3, 3] = rng.normal(0, 2, 3)
        if i % 10 == 0: w2[i] = w1[i]
        L.ref_slerp_pose(p(w1[i]), p(w2[i]), C.c_double(ratio[i]), p(so[i]))
    out.update(slerp_w1=w1, slerp_w2=w2, slerp_ratio=ratio, slerp_out=so)
    # projection
    rows, cols = 2880, 5760
    cam_f = np.concatenate([rng.normal(0, 5, (4000, 3)), [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [1e-8, 0, -1], [-1e-8, 0, -1]]]).astype(np.float32)
    px_f = np.zeros((len(cam_f), 2), np.float32); L.ref_cam_to_image_f(C.c_int(rows), C.c_int(cols), C.c_long(len(cam_f)), p(cam_f), p(px_f))
    cam_d = cam_f.astype(np.float64) + rng.normal(0, 1e-3, cam_f.shape)
    px_d, px_e = np.zeros((len(cam_d), 2)), np.zeros((len(cam_d), 2))
    L.ref_cam_to_image_d(C.c_int(rows), C.c_int(cols), C.c_long(len(cam_d)), p(cam_d), p(px_d))
    L.ref_cam_to_image_eigen_d(C.c_int(rows), C.c_int(cols), C.c_long(len(cam_d)), p(cam_d), p(px_e))
    pix = np.stack([rng.uniform(0, cols, 2000), rng.uniform(0, rows, 2000)], axis=1)
    pix[:4] = [[0, 0], [cols, rows], [cols / 2, rows / 2], [cols - 1, 0]]
    i2c_d, i2c_e = np.zeros((len(pix), 3)), np.zeros((len(pix), 3))
    L.ref_image_to_cam_d(C.c_int(rows), C.c_int(cols), C.c_long(len(pix)), p(pix), C.c_double(1.0), p(i2c_d))
    L.ref_image_to_cam_eigen_d(C.c_int(rows), C.c_int(cols), C.c_long(len(pix)), p(pix), C.c_double(1.0), p(i2c_e))
    pix_f = pix.astype(np.float32); i2c_f = np.zeros((len(pix), 3), np.float32)
    L.ref_image_to_cam_f(C.c_int(rows), C.c_int(cols), C.c_long(len(pix)), p(pix_f), C.c_float(5.0), p(i2c_f))
    n_l = 300
    ln = np.stack([rng.uniform(0, cols, n_l), rng.uniform(0, rows, n_l), rng.uniform(0, cols, n_l), rng.uniform(0, rows, n_l)], axis=1).astype(np.float32)
    ln[::5, 0] = rng.uniform(0, 300, len(ln[::5])); ln[::5, 2] = rng.uniform(cols - 300, cols, len(ln[::5]))     # seam crossings
    seg_len = np.where(np.arange(n_l) % 2 == 0, 70.0, 100.0).astype(np.float32)
    seg_off, seg_xy = [0], []
    for i in range(n_l):
        buf = np.zeros((256, 2), np.float32)
        k = L.ref_break_to_segments(C.c_int(rows), C.c_int(cols), p(ln[i]), C.c_float(seg_len[i]), C.c_int(256), p(buf))
        assert k > 0
        seg_xy.append(buf[:k].copy()); seg_off.append(seg_off[-1] + k)
    out.update(rows=rows, cols=cols, cam_f=cam_f, px_f=px_f, cam_d=cam_d, px_d=px_d, px_eigen_d=px_e, pix=pix, i2c_d=i2c_d, i2c_eigen_d=i2c_e, pix_f=pix_f, i2c_f5=i2c_f,
               bts_lines=ln, bts_seg_len=seg_len, bts_off=np.array(seg_off, np.int32), bts_xy=np.concatenate(seg_xy))
    np.savez_compressed(os.path.join(OUT, "ref_geometry.npz"), **out)


def golden_ref_assoc():
    """tests/golden/ref_assoc.npz: outputs of the reference's own lidar_mapping/LidarFeatureAssociate.cpp (oracle/_ref/libpvo_ref_assoc.so) on the synthetic
    pairs of tests/test_reference_pinning.py: ASSOC_CASES, plus FindNeighbors on pose sets with loops and missing poses, and TransformLines."""
    from oracle import pvo
    if pvo.ref_assoc_lib() is None:
        print("oracle/_ref/libpvo_ref_assoc.so not built (no /root/reference here): ref_assoc.npz left as committed")
        return
    import test_reference_pinning as trp
    out = {}
    for ci, (seed, n_az, perturb, tol, thr_p, thr_l) in enumerate(trp.ASSOC_CASES):
        A, B, RB, tB = trp.assoc_case(pvo, seed, n_az, perturb)
        for k, v in trp.reference_associations(pvo, A, B, RB, tB, tol, thr_p, thr_l).items():
            out[f"c{ci}_{k}"] = v
            print(f"  case {ci} {k}: {len(v)}")
    rng = np.random.default_rng(20261025)
    fn = []
    # a straight walk, a loop that closes after > 200 frames, and the same with poses / frames missing
    s = np.arange(60) * 0.4
    fn.append((np.stack([s, 0.1 * np.sin(s), np.zeros_like(s)], 1), np.ones(60, np.uint8), np.ones(60, np.uint8), 6))
    a = np.linspace(0, 2 * np.pi, 520)
    loop = np.stack([30 * np.cos(a), 30 * np.sin(a), 0.05 * rng.normal(size=520)], 1)
    fn.append((loop, np.ones(520, np.uint8), np.ones(520, np.uint8), 6))
    pv = (rng.random(520) > 0.05).astype(np.uint8); va = (rng.random(520) > 0.03).astype(np.uint8); pv[:3] = 1; va[:3] = 1
    fn.append((loop + rng.normal(0, 0.3, loop.shape), pv, va, 6))
    fn.append((rng.uniform(-15, 15, (300, 3)), np.ones(300, np.uint8), np.ones(300, np.uint8), 4))
    for ci, (t, pv, va, k) in enumerate(fn):
        R = np.tile(np.eye(3).reshape(1, 9), (len(t), 1))
        nb = pvo.ref_find_neighbors(R, t, pv, va, k)
        off = np.zeros(len(t) + 1, np.int32); off[1:] = np.cumsum([len(x) for x in nb])
        out.update({f"fn{ci}_t": t, f"fn{ci}_pose_valid": pv, f"fn{ci}_valid": va, f"fn{ci}_k": k, f"fn{ci}_off": off, f"fn{ci}_ids": np.concatenate(nb).astype(np.int32)})
    out["fn_cases"] = len(fn)
    from scipy.spatial.transform import Rotation
    import ctypes as C
    T = np.eye(4); T[:3, :3] = Rotation.from_rotvec([0.3, -1.1, 0.7]).as_matrix(); T[:3, 3] = [1.5, -2.0, 0.3]
    lin = rng.normal(0, 3, (40, 6)); lout = np.zeros_like(lin)
    pvo.ref_assoc_lib().ref_transform_lines(T.ctypes.data_as(C.c_void_p), C.c_int(40), lin.ctypes.data_as(C.c_void_p), lout.ctypes.data_as(C.c_void_p))
    out.update(tl_T=T, tl_in=lin, tl_out=lout)
    for ci, (nf, n_az, k, min_len, no_pose) in enumerate(trp.TRACK_CASES):          # LidarLineMatch::GenerateTracks
        tr = trp.reference_line_tracks(pvo, trp.track_case(nf, n_az), k, min_len, no_pose)
        off = np.zeros(len(tr) + 1, np.int32); off[1:] = np.cumsum([len(t) for t in tr])
        out[f"tr{ci}_off"] = off; out[f"tr{ci}_feat"] = np.concatenate(tr).astype(np.int32)
        print(f"  tracks case {ci}: {len(tr)} tracks, {off[-1]} features")
        frames_ci = trp.track_case(nf, n_az)                                              # CameraLidarOptimizer::LidarMaskByTrack with the same parameters
        rf = [pvo.RefFrame(f["R_wl"], f["t_wl"], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], id=i, pose_valid=(i != no_pose), local="keep")
              for i, f in enumerate(frames_ci)]
        masks = pvo.ref_lidar_mask_by_track(rf, min_len, k)
        moff = np.zeros(len(masks) + 1, np.int32); moff[1:] = np.cumsum([len(m) for m in masks])
        out[f"lm{ci}_off"] = moff; out[f"lm{ci}_mask"] = np.concatenate(masks).astype(np.uint8)
    t_wc, fpv, t_wl, lpv, lva = trp.neighbor_each_frame_case()                             # CameraLidarOptimizer::NeighborEachFrame
    I = np.tile(np.eye(3).reshape(1, 9), (len(t_wc), 1))
    for ci, (k, temporal) in enumerate(((3, True), (1, True), (4, False), (2, False))):
        nb = pvo.ref_neighbor_each_frame(I, t_wc, fpv, I, t_wl, lpv, lva, k, temporal)
        off = np.zeros(len(nb) + 1, np.int32); off[1:] = np.cumsum([len(x) for x in nb])
        out[f"nef{ci}_off"] = off; out[f"nef{ci}_ids"] = np.concatenate([np.asarray(x, np.int32) for x in nb]) if off[-1] else np.zeros(0, np.int32)
    np.savez_compressed(os.path.join(OUT, "ref_assoc.npz"), **out)


def golden_ref_camlidar():
    """tests/golden/ref_camlidar.npz: pair lists of the reference's own CameraLidarLineAssociate::AssociateByAngle (+ Filter + UniqueLinePair) and depth images of
    its ProjectLidar2PanoramaDepth (oracle/_ref/libpvo_ref_camlidar.so) on the case of tests/test_reference_pinning.py: camlidar_case."""
    from oracle import pvo
    if pvo.ref_camlidar_lib() is None:
        print("oracle/_ref/libpvo_ref_camlidar.so not built (no /root/reference here): ref_camlidar.npz left as committed")
        return
    import test_reference_pinning as trp
    A, rows, cols, T, lines = trp.camlidar_case()
    out = dict(lines=lines, T_cl=T)
    n_seg = len(A["segment_coeffs"])
    for name, multi, masked in trp.CAMLIDAR_VARIANTS:
        im, lm = trp._camlidar_masks(out, len(lines), n_seg) if masked else (None, None)
        r = pvo.ref_associate_by_angle(rows, cols, lines, A["cornerLessSharp"], A["p2s_off"], A["p2s_ids"], A["segment_coeffs"], A["end_points"], T, multi, im, lm)
        for k, v in zip(("image", "lidar", "start", "end", "score"), r):
            out[f"{name}_{k}"] = v
        print(f"  {name}: {len(r[0])} pairs")
    cand = pvo.ref_pixel_associate_candidates(rows, cols, lines, A["cloud"][::4], T)          # first stage of the pixel-space Associate()
    out["px_off"] = np.concatenate([[0], np.cumsum([len(c) for c in cand])]).astype(np.int32); out["px_xyz"] = np.concatenate(cand)
    print(f"  pixel-space candidates: {len(cand)} lines, {out['px_off'][-1]} points")
    out["depth_720"] = pvo.ref_project_depth(A["cloud"], 720, 1440, T, 3)
    out["depth_360"] = pvo.ref_project_depth(A["cloud"], 360, 720, T, 4)
    np.savez_compressed(os.path.join(OUT, "ref_camlidar.npz"), **out)


def golden_ref_builders():
    """tests/golden/ref_builders.npz: the residual blocks registered by the reference's own util/Optimization.cpp builders (oracle/_ref/libpvo_ref_assoc.so; ceres::Problem
    replaced by a recorder) for the RefinePose configurations of tests/test_reference_pinning.py: BUILDER_CASES, each block evaluated once through
    ceres::CostFunction::Evaluate; plus AddCameraLidarResidual for one (image, LiDAR) pair of frames."""
    from oracle import pvo
    if pvo.ref_assoc_lib() is None:
        print("oracle/_ref/libpvo_ref_assoc.so not built (no /root/reference here): ref_builders.npz left as committed")
        return
    import test_reference_pinning as trp
    from scipy.spatial.transform import Rotation
    frames, Rs, ts = trp.builder_case()
    out = {}
    for ci, kw in enumerate(trp.BUILDER_CASES):
        r = trp.reference_refine_blocks(pvo, frames, Rs, ts, **kw)
        for k, v in r.items():
            out[f"b{ci}_{k}"] = v.astype(np.int16) if k in ("ref", "nei") else (v[::trp.JAC_STRIDE] if k == "jacobian" else v)   # every 8th Jacobian row keeps the file small
        print(f"  builders case {ci}: {len(r['residual'])} blocks, {int((r['huber'] == 0).sum())} without loss")
    rng = np.random.default_rng(20261031)
    rows, cols, n = 2880, 5760, 80
    lines = np.stack([rng.uniform(0, cols, n), rng.uniform(200, rows - 200, n), rng.uniform(0, cols, n), rng.uniform(200, rows - 200, n)], axis=1).astype(np.float32)
    start, end = rng.normal(0, 3, (n, 3)), rng.normal(0, 3, (n, 3))
    pw = rng.uniform(0.5, 2, n).astype(np.float32)
    R_wc, t_wc = Rotation.from_rotvec([0.2, -0.4, 0.1]).as_matrix(), np.array([1.0, -0.5, 0.3])
    R_wl, t_wl = Rotation.from_rotvec([0.25, -0.35, 0.12]).as_matrix(), np.array([1.1, -0.45, 0.2])
    r, J, poses = pvo.ref_camera_lidar_blocks(rows, cols, lines, start, end, pw, R_wc, t_wc, R_wl, t_wl, 25.0)
    out.update(cl_rows=rows, cl_cols=cols, cl_lines=lines, cl_start=start, cl_end=end, cl_pair_weight=pw, cl_weight=25.0, cl_poses=poses, cl_residual=r, cl_jacobian=J)
    rows, cols, R, t, pv, kp_off, xy, track_off, ff, fi, pts = trp.camera_residual_case()          # AddCameraResidual (ANGLE_RESIDUAL_1)
    cr = pvo.ref_camera_residual_blocks(rows, cols, R, t, pv, kp_off, xy, track_off, ff, fi, pts, 0.7)
    out.update(cr_cam=cr["cam"], cr_track=cr["track"], cr_residual=cr["residual"], cr_jacobian=cr["jacobian"], cr_cams=cr["cams"], cr_weight=0.7)
    print(f"  camera residual blocks: {len(cr['cam'])} of {len(ff)} features")
    np.savez_compressed(os.path.join(OUT, "ref_builders.npz"), **out)


def golden_ref_velodyne():
    """tests/golden/ref_velodyne.npz: clouds moved by the reference's own Velodyne::Transform2LidarWorld / Transform2Local and undistorted by its UndistortCloud
    (sensors/Velodyne.cpp compiled where it lies, oracle/_ref/libpvo_ref_assoc.so)."""
    from oracle import pvo
    if pvo.ref_assoc_lib() is None:
        print("oracle/_ref/libpvo_ref_assoc.so not built (no /root/reference here): ref_velodyne.npz left as committed")
        return
    import test_reference_pinning as trp
    cloud, R_wl, t_wl, sweeps = trp.velodyne_case()
    f = pvo.RefFrame(R_wl, t_wl, surf_less_flat_world=cloud, local=True)
    out = dict(world=f.cloud("less_flat"))
    f.to_local()
    out["local_again"] = f.cloud("less_flat")
    for k, (R_we, t_we) in enumerate(sweeps):
        ok, und = pvo.ref_undistort_cloud(R_wl, t_wl, R_we, t_we, cloud)
        assert ok
        out[f"undistorted{k}"] = und
    T, pv, va, off, clouds = trp.undistort_lidars_case()                                  # LidarOdometry::UndistortLidars
    for gi, gap in enumerate((0.0, 0.02)):
        out[f"ul_out{gi}"] = pvo.ref_undistort_lidars(T[:, :3, :3], T[:, :3, 3], pv, va, off, clouds, gap)
        print(f"  UndistortLidars gap {gap}: {int((out[f'ul_out{gi}'] != clouds).any(1).sum())} of {len(clouds)} points moved")
    # pose text files (util/FileIO.cpp: ExportPoseT / ReadPoseT)
    import tempfile
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(20261105)
    n = 9
    R = np.stack([Rotation.from_rotvec(rng.normal(size=3)).as_matrix() for _ in range(n)]); t = rng.normal(size=(n, 3)) * 100
    t[3] = np.inf; R[3] = 0; t[6, 1] = 1e-7; t[7] = [123456.789, -0.000012345, 1e10]
    names = [f"frame_{i:04d}.pcd" for i in range(n)]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "poses.txt")
        pvo.ref_export_pose_t(path, R, t, names)
        text = open(path, "rb").read()
        ra = pvo.ref_read_pose_t(path, True); rv = pvo.ref_read_pose_t(path, False)
    out.update(pt_R=R, pt_t=t, pt_names=np.array(names), pt_text=np.frombuffer(text, np.uint8), pt_read_all_R=ra[0], pt_read_all_t=ra[1], pt_read_all_names=np.array(ra[2]),
               pt_read_valid_R=rv[0], pt_read_valid_t=rv[1], pt_read_valid_names=np.array(rv[2]))
    print(text.decode().splitlines()[3], "|", len(ra[2]), "read with invalid,", len(rv[2]), "without")
    np.savez_compressed(os.path.join(OUT, "ref_velodyne.npz"), **out)


def golden_ref_joint():
    """tests/golden/ref_joint.npz: the problem the reference's own CameraLidarOptimizer::AssociateLineMulti + mapping-mode Optimize hand to ceres::Solve (recorded by the
    stand-in's solve hook, every block evaluated once) for tests/test_reference_pinning.py: JOINT_CASES."""
    from oracle import pvo
    if pvo.ref_assoc_lib() is None:
        print("oracle/_ref/libpvo_ref_assoc.so not built (no /root/reference here): ref_joint.npz left as committed")
        return
    import test_reference_pinning as trp
    d = trp.joint_case()
    out = {}
    for ci, kw in enumerate(trp.JOINT_CASES):
        r = trp.reference_joint_blocks(pvo, d, **kw)
        for k, v in r.items():
            out[f"j{ci}_{k}"] = (np.asarray(v).astype(np.int16) if k in ("a", "b", "n_params") else (v[::trp.JAC_STRIDE] if k == "jacobian" else v))
        print(f"  joint case {ci}: {len(r['residual'])} blocks ({int((r['n_params'] == 3).sum())} reprojection), {r['n_line_pairs']} line pairs, constant parts {r['const_part'][:6].tolist()}...")
    T = d["T_cl"].copy(); T[:3, 3] += [0.02, -0.01, 0.015]                               # calibration mode: AssociateLineSingle + Optimize(line_pairs, T_cl)
    rf = [pvo.RefFrame(np.eye(3), np.zeros(3), f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], id=i, local="keep", end_points=f["end_points"])
          for i, f in enumerate(d["frames"])]
    c = pvo.ref_calibration_blocks(d["rows"], d["cols"], d["image_lines"], rf, T)
    out.update({f"cal_{k}": v for k, v in c.items()})
    print(f"  calibration: {len(c['residual'])} blocks from {c['info'][0]} line pairs, options {c['info'][1:].tolist()}")
    rf3 = rf[:3]                                                                          # calibration-mode JointOptimize with a scripted solver
    res = [pvo.ref_calibration_loop(d["rows"], d["cols"], d["image_lines"][:3], rf3, T, sc) for sc in trp.CALIB_SCRIPTS]
    out["cal_loop_calls"] = np.array([r_[0] for r_ in res], np.int32); out["cal_loop_T"] = np.stack([r_[1] for r_ in res])
    print("  calibration loop solver calls per script:", out["cal_loop_calls"].tolist())
    rf2 = [pvo.RefFrame(d["Rs"][i], d["ts"][i], f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["surfFlat"], f["surfLessFlat"], id=i, local="keep",
                        end_points=f["end_points"]) for i, f in enumerate(d["frames"])]
    out["loop_counts"] = np.array([pvo.ref_joint_optimize_loop(d["rows"], d["cols"], d["R_wc"], d["t_wc"], d["image_lines"], rf2, d["T_cl"], 7, [x[0] for x in sc],
                                                               [x[1] for x in sc]) for sc in trp.LOOP_SCRIPTS], np.int32)     # JointOptimize with a scripted solver
    print("  JointOptimize solver calls per script:", out["loop_counts"].tolist())
    np.savez_compressed(os.path.join(OUT, "ref_joint.npz"), **out)


def golden_ref_dense():
    """tests/golden/ref_dense.npz: per-frame normal equations of the dense ICP evaluation (configs[4] shape, small) rebuilt from the reference's own pieces
    (tests/test_reference_pinning.py: dense_systems_from_reference_pieces)."""
    from oracle import pvo
    if pvo.ref_assoc_lib() is None or pvo.ref_path_lib() is None:
        print("oracle/_ref not built (no /root/reference here): ref_dense.npz left as committed")
        return
    import test_reference_pinning as trp
    c = trp.DENSE_CASE
    d = synth.make_dense_sweep(n_target=c["n_target"], n_frames=c["n_frames"], pts_per_frame=c["pts_per_frame"], seed=c["seed"])
    sysm = trp.dense_systems_from_reference_pieces(pvo, d, c["plane_tol"], c["dist_thr"], c["huber"], c["weight"])
    print("  dense systems:", sysm[:, 28].astype(int).tolist(), "accepted queries per frame")
    np.savez_compressed(os.path.join(OUT, "ref_dense.npz"), systems=sysm)


def pixel_fit_script(pvo, rows, cols, lines, cloud, T):
    """The inlier lists the product's pvb_pixel_fit_line reports for the candidate lists of the pixel-space Associate() (empty when it finds no line)."""
    from panovlm_b200 import Context
    line3, _, _ = pvo.pixel_line_neighbors(rows, cols, lines, cloud, T)
    off, idx = Context.pixel_line_candidates(len(lines), line3, 6)
    cam = pvo.transform_cloud(T[:3, :3], T[:3, 3], cloud)
    script = []
    for li in range(len(lines)):
        if off[li + 1] > off[li]:
            fit = Context.pixel_fit_line(cam[idx[off[li]:off[li + 1]]])
            script.append(np.zeros(0, np.int32) if fit is None else fit[1])
    return script


def segments_of(A):
    """The LiDAR segments of a synthetic frame as separate clouds (edge_segmented of the reference): points of cornerLessSharp by their first segment id."""
    off, ids = A["p2s_off"], A["p2s_ids"]
    n_seg = len(A["segment_coeffs"])
    pts = A["cornerLessSharp"]
    first = np.array([ids[off[i]] if off[i + 1] > off[i] else -1 for i in range(len(pts))])
    return [np.ascontiguousarray(pts[first == k]) for k in range(n_seg)]


def pixel_fit_script_segmented(pvo, rows, cols, lines, segments, T):
    """The inlier lists pvb_pixel_fit_line reports for the fits of the SEGMENTED pixel-space Associate() (one per image line that passes the 6-point and 70 % tests: the
    whole majority segment, LiDAR frame)."""
    from panovlm_b200 import Context
    segs = [np.asarray(x, np.float32).reshape(-1, 4) for x in segments]
    seg_off = np.concatenate([[0], np.cumsum([len(x) for x in segs])]).astype(np.int32)
    base = np.concatenate(segs)
    seg_of_point = np.repeat(np.arange(len(segs), dtype=np.int32), [len(x) for x in segs])
    cloud = base.copy(); cloud[:, 3] = seg_of_point
    line3, _, _ = pvo.pixel_line_neighbors(rows, cols, lines, cloud, T)
    off, idx = Context.pixel_line_candidates(len(lines), line3, 6)
    ids, _, f_off, f_idx = Context.segmented_fit_lists(len(lines), off, idx, seg_of_point, seg_off)
    script = []
    for k in range(len(ids)):
        fit = Context.pixel_fit_line(base[f_idx[f_off[k]:f_off[k + 1]]])
        script.append(np.zeros(0, np.int32) if fit is None else fit[1])
    return script


def pixel_calibration_from_reference(pvo, rows, cols, lines, cloud, T, script):
    """The reference's own AssociateLineSingle + Optimize(line_pairs, T_cl) (recorded at ceres::Solve) on ONE frame without LiDAR segments, RANSAC answers scripted."""
    rf = pvo.RefFrame(np.eye(3), np.zeros(3), cloud, np.zeros(len(cloud) + 1, np.int32), np.zeros(0, np.int32), np.zeros((0, 6)), id=0, local="keep", end_points=np.zeros((0, 6)))
    pvo.ref_set_sac_script(script)
    try:
        cal = pvo.ref_calibration_blocks(rows, cols, [lines], [rf], T)
    finally:
        answered = pvo.ref_set_sac_script(None)
    assert answered == len(script), (answered, len(script))
    return cal


def golden_ref_pixel_fit():
    """tests/golden/ref_pixel_fit.npz: the reference's own pixel-space Associate() (CameraLidarLineAssociate.cpp:22-188) run with the RANSAC's inliers SCRIPTED to what
    the product's pvb_pixel_fit_line reports (PCL is not available; oracle/shim's SACSegmentation replays the script) - everything after the RANSAC is the reference's code."""
    from oracle import pvo
    if pvo.ref_camlidar_lib() is None:
        print("oracle/_ref not built (no /root/reference here): ref_pixel_fit.npz left as committed")
        return
    import test_reference_pinning as trp
    A, rows, cols, T, lines = trp.camlidar_case()
    cloud = A["cloud"][::4]
    script = pixel_fit_script(pvo, rows, cols, lines, cloud, T)
    il, s, e, ang = pvo.ref_pixel_associate_scripted(rows, cols, lines, cloud, T, script)
    print(f"  scripted pixel-space Associate: {len(script)} candidate lists, {sum(len(x) >= 3 for x in script)} fits, {len(il)} pairs after Filter(true, true)")
    # the same frame through the reference's calibration mode: AssociateLineSingle takes Associate(lines, cornerLessSharp, T_cl) when edge_segmented is empty (:313-314)
    cal = pixel_calibration_from_reference(pvo, rows, cols, lines, cloud, T, script)
    print(f"  calibration mode over the pixel path: {int(cal['info'][0])} pairs, {len(cal['residual'])} residual blocks")
    # the segmented overload (:191-338) on the same frame, its segments as separate clouds
    segs = segments_of(A)
    script_seg = pixel_fit_script_segmented(pvo, rows, cols, lines, segs, T)
    out_seg = pvo.ref_pixel_associate_segmented_scripted(rows, cols, lines, segs, T, script_seg)
    assert out_seg is not None, "the script of the segmented Associate() does not match the reference's number of fits"
    print(f"  scripted segmented Associate: {len(segs)} segments, {len(script_seg)} fits, {len(out_seg[0])} pairs after Filter(true, true)")
    np.savez_compressed(os.path.join(OUT, "ref_pixel_fit.npz"), seg_inl_off=np.concatenate([[0], np.cumsum([len(x) for x in script_seg])]).astype(np.int32),
                        seg_inl_idx=np.concatenate(script_seg + [np.zeros(0, np.int32)]).astype(np.int32), seg_image_line=out_seg[0], seg_start=out_seg[1], seg_end=out_seg[2],
                        seg_angle=out_seg[3], cal_residual=cal["residual"], cal_jacobian=cal["jacobian"], cal_huber=cal["huber"], cal_pose=cal["pose"],
                        cal_info=cal["info"], inl_off=np.concatenate([[0], np.cumsum([len(x) for x in script])]).astype(np.int32),
                        inl_idx=np.concatenate(script).astype(np.int32), image_line=il, start=s, end=e, angle=ang)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_functors(); golden_functors_f6(); golden_rotations(); golden_assoc(); golden_atan2(); golden_reproj(); golden_ref_math(); golden_ref_path(); golden_ref_assoc(); golden_ref_camlidar(); golden_ref_builders(); golden_ref_velodyne(); golden_ref_joint(); golden_ref_dense(); golden_ref_pixel_fit()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
