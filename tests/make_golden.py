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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_functors(); golden_functors_f6(); golden_rotations(); golden_assoc(); golden_atan2(); golden_reproj(); golden_ref_math()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
