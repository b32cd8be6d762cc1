"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/libpvo_oracle.so`` (the CPU restatement of PanoVLM's hot path, see
``pvo_math.hpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; nothing under ``panovlm_b200/`` does.
Parity status: unpinned by the reference's own tests (it has none) — pinned against scipy / numpy /
torch-autograd in ``tests/test_oracle_*.py`` and, for the functor / geometry / projection layers, against the reference's own
source compiled with stand-in container types (``oracle/_ref``, ``oracle/shim``, ``tests/test_reference_pinning.py``).
The RANSAC line fit of the pixel-space fallback (pcl::SACSegmentation) has no restatement here - PARITY UNPINNED: PCL is not available; the product's
``pvb_pixel_fit_line`` is checked against a numpy twin in ``tests/test_pixel_fit_line.py`` and everything after the RANSAC against the reference's own code
run with scripted inliers (``ref_pixel_associate_scripted``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P2PLANE_METER, P2PLANE_ANGLE, P2LINE_METER, P2LINE_ANGLE, PLANE2PLANE_GLOBAL, PLANE_IOU = range(6)
PLANE2PLANE_RELATIVE, PLANE_RELATIVE_IOU, LINE2LINE_ANGLE = 6, 7, 8


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpvo_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.pvo_normal_equations.restype = C.c_double
        _LIB.pvo_reproj_normal_equations.restype = C.c_double
        _LIB.pvo_kdtree_build.restype = C.c_void_p
    return _LIB


def ref_lib():
    """oracle/_ref/libpvo_ref.so: the reference's own base/Math.h compiled where it lies (make -C oracle ref); None when it has not been built."""
    path = os.path.join(_HERE, "_ref", "libpvo_ref.so")
    return C.CDLL(path) if os.path.exists(path) else None


def ref_fast_atan2(y, x):
    """FastAtan2 of the REFERENCE (base/Math.h:15-29) through oracle/_ref."""
    L = ref_lib()
    y, x = np.ascontiguousarray(y), np.ascontiguousarray(x)
    out = np.empty_like(y)
    (L.ref_fast_atan2_f if y.dtype == np.float32 else L.ref_fast_atan2_d)(C.c_long(y.size), _p(y), _p(x), _p(out))
    return out


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return lib().pvo_num_threads()


def set_num_threads(n):
    lib().pvo_set_num_threads(C.c_int(int(n)))


def aa_to_R(aa):
    """Rotation matrix (row-major numpy 3x3) of an angle-axis vector."""
    out = np.zeros(9)
    lib().pvo_aa_to_R(_p(_f64(aa)), _p(out))
    return out.reshape(3, 3).T.copy()  # library buffer is column-major


def R_to_aa(R):
    out = np.zeros(3)
    lib().pvo_R_to_aa(_p(_f64(np.asarray(R).T)), _p(out))
    return out


def aa_rotate(aa, p):
    out = np.zeros(3)
    lib().pvo_aa_rotate(_p(_f64(aa)), _p(_f64(p)), _p(out))
    return out


def form_plane(pts, tol):
    pts = _f64(pts)
    out = np.zeros(4)
    lib().pvo_form_plane(C.c_int(len(pts)), _p(pts), C.c_double(tol), _p(out))
    return out


def form_line(pts, tol, dis_thr=0.0):
    pts = _f64(pts)
    out = np.zeros(6)
    ok = lib().pvo_form_line(C.c_int(len(pts)), _p(pts), C.c_double(tol), C.c_double(dis_thr), _p(out))
    return bool(ok), out


def sym_eig3(A):
    ev, vec = np.zeros(3), np.zeros(9)
    lib().pvo_sym_eig3(_p(_f64(A)), _p(ev), _p(vec))
    return ev, vec.reshape(3, 3)


def fast_atan2(y, x):
    y = np.ascontiguousarray(y)
    x = np.ascontiguousarray(x)
    out = np.empty_like(y)
    fn = lib().pvo_fast_atan2_f if y.dtype == np.float32 else lib().pvo_fast_atan2_d
    fn(C.c_long(y.size), _p(y), _p(x), _p(out))
    return out


def image_to_cam(rows, cols, px):
    px = _f64(px).reshape(-1, 2)
    out = np.zeros((len(px), 3))
    lib().pvo_image_to_cam_d(C.c_int(rows), C.c_int(cols), C.c_long(len(px)), _p(px), _p(out))
    return out


def cam_to_image(rows, cols, cam):
    cam = np.ascontiguousarray(cam).reshape(-1, 3)
    out = np.zeros((len(cam), 2), dtype=cam.dtype)
    fn = lib().pvo_cam_to_image_f if cam.dtype == np.float32 else lib().pvo_cam_to_image_d
    fn(C.c_int(rows), C.c_int(cols), C.c_long(len(cam)), _p(cam), _p(out))
    return out


class Blocks:
    """Parallel-array residual-block list (mirrors util/Optimization.cpp's AddResidualBlock calls)."""

    def __init__(self, type, ref, nei, consts, huber, normalize=None):
        self.type = _i32(type)
        n = len(self.type)
        self.ref = _i32(np.broadcast_to(ref, n))
        self.nei = _i32(np.broadcast_to(nei, n))
        self.consts = _f64(consts).reshape(n, 12)
        self.huber = _f64(np.broadcast_to(huber, n))
        self.normalize = _i32(np.broadcast_to(1 if normalize is None else normalize, n))
        self.n = n

    def _args(self):
        return (C.c_long(self.n), _p(self.type), _p(self.ref), _p(self.nei), _p(self.normalize), _p(self.huber), _p(self.consts))

    def evaluate(self, poses, apply_loss=True, jac=True):
        poses = _f64(poses)
        r, cost = np.zeros(self.n), np.zeros(self.n)
        J = np.zeros((self.n, 12)) if jac else None
        lib().pvo_eval_blocks(*self._args(), _p(poses), C.c_int(int(apply_loss)), _p(r), _p(J), _p(cost))
        return r, J, cost

    def normal_equations(self, poses):
        poses = _f64(poses)
        nb = poses.size // 6
        H, g = np.zeros((6 * nb, 6 * nb)), np.zeros(6 * nb)
        cost = lib().pvo_normal_equations(*self._args(), _p(poses), C.c_int(nb), _p(H), _p(g))
        return H, g, cost

    def solve_lm(self, poses, is_const=None, max_iter=20):
        poses = _f64(poses).copy()
        nb = poses.size // 6
        mask = np.zeros(nb, dtype=np.uint8) if is_const is None else np.ascontiguousarray(is_const, dtype=np.uint8)
        summ = np.zeros(6)
        lib().pvo_solve_lm(*self._args(), _p(poses), C.c_int(nb), _p(mask), C.c_int(max_iter), _p(summ))
        keys = ["initial_cost", "final_cost", "iterations", "successful", "unsuccessful", "termination"]
        return poses, dict(zip(keys, summ.tolist()))


def transform_cloud(R, t, cloud):
    cloud = _f32(cloud).reshape(-1, 4)
    out = np.empty_like(cloud)
    lib().pvo_transform_cloud(_p(_f64(R)), _p(_f64(t)), _p(cloud), C.c_int(len(cloud)), _p(out))
    return out


class Reproj:
    """Observation list of the camera-camera term (AddCameraResidual, util/Optimization.cpp:172-222, ANGLE_RESIDUAL_1): observation i sees
    point[i] from camera cam[i] along bearing[i] (unit-sphere direction of the key point)."""

    def __init__(self, cam, point, bearing, weight=1.0, huber=4.0 * np.pi / 180.0):
        self.cam, self.point, self.bearing = _i32(cam), _i32(point), _f64(bearing).reshape(-1, 3)
        self.weight, self.huber, self.n = float(weight), float(huber), len(self.cam)

    def _args(self):
        return C.c_long(self.n), _p(self.cam), _p(self.point), _p(self.bearing), C.c_double(self.weight), C.c_double(self.huber)

    def evaluate(self, cams, points, apply_loss=True, want_jacobian=True):
        r, J, cost = np.zeros(self.n), np.zeros((self.n, 9)), np.zeros(self.n)
        lib().pvo_reproj_eval(*self._args(), _p(_f64(cams)), _p(_f64(points)), C.c_int(int(apply_loss)), _p(r), _p(J) if want_jacobian else None, _p(cost))
        return r, J, cost

    def normal_equations(self, cams, points):
        cams, points = _f64(cams).reshape(-1, 6), _f64(points).reshape(-1, 3)
        x = np.concatenate([cams.ravel(), points.ravel()])
        D = x.size
        H, g = np.zeros((D, D)), np.zeros(D)
        cost = lib().pvo_reproj_normal_equations(*self._args(), _p(x), C.c_int(len(cams)), C.c_long(len(points)), _p(H), _p(g))
        return H, g, cost

    def solve_lm(self, cams, points, param_const=None, max_iter=50):
        cams, points = _f64(cams).reshape(-1, 6), _f64(points).reshape(-1, 3)
        x = np.concatenate([cams.ravel(), points.ravel()])
        mask = np.zeros(x.size, np.uint8) if param_const is None else np.ascontiguousarray(param_const, dtype=np.uint8)
        summ = np.zeros(6)
        lib().pvo_reproj_solve_lm(*self._args(), _p(x), C.c_int(len(cams)), C.c_long(len(points)), _p(mask), C.c_int(max_iter), _p(summ))
        keys = ["initial_cost", "final_cost", "iterations", "successful", "unsuccessful", "termination"]
        return x[:cams.size].reshape(-1, 6), x[cams.size:].reshape(-1, 3), dict(zip(keys, summ.tolist()))


def joint_solve_lm(blocks, reproj, poses, points, param_const=None, max_iter=20):
    """One LM over the pose blocks [cameras | LiDARs] and the points with both residual families (CameraLidarOptimizer::Optimize)."""
    poses, points = _f64(poses).reshape(-1, 6), _f64(points).reshape(-1, 3)
    x = np.concatenate([poses.ravel(), points.ravel()])
    mask = np.zeros(x.size, np.uint8) if param_const is None else np.ascontiguousarray(param_const, dtype=np.uint8)
    summ = np.zeros(6)
    lib().pvo_joint_solve_lm(*blocks._args(), *reproj._args(), _p(x), C.c_int(len(poses)), C.c_long(len(points)), _p(mask), C.c_int(max_iter), _p(summ))
    keys = ["initial_cost", "final_cost", "iterations", "successful", "unsuccessful", "termination"]
    return x[:poses.size].reshape(-1, 6), x[poses.size:].reshape(-1, 3), dict(zip(keys, summ.tolist()))


def build_calibration_blocks(rows, cols, image_lines, start, end):
    """The two residual blocks per line pair of the calibration-mode Optimize (joint_optimization/CameraLidarOptimizer.cpp:32-64)."""
    ln, s, e = _f32(image_lines).reshape(-1, 4), _f64(start).reshape(-1, 3), _f64(end).reshape(-1, 3)
    n = len(ln)
    typ, hub, consts = np.zeros(2 * n, np.int32), np.zeros(2 * n), np.zeros((2 * n, 12))
    lib().pvo_build_calibration_blocks(C.c_int(rows), C.c_int(cols), C.c_int(n), _p(ln), _p(s), _p(e), _p(typ), _p(hub), _p(consts))
    return typ, hub, consts


def filter_line_pairs(rows, cols, image_lines, start, end, by_angle, by_length):
    """CameraLidarLineAssociate::Filter (joint_optimization/CameraLidarLineAssociate.cpp:628-715) on camera-frame pairs."""
    ln, s, e = _f32(image_lines).reshape(-1, 4), _f64(start).reshape(-1, 3), _f64(end).reshape(-1, 3)
    keep, ang = np.zeros(len(ln), np.uint8), np.full(len(ln), np.float32(3.4028235e38), np.float32)
    lib().pvo_filter_line_pairs(C.c_int(rows), C.c_int(cols), C.c_int(len(ln)), _p(ln), _p(s), _p(e), C.c_int(int(by_angle)), C.c_int(int(by_length)), _p(keep), _p(ang))
    return keep.astype(bool), ang


def pixel_line_neighbors(rows, cols, lines, cloud_local, T_cl):
    """First stage of the pixel-space CameraLidarLineAssociate::Associate (joint_optimization/CameraLidarLineAssociate.cpp:22-91)."""
    lines, cloud = _f32(lines).reshape(-1, 4), _f32(cloud_local).reshape(-1, 4)
    n = len(cloud)
    line3, d2, px = np.full((n, 3), -1, np.int32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
    lib().pvo_pixel_line_neighbors(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(n), _p(_f64(T_cl)), _p(line3), _p(d2), _p(px))
    return line3, d2, px


def pixel_sub_lines(rows, cols, lines):
    lines = _f32(lines).reshape(-1, 4)
    cap = 64 + int(sum(np.hypot(l[0] - l[2], l[1] - l[3]) / 70 + 4 for l in lines))
    mid, s2l = np.zeros((cap, 2), np.float32), np.zeros(cap, np.int32)
    m = lib().pvo_pixel_sub_lines(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), C.c_int(cap), _p(mid), _p(s2l))
    assert m >= 0
    return mid[:m].copy(), s2l[:m].copy()


def slerp_pose(pose_w1, pose_w2, ratio):
    """SlerpPose (base/Geometry.hpp:572-583); poses 4x4 row-major."""
    out = np.empty((4, 4))
    lib().pvo_slerp_pose(_p(_f64(pose_w1)), _p(_f64(pose_w2)), C.c_double(ratio), _p(out))
    return out


def undistort_cloud(R_wl, t_wl, R_we, t_we, cloud):
    """Velodyne::UndistortCloud (sensors/Velodyne.cpp:1642-1674); cloud n x 4 float32 in scan order."""
    cloud = _f32(cloud).reshape(-1, 4)
    out = np.empty_like(cloud)
    lib().pvo_undistort_cloud(_p(_f64(R_wl)), _p(_f64(t_wl)), _p(_f64(R_we)), _p(_f64(t_we)), _p(cloud), C.c_long(len(cloud)), _p(out))
    return out


def undistort_end_poses(poses, pose_valid, frame_valid, gap_time):
    """Sweep-end pose per frame as LidarOdometry::UndistortLidars picks it (lidar_mapping/LidarOdometry.cpp:203-243)."""
    poses = _f64(poses).reshape(-1, 16)
    n = len(poses)
    pv, fv = np.ascontiguousarray(pose_valid, dtype=np.uint8), np.ascontiguousarray(frame_valid, dtype=np.uint8)
    out, has = np.zeros((n, 16)), np.zeros(n, dtype=np.uint8)
    lib().pvo_undistort_end_poses(C.c_int(n), _p(poses), _p(pv), _p(fv), C.c_float(gap_time), _p(out), _p(has))
    return out.reshape(n, 4, 4), has.astype(bool)


def world2local(R, t, pw):
    pw = _f64(pw).reshape(-1, 3)
    out = np.empty_like(pw)
    lib().pvo_world2local(_p(_f64(R)), _p(_f64(t)), C.c_long(len(pw)), _p(pw), _p(out))
    return out


def knn(pts, queries, k, use_kdtree=True):
    pts, queries = _f32(pts).reshape(-1, 4), _f32(queries).reshape(-1, 4)
    idx = np.empty((len(queries), k), dtype=np.int32)
    d2 = np.empty((len(queries), k), dtype=np.float32)
    lib().pvo_knn(_p(pts), C.c_int(len(pts)), _p(queries), C.c_int(len(queries)), C.c_int(k), C.c_int(int(use_kdtree)), _p(idx), _p(d2))
    return idx, d2


def associate_p2plane(ref_world, R_ref, t_ref, nei_world, R_nei, t_nei, plane_tol, dist_thr, k=10, use_kdtree=True):
    ref_world, nei_world = _f32(ref_world).reshape(-1, 4), _f32(nei_world).reshape(-1, 4)
    n = len(nei_world)
    q, pt, pl = np.empty(n, dtype=np.int32), np.empty((n, 3)), np.empty((n, 4))
    m = lib().pvo_associate_p2plane(_p(ref_world), C.c_int(len(ref_world)), _p(_f64(R_ref)), _p(_f64(t_ref)),
                                    _p(nei_world), C.c_int(n), _p(_f64(R_nei)), _p(_f64(t_nei)),
                                    C.c_double(plane_tol), C.c_float(dist_thr), C.c_int(k), C.c_int(int(use_kdtree)), _p(q), _p(pt), _p(pl))
    return q[:m].copy(), pt[:m].copy(), pl[:m].copy()


def associate_p2line(ref_world, R_ref, t_ref, nei_world, R_nei, t_nei, dist_thr, use_kdtree=True):
    ref_world, nei_world = _f32(ref_world).reshape(-1, 4), _f32(nei_world).reshape(-1, 4)
    n = len(nei_world)
    q, pt, a, b = np.empty(n, dtype=np.int32), np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
    m = lib().pvo_associate_p2line(_p(ref_world), C.c_int(len(ref_world)), _p(_f64(R_ref)), _p(_f64(t_ref)), _p(nei_world), C.c_int(n), _p(_f64(R_nei)), _p(_f64(t_nei)),
                                   C.c_float(dist_thr), C.c_int(int(use_kdtree)), _p(q), _p(pt), _p(a), _p(b))
    return q[:m].copy(), pt[:m].copy(), a[:m].copy(), b[:m].copy()


def associate_p2line_segment_knn(ref_world, ref_p2s_off, ref_p2s_ids, ref_coeffs_local, nei_world, R_nei, t_nei, dist_thr, use_kdtree=True):
    ref_world, nei_world = _f32(ref_world).reshape(-1, 4), _f32(nei_world).reshape(-1, 4)
    cap = max(1, len(nei_world) * 4)
    q, ln, pt, a, b = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32), np.empty((cap, 3)), np.empty((cap, 3)), np.empty((cap, 3))
    m = lib().pvo_associate_p2line_segment_knn(_p(ref_world), C.c_int(len(ref_world)), _p(_i32(ref_p2s_off)), _p(_i32(ref_p2s_ids)), _p(_f64(ref_coeffs_local)),
                                               _p(nei_world), C.c_int(len(nei_world)), _p(_f64(R_nei)), _p(_f64(t_nei)), C.c_float(dist_thr), C.c_int(int(use_kdtree)),
                                               _p(q), _p(ln), _p(pt), _p(a), _p(b))
    return q[:m].copy(), ln[:m].copy(), pt[:m].copy(), a[:m].copy(), b[:m].copy()


def associate_p2line_segment(ref_lines_world, ref_coeffs_local, nei_world, R_nei, t_nei, dist_thr):
    nei_world = _f32(nei_world).reshape(-1, 4)
    lw = _f64(ref_lines_world).reshape(-1, 6)
    cap = max(1, len(nei_world))
    q, ln, pt, a, b = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32), np.empty((cap, 3)), np.empty((cap, 3)), np.empty((cap, 3))
    m = lib().pvo_associate_p2line_segment(_p(lw), _p(_f64(ref_coeffs_local)), C.c_int(len(lw)), _p(nei_world), C.c_int(len(nei_world)), _p(_f64(R_nei)), _p(_f64(t_nei)),
                                           C.c_float(dist_thr), _p(q), _p(ln), _p(pt), _p(a), _p(b))
    return q[:m].copy(), ln[:m].copy(), pt[:m].copy(), a[:m].copy(), b[:m].copy()


def line2line_knn_votes(ref_world, ref_p2s_off, ref_p2s_ids, S_ref, nei_world, nei_p2s_off, nei_p2s_ids, S_nei, dist_thr, use_kdtree=True):
    ref_world, nei_world = _f32(ref_world).reshape(-1, 4), _f32(nei_world).reshape(-1, 4)
    M = np.zeros((S_nei, S_ref), dtype=np.int32)
    lib().pvo_line2line_knn_votes(_p(ref_world), C.c_int(len(ref_world)), _p(_i32(ref_p2s_off)), _p(_i32(ref_p2s_ids)), C.c_int(S_ref), _p(nei_world), C.c_int(len(nei_world)),
                                  _p(_i32(nei_p2s_off)), _p(_i32(nei_p2s_ids)), C.c_int(S_nei), C.c_float(dist_thr), C.c_int(int(use_kdtree)), _p(M))
    return M


def line_tracks(pair_a, pair_b, match_off, match_a, match_b, min_length=3, allow_multiple_map=True):
    """TrackBuilder::Build + Filter + ExportTracks (util/Tracks.cpp:58-186): returns a list of tracks, each an (m, 2) array of (frame, line)."""
    pair_a, pair_b, match_off, match_a, match_b = (_i32(x) for x in (pair_a, pair_b, match_off, match_a, match_b))
    cap = 2 * len(match_a) + 1
    off, ff, fl = np.zeros(cap + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n = lib().pvo_line_tracks(C.c_int(len(pair_a)), _p(pair_a), _p(pair_b), _p(match_off), _p(match_a), _p(match_b), C.c_int(min_length), C.c_int(int(allow_multiple_map)),
                              _p(off), _p(ff), _p(fl))
    return [np.stack([ff[off[t]:off[t + 1]], fl[off[t]:off[t + 1]]], axis=1) for t in range(n)]


def line_track_gate(tracks, ref_frame, nei_frame, ref_line, nei_line):
    off = np.zeros(len(tracks) + 1, np.int32)
    off[1:] = np.cumsum([len(t) for t in tracks])
    feats = np.concatenate(tracks) if tracks else np.zeros((0, 2), np.int32)
    ff, fl = _i32(feats[:, 0]), _i32(feats[:, 1])
    ref_line, nei_line = _i32(ref_line), _i32(nei_line)
    keep = np.zeros(len(ref_line), np.uint8)
    lib().pvo_line_track_gate(C.c_int(len(tracks)), _p(off), _p(ff), _p(fl), C.c_int(ref_frame), C.c_int(nei_frame), C.c_int(len(ref_line)), _p(ref_line), _p(nei_line), _p(keep))
    return keep.astype(bool)


def transform_lines(R, t, lines):
    lines = _f64(lines).reshape(-1, 6)
    out = np.empty_like(lines)
    lib().pvo_transform_lines(_p(_f64(R)), _p(_f64(t)), C.c_int(len(lines)), _p(lines), _p(out))
    return out


def line_votes(ref_lines_world, nei_corner_world, p2s_off, p2s_ids, S_nei, dist_thr):
    ref_lines_world = _f64(ref_lines_world).reshape(-1, 6)
    pts = _f32(nei_corner_world).reshape(-1, 4)
    M = np.zeros((S_nei, len(ref_lines_world)), dtype=np.int32)
    lib().pvo_line_votes(_p(ref_lines_world), C.c_int(len(ref_lines_world)), _p(pts), C.c_int(len(pts)), _p(_i32(p2s_off)), _p(_i32(p2s_ids)),
                         C.c_int(S_nei), C.c_double(dist_thr), _p(M))
    return M


def find_associations(ref_coeffs_local, ref_lines_world, nei_lines_world, seg_sizes_nei, M):
    S_ref, S_nei = len(ref_lines_world), len(nei_lines_world)
    on, orf = np.empty(S_nei, dtype=np.int32), np.empty(S_nei, dtype=np.int32)
    oa, ob = np.empty((S_nei, 3)), np.empty((S_nei, 3))
    m = lib().pvo_find_associations(_p(_f64(ref_coeffs_local)), _p(_f64(ref_lines_world)), C.c_int(S_ref), _p(_f64(nei_lines_world)), C.c_int(S_nei),
                                    _p(_i32(seg_sizes_nei)), _p(_i32(M)), _p(on), _p(orf), _p(oa), _p(ob))
    return on[:m].copy(), orf[:m].copy(), oa[:m].copy(), ob[:m].copy()


def angle_votes(rows, cols, lines, cloud_local, p2s_off, p2s_ids, S, T_cl):
    lines, cloud = _f32(lines).reshape(-1, 4), _f32(cloud_local).reshape(-1, 4)
    counts = np.zeros((len(lines), S), dtype=np.int32)
    lib().pvo_angle_votes(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(len(cloud)), _p(_i32(p2s_off)), _p(_i32(p2s_ids)),
                          C.c_int(S), _p(_f64(T_cl)), _p(counts))
    return counts


def unique_line_pairs(image_line, lidar_line, score):
    il, ll, sc = _i32(image_line), _i32(lidar_line), _f32(score)
    oi, ol, os_ = np.empty(max(1, len(il)), np.int32), np.empty(max(1, len(il)), np.int32), np.empty(max(1, len(il)), np.float32)
    m = lib().pvo_unique_line_pairs(C.c_int(len(il)), _p(il), _p(ll), _p(sc), _p(oi), _p(ol), _p(os_))
    return oi[:m].copy(), ol[:m].copy(), os_[:m].copy()


def associate_by_angle(rows, cols, lines, cloud_local, p2s_off, p2s_ids, seg_sizes, end_points, T_cl, filter_by_length=True, multiple_association=True,
                       image_mask=None, lidar_mask=None):
    lines, cloud = _f32(lines).reshape(-1, 4), _f32(cloud_local).reshape(-1, 4)
    S = len(seg_sizes)
    cap = max(1, len(lines) * S)
    oi, ol = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
    os_, oe, oa = np.empty((cap, 3)), np.empty((cap, 3)), np.empty(cap, dtype=np.float32)
    m = lib().pvo_associate_by_angle(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(len(cloud)), _p(_i32(p2s_off)), _p(_i32(p2s_ids)),
                                     C.c_int(S), _p(_i32(seg_sizes)), _p(_f64(end_points)), _p(_f64(T_cl)), C.c_int(int(filter_by_length)), C.c_int(cap),
                                     _p(oi), _p(ol), _p(os_), _p(oe), _p(oa), C.c_int(int(multiple_association)),
                                     _p(np.ascontiguousarray(image_mask, np.uint8)) if image_mask is not None else None,
                                     _p(np.ascontiguousarray(lidar_mask, np.uint8)) if lidar_mask is not None else None)
    return oi[:m].copy(), ol[:m].copy(), os_[:m].copy(), oe[:m].copy(), oa[:m].copy()


def project_depth(cloud, rows, cols, T_cl, size=3, want_image=True):
    cloud = _f32(cloud).reshape(-1, 4)
    img = np.zeros((rows, cols), dtype=np.uint16) if want_image else None
    uvd = np.zeros((len(cloud), 3), dtype=np.float32)
    lib().pvo_project_depth(_p(cloud), C.c_int(len(cloud)), C.c_int(rows), C.c_int(cols), _p(_f64(T_cl)), C.c_int(size), _p(img), _p(uvd))
    return img, uvd


class KdTreeHandle:
    def __init__(self, pts):
        self.pts = _f32(pts).reshape(-1, 4)
        self.h = C.c_void_p(lib().pvo_kdtree_build(_p(self.pts), C.c_int(len(self.pts))))

    def __del__(self):
        if getattr(self, "h", None):
            lib().pvo_kdtree_free(self.h)
            self.h = None


def dense_icp_eval(target_world, src_local, src_off, poses_lw, plane_tol, dist_thr, k=10, huber=0.2, weight=1.0, mode=1, tree=None):
    """One Gauss-Newton evaluation of the dense ICP sweep (configs[4]); returns (sys[n_frames,29], times[3], n_assoc)."""
    target_world, src_local = _f32(target_world).reshape(-1, 4), _f32(src_local).reshape(-1, 4)
    src_off = _i32(src_off)
    nf = len(src_off) - 1
    poses_lw = _f64(poses_lw).reshape(nf, 6)
    out, times, nassoc = np.zeros((nf, 29)), np.zeros(3), C.c_long(0)
    lib().pvo_dense_icp_eval(_p(target_world), C.c_int(len(target_world)), _p(src_local), _p(src_off), C.c_int(nf), _p(poses_lw),
                             C.c_double(plane_tol), C.c_float(dist_thr), C.c_int(k), C.c_double(huber), C.c_double(weight), C.c_int(mode),
                             tree.h if tree is not None else None, _p(out), _p(times), C.byref(nassoc))
    return out, times, nassoc.value


# ---- oracle/_ref/libpvo_ref_path.so: the reference's own CostFunction.h / Geometry.hpp / Equirectangular.{h,cpp} compiled where they lie, with the
# stand-in container types of oracle/shim (make -C oracle ref).  Checker of the checker: used by tests/ and tests/make_golden.py only. ----
REF_PAIRWISE_P2PLANE, REF_PAIRWISE_P2LINE, REF_REPROJ_1ANGLE, REF_PLANE_IOU_CAMERA = 9, 10, 11, 12


def ref_path_lib():
    path = os.path.join(_HERE, "_ref", "libpvo_ref_path.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_eval_functors.restype = C.c_long
    for f in ("ref_point_to_line_distance3d", "ref_point_to_plane_distance", "ref_vector_angle3d", "ref_plane_angle"):
        getattr(L, f).restype = C.c_double
    return L


def ref_eval_functors(type, normalize, raw, params, jac=True, consts=True):
    """The reference's functors through `Functor::Create(...)->Evaluate()`.  raw: n x 16 constructor arguments, params: n x 12 (blocks of 3 in call
    order).  Returns r[n], J[n x 12] (blocks in call order), consts[n x 12] (members after the constructor, oracle block layout)."""
    type = _i32(type)
    n = len(type)
    normalize = _i32(np.broadcast_to(normalize, n))
    raw, params = _f64(raw).reshape(n, 16), _f64(params).reshape(n, 12)
    r = np.zeros(n)
    J = np.zeros((n, 12)) if jac else None
    c = np.zeros((n, 12)) if consts else None
    rc = ref_path_lib().ref_eval_functors(C.c_long(n), _p(type), _p(normalize), _p(raw), _p(params), _p(r), _p(J), _p(c))
    assert rc == 0, f"reference functor evaluation failed at row {-1 - rc}"
    return r, J, c


def image_to_cam_f(rows, cols, px, r):
    px = _f32(px).reshape(-1, 2)
    out = np.zeros((len(px), 3), np.float32)
    lib().pvo_image_to_cam_f(C.c_int(rows), C.c_int(cols), C.c_long(len(px)), _p(px), C.c_float(r), _p(out))
    return out


def break_to_segments(rows, cols, line4, seg_length):
    """Equirectangular::BreakToSegments (sensors/Equirectangular.cpp:20-58) of one image line."""
    buf = np.zeros((512, 2), np.float32)
    k = lib().pvo_break_to_segments(C.c_int(rows), C.c_int(cols), _p(_f32(line4)), C.c_float(seg_length), C.c_int(512), _p(buf))
    assert k >= 0
    return buf[:k].copy()


def form_plane3(p1, p2, p3):
    out = np.zeros(4)
    lib().pvo_form_plane3(_p(_f64(p1)), _p(_f64(p2)), _p(_f64(p3)), _p(out))
    return out


def geometry_helpers(point, line6, plane4):
    out, proj = np.zeros(5), np.zeros((2, 3))
    lib().pvo_geometry_helpers(_p(_f64(point)), _p(_f64(line6)), _p(_f64(plane4)), _p(out), _p(proj))
    return out, proj


# ---- oracle/_ref/libpvo_ref_assoc.so: the reference's own lidar_mapping/LidarFeatureAssociate.cpp compiled where it lies (make -C oracle ref) ----
def ref_assoc_lib():
    path = os.path.join(_HERE, "_ref", "libpvo_ref_assoc.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_frame_create.restype = C.c_void_p
    return L


class RefFrame:
    """A `Velodyne` object of the reference as its association functions see it (world-frame feature clouds, segment tables, pose)."""

    def __init__(self, R_wl, t_wl, corner_world=None, p2s_off=None, p2s_ids=None, coeffs_local=None, surf_flat_world=None, surf_less_flat_world=None, id=0,
                 valid=True, pose_valid=True, local=False, end_points=None):
        """local=True: the clouds are given in the SENSOR frame and the reference's own Transform2LidarWorld() moves them; local="keep": sensor-frame clouds
        that stay there (for the entry points that run the reference's own pipeline stages, which transform the frames themselves)."""
        self.L = ref_assoc_lib()
        z = np.zeros((0, 4), np.float32)
        cw = _f32(z if corner_world is None else corner_world).reshape(-1, 4)
        sf = _f32(z if surf_flat_world is None else surf_flat_world).reshape(-1, 4)
        sl = _f32(z if surf_less_flat_world is None else surf_less_flat_world).reshape(-1, 4)
        co = _f64(np.zeros((0, 6)) if coeffs_local is None else coeffs_local).reshape(-1, 6)
        off = None if p2s_off is None else _i32(p2s_off)
        ids = None if p2s_ids is None else _i32(p2s_ids)
        self.h = self.L.ref_frame_create(C.c_int(id), C.c_int(int(valid)), C.c_int(int(pose_valid)), _p(_f64(R_wl)), _p(_f64(t_wl)), _p(cw), C.c_int(len(cw)), _p(off), _p(ids),
                                         C.c_int(len(co)), _p(co), None, _p(sf), C.c_int(len(sf)), _p(sl), C.c_int(len(sl)), C.c_int(2 if local == "keep" else (0 if local else 1)), _p(None if end_points is None else _f64(end_points)))
        assert self.h
        self.n_corner, self.n_flat = len(cw), len(sf)

    def cloud(self, which):
        """which: "corner" / "flat" / "less_flat": the frame's cloud as the reference holds it now (n x 4 float32)."""
        k = {"corner": 0, "flat": 1, "less_flat": 2}[which]
        out = np.zeros((1 << 16, 4), np.float32)
        m = self.L.ref_frame_get_cloud(C.c_void_p(self.h), C.c_int(k), C.c_int(len(out)), _p(out))
        assert m >= 0
        return out[:m].copy()

    def to_local(self):
        self.L.ref_frame_to_local(C.c_void_p(self.h))

    def to_world(self):
        self.L.ref_frame_to_world(C.c_void_p(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_frame_destroy(C.c_void_p(self.h))
            self.h = None


def ref_associate_point2plane(ref, nei, plane_tol, dist_thr):
    cap = max(1, nei.n_flat)
    pt, pl = np.empty((cap, 3)), np.empty((cap, 4))
    m = ref.L.ref_associate_point2plane(C.c_void_p(ref.h), C.c_void_p(nei.h), C.c_double(plane_tol), C.c_float(dist_thr), C.c_int(cap), _p(pt), _p(pl))
    assert m >= 0
    return pt[:m].copy(), pl[:m].copy()


def ref_associate_point2line(ref, nei, dist_thr, variant=""):
    """variant: "" (AssociatePoint2Line), "_segment_knn", "_segment"."""
    cap = max(1, 4 * nei.n_corner)
    pt, a, b = np.empty((cap, 3)), np.empty((cap, 3)), np.empty((cap, 3))
    m = getattr(ref.L, "ref_associate_point2line" + variant)(C.c_void_p(ref.h), C.c_void_p(nei.h), C.c_float(dist_thr), C.c_int(cap), _p(pt), _p(a), _p(b))
    assert m >= 0
    return pt[:m].copy(), a[:m].copy(), b[:m].copy()


def ref_associate_line2line(ref, nei, dist_thr, knn=False):
    cap = 4096
    ni, ri, a, b = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty((cap, 3)), np.empty((cap, 3))
    m = getattr(ref.L, "ref_associate_line2line_knn" if knn else "ref_associate_line2line")(C.c_void_p(ref.h), C.c_void_p(nei.h), C.c_float(dist_thr), C.c_int(cap), _p(ni), _p(ri), _p(a), _p(b))
    assert m >= 0
    return ni[:m].copy(), ri[:m].copy(), a[:m].copy(), b[:m].copy()


def ref_find_neighbors(R_wl, t_wl, pose_valid, valid, neighbor_size):
    R_wl, t_wl = _f64(R_wl).reshape(-1, 9), _f64(t_wl).reshape(-1, 3)
    n = len(t_wl)
    pv, va = np.ascontiguousarray(pose_valid, np.uint8), np.ascontiguousarray(valid, np.uint8)
    cap = n * (neighbor_size + 64) + 64
    off, ids = np.zeros(n + 1, np.int32), np.zeros(cap, np.int32)
    m = ref_assoc_lib().ref_find_neighbors(C.c_int(n), _p(R_wl), _p(t_wl), _p(pv), _p(va), C.c_int(neighbor_size), C.c_int(cap), _p(off), _p(ids))
    assert m >= 0
    return [ids[off[i]:off[i + 1]].tolist() for i in range(n)]


# ---- oracle/_ref/libpvo_ref_camlidar.so: the reference's own CameraLidarLineAssociate.cpp + ProjectLidar2PanoramaDepth (make -C oracle ref) ----
def ref_camlidar_lib():
    path = os.path.join(_HERE, "_ref", "libpvo_ref_camlidar.so")
    return C.CDLL(path) if os.path.exists(path) else None


def ref_associate_by_angle(rows, cols, lines, cloud_local, p2s_off, p2s_ids, coeffs, end_points, T_cl, multiple_association=True, image_mask=None, lidar_mask=None):
    lines, cloud = _f32(lines).reshape(-1, 4), _f32(cloud_local).reshape(-1, 4)
    coeffs = _f64(coeffs).reshape(-1, 6)
    S = len(coeffs)
    cap = max(1, len(lines) * S)
    oi, ol = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
    os_, oe, oa = np.empty((cap, 3)), np.empty((cap, 3)), np.empty(cap, dtype=np.float32)
    m = ref_camlidar_lib().ref_associate_by_angle(
        C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(len(cloud)), _p(_i32(p2s_off)), _p(_i32(p2s_ids)), C.c_int(S), _p(coeffs),
        _p(_f64(end_points)), _p(_f64(T_cl)), C.c_int(int(multiple_association)), _p(np.ascontiguousarray(image_mask, np.uint8)) if image_mask is not None else None,
        _p(np.ascontiguousarray(lidar_mask, np.uint8)) if lidar_mask is not None else None, C.c_int(cap), _p(oi), _p(ol), _p(os_), _p(oe), _p(oa))
    assert m >= 0
    return oi[:m].copy(), ol[:m].copy(), os_[:m].copy(), oe[:m].copy(), oa[:m].copy()


def ref_project_depth(cloud, rows, cols, T_cl, size=3):
    cloud = _f32(cloud).reshape(-1, 4)
    img = np.zeros((rows, cols), dtype=np.uint16)
    ref_camlidar_lib().ref_project_lidar2panorama_depth(_p(cloud), C.c_long(len(cloud)), C.c_int(rows), C.c_int(cols), _p(_f64(T_cl)), C.c_int(size), _p(img))
    return img


def ref_generate_line_tracks(frames, neighbor_size, min_track_length):
    """LidarLineMatch::GenerateTracks of the reference over a list of RefFrame; returns a list of (m, 2) arrays of (frame, line)."""
    n = len(frames)
    arr = (C.c_void_p * n)(*[f.h for f in frames])
    cap = 1 << 20
    off, ff, fl = np.zeros(cap + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    m = frames[0].L.ref_generate_line_tracks(C.c_int(n), arr, C.c_int(neighbor_size), C.c_int(min_track_length), C.c_int(cap), _p(off), _p(ff), _p(fl))
    assert m >= 0
    return [np.stack([ff[off[t]:off[t + 1]], fl[off[t]:off[t + 1]]], axis=1) for t in range(m)]


def ref_refine_pose_blocks(frames, point_to_plane=True, line_to_line=True, point_to_line=False, use_segment=True, angle_residual=True, normalize_distance=True,
                           plane_dis_threshold=1.0, line_dis_threshold=0.3, plane_tolerance=0.05):
    """The problem that the reference's own LidarOdometry::RefinePose hands to ceres::Solve (recorded by the stand-in's solve hook) over a list of RefFrame, every block
    evaluated once (raw residual / 1x12 Jacobian).  Returns dict(ref, nei, huber, residual, jacobian, poses, info = [max_num_iterations, linear_solver_type,
    number of constant blocks, first valid frame])."""
    n = len(frames)
    arr = (C.c_void_p * n)(*[f.h for f in frames])
    cap = 1 << 21
    rf, nf, hb, r, J, poses = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap), np.zeros(cap), np.zeros((cap, 12)), np.zeros((n, 6))
    info = np.zeros(4, np.int32)
    L = frames[0].L
    L.ref_refine_pose_blocks.restype = C.c_long
    m = L.ref_refine_pose_blocks(C.c_int(n), arr, C.c_int(int(point_to_plane)), C.c_int(int(line_to_line)), C.c_int(int(point_to_line)), C.c_int(int(use_segment)),
                                 C.c_int(int(angle_residual)), C.c_int(int(normalize_distance)), C.c_double(plane_dis_threshold), C.c_double(line_dis_threshold),
                                 C.c_double(plane_tolerance), C.c_long(cap), _p(rf), _p(nf), _p(hb), _p(r), _p(J), _p(poses), _p(info))
    assert m >= 0, m
    return dict(ref=rf[:m].copy(), nei=nf[:m].copy(), huber=hb[:m].copy(), residual=r[:m].copy(), jacobian=J[:m].copy(), poses=poses, info=info)


def ref_camera_lidar_blocks(rows, cols, image_lines, start, end, pair_weight, R_wc, t_wc, R_wl, t_wl, weight):
    """AddCameraLidarResidual of the reference for one (image, LiDAR) frame pair: raw residuals / Jacobians of the 2 n blocks + the two pose blocks."""
    ln, s, e = _f32(image_lines).reshape(-1, 4), _f64(start).reshape(-1, 3), _f64(end).reshape(-1, 3)
    n = len(ln)
    pw = None if pair_weight is None else _f32(pair_weight)
    r, J, poses = np.zeros(2 * n), np.zeros((2 * n, 12)), np.zeros((2, 6))
    L = ref_assoc_lib()
    L.ref_camera_lidar_blocks.restype = C.c_long
    m = L.ref_camera_lidar_blocks(C.c_int(rows), C.c_int(cols), C.c_int(n), _p(ln), _p(s), _p(e), _p(pw), _p(_f64(R_wc)), _p(_f64(t_wc)), _p(_f64(R_wl)), _p(_f64(t_wl)),
                                  C.c_double(weight), C.c_long(2 * n), _p(r), _p(J), _p(poses))
    assert m == 2 * n, m
    return r, J, poses


def ref_camera_residual_blocks(rows, cols, R_wc, t_wc, pose_valid, kp_off, kp_xy, track_off, feat_frame, feat_index, points3, weight=1.0):
    """AddCameraResidual (ANGLE_RESIDUAL_1) of the reference: returns dict(cam, track, residual, jacobian (n x 9), cams (n_frames x 6 pose blocks))."""
    R_wc, t_wc = _f64(R_wc).reshape(-1, 9), _f64(t_wc).reshape(-1, 3)
    nf = len(t_wc)
    pv = np.ascontiguousarray(pose_valid, np.uint8)
    kp_off, track_off, feat_frame, feat_index = _i32(kp_off), _i32(track_off), _i32(feat_frame), _i32(feat_index)
    kp_xy, points3 = _f32(kp_xy).reshape(-1, 2), _f64(points3).reshape(-1, 3)
    cap = len(feat_frame)
    cam, trk, r, J, cams = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap), np.zeros((cap, 9)), np.zeros((nf, 6))
    L = ref_assoc_lib()
    L.ref_camera_residual_blocks.restype = C.c_long
    m = L.ref_camera_residual_blocks(C.c_int(rows), C.c_int(cols), C.c_int(nf), _p(R_wc), _p(t_wc), _p(pv), _p(kp_off), _p(kp_xy), C.c_int(len(track_off) - 1), _p(track_off),
                                     _p(feat_frame), _p(feat_index), _p(points3), C.c_double(weight), C.c_long(cap), _p(cam), _p(trk), _p(r), _p(J), _p(cams))
    assert m >= 0, m
    return dict(cam=cam[:m].copy(), track=trk[:m].copy(), residual=r[:m].copy(), jacobian=J[:m].copy(), cams=cams)


def ref_undistort_cloud(R_wl, t_wl, R_we, t_we, cloud):
    """Velodyne::UndistortCloud of the reference on one raw sweep (n x 4 float32)."""
    cloud = _f32(cloud).reshape(-1, 4)
    out = np.empty_like(cloud)
    ok = ref_assoc_lib().ref_undistort_cloud(_p(_f64(R_wl)), _p(_f64(t_wl)), _p(_f64(R_we)), _p(_f64(t_we)), _p(cloud), C.c_long(len(cloud)), _p(out))
    return bool(ok), out


def ref_undistort_lidars(R_wl, t_wl, pose_valid, valid, off, clouds, gap_time=0.0):
    """LidarOdometry::UndistortLidars of the reference: returns the sweeps (concatenated, n x 4 float32) as it leaves them."""
    R_wl, t_wl = _f64(R_wl).reshape(-1, 9), _f64(t_wl).reshape(-1, 3)
    clouds, off = _f32(clouds).reshape(-1, 4), _i32(off)
    out = np.empty_like(clouds)
    ref_assoc_lib().ref_undistort_lidars(C.c_int(len(t_wl)), _p(R_wl), _p(t_wl), _p(np.ascontiguousarray(pose_valid, np.uint8)), _p(np.ascontiguousarray(valid, np.uint8)),
                                         _p(off), _p(clouds), C.c_float(gap_time), _p(out))
    return out


def ref_export_pose_t(path, R, t, names=None):
    R, t = _f64(R).reshape(-1, 9), _f64(t).reshape(-1, 3)
    arr = None if names is None else (C.c_char_p * len(R))(*[nm.encode() for nm in names])
    ref_assoc_lib().ref_export_pose_t(str(path).encode(), C.c_int(len(R)), _p(R), _p(t), arr)


def ref_read_pose_t(path, with_invalid=False, cap=4096):
    R, t, names = np.zeros((cap, 9)), np.zeros((cap, 3)), C.create_string_buffer(256 * cap)
    m = ref_assoc_lib().ref_read_pose_t(str(path).encode(), C.c_int(int(with_invalid)), C.c_int(cap), _p(R), _p(t), names)
    assert m >= 0, m
    return R[:m].reshape(-1, 3, 3).copy(), t[:m].copy(), [names.raw[256 * i:256 * (i + 1)].split(b"\0")[0].decode() for i in range(m)]


def ref_neighbor_each_frame(R_wc, t_wc, frame_pose_valid, R_wl, t_wl, lidar_pose_valid, lidar_valid, neighbor_size, temporal):
    R_wc, t_wc, R_wl, t_wl = _f64(R_wc).reshape(-1, 9), _f64(t_wc).reshape(-1, 3), _f64(R_wl).reshape(-1, 9), _f64(t_wl).reshape(-1, 3)
    nf, nl = len(t_wc), len(t_wl)
    cap = nf * (neighbor_size + 4) + 16
    off, ids = np.zeros(nf + 1, np.int32), np.zeros(cap, np.int32)
    m = ref_assoc_lib().ref_neighbor_each_frame(C.c_int(nf), _p(R_wc), _p(t_wc), _p(np.ascontiguousarray(frame_pose_valid, np.uint8)), C.c_int(nl), _p(R_wl), _p(t_wl),
                                                _p(np.ascontiguousarray(lidar_pose_valid, np.uint8)), _p(np.ascontiguousarray(lidar_valid, np.uint8)), C.c_int(neighbor_size),
                                                C.c_int(int(temporal)), C.c_int(cap), _p(off), _p(ids))
    assert m >= 0
    return [ids[off[i]:off[i + 1]].tolist() for i in range(nf)]


def ref_lidar_mask_by_track(frames, min_track_length=3, neighbor_size=3):
    n = len(frames)
    arr = (C.c_void_p * n)(*[f.h for f in frames])
    cap = 1 << 16
    off, mask = np.zeros(n + 1, np.int32), np.zeros(cap, np.uint8)
    m = frames[0].L.ref_lidar_mask_by_track(C.c_int(n), arr, C.c_int(min_track_length), C.c_int(neighbor_size), C.c_int(cap), _p(off), _p(mask))
    assert m >= 0
    return [mask[off[i]:off[i + 1]].astype(bool) for i in range(n)]


def ref_joint_optimize_blocks(rows, cols, R_wc, t_wc, image_lines, keypoints, lidar_frames, track_off, feat_frame, feat_index, points3, T_cl_init, neighbor_size_joint=1,
                              camera_weight=1.0, lidar_weight=0.01, camera_lidar_weight=25.0, point_to_plane=True, line_to_line=True, point_to_line=False, angle_residual=True,
                              normalize_distance=True, plane_dis_threshold=1.0, line_dis_threshold=0.3, plane_tolerance=0.05, refine=(True, True, True, True, True)):
    """The problem the reference's own mapping-mode CameraLidarOptimizer::Optimize hands to ceres::Solve (line pairs from its AssociateLineMulti).  image_lines / keypoints:
    one array per camera frame.  Returns dict(n_params, a, b, huber, residual, jacobian, const_part, poses, n_line_pairs)."""
    R_wc, t_wc = _f64(R_wc).reshape(-1, 9), _f64(t_wc).reshape(-1, 3)
    nc, nl = len(t_wc), len(lidar_frames)
    line_off = np.concatenate([[0], np.cumsum([len(x) for x in image_lines])]).astype(np.int32)
    lines = _f32(np.concatenate([np.asarray(x, np.float32).reshape(-1, 4) for x in image_lines]))
    kp_off = np.concatenate([[0], np.cumsum([len(x) for x in keypoints])]).astype(np.int32)
    kp = _f32(np.concatenate([np.asarray(x, np.float32).reshape(-1, 2) for x in keypoints]))
    arr = (C.c_void_p * nl)(*[f.h for f in lidar_frames])
    track_off, feat_frame, feat_index, points3 = _i32(track_off), _i32(feat_frame), _i32(feat_index), _f64(points3).reshape(-1, 3)
    cap, ccap = 1 << 21, 1 << 16
    npar, a, b, hb, r, J = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap), np.zeros(cap), np.zeros((cap, 12))
    cpart, ncst, poses, nlp = np.zeros(ccap, np.int32), C.c_int(), np.zeros((nc + nl, 6)), C.c_int()
    L = lidar_frames[0].L
    L.ref_joint_optimize_blocks.restype = C.c_long
    m = L.ref_joint_optimize_blocks(C.c_int(rows), C.c_int(cols), C.c_int(nc), _p(R_wc), _p(t_wc), _p(line_off), _p(lines), _p(kp_off), _p(kp), C.c_int(nl), arr,
                                    C.c_int(len(track_off) - 1), _p(track_off), _p(feat_frame), _p(feat_index), _p(points3), _p(_f64(T_cl_init)), C.c_int(neighbor_size_joint),
                                    C.c_double(camera_weight), C.c_double(lidar_weight), C.c_double(camera_lidar_weight), C.c_int(int(point_to_plane)), C.c_int(int(line_to_line)),
                                    C.c_int(int(point_to_line)), C.c_int(int(angle_residual)), C.c_int(int(normalize_distance)), C.c_double(plane_dis_threshold),
                                    C.c_double(line_dis_threshold), C.c_double(plane_tolerance), *[C.c_int(int(x)) for x in refine], C.c_long(cap), _p(npar), _p(a), _p(b), _p(hb),
                                    _p(r), _p(J), C.c_int(ccap), _p(cpart), C.byref(ncst), _p(poses), C.byref(nlp))
    assert m >= 0, m
    return dict(n_params=npar[:m].copy(), a=a[:m].copy(), b=b[:m].copy(), huber=hb[:m].copy(), residual=r[:m].copy(), jacobian=J[:m].copy(), const_part=cpart[:ncst.value].copy(),
                poses=poses, n_line_pairs=nlp.value)


def ref_calibration_blocks(rows, cols, image_lines, lidar_frames, T_cl):
    """Calibration mode of the reference: AssociateLineSingle(T_cl) + Optimize(line_pairs, T_cl) recorded at ceres::Solve.  Returns dict(huber, residual, jacobian (n x 6),
    pose (aa_cl, t_cl), info = [line pairs, max_num_iterations, linear_solver_type])."""
    n = len(lidar_frames)
    line_off = np.concatenate([[0], np.cumsum([len(x) for x in image_lines])]).astype(np.int32)
    lines = _f32(np.concatenate([np.asarray(x, np.float32).reshape(-1, 4) for x in image_lines]))
    arr = (C.c_void_p * n)(*[f.h for f in lidar_frames])
    cap = 1 << 16
    hb, r, J, pose, info = np.zeros(cap), np.zeros(cap), np.zeros((cap, 6)), np.zeros(6), np.zeros(3, np.int32)
    L = lidar_frames[0].L
    L.ref_calibration_blocks.restype = C.c_long
    m = L.ref_calibration_blocks(C.c_int(rows), C.c_int(cols), C.c_int(n), _p(line_off), _p(lines), arr, _p(_f64(T_cl)), C.c_long(cap), _p(hb), _p(r), _p(J), _p(pose), _p(info))
    assert m >= 0, m
    return dict(huber=hb[:m].copy(), residual=r[:m].copy(), jacobian=J[:m].copy(), pose=pose, info=info)


def ref_pixel_associate_candidates(rows, cols, lines, cloud_local, T_cl):
    """First stage of the reference's pixel-space Associate(): the candidate point lists (camera-frame float32 x, y, z) handed to the RANSAC fit, one per image line with
    at least 6 candidates, in ascending line order."""
    lines, cloud = _f32(lines).reshape(-1, 4), _f32(cloud_local).reshape(-1, 4)
    cap_l, cap_p = len(lines) + 1, 3 * len(cloud) + 16
    off, xyz = np.zeros(cap_l + 1, np.int32), np.zeros((cap_p, 3), np.float32)
    m = ref_camlidar_lib().ref_pixel_associate_candidates(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(len(cloud)), _p(_f64(T_cl)),
                                                          C.c_int(cap_l), C.c_long(cap_p), _p(off), _p(xyz))
    assert m >= 0
    return [xyz[off[k]:off[k + 1]].copy() for k in range(m)]


def ref_set_sac_script(inlier_lists):
    """Scripts (None: clears) the RANSAC answers inside libpvo_ref_assoc.so - see oracle/ref_assoc_wrap.cpp; returns the number of calls answered by the previous script."""
    L = ref_assoc_lib()
    if inlier_lists is None:
        return L.ref_set_sac_script(C.c_int(-1), None, None)
    off = np.concatenate([[0], np.cumsum([len(x) for x in inlier_lists])]).astype(np.int32)
    idx = _i32(np.concatenate([np.asarray(x, np.int32) for x in inlier_lists] + [np.zeros(0, np.int32)]))
    return L.ref_set_sac_script(C.c_int(len(inlier_lists)), _p(off), _p(idx))


def ref_pixel_associate_scripted(rows, cols, lines, cloud_local, T_cl, inlier_lists):
    """The reference's whole pixel-space Associate() with the RANSAC's inliers scripted (inlier_lists[k] = indices into the k-th candidate list, in the order of
    ref_pixel_associate_candidates).  Returns (image_line4 float32, start (n,3), end (n,3), angle float32) of the surviving pairs."""
    lines, cloud = _f32(lines).reshape(-1, 4), _f32(cloud_local).reshape(-1, 4)
    off = np.concatenate([[0], np.cumsum([len(x) for x in inlier_lists])]).astype(np.int32)
    idx = _i32(np.concatenate([np.asarray(x, np.int32) for x in inlier_lists] + [np.zeros(0, np.int32)]))
    cap = len(lines) + 1
    il, s, e, ang = np.zeros((cap, 4), np.float32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(cap, np.float32)
    m = ref_camlidar_lib().ref_pixel_associate_scripted(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(len(cloud)), _p(_f64(T_cl)),
                                                        C.c_int(len(inlier_lists)), _p(off), _p(idx), C.c_int(cap), _p(il), _p(s), _p(e), _p(ang))
    assert m >= 0, m
    return il[:m].copy(), s[:m].copy(), e[:m].copy(), ang[:m].copy()


def ref_pixel_associate_segmented_scripted(rows, cols, lines, segments, T_cl, inlier_lists):
    """The reference's segmented pixel-space Associate(lines, segmented_cloud, T_cl) (CameraLidarLineAssociate.cpp:191-338) with the RANSAC's inliers scripted
    (inlier_lists[k] = indices into the majority segment of the k-th image line that reaches the fit).  Returns (image_line4, start, end, angle) of the surviving pairs,
    or None when the script does not match the number of fits."""
    lines = _f32(lines).reshape(-1, 4)
    segs = [_f32(x).reshape(-1, 4) for x in segments]
    seg_off = np.concatenate([[0], np.cumsum([len(x) for x in segs])]).astype(np.int32)
    cloud = _f32(np.concatenate(segs + [np.zeros((0, 4), np.float32)]))
    off = np.concatenate([[0], np.cumsum([len(x) for x in inlier_lists])]).astype(np.int32)
    idx = _i32(np.concatenate([np.asarray(x, np.int32) for x in inlier_lists] + [np.zeros(0, np.int32)]))
    cap = len(lines) + 1
    il, s, e, ang = np.zeros((cap, 4), np.float32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(cap, np.float32)
    m = ref_camlidar_lib().ref_pixel_associate_segmented_scripted(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(len(segs)), _p(seg_off),
                                                                  _p(_f64(T_cl)), C.c_int(len(inlier_lists)), _p(off), _p(idx), C.c_int(cap), _p(il), _p(s), _p(e), _p(ang))
    if m == -2:
        return None
    assert m >= 0, m
    return il[:m].copy(), s[:m].copy(), e[:m].copy(), ang[:m].copy()


def ref_joint_optimize_loop(rows, cols, R_wc, t_wc, image_lines, lidar_frames, T_cl_init, num_iteration_joint, script_cost, script_steps):
    """The reference's own mapping-mode JointOptimize with a scripted solver (k-th solve reports script_cost[k] / script_steps[k]); returns the number of solver calls."""
    R_wc, t_wc = _f64(R_wc).reshape(-1, 9), _f64(t_wc).reshape(-1, 3)
    line_off = np.concatenate([[0], np.cumsum([len(x) for x in image_lines])]).astype(np.int32)
    lines = _f32(np.concatenate([np.asarray(x, np.float32).reshape(-1, 4) for x in image_lines]))
    arr = (C.c_void_p * len(lidar_frames))(*[f.h for f in lidar_frames])
    sc, ss = _f64(script_cost), _i32(script_steps)
    return lidar_frames[0].L.ref_joint_optimize_loop(C.c_int(rows), C.c_int(cols), C.c_int(len(t_wc)), _p(R_wc), _p(t_wc), _p(line_off), _p(lines), C.c_int(len(lidar_frames)), arr,
                                                     _p(_f64(T_cl_init)), C.c_int(num_iteration_joint), C.c_int(len(sc)), _p(sc), _p(ss))


def ref_calibration_loop(rows, cols, image_lines, lidar_frames, T_cl, deltas6):
    """The reference's own calibration-mode JointOptimize with a scripted solver (the k-th solve adds deltas6[k] to (aa_cl, t_cl)); returns (solver calls, T_cl result)."""
    n = len(lidar_frames)
    line_off = np.concatenate([[0], np.cumsum([len(x) for x in image_lines])]).astype(np.int32)
    lines = _f32(np.concatenate([np.asarray(x, np.float32).reshape(-1, 4) for x in image_lines]))
    arr = (C.c_void_p * n)(*[f.h for f in lidar_frames])
    d6 = _f64(deltas6).reshape(-1, 6)
    T_out = np.zeros((4, 4))
    m = lidar_frames[0].L.ref_calibration_loop(C.c_int(rows), C.c_int(cols), C.c_int(n), _p(line_off), _p(lines), arr, _p(_f64(T_cl)), C.c_int(len(d6)), _p(d6), _p(T_out))
    return m, T_out
