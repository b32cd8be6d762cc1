"""ctypes mirror of include/panovlm_b200.h (same names, argument meaning and error behaviour).

Every method forwards to one ``pvb_*`` entry point; errors (negative return codes) raise :class:`PvbError` with
``pvb_last_error``.  numpy arrays are the host buffers; nothing here computes on the CPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

P2PLANE_METER, P2PLANE_ANGLE, P2LINE_METER, P2LINE_ANGLE, PLANE2PLANE_GLOBAL, PLANE_IOU = range(6)
PLANE2PLANE_RELATIVE, PLANE_RELATIVE_IOU, LINE2LINE_ANGLE = 6, 7, 8
SOLVER_AUTO, SOLVER_HOST, SOLVER_DEVICE, SOLVER_PCG = 0, 1, 2, 3   # pvb_blocks_set_linear_solver

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PvbError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libpanovlm_b200.so")


def build_library():
    """nvcc -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])


class _Frame(C.Structure):
    _fields_ = [("surf_target", C.c_void_p), ("n_target", C.c_int), ("surf_query", C.c_void_p), ("n_query", C.c_int)]


class _LineFrame(C.Structure):
    _fields_ = [("corner_local", C.c_void_p), ("n_corner", C.c_int), ("p2s_off", C.c_void_p), ("p2s_ids", C.c_void_p), ("n_segments", C.c_int),
                ("segment_coeffs", C.c_void_p), ("end_points", C.c_void_p), ("R_wl", C.c_void_p), ("t_wl", C.c_void_p)]


class LineFrame:
    """Host arrays of one frame's corner features (pvb_line_frame); keeps the numpy buffers alive."""

    def __init__(self, corner_local, p2s_off, p2s_ids, segment_coeffs, end_points, R_wl, t_wl):
        self.corner = _arr(corner_local, np.float32).reshape(-1, 4)
        self.p2s_off, self.p2s_ids = _arr(p2s_off, np.int32), _arr(p2s_ids, np.int32)
        self.coeffs = _arr(segment_coeffs, np.float64).reshape(-1, 6)
        self.ends = None if end_points is None else _arr(end_points, np.float64).reshape(-1, 6)
        self.R, self.t = _arr(R_wl, np.float64).reshape(3, 3), _arr(t_wl, np.float64).reshape(3)
        self.c = _LineFrame(self.corner.ctypes.data, len(self.corner), self.p2s_off.ctypes.data, self.p2s_ids.ctypes.data if len(self.p2s_ids) else None,
                            len(self.coeffs), self.coeffs.ctypes.data if len(self.coeffs) else None, None if self.ends is None else self.ends.ctypes.data,
                            self.R.ctypes.data, self.t.ctypes.data)


class BlockList:
    """Growable parallel arrays of residual blocks filled by the pvb_build_* entry points."""

    def __init__(self, cap):
        self.cap, self.n = cap, 0
        self.type, self.ref, self.nei, self.normalize = (np.zeros(cap, np.int32) for _ in range(4))
        self.huber, self.consts = np.zeros(cap), np.zeros((cap, 12))

    def args(self):
        return (C.c_long(self.n), C.c_long(self.cap), _p(self.type), _p(self.ref), _p(self.nei), _p(self.normalize), _p(self.huber), _p(self.consts))

    def view(self):
        n = self.n
        return dict(type=self.type[:n], ref=self.ref[:n], nei=self.nei[:n], normalize=self.normalize[:n], huber=self.huber[:n], consts=self.consts[:n])

    def extend(self, blocks):
        """Append blocks given as a dict of parallel arrays (the layout view() returns)."""
        m = len(blocks["type"])
        if self.n + m > self.cap:
            raise PvbError("BlockList.extend: capacity exceeded")
        for k in ("type", "ref", "nei", "normalize", "huber", "consts"):
            getattr(self, k)[self.n:self.n + m] = blocks[k]
        self.n += m


class _AssocParams(C.Structure):
    _fields_ = [("plane_tolerance", C.c_double), ("dist_threshold", C.c_float), ("k", C.c_int), ("cell_size", C.c_double)]


class _DenseParams(C.Structure):
    _fields_ = [("plane_tolerance", C.c_double), ("dist_threshold", C.c_float), ("k", C.c_int), ("residual_type", C.c_int),
                ("normalize", C.c_int), ("huber", C.c_double), ("weight", C.c_double)]


def load_library():
    """Loads libpanovlm_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PvbError(f"{path} is missing: build it with panovlm_b200.build_library() / __graft_entry__.build() "
                       "(there is no CPU fallback)")
    L = C.CDLL(path)
    L.pvb_last_error.restype = C.c_char_p
    L.pvb_kernel_launches.restype = C.c_long
    L.pvb_stream.restype = C.c_void_p
    L.pvb_blocks_edge_systems_ptr.restype = C.POINTER(C.c_double)
    L.pvb_blocks_residuals.restype = C.POINTER(C.c_double)
    L.pvb_blocks_jacobians.restype = C.POINTER(C.c_double)
    L.pvb_reproj_residuals.restype = C.POINTER(C.c_double)
    L.pvb_reproj_jacobians.restype = C.POINTER(C.c_double)
    _LIB = L
    return L


EXPORTS = [
    "pvb_create", "pvb_destroy", "pvb_last_error", "pvb_set_stream", "pvb_synchronize", "pvb_kernel_launches", "pvb_stream",
    "pvb_blocks_set", "pvb_blocks_evaluate", "pvb_blocks_residuals", "pvb_blocks_jacobians", "pvb_blocks_cost", "pvb_blocks_kernel_time_ms", "pvb_blocks_num_edges",
    "pvb_blocks_edges", "pvb_blocks_edge_systems", "pvb_blocks_edge_systems_ptr", "pvb_blocks_dense_system", "pvb_blocks_solve_lm", "pvb_blocks_pcg_stats",
    "pvb_frames_set", "pvb_frames_associate_point2plane", "pvb_frames_get_point2plane", "pvb_frames_knn", "pvb_frames_set_corners", "pvb_frames_associate_point2line", "pvb_frames_get_point2line",
    "pvb_dense_set_target", "pvb_dense_set_sources", "pvb_dense_evaluate", "pvb_dense_evaluate_device", "pvb_dense_gauss_newton_step",
    "pvb_dense_get_rows", "pvb_dense_set_hints", "pvb_dense_reset_hints", "pvb_debug_counters", "pvb_dense_order_stats", "pvb_dense_kernel_time_ms", "pvb_project_equirect", "pvb_project_depth_image", "pvb_line_votes", "pvb_angle_votes",
    "pvb_find_neighbors", "pvb_line2line_associate", "pvb_camera_lidar_associate", "pvb_build_point2plane_blocks", "pvb_build_point2line_blocks", "pvb_build_line2line_blocks",
    "pvb_build_camera_lidar_blocks", "pvb_transform_cloud",
    "pvb_pair_knn5", "pvb_nearest_line", "pvb_point2line_segment_knn_associate", "pvb_point2line_segment_knn_tail", "pvb_point2line_segment_associate",
    "pvb_line2line_knn_associate", "pvb_line2line_knn_tail", "pvb_line_tracks_build", "pvb_line_tracks_gate", "pvb_generate_line_tracks",
    "pvb_line_votes_batch", "pvb_frames_line2line_blocks", "pvb_frames_line2line_blocks_device",
    "pvb_blocks_set_linear_solver", "pvb_cholesky_solve", "pvb_unique_line_pairs",
    "pvb_slerp_pose", "pvb_undistort_end_poses", "pvb_undistort_clouds",
    "pvb_reproj_set", "pvb_reproj_evaluate", "pvb_reproj_residuals", "pvb_reproj_jacobians", "pvb_reproj_cost", "pvb_reproj_blocks", "pvb_reproj_kernel_time_ms",
    "pvb_reproj_solve_lm", "pvb_build_reproj_observations", "pvb_joint_solve_lm",
    "pvb_blocks_set_edge_list", "pvb_blocks_set_reduce_hook", "pvb_write_poses_text", "pvb_read_poses_text", "pvb_build_point2plane_blocks_edges", "pvb_filter_line_pairs", "pvb_frames_point2plane_blocks", "pvb_neighbor_each_frame", "pvb_lidar_mask_by_track", "pvb_build_calibration_blocks",
    "pvb_pixel_sub_lines", "pvb_pixel_knn3", "pvb_pixel_line_neighbors", "pvb_pixel_line_candidates", "pvb_pixel_fit_line", "pvb_pixel_fit_lines",
]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Context:
    """One GPU + one stream (pvb_create / pvb_destroy)."""

    def __init__(self, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        rc = self._L.pvb_create(C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise PvbError(f"pvb_create(device={device}) failed with code {rc}: a CUDA device is required (no CPU fallback)")
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.pvb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PvbError(f"code {rc}: {self._L.pvb_last_error(self._h).decode()}")

    # ---- lifecycle
    def set_stream(self, cuda_stream):
        self._ck(self._L.pvb_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self._L.pvb_synchronize(self._h))

    @property
    def kernel_launches(self):
        return self._L.pvb_kernel_launches(self._h)

    @property
    def stream(self):
        return self._L.pvb_stream(self._h)

    # ---- A. correspondence-list mode
    def blocks_set(self, type, ref, nei, consts, huber, normalize, n_pose_blocks):
        t = _arr(type, np.int32)
        n = len(t)
        r = _arr(np.broadcast_to(ref, n), np.int32)
        m = _arr(np.broadcast_to(nei, n), np.int32)
        nz = _arr(np.broadcast_to(normalize, n), np.int32)
        h = _arr(np.broadcast_to(huber, n), np.float64)
        c = _arr(consts, np.float64).reshape(n, 12)
        self._n_blocks, self._nb = n, int(n_pose_blocks)
        self._ck(self._L.pvb_blocks_set(self._h, C.c_long(n), _p(t), _p(r), _p(m), _p(nz), _p(h), _p(c), C.c_int(n_pose_blocks)))

    def blocks_set_edge_list(self, ref=None, nei=None):
        """Global edge list of a sharded pose graph (sorted by (ref, nei), unique); None / empty: back to single-GPU behaviour."""
        if ref is None or len(ref) == 0:
            self._ck(self._L.pvb_blocks_set_edge_list(self._h, C.c_int(0), None, None))
            return
        ref, nei = _arr(ref, np.int32), _arr(nei, np.int32)
        self._ck(self._L.pvb_blocks_set_edge_list(self._h, C.c_int(len(ref)), _p(ref), _p(nei)))

    def blocks_evaluate(self, poses, want_rows=True, want_system=True):
        poses = _arr(poses, np.float64)
        self._ck(self._L.pvb_blocks_evaluate(self._h, _p(poses), C.c_int(int(want_rows)), C.c_int(int(want_system))))

    def blocks_rows(self):
        n = self._n_blocks
        rp, jp = self._L.pvb_blocks_residuals(self._h), self._L.pvb_blocks_jacobians(self._h)
        if not rp or not jp:
            raise PvbError("no rows: call blocks_evaluate(want_rows=True) first")
        return np.ctypeslib.as_array(rp, shape=(n,)).copy(), np.ctypeslib.as_array(jp, shape=(n, 12)).copy()

    def blocks_cost(self):
        c, n = C.c_double(), C.c_long()
        self._ck(self._L.pvb_blocks_cost(self._h, C.byref(c), C.byref(n)))
        return c.value, n.value

    def blocks_kernel_time_ms(self):
        ms = C.c_float()
        self._ck(self._L.pvb_blocks_kernel_time_ms(self._h, C.byref(ms)))
        return ms.value

    def blocks_edges(self):
        ne = self._L.pvb_blocks_num_edges(self._h)
        r, m = np.zeros(ne, np.int32), np.zeros(ne, np.int32)
        self._ck(self._L.pvb_blocks_edges(self._h, _p(r), _p(m)))
        s = np.zeros((ne, 92))
        self._ck(self._L.pvb_blocks_edge_systems(self._h, _p(s)))
        return r, m, s

    def blocks_dense_system(self):
        D = 6 * self._nb
        H, g, c = np.zeros((D, D)), np.zeros(D), C.c_double()
        self._ck(self._L.pvb_blocks_dense_system(self._h, _p(H), _p(g), C.byref(c)))
        return H, g, c.value

    def blocks_solve_lm(self, poses, is_const=None, max_iterations=20):
        poses = _arr(poses, np.float64).copy()
        mask = np.zeros(self._nb, np.uint8) if is_const is None else _arr(is_const, np.uint8)
        s = np.zeros(6)
        self._ck(self._L.pvb_blocks_solve_lm(self._h, _p(poses), _p(mask), C.c_int(max_iterations), _p(s)))
        keys = ["initial_cost", "final_cost", "iterations", "successful", "unsuccessful", "termination"]
        return poses, dict(zip(keys, s.tolist()))

    def blocks_pcg_stats(self):
        """(number of PCG solves, total CG iterations) since the context was created"""
        a, b = C.c_long(0), C.c_long(0)
        self._ck(self._L.pvb_blocks_pcg_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def blocks_set_linear_solver(self, kind):
        self._ck(self._L.pvb_blocks_set_linear_solver(self._h, C.c_int(kind)))

    def cholesky_solve(self, A, b):
        A, b = _arr(A, np.float64), _arr(b, np.float64)
        x, ms = np.zeros(len(b)), C.c_float()
        self._ck(self._L.pvb_cholesky_solve(self._h, _p(A), C.c_int(len(b)), _p(b), _p(x), C.byref(ms)))
        return x, ms.value

    # ---- B. frames
    def frames_set(self, targets, queries):
        """targets[f] / queries[f]: (n,4) float32 arrays in the sensor frame (surfLessFlat / surfFlat)."""
        self._keep = [[_arr(t, np.float32).reshape(-1, 4) for t in targets], [_arr(q, np.float32).reshape(-1, 4) for q in queries]]
        n = len(targets)
        arr = (_Frame * n)()
        for f in range(n):
            t, q = self._keep[0][f], self._keep[1][f]
            arr[f] = _Frame(t.ctypes.data if len(t) else None, len(t), q.ctypes.data if len(q) else None, len(q))
        self._n_frames = n
        self._ck(self._L.pvb_frames_set(self._h, C.c_int(n), arr))

    def frames_associate_point2plane(self, poses, ref, nei, plane_tolerance, dist_threshold, k=10, cell_size=0.0):
        poses = _arr(poses, np.float64)
        ref, nei = _arr(ref, np.int32), _arr(nei, np.int32)
        prm = _AssocParams(plane_tolerance, dist_threshold, k, cell_size)
        n = C.c_long()
        self._ck(self._L.pvb_frames_associate_point2plane(self._h, _p(poses), C.c_int(len(ref)), _p(ref), _p(nei), C.byref(prm), C.byref(n)))
        m = n.value
        e, q, pt, pl = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros((m, 3)), np.zeros((m, 4))
        self._ck(self._L.pvb_frames_get_point2plane(self._h, C.c_long(m), _p(e), _p(q), _p(pt), _p(pl)))
        return e, q, pt, pl

    def frames_point2plane_blocks(self, poses, ref, nei, plane_tolerance, dist_threshold, angle_residual, normalize_distance, weight, n_pose_blocks, block_offset=0,
                                  extra=None, k=10, cell_size=0.0):
        """Association + residual blocks of all edges on the device (AddLidarPointToPlaneResidual); `extra`: a BlockList.view() of host-built blocks."""
        prm = _AssocParams(plane_tolerance, dist_threshold, k, cell_size)
        ref, nei = _arr(ref, np.int32), _arr(nei, np.int32)
        n = C.c_long()
        if extra is not None and len(extra["type"]):
            xt, xr, xn, xz = (_arr(extra[key], np.int32) for key in ("type", "ref", "nei", "normalize"))
            xh, xc = _arr(extra["huber"], np.float64), _arr(extra["consts"], np.float64).reshape(-1, 12)
            args = (C.c_long(len(xt)), _p(xt), _p(xr), _p(xn), _p(xz), _p(xh), _p(xc))
        else:
            args = (C.c_long(0), None, None, None, None, None, None)
        self._ck(self._L.pvb_frames_point2plane_blocks(self._h, _p(_arr(poses, np.float64)), C.c_int(len(ref)), _p(ref), _p(nei), C.byref(prm), C.c_int(int(angle_residual)),
                                                       C.c_int(int(normalize_distance)), C.c_double(weight), C.c_int(block_offset), C.c_int(n_pose_blocks), *args, C.byref(n)))
        self._n_blocks, self._nb = n.value, int(n_pose_blocks)
        return n.value

    def frames_knn(self, poses, ref, nei, n_query, dist_threshold, k=10, cell_size=0.0):
        poses = _arr(poses, np.float64)
        prm = _AssocParams(0.05, dist_threshold, k, cell_size)
        idx, d2 = np.zeros((n_query, k), np.int32), np.zeros((n_query, k), np.float32)
        self._ck(self._L.pvb_frames_knn(self._h, _p(poses), C.c_int(ref), C.c_int(nei), C.byref(prm), _p(idx), _p(d2)))
        return idx, d2

    def frames_set_corners(self, corners):
        self._keep_c = [_arr(c, np.float32).reshape(-1, 4) for c in corners]
        n = len(corners)
        ptrs = (C.c_void_p * n)(*[c.ctypes.data if len(c) else None for c in self._keep_c])
        cnt = np.array([len(c) for c in self._keep_c], np.int32)
        self._ck(self._L.pvb_frames_set_corners(self._h, C.c_int(n), ptrs, _p(cnt)))

    def frames_associate_point2line(self, poses, ref, nei, dist_threshold, cell_size=0.0):
        poses = _arr(poses, np.float64)
        ref, nei = _arr(ref, np.int32), _arr(nei, np.int32)
        n = C.c_long()
        self._ck(self._L.pvb_frames_associate_point2line(self._h, _p(poses), C.c_int(len(ref)), _p(ref), _p(nei), C.c_float(dist_threshold), C.c_double(cell_size), C.byref(n)))
        m = n.value
        e, q, pt, a, b = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros((m, 3)), np.zeros((m, 3)), np.zeros((m, 3))
        self._ck(self._L.pvb_frames_get_point2line(self._h, C.c_long(m), _p(e), _p(q), _p(pt), _p(a), _p(b)))
        return e, q, pt, a, b

    # ---- C. dense
    @staticmethod
    def dense_params(plane_tolerance=0.05, dist_threshold=1.0, k=10, residual_type=P2PLANE_METER, normalize=1, huber=0.2, weight=1.0):
        return _DenseParams(plane_tolerance, dist_threshold, k, residual_type, normalize, huber, weight)

    def dense_set_target(self, xyzc, cell_size=0.0):
        a = _arr(xyzc, np.float32).reshape(-1, 4)
        self._ck(self._L.pvb_dense_set_target(self._h, _p(a), C.c_long(len(a)), C.c_double(cell_size)))

    def dense_set_sources(self, xyzc, offsets):
        a = xyzc if (isinstance(xyzc, np.ndarray) and xyzc.dtype == np.float32 and xyzc.flags.c_contiguous) else _arr(xyzc, np.float32)
        off = _arr(offsets, np.int32)
        self._dense_frames, self._dense_n = len(off) - 1, int(off[-1])
        self._ck(self._L.pvb_dense_set_sources(self._h, _p(a), _p(off), C.c_int(len(off) - 1)))

    def dense_set_sources_ptr(self, ptr, offsets):
        """Same, from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        off = _arr(offsets, np.int32)
        self._dense_frames, self._dense_n = len(off) - 1, int(off[-1])
        self._ck(self._L.pvb_dense_set_sources(self._h, C.c_void_p(ptr), _p(off), C.c_int(len(off) - 1)))

    def dense_evaluate(self, poses_lw, prm):
        poses = _arr(poses_lw, np.float64)
        out = np.zeros((self._dense_frames, 29))
        self._ck(self._L.pvb_dense_evaluate(self._h, _p(poses), C.byref(prm), _p(out)))
        return out

    def dense_evaluate_device(self, poses_lw, prm, out_ptr=None):
        """Enqueues the evaluation; returns the device pointer of the (n_frames x 29) reduced systems
        (out_ptr: caller-owned device buffer to write them to, e.g. a slice of an allreduce buffer)."""
        poses = _arr(poses_lw, np.float64)
        ptr = C.c_void_p(out_ptr)
        self._ck(self._L.pvb_dense_evaluate_device(self._h, _p(poses), C.byref(prm), C.byref(ptr)))
        return ptr.value

    def dense_set_hints(self, enable=True):
        self._ck(self._L.pvb_dense_set_hints(self._h, C.c_int(1 if enable else 0)))

    def dense_order_stats(self):
        """(sorts of the queries by target cell, fresh uploads that re-used the previous permutation)"""
        a, b = C.c_long(0), C.c_long(0)
        self._ck(self._L.pvb_dense_order_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def dense_reset_hints(self):
        self._ck(self._L.pvb_dense_reset_hints(self._h))

    def debug_counters(self):
        out = (C.c_ulonglong * 2)()
        self._ck(self._L.pvb_debug_counters(self._h, out))
        return int(out[0]), int(out[1])

    def dense_kernel_time_ms(self):
        ms = C.c_float()
        self._ck(self._L.pvb_dense_kernel_time_ms(self._h, C.byref(ms)))
        return ms.value

    def dense_gauss_newton_step(self, sys29, poses_lw, lam=0.0):
        poses = _arr(poses_lw, np.float64).copy()
        s = _arr(sys29, np.float64)
        rc = self._L.pvb_dense_gauss_newton_step(_p(s), C.c_int(len(s)), C.c_double(lam), _p(poses))
        if rc != 0:
            raise PvbError(f"pvb_dense_gauss_newton_step: code {rc}")
        return poses

    def dense_get_rows(self, poses_lw, prm):
        poses = _arr(poses_lw, np.float64)
        n = self._dense_n
        v, pt, pl, r, j = np.zeros(n, np.uint8), np.zeros((n, 3)), np.zeros((n, 4)), np.zeros(n), np.zeros((n, 6))
        self._ck(self._L.pvb_dense_get_rows(self._h, _p(poses), C.byref(prm), _p(v), _p(pt), _p(pl), _p(r), _p(j)))
        return v.astype(bool), pt, pl, r, j

    # ---- G. host builders
    @staticmethod
    def find_neighbors(t_wl, pose_valid=None, frame_valid=None, neighbor_size=6):
        t = _arr(t_wl, np.float64).reshape(-1, 3)
        n = len(t)
        pv = None if pose_valid is None else _arr(pose_valid, np.uint8)
        fv = None if frame_valid is None else _arr(frame_valid, np.uint8)
        off, out = np.zeros(n + 1, np.int32), np.zeros(n * (neighbor_size + 64), np.int32)
        m = load_library().pvb_find_neighbors(C.c_int(n), _p(t), _p(pv), _p(fv), C.c_int(neighbor_size), _p(off), _p(out), C.c_int(len(out)))
        if m < 0:
            raise PvbError(f"pvb_find_neighbors: code {m}")
        return [out[off[i]:off[i + 1]].tolist() for i in range(n)]

    @staticmethod
    def neighbor_each_frame(n_frames, n_lidars, neighbor_size, temporal, t_wc=None, frame_pose_valid=None, t_wl=None, lidar_pose_valid=None, lidar_valid=None):
        tc = None if t_wc is None else _arr(t_wc, np.float64).reshape(-1, 3)
        tl = None if t_wl is None else _arr(t_wl, np.float64).reshape(-1, 3)
        fpv = None if frame_pose_valid is None else _arr(frame_pose_valid, np.uint8)
        lpv = None if lidar_pose_valid is None else _arr(lidar_pose_valid, np.uint8)
        lv = None if lidar_valid is None else _arr(lidar_valid, np.uint8)
        off, out = np.zeros(n_frames + 1, np.int32), np.zeros(max(1, n_frames * (neighbor_size + 2)), np.int32)
        m = load_library().pvb_neighbor_each_frame(C.c_int(n_frames), C.c_int(n_lidars), C.c_int(neighbor_size), C.c_int(int(temporal)), _p(tc), _p(fpv), _p(tl), _p(lpv),
                                                   _p(lv), _p(off), _p(out), C.c_int(len(out)))
        if m < 0:
            raise PvbError(f"pvb_neighbor_each_frame: code {m}")
        return [out[off[i]:off[i + 1]].tolist() for i in range(n_frames)]

    @staticmethod
    def lidar_mask_by_track(tracks, n_segments_per_lidar):
        """tracks: list of (m, 2) arrays of (frame, line) features as returned by generate_line_tracks; returns one bool mask per LiDAR frame."""
        track_off = np.concatenate([[0], np.cumsum([len(t) for t in tracks])]).astype(np.int32)
        feats = np.concatenate([np.asarray(t, np.int32).reshape(-1, 2) for t in tracks]) if len(tracks) else np.zeros((0, 2), np.int32)
        ff, fl = _arr(feats[:, 0], np.int32), _arr(feats[:, 1], np.int32)
        seg_off = np.concatenate([[0], np.cumsum(n_segments_per_lidar)]).astype(np.int32)
        mask = np.zeros(max(1, int(seg_off[-1])), np.uint8)
        rc = load_library().pvb_lidar_mask_by_track(C.c_int(len(track_off) - 1), _p(track_off), _p(ff), _p(fl), C.c_int(len(seg_off) - 1), _p(seg_off), _p(mask))
        if rc < 0:
            raise PvbError(f"pvb_lidar_mask_by_track: code {rc}")
        return [mask[seg_off[i]:seg_off[i + 1]].astype(bool) for i in range(len(seg_off) - 1)]

    def transform_cloud(self, xyzi, R, t):
        a = _arr(xyzi, np.float32).reshape(-1, 4)
        out = np.empty_like(a)
        self._ck(self._L.pvb_transform_cloud(self._h, _p(a), C.c_long(len(a)), _p(_arr(R, np.float64)), _p(_arr(t, np.float64)), _p(out)))
        return out

    # ---- camera-camera reprojection residuals / bundle adjustment
    def reproj_set(self, cam, point, bearing, n_cams, n_points, weight=1.0, huber=4.0 * np.pi / 180.0):
        cam, point, bearing = _arr(cam, np.int32), _arr(point, np.int32), _arr(bearing, np.float64).reshape(-1, 3)
        self._reproj_n, self._reproj_nc, self._reproj_np = len(cam), int(n_cams), int(n_points)
        self._ck(self._L.pvb_reproj_set(self._h, C.c_long(len(cam)), _p(cam), _p(point), _p(bearing), C.c_double(weight), C.c_double(huber), C.c_int(n_cams), C.c_long(n_points)))

    def reproj_evaluate(self, cams, points, want_rows=True, want_system=True):
        self._ck(self._L.pvb_reproj_evaluate(self._h, _p(_arr(cams, np.float64)), _p(_arr(points, np.float64)), C.c_int(int(want_rows)), C.c_int(int(want_system))))

    def reproj_rows(self):
        n = self._reproj_n
        r = np.ctypeslib.as_array(self._L.pvb_reproj_residuals(self._h), shape=(n,)).copy() if n else np.zeros(0)
        J = np.ctypeslib.as_array(self._L.pvb_reproj_jacobians(self._h), shape=(n, 9)).copy() if n else np.zeros((0, 9))
        return r, J

    def reproj_cost(self):
        c = C.c_double()
        self._ck(self._L.pvb_reproj_cost(self._h, C.byref(c)))
        return c.value

    def reproj_blocks(self):
        nc, npt, n = self._reproj_nc, self._reproj_np, self._reproj_n
        B, gc, Cp, gp, E = np.zeros((nc, 21)), np.zeros((nc, 6)), np.zeros((npt, 6)), np.zeros((npt, 3)), np.zeros((n, 18))
        self._ck(self._L.pvb_reproj_blocks(self._h, _p(B), _p(gc), _p(Cp), _p(gp), _p(E)))
        return B, gc, Cp, gp, E.reshape(n, 6, 3)

    def reproj_kernel_time_ms(self):
        ms = C.c_float()
        self._ck(self._L.pvb_reproj_kernel_time_ms(self._h, C.byref(ms)))
        return ms.value

    def reproj_solve_lm(self, cams, points, cam_param_const=None, point_const=None, max_iterations=50):
        cams, points = _arr(cams, np.float64).copy().reshape(-1, 6), _arr(points, np.float64).copy().reshape(-1, 3)
        cc = None if cam_param_const is None else _arr(cam_param_const, np.uint8)
        pc = None if point_const is None else _arr(point_const, np.uint8)
        summ = np.zeros(6)
        self._ck(self._L.pvb_reproj_solve_lm(self._h, _p(cams), _p(points), _p(cc), _p(pc), C.c_int(max_iterations), _p(summ)))
        keys = ["initial_cost", "final_cost", "iterations", "successful", "unsuccessful", "termination"]
        return cams, points, dict(zip(keys, summ.tolist()))

    # ---- pixel-space camera-LiDAR association (first stage)
    @staticmethod
    def pixel_sub_lines(rows, cols, lines):
        lines = _arr(lines, np.float32).reshape(-1, 4)
        cap = 64 + int(sum(np.hypot(float(l[0]) - float(l[2]), float(l[1]) - float(l[3])) / 70 + 4 for l in lines))
        mid, s2l = np.zeros((cap, 2), np.float32), np.zeros(cap, np.int32)
        m = load_library().pvb_pixel_sub_lines(C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), C.c_int(cap), _p(mid), _p(s2l))
        if m < 0:
            raise PvbError(f"pvb_pixel_sub_lines: code {m}")
        return mid[:m].copy(), s2l[:m].copy()

    def pixel_line_neighbors(self, rows, cols, lines, cloud_local, T_cl):
        lines, cloud = _arr(lines, np.float32).reshape(-1, 4), _arr(cloud_local, np.float32).reshape(-1, 4)
        n = len(cloud)
        line3, d2, px = np.full((n, 3), -1, np.int32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
        self._ck(self._L.pvb_pixel_line_neighbors(self._h, C.c_int(rows), C.c_int(cols), _p(lines), C.c_int(len(lines)), _p(cloud), C.c_int(n), _p(_arr(T_cl, np.float64)),
                                                  _p(line3), _p(d2), _p(px)))
        return line3, d2, px

    @staticmethod
    def filter_line_pairs(rows, cols, image_lines, start, end, by_angle, by_length):
        ln, s_, e_ = _arr(image_lines, np.float32).reshape(-1, 4), _arr(start, np.float64).reshape(-1, 3), _arr(end, np.float64).reshape(-1, 3)
        keep, ang = np.zeros(len(ln), np.uint8), np.full(len(ln), np.float32(3.4028235e38), np.float32)
        rc = load_library().pvb_filter_line_pairs(C.c_int(rows), C.c_int(cols), C.c_int(len(ln)), _p(ln), _p(s_), _p(e_), C.c_int(int(by_angle)), C.c_int(int(by_length)),
                                                  _p(keep), _p(ang))
        if rc < 0:
            raise PvbError(f"pvb_filter_line_pairs: code {rc}")
        return keep.astype(bool), ang

    @staticmethod
    def pixel_line_candidates(n_lines, line3, min_points=6):
        line3 = _arr(line3, np.int32).reshape(-1, 3)
        off, idx = np.zeros(n_lines + 1, np.int32), np.zeros(max(1, line3.size), np.int32)
        m = load_library().pvb_pixel_line_candidates(C.c_int(n_lines), C.c_int(len(line3)), _p(line3), C.c_int(min_points), C.c_int(len(idx)), _p(off), _p(idx))
        if m < 0:
            raise PvbError(f"pvb_pixel_line_candidates: code {m}")
        return off, idx[:m].copy()

    @staticmethod
    def pixel_fit_line(xyz, dist_threshold=0.1, max_iterations=50, probability=0.99):
        """FitLineRANSAC + end points for one candidate list (CameraLidarLineAssociate.cpp:105-144, 717-752; the RANSAC restates PCL 1.10, parity unpinned).
        Returns None when fewer than 3 inliers, else (coeff6 float32, inlier indices, start, end)."""
        a = _arr(xyz, np.float32)
        a = a.reshape(-1, a.shape[-1] if a.ndim == 2 else 3)
        coeff, inl, s_, e_ = np.zeros(6, np.float32), np.zeros(max(1, len(a)), np.int32), np.zeros(3), np.zeros(3)
        m = load_library().pvb_pixel_fit_line(_p(a), C.c_int(len(a)), C.c_int(a.shape[1]), C.c_double(dist_threshold), C.c_int(max_iterations), C.c_double(probability),
                                              _p(coeff), C.c_int(len(inl)), _p(inl), _p(s_), _p(e_))
        if m < 0:
            raise PvbError(f"pvb_pixel_fit_line: code {m}")
        return None if m < 3 else (coeff, inl[:m].copy(), s_, e_)

    @staticmethod
    def pixel_fit_lines(cloud_cam, line_off, lidar_idx, dist_threshold=0.1, max_iterations=50, probability=0.99):
        """pixel_fit_line for every candidate list of one image (CSR from pixel_line_candidates) on all host threads: (inlier counts, start (L,3), end (L,3))."""
        cam, off, idx = _arr(cloud_cam, np.float32).reshape(-1, 4), _arr(line_off, np.int32), _arr(lidar_idx, np.int32)
        L = len(off) - 1
        n_in, coeff, s_, e_ = np.zeros(L, np.int32), np.zeros((L, 6), np.float32), np.zeros((L, 3)), np.zeros((L, 3))
        rc = load_library().pvb_pixel_fit_lines(_p(cam), C.c_int(len(cam)), C.c_int(L), _p(off), _p(idx), C.c_double(dist_threshold), C.c_int(max_iterations), C.c_double(probability),
                                                _p(n_in), _p(coeff), _p(s_), _p(e_))
        if rc < 0:
            raise PvbError(f"pvb_pixel_fit_lines: code {rc}")
        return n_in, s_, e_

    def pixel_associate(self, rows, cols, lines, cloud_local, T_cl):
        """CameraLidarLineAssociate::Associate(lines, cloud, T_cl) (CameraLidarLineAssociate.cpp:22-188), the fallback for frames without LiDAR segments: projected LiDAR
        points -> 3 nearest image sub-lines (device) -> per image line the candidate points (>= 6) -> line fit (pixel_fit_line; its RANSAC is parity-unpinned) ->
        Filter(true, true) -> end points back in the LiDAR frame.  Returns (image line ids, start (n,3), end (n,3), angle float32) in ascending image line order."""
        lines = _arr(lines, np.float32).reshape(-1, 4)
        T = _arr(T_cl, np.float64).reshape(4, 4)
        cloud = _arr(cloud_local, np.float32).reshape(-1, 4)
        line3, _, _ = self.pixel_line_neighbors(rows, cols, lines, cloud, T)
        off, idx = Context.pixel_line_candidates(len(lines), line3, 6)
        if off[-1] == 0:                                                              # no image line collected 6 candidates
            return np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.float32)
        cam = self.transform_cloud(cloud, T[:3, :3], T[:3, 3])                       # pcl::transformPointCloud(point_cloud, cloud, T_cl) (:26)
        n_in, s_, e_ = Context.pixel_fit_lines(cam, off, idx)
        ids = np.flatnonzero(n_in >= 3).astype(np.int32)
        s_, e_ = s_[ids], e_[ids]
        keep, ang = Context.filter_line_pairs(rows, cols, lines[ids], s_, e_, True, True)
        T_lc = np.linalg.inv(T)                                                       # :182-187

        def back(p):                                                                  # (T_lc * p.homogeneous()).hnormalized()
            h = np.c_[p, np.ones(len(p))] @ T_lc.T
            return h[:, :3] / h[:, 3:]
        return ids[keep], back(s_[keep]), back(e_[keep]), ang[keep]

    @staticmethod
    def segmented_fit_lists(n_lines, cand_off, cand_idx, seg_of_point, seg_off):
        """The per-line tests of the segmented Associate() between its k-NN stage and its line fit (CameraLidarLineAssociate.cpp:265-292): an image line with at least 6
        candidate points goes on when at least 70 % of them belong to ONE LiDAR segment (the first segment with the largest count); the fit then takes ALL points of that
        segment.  Returns (image line ids, majority segment per line, CSR offsets, point indices) of the lists to fit."""
        n_seg = len(seg_off) - 1
        ids, segs, off, idx = [], [], [0], []
        for li in range(n_lines):
            c = cand_idx[cand_off[li]:cand_off[li + 1]]
            if len(c) == 0:
                continue
            cnt = np.bincount(seg_of_point[c], minlength=n_seg)
            mp = int(np.argmax(cnt))                                                  # max_element: the first maximum
            if cnt[mp] < 0.7 * len(c):
                continue
            ids.append(li); segs.append(mp)
            idx.append(np.arange(seg_off[mp], seg_off[mp + 1], dtype=np.int32)); off.append(off[-1] + int(seg_off[mp + 1] - seg_off[mp]))
        return (np.array(ids, np.int32), np.array(segs, np.int32), np.array(off, np.int32),
                np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32))

    def pixel_associate_segmented(self, rows, cols, lines, segments, T_cl):
        """CameraLidarLineAssociate::Associate(lines, segmented_cloud, T_cl) (CameraLidarLineAssociate.cpp:191-338; the reference's call sites of this overload are commented
        out, it is built for completeness): every point carries its segment index, projected points -> 3 nearest image sub-lines (device) -> per image line the candidate
        points (>= 6) -> 70 % single-segment test -> line fit of the WHOLE majority segment -> Filter(true, true) -> T_lc.  As written in the reference the fit runs on the
        segment in the LiDAR frame (`segmented_cloud[max_position]`, not the transformed copy), so Filter sees LiDAR-frame end points as if they were camera-frame ones and
        the final T_lc moves them once more; this is reproduced, not corrected.  Returns (image line ids, start (n,3), end (n,3), angle float32)."""
        lines = _arr(lines, np.float32).reshape(-1, 4)
        T = _arr(T_cl, np.float64).reshape(4, 4)
        segs = [_arr(x, np.float32).reshape(-1, 4) for x in segments]
        empty = (np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0, np.float32))
        if not segs or sum(len(x) for x in segs) == 0:
            return empty
        seg_off = np.concatenate([[0], np.cumsum([len(x) for x in segs])]).astype(np.int32)
        base = np.ascontiguousarray(np.concatenate(segs))
        seg_of_point = np.repeat(np.arange(len(segs), dtype=np.int32), [len(x) for x in segs])
        cloud = base.copy(); cloud[:, 3] = seg_of_point                                # :196-204 p.intensity = i
        line3, _, _ = self.pixel_line_neighbors(rows, cols, lines, cloud, T)
        off, idx = Context.pixel_line_candidates(len(lines), line3, 6)
        ids, _, f_off, f_idx = Context.segmented_fit_lists(len(lines), off, idx, seg_of_point, seg_off)
        if len(ids) == 0:
            return empty
        n_in, s_, e_ = Context.pixel_fit_lines(base, f_off, f_idx)
        ok = n_in >= 3
        ids, s_, e_ = ids[ok], s_[ok], e_[ok]
        keep, ang = Context.filter_line_pairs(rows, cols, lines[ids], s_, e_, True, True)
        T_lc = np.linalg.inv(T)                                                       # :331-336

        def back(p):
            h = np.c_[p, np.ones(len(p))] @ T_lc.T
            return h[:, :3] / h[:, 3:]
        return ids[keep], back(s_[keep]), back(e_[keep]), ang[keep]

    def joint_solve_lm(self, poses, points, pose_param_const=None, point_const=None, max_iterations=20):
        poses, points = _arr(poses, np.float64).copy().reshape(-1, 6), _arr(points, np.float64).copy().reshape(-1, 3)
        cc = None if pose_param_const is None else _arr(pose_param_const, np.uint8)
        pc = None if point_const is None else _arr(point_const, np.uint8)
        summ = np.zeros(6)
        self._ck(self._L.pvb_joint_solve_lm(self._h, _p(poses), _p(points), _p(cc), _p(pc), C.c_int(max_iterations), _p(summ)))
        keys = ["initial_cost", "final_cost", "iterations", "successful", "unsuccessful", "termination"]
        return poses, points, dict(zip(keys, summ.tolist()))

    @staticmethod
    def build_reproj_observations(rows, cols, track_off, feat_frame, feat_xy, pose_valid=None):
        track_off, feat_frame, feat_xy = _arr(track_off, np.int32), _arr(feat_frame, np.int32), _arr(feat_xy, np.float32).reshape(-1, 2)
        cap = len(feat_frame)
        cam, point, bearing = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros((cap, 3))
        pv = None if pose_valid is None else _arr(pose_valid, np.uint8)
        m = load_library().pvb_build_reproj_observations(C.c_int(rows), C.c_int(cols), C.c_long(len(track_off) - 1), _p(track_off), _p(feat_frame), _p(feat_xy), _p(pv),
                                                         C.c_long(cap), _p(cam), _p(point), _p(bearing))
        if m < 0:
            raise PvbError(f"pvb_build_reproj_observations: code {m}")
        return cam[:m].copy(), point[:m].copy(), bearing[:m].copy()

    @staticmethod
    def write_poses_text(path, R, t, names=None):
        R, t = _arr(R, np.float64).reshape(-1, 9), _arr(t, np.float64).reshape(-1, 3)
        arr = None
        if names is not None:
            arr = (C.c_char_p * len(R))(*[nm.encode() for nm in names])
        rc = load_library().pvb_write_poses_text(str(path).encode(), C.c_int(len(R)), _p(R), _p(t), arr)
        if rc < 0:
            raise PvbError(f"pvb_write_poses_text: code {rc}")

    @staticmethod
    def read_poses_text(path, with_invalid=True, cap=100000, name_len=256):
        R, t, valid = np.zeros((cap, 9)), np.zeros((cap, 3)), np.zeros(cap, np.uint8)
        names = C.create_string_buffer(cap * name_len)
        n = load_library().pvb_read_poses_text(str(path).encode(), C.c_int(int(with_invalid)), C.c_int(cap), _p(R), _p(t), _p(valid), names, C.c_int(name_len))
        if n < 0:
            raise PvbError(f"pvb_read_poses_text: code {n}")
        nm = [names.raw[i * name_len:(i + 1) * name_len].split(b"\0", 1)[0].decode() for i in range(n)]
        return R[:n].reshape(n, 3, 3).copy(), t[:n].copy(), valid[:n].astype(bool), nm

    @staticmethod
    def slerp_pose(pose_w1, pose_w2, ratio):
        out = np.empty((4, 4))
        rc = load_library().pvb_slerp_pose(_p(_arr(pose_w1, np.float64)), _p(_arr(pose_w2, np.float64)), C.c_double(ratio), _p(out))
        if rc < 0:
            raise PvbError(f"pvb_slerp_pose: code {rc}")
        return out

    @staticmethod
    def undistort_end_poses(poses, pose_valid, frame_valid, gap_time):
        P = _arr(poses, np.float64).reshape(-1, 16)
        n = len(P)
        out, has = np.zeros((n, 16)), np.zeros(n, np.uint8)
        rc = load_library().pvb_undistort_end_poses(C.c_int(n), _p(P), _p(_arr(pose_valid, np.uint8)), _p(_arr(frame_valid, np.uint8)), C.c_float(gap_time), _p(out), _p(has))
        if rc < 0:
            raise PvbError(f"pvb_undistort_end_poses: code {rc}")
        return out.reshape(n, 4, 4), has.astype(bool)

    def undistort_clouds(self, xyzi, offsets, T_wl, T_we, has_end=None):
        a = _arr(xyzi, np.float32).reshape(-1, 4)
        off = _arr(offsets, np.int32)
        out = np.empty_like(a)
        he = None if has_end is None else _arr(has_end, np.uint8)
        self._ck(self._L.pvb_undistort_clouds(self._h, _p(a), _p(off), C.c_int(len(off) - 1), _p(_arr(T_wl, np.float64)), _p(_arr(T_we, np.float64)), _p(he), _p(out)))
        return out

    def line2line_associate(self, ref, nei, dist_threshold):
        S = max(1, nei.c.n_segments)
        nl, rl, a, b, n = np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros((S, 3)), np.zeros((S, 3)), C.c_int()
        self._ck(self._L.pvb_line2line_associate(self._h, C.byref(ref.c), C.byref(nei.c), C.c_double(dist_threshold), C.byref(n), _p(nl), _p(rl), _p(a), _p(b)))
        m = n.value
        return nl[:m].copy(), rl[:m].copy(), a[:m].copy(), b[:m].copy()

    def pair_knn5(self, ref_local, R_ref, t_ref, nei_local, R_nei, t_nei, dist_threshold, cell_size=0.0):
        r, q = _arr(ref_local, np.float32).reshape(-1, 4), _arr(nei_local, np.float32).reshape(-1, 4)
        idx = np.full((len(q), 5), -1, np.int32)
        self._ck(self._L.pvb_pair_knn5(self._h, _p(r), C.c_int(len(r)), _p(_arr(R_ref, np.float64)), _p(_arr(t_ref, np.float64)), _p(q), C.c_int(len(q)),
                                       _p(_arr(R_nei, np.float64)), _p(_arr(t_nei, np.float64)), C.c_float(dist_threshold), C.c_double(cell_size), _p(idx)))
        return idx

    def nearest_line(self, lines_world, points_world):
        lw, pw = _arr(lines_world, np.float64).reshape(-1, 6), _arr(points_world, np.float32).reshape(-1, 4)
        line, dist = np.full(len(pw), -1, np.int32), np.zeros(len(pw))
        self._ck(self._L.pvb_nearest_line(self._h, _p(lw), C.c_int(len(lw)), _p(pw), C.c_int(len(pw)), _p(line), _p(dist)))
        return line, dist

    def _segment_assoc(self, fn, ref, nei, dist_threshold):
        cap = max(1, nei.c.n_corner * 4)
        q, ln, pt, a, b, n = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3)), C.c_long()
        self._ck(fn(self._h, C.byref(ref.c), C.byref(nei.c), C.c_float(dist_threshold), C.c_long(cap), C.byref(n), _p(q), _p(ln), _p(pt), _p(a), _p(b)))
        m = n.value
        return q[:m].copy(), ln[:m].copy(), pt[:m].copy(), a[:m].copy(), b[:m].copy()

    def point2line_segment_knn_associate(self, ref, nei, dist_threshold):
        return self._segment_assoc(self._L.pvb_point2line_segment_knn_associate, ref, nei, dist_threshold)

    def point2line_segment_associate(self, ref, nei, dist_threshold):
        return self._segment_assoc(self._L.pvb_point2line_segment_associate, ref, nei, dist_threshold)

    def line2line_knn_associate(self, ref, nei, dist_threshold):
        S = max(1, nei.c.n_segments)
        nl, rl, a, b, n = np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros((S, 3)), np.zeros((S, 3)), C.c_int()
        self._ck(self._L.pvb_line2line_knn_associate(self._h, C.byref(ref.c), C.byref(nei.c), C.c_float(dist_threshold), C.byref(n), _p(nl), _p(rl), _p(a), _p(b)))
        m = n.value
        return nl[:m].copy(), rl[:m].copy(), a[:m].copy(), b[:m].copy()

    @staticmethod
    def line_tracks_build(pair_a, pair_b, match_off, match_a, match_b, min_track_length=3, allow_multiple_map=True):
        pa, pb, mo, ma, mb = (_arr(x, np.int32) for x in (pair_a, pair_b, match_off, match_a, match_b))
        cap = 2 * len(ma) + 1
        off, ff, fl, n = np.zeros(cap + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), C.c_int()
        rc = load_library().pvb_line_tracks_build(C.c_int(len(pa)), _p(pa), _p(pb), _p(mo), _p(ma), _p(mb), C.c_int(min_track_length), C.c_int(int(allow_multiple_map)),
                                                  C.c_int(cap), C.byref(n), _p(off), _p(ff), _p(fl))
        if rc:
            raise PvbError(f"pvb_line_tracks_build: code {rc}")
        return [np.stack([ff[off[t]:off[t + 1]], fl[off[t]:off[t + 1]]], axis=1) for t in range(n.value)]

    @staticmethod
    def line_tracks_gate(tracks, ref_frame, nei_frame, ref_line, nei_line):
        off = np.zeros(len(tracks) + 1, np.int32)
        off[1:] = np.cumsum([len(t) for t in tracks])
        feats = np.concatenate(tracks) if tracks else np.zeros((0, 2), np.int32)
        ff, fl = _arr(feats[:, 0], np.int32), _arr(feats[:, 1], np.int32)
        rl, nl = _arr(ref_line, np.int32), _arr(nei_line, np.int32)
        keep = np.zeros(len(rl), np.uint8)
        rc = load_library().pvb_line_tracks_gate(C.c_int(len(tracks)), _p(off), _p(ff), _p(fl), C.c_int(ref_frame), C.c_int(nei_frame), C.c_int(len(rl)), _p(rl), _p(nl), _p(keep))
        if rc:
            raise PvbError(f"pvb_line_tracks_gate: code {rc}")
        return keep.astype(bool)

    def generate_line_tracks(self, frames, neighbors, pose_valid=None, dist_threshold=0.3, min_track_length=3):
        """frames: list of LineFrame; neighbors: list of lists (pvb_find_neighbors)."""
        arr = (_LineFrame * len(frames))(*[f.c for f in frames])
        off = np.zeros(len(frames) + 1, np.int32)
        off[1:] = np.cumsum([len(nb) for nb in neighbors])
        ids = _arr(np.concatenate([np.asarray(nb, np.int32) for nb in neighbors]) if off[-1] else np.zeros(0, np.int32), np.int32)
        cap = 2 * sum(max(1, f.c.n_segments) for f in frames) * max(1, max((len(nb) for nb in neighbors), default=1)) + 1
        toff, ff, fl, n = np.zeros(cap + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), C.c_int()
        pv = None if pose_valid is None else _arr(pose_valid, np.uint8)
        self._ck(self._L.pvb_generate_line_tracks(self._h, C.c_int(len(frames)), arr, _p(pv) if pv is not None else None, _p(off), _p(ids), C.c_double(dist_threshold),
                                                  C.c_int(min_track_length), C.c_int(cap), C.byref(n), _p(toff), _p(ff), _p(fl)))
        return [np.stack([ff[toff[t]:toff[t + 1]], fl[toff[t]:toff[t + 1]]], axis=1) for t in range(n.value)]

    def frames_line2line_blocks(self, bl, frames, edges, dist_threshold, tracks, angle_residual, normalize_distance, weight):
        """AddLidarLineToLineResidual2 for all pose-graph edges in one call (batched votes on the device, tails + blocks in C++ on the host cores).
        frames: list of LineFrame; edges: list of (ref, nei); tracks: list of (n, 2) arrays (generate_line_tracks) or None = no gate."""
        arr = (_LineFrame * len(frames))(*[f.c for f in frames])
        ref, nei = _arr([e[0] for e in edges], np.int32), _arr([e[1] for e in edges], np.int32)
        if tracks is None:
            nt, toff, ff, fl = -1, None, None, None
        else:
            nt = len(tracks)
            toff = np.zeros(nt + 1, np.int32)
            toff[1:] = np.cumsum([len(t) for t in tracks])
            feats = np.concatenate(tracks) if tracks else np.zeros((0, 2), np.int32)
            ff, fl = _arr(feats[:, 0], np.int32), _arr(feats[:, 1], np.int32)
        n, cap, *arrs = bl.args()
        m = self._L.pvb_frames_line2line_blocks(self._h, C.c_int(len(frames)), arr, C.c_int(len(edges)), _p(ref), _p(nei), C.c_double(dist_threshold), C.c_int(nt),
                                                _p(toff) if toff is not None else None, _p(ff) if ff is not None else None, _p(fl) if fl is not None else None,
                                                C.c_int(int(angle_residual)), C.c_int(int(normalize_distance)), C.c_double(weight), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_frames_line2line_blocks: code {m}: {self._L.pvb_last_error(self._h).decode()}")
        bl.n = m

    def frames_line2line_blocks_device(self, frames, edges, dist_threshold, tracks, angle_residual, normalize_distance, weight):
        """AddLidarLineToLineResidual2 with the tails on the device too: the blocks wait in HBM for the next frames_point2plane_blocks call, which places
        them in the block arrays.  Returns the number of waiting line blocks."""
        arr = (_LineFrame * len(frames))(*[f.c for f in frames])
        ref, nei = _arr([e[0] for e in edges], np.int32), _arr([e[1] for e in edges], np.int32)
        if tracks is None:
            nt, toff, ff, fl = -1, None, None, None
        else:
            nt = len(tracks)
            toff = np.zeros(nt + 1, np.int32)
            toff[1:] = np.cumsum([len(t) for t in tracks])
            feats = np.concatenate(tracks) if tracks else np.zeros((0, 2), np.int32)
            ff, fl = _arr(feats[:, 0], np.int32), _arr(feats[:, 1], np.int32)
        n = C.c_long()
        self._ck(self._L.pvb_frames_line2line_blocks_device(self._h, C.c_int(len(frames)), arr, C.c_int(len(edges)), _p(ref), _p(nei), C.c_double(dist_threshold), C.c_int(nt),
                                                            _p(toff) if toff is not None else None, _p(ff) if ff is not None else None, _p(fl) if fl is not None else None,
                                                            C.c_int(int(angle_residual)), C.c_int(int(normalize_distance)), C.c_double(weight), C.byref(n)))
        return n.value

    @staticmethod
    def point2line_segment_knn_tail(ref, nei, idx5):
        idx = _arr(idx5, np.int32).reshape(-1, 5)
        cap = max(1, len(idx) * 4)
        q, ln, pt, a, b, n = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3)), C.c_long()
        rc = load_library().pvb_point2line_segment_knn_tail(C.byref(ref.c), C.byref(nei.c), _p(idx), C.c_long(cap), C.byref(n), _p(q), _p(ln), _p(pt), _p(a), _p(b))
        if rc:
            raise PvbError(f"pvb_point2line_segment_knn_tail: code {rc}")
        m = n.value
        return q[:m].copy(), ln[:m].copy(), pt[:m].copy(), a[:m].copy(), b[:m].copy()

    @staticmethod
    def line2line_knn_tail(ref, nei, idx5):
        idx = _arr(idx5, np.int32).reshape(-1, 5)
        S = max(1, nei.c.n_segments)
        nl, rl, a, b, n = np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros((S, 3)), np.zeros((S, 3)), C.c_int()
        rc = load_library().pvb_line2line_knn_tail(C.byref(ref.c), C.byref(nei.c), _p(idx), C.byref(n), _p(nl), _p(rl), _p(a), _p(b))
        if rc:
            raise PvbError(f"pvb_line2line_knn_tail: code {rc}")
        m = n.value
        return nl[:m].copy(), rl[:m].copy(), a[:m].copy(), b[:m].copy()

    @staticmethod
    def unique_line_pairs(image_line, lidar_line, score):
        il, ll, sc = _arr(image_line, np.int32), _arr(lidar_line, np.int32), _arr(score, np.float32)
        m = max(1, len(il))
        oi, ol, os_, n = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.float32), C.c_int()
        rc = load_library().pvb_unique_line_pairs(C.c_int(len(il)), _p(il), _p(ll), _p(sc), C.byref(n), _p(oi), _p(ol), _p(os_))
        if rc:
            raise PvbError(f"pvb_unique_line_pairs: code {rc}")
        return oi[:n.value].copy(), ol[:n.value].copy(), os_[:n.value].copy()

    def camera_lidar_associate(self, rows, cols, lines, lidar, T_cl, filter_by_length=True, multiple_association=True, image_mask=None, lidar_mask=None):
        ln = _arr(lines, np.float32).reshape(-1, 4)
        cap = max(1, len(ln) * max(1, lidar.c.n_segments))
        il, ll, s, e, ang, n = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(cap, np.float32), C.c_int()
        im = None if image_mask is None else _arr(image_mask, np.uint8)
        lm = None if lidar_mask is None else _arr(lidar_mask, np.uint8)
        self._ck(self._L.pvb_camera_lidar_associate(self._h, C.c_int(rows), C.c_int(cols), _p(ln), C.c_int(len(ln)), C.byref(lidar.c), _p(_arr(T_cl, np.float64)),
                                                     C.c_int(int(filter_by_length)), C.c_int(int(multiple_association)), _p(im), _p(lm),
                                                     C.c_int(cap), C.byref(n), _p(il), _p(ll), _p(s), _p(e), _p(ang)))
        m = n.value
        return il[:m].copy(), ll[:m].copy(), s[:m].copy(), e[:m].copy(), ang[:m].copy()

    @staticmethod
    def build_point2plane_blocks(bl, point, plane, ref_block, nei_block, angle_residual, normalize_distance, weight):
        pt, pl = _arr(point, np.float64).reshape(-1, 3), _arr(plane, np.float64).reshape(-1, 4)
        n, cap, *arrs = bl.args()
        m = load_library().pvb_build_point2plane_blocks(C.c_long(len(pt)), _p(pt), _p(pl), C.c_int(ref_block), C.c_int(nei_block), C.c_int(int(angle_residual)),
                                                        C.c_int(int(normalize_distance)), C.c_double(weight), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_build_point2plane_blocks: code {m}")
        bl.n = m

    @staticmethod
    def build_point2plane_blocks_edges(bl, edge, point, plane, edge_ref_block, edge_nei_block, angle_residual, normalize_distance, weight):
        e, pt, pl = _arr(edge, np.int32), _arr(point, np.float64).reshape(-1, 3), _arr(plane, np.float64).reshape(-1, 4)
        er, en = _arr(edge_ref_block, np.int32), _arr(edge_nei_block, np.int32)
        n, cap, *arrs = bl.args()
        m = load_library().pvb_build_point2plane_blocks_edges(C.c_long(len(e)), _p(e), _p(pt), _p(pl), C.c_int(len(er)), _p(er), _p(en), C.c_int(int(angle_residual)),
                                                              C.c_int(int(normalize_distance)), C.c_double(weight), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_build_point2plane_blocks_edges: code {m}")
        bl.n = m

    @staticmethod
    def build_point2line_blocks(bl, point, a, b, ref_block, nei_block, angle_residual, normalize_distance, weight):
        pt, pa, pb = (_arr(x, np.float64).reshape(-1, 3) for x in (point, a, b))
        n, cap, *arrs = bl.args()
        m = load_library().pvb_build_point2line_blocks(C.c_long(len(pt)), _p(pt), _p(pa), _p(pb), C.c_int(ref_block), C.c_int(nei_block), C.c_int(int(angle_residual)),
                                                       C.c_int(int(normalize_distance)), C.c_double(weight), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_build_point2line_blocks: code {m}")
        bl.n = m

    @staticmethod
    def build_line2line_blocks(bl, nei, nei_corner_world, nei_line, a, b, ref_block, nei_block, angle_residual, normalize_distance, weight):
        w = _arr(nei_corner_world, np.float32).reshape(-1, 4)
        n, cap, *arrs = bl.args()
        m = load_library().pvb_build_line2line_blocks(C.byref(nei.c), _p(w), C.c_int(int(nei_line)), _p(_arr(a, np.float64)), _p(_arr(b, np.float64)), C.c_int(ref_block),
                                                      C.c_int(nei_block), C.c_int(int(angle_residual)), C.c_int(int(normalize_distance)), C.c_double(weight), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_build_line2line_blocks: code {m}")
        bl.n = m

    @staticmethod
    def build_camera_lidar_blocks(bl, rows, cols, image_lines, start, end, pair_weight, cam_block, lidar_block, weight):
        ln, s, e = _arr(image_lines, np.float32).reshape(-1, 4), _arr(start, np.float64).reshape(-1, 3), _arr(end, np.float64).reshape(-1, 3)
        pw = None if pair_weight is None else _arr(pair_weight, np.float32)
        n, cap, *arrs = bl.args()
        m = load_library().pvb_build_camera_lidar_blocks(C.c_int(rows), C.c_int(cols), C.c_int(len(ln)), _p(ln), _p(s), _p(e), _p(pw), C.c_int(cam_block), C.c_int(lidar_block),
                                                         C.c_double(weight), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_build_camera_lidar_blocks: code {m}")
        bl.n = m

    @staticmethod
    def build_calibration_blocks(bl, rows, cols, image_lines, start, end, pose_block=0):
        ln, s, e = _arr(image_lines, np.float32).reshape(-1, 4), _arr(start, np.float64).reshape(-1, 3), _arr(end, np.float64).reshape(-1, 3)
        n, cap, *arrs = bl.args()
        m = load_library().pvb_build_calibration_blocks(C.c_int(rows), C.c_int(cols), C.c_int(len(ln)), _p(ln), _p(s), _p(e), C.c_int(pose_block), n, cap, *arrs)
        if m < 0:
            raise PvbError(f"pvb_build_calibration_blocks: code {m}")
        bl.n = m

    # ---- D/E/F
    def project_equirect(self, xyzi, T_cl, rows, cols):
        a = _arr(xyzi, np.float32).reshape(-1, 4)
        out = np.zeros((len(a), 3), np.float32)
        self._ck(self._L.pvb_project_equirect(self._h, _p(a), C.c_long(len(a)), _p(_arr(T_cl, np.float64)), C.c_int(rows), C.c_int(cols), _p(out)))
        return out

    def project_depth_image(self, xyzi, T_cl, rows, cols, size=3):
        a = _arr(xyzi, np.float32).reshape(-1, 4)
        img = np.zeros((rows, cols), np.uint16)
        self._ck(self._L.pvb_project_depth_image(self._h, _p(a), C.c_long(len(a)), _p(_arr(T_cl, np.float64)), C.c_int(rows), C.c_int(cols), C.c_int(size), _p(img)))
        return img

    def line_votes(self, ref_lines_world, nei_corner_world, p2s_off, p2s_ids, n_nei_lines, dist_threshold):
        rl = _arr(ref_lines_world, np.float64).reshape(-1, 6)
        pts = _arr(nei_corner_world, np.float32).reshape(-1, 4)
        M = np.zeros((n_nei_lines, len(rl)), np.int32)
        self._ck(self._L.pvb_line_votes(self._h, _p(rl), C.c_int(len(rl)), _p(pts), C.c_int(len(pts)), _p(_arr(p2s_off, np.int32)), _p(_arr(p2s_ids, np.int32)),
                                        C.c_int(n_nei_lines), C.c_double(dist_threshold), _p(M)))
        return M

    def angle_votes(self, rows, cols, lines, cloud_local, p2s_off, p2s_ids, n_segments, T_cl):
        ln = _arr(lines, np.float32).reshape(-1, 4)
        cl = _arr(cloud_local, np.float32).reshape(-1, 4)
        counts = np.zeros((len(ln), n_segments), np.int32)
        self._ck(self._L.pvb_angle_votes(self._h, C.c_int(rows), C.c_int(cols), _p(ln), C.c_int(len(ln)), _p(cl), C.c_int(len(cl)), _p(_arr(p2s_off, np.int32)),
                                         _p(_arr(p2s_ids, np.int32)), C.c_int(n_segments), _p(_arr(T_cl, np.float64)), _p(counts)))
        return counts
