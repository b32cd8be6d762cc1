"""Dry run of tests/test_zz_gpu_reference_fixtures.py on the CPU: the oracle stands in for the device behind an object with the Context methods those tests
call.  It checks the TEST code (arguments, fixture keys, shapes, tolerances) wherever there is no GPU; on the B200 the same tests run against the kernels."""
import numpy as np

import panovlm_b200
from oracle import pvo


class FakeLF:
    def __init__(self, corner, p2s_off, p2s_ids, coeffs, end_points, R, t):
        self.corner, self.off, self.ids, self.coeffs, self.ends, self.R, self.t = corner, p2s_off, p2s_ids, coeffs, end_points, np.asarray(R, float), np.asarray(t, float)
        self.sizes = np.bincount(np.asarray(p2s_ids), minlength=len(coeffs))
    def world(self): return pvo.transform_cloud(self.R, self.t, self.corner)
    def lines_w(self): return pvo.transform_lines(self.R, self.t, self.coeffs)

class FakeCtx:
    def frames_set(self, tg, q): self.tg, self.q = tg, q
    def _Rt(self, p):
        R = pvo.aa_to_R(p[:3]).T; return R, -R @ p[3:]
    def frames_associate_point2plane(self, poses, ref, nei, tol, thr, k):
        i, j = ref[0], nei[0]; Ri, ti = self._Rt(poses[i]); Rj, tj = self._Rt(poses[j])
        q, pt, pl = pvo.associate_p2plane(pvo.transform_cloud(Ri, ti, self.tg[i]), Ri, ti, pvo.transform_cloud(Rj, tj, self.q[j]), Rj, tj, tol, thr, k, True)
        return np.zeros(len(q), np.int32), q, pt, pl
    def line2line_associate(self, a, b, thr):
        M = pvo.line_votes(a.lines_w(), b.world(), b.off, b.ids, len(b.coeffs), thr)
        return pvo.find_associations(a.coeffs, a.lines_w(), b.lines_w(), b.sizes, M)
    def line2line_knn_associate(self, a, b, thr):
        M = pvo.line2line_knn_votes(a.world(), a.off, a.ids, len(a.coeffs), b.world(), b.off, b.ids, len(b.coeffs), thr)
        return pvo.find_associations(a.coeffs, a.lines_w(), b.lines_w(), b.sizes, M)
    def point2line_segment_knn_associate(self, a, b, thr): return pvo.associate_p2line_segment_knn(a.world(), a.off, a.ids, a.coeffs, b.world(), b.R, b.t, thr)
    def point2line_segment_associate(self, a, b, thr): return pvo.associate_p2line_segment(a.lines_w(), a.coeffs, b.world(), b.R, b.t, thr)
    def frames_set_corners(self, c): self.c = c
    def frames_associate_point2line(self, poses, ref, nei, thr):
        i, j = ref[0], nei[0]; Ri, ti = self._Rt(poses[i]); Rj, tj = self._Rt(poses[j])
        q, pt, a, b = pvo.associate_p2line(pvo.transform_cloud(Ri, ti, self.c[i]), Ri, ti, pvo.transform_cloud(Rj, tj, self.c[j]), Rj, tj, thr)
        return np.zeros(len(q), np.int32), q, pt, a, b
    def project_equirect(self, cloud, T, rows, cols): return pvo.project_depth(cloud, rows, cols, T, 3, want_image=False)[1]
    def project_depth_image(self, cloud, T, rows, cols, size): return pvo.project_depth(cloud, rows, cols, T, size)[0]
    def camera_lidar_associate(self, rows, cols, lines, lf, T, flt, multi, im, lm):
        return pvo.associate_by_angle(rows, cols, lines, lf.corner, lf.off, lf.ids, lf.sizes, lf.ends, T, flt, multi, im, lm)
    def generate_line_tracks(self, lfs, nbrs, pv, thr, min_len):
        pa, pb, off, ma, mb = [], [], [0], [], []
        for i in range(len(lfs)):
            if pv is not None and not pv[i]: continue
            for nb in nbrs[i]:
                M = pvo.line_votes(lfs[nb].lines_w(), lfs[i].world(), lfs[i].off, lfs[i].ids, len(lfs[i].coeffs), thr)
                on, orf, _, _ = pvo.find_associations(lfs[nb].coeffs, lfs[nb].lines_w(), lfs[i].lines_w(), lfs[i].sizes, M)
                for x, y in sorted(set(zip(on.tolist(), orf.tolist()))): ma.append(x); mb.append(y)
                pa.append(i); pb.append(nb); off.append(len(ma))
        return pvo.line_tracks(pa, pb, off, ma, mb, min_len, True)
    def dense_set_target(self, t): self.dt = t
    def dense_set_sources(self, s, off): self.ds, self.doff = s, off
    def dense_params(self, tol, thr, k, typ, normalize, huber, weight): return (tol, thr, k, huber, weight)
    def dense_evaluate(self, poses, prm): return pvo.dense_icp_eval(self.dt, self.ds, self.doff, poses, prm[0], prm[1], prm[2], prm[3], prm[4], 1)[0]
    def transform_cloud(self, c, R, t): return pvo.transform_cloud(R, t, c)
    def pixel_line_neighbors(self, rows, cols, lines, cloud, T): return pvo.pixel_line_neighbors(rows, cols, lines, cloud, T)
    def undistort_clouds(self, cloud, off, T_wl, T_we, has=None):
        out = cloud.copy()
        for f in range(len(off) - 1):
            out[off[f]:off[f + 1]] = pvo.undistort_cloud(T_wl[f][:3, :3], T_wl[f][:3, 3], T_we[f][:3, :3], T_we[f][:3, 3], cloud[off[f]:off[f + 1]])
        return out



def test_gpu_fixture_tests_dry_run(monkeypatch):
    import test_zz_gpu_reference_fixtures as z
    monkeypatch.setattr(panovlm_b200, "LineFrame", FakeLF)
    ctx = FakeCtx()
    names = [n for n in dir(z) if n.startswith("test_") and n != "test_ceres_bridge_serves_the_device_rows_through_the_ceres_surface"]   # that one needs the real library
    assert len(names) == 9
    for name in names:
        getattr(z, name)(ctx)
