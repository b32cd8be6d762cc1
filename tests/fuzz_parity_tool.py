"""TEST INFRASTRUCTURE (lives under tests/ because it calls the oracle; not collected by pytest - run it by hand on a GPU box: python tests/fuzz_parity_tool.py [n]).
Randomised GPU-vs-oracle parity sweep of the fused association path (development / validation tool).
Random scene sizes, rigid placements of the world (negative coordinates, clamped grids), cell sizes (ring-2+ paths),
thresholds, k, tolerances, residual types, with and without TMA staging.  Exits non-zero on the first mismatch."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from oracle import pvo  # noqa: E402
from panovlm_b200 import synth  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ctxs = {}
for stage in (0, 1):
    os.environ["PVB_STAGE"] = str(stage)
    ctxs[stage] = panovlm_b200.Context(0)
rng = np.random.default_rng(12345)
worst = {"plane": 0.0, "sys": 0.0}
n_assoc_total = 0
for case in range(n_cases):
    n_t = int(rng.integers(3000, 150000)); nf = int(rng.integers(1, 4)); per = int(rng.integers(100, 2500))
    d = synth.make_dense_sweep(n_target=n_t, n_frames=nf, pts_per_frame=per, seed=1000 + case)
    # place the whole world somewhere else (rigid): target points move, source poses are composed accordingly
    Rw = Rotation.from_rotvec(rng.normal(0, 0.6, 3)).as_matrix(); tw = rng.normal(0, 30, 3)
    tgt = d["target"].copy(); tgt[:, :3] = (d["target"][:, :3].astype(np.float64) @ Rw.T + tw).astype(np.float32)
    if case % 5 == 0:
        tgt[:, 3] = np.where(rng.random(n_t) < 0.3, 16.0, 1.0)                 # mixed classes: the same-class test matters
    poses = []
    for f in range(nf):
        R_lw = Rotation.from_rotvec(d["poses_lw_init"][f, :3]).as_matrix(); t_lw = d["poses_lw_init"][f, 3:]
        R_wl, t_wl = R_lw.T, -R_lw.T @ t_lw
        R2, t2 = Rw @ R_wl, Rw @ t_wl + tw                                     # new T_wl
        poses.append(np.concatenate([Rotation.from_matrix(R2.T).as_rotvec(), -R2.T @ t2]))
    poses = np.array(poses)
    cell = float(rng.choice([0.0, 0.0, 0.07, 0.15, 0.3, 0.6, 1.5]))
    thr = float(rng.choice([0.2, 0.3, 1.0, 2.0])); k = int(rng.choice([5, 10])); tol = float(rng.choice([0.01, 0.05]))
    rtype = int(rng.choice([0, 1])); hub = 0.2 if rtype == 0 else 2 * np.pi / 180
    exp_sys = []
    for f in range(nf):
        lo, hi = d["src_off"][f], d["src_off"][f + 1]
        R_wl = pvo.aa_to_R(poses[f, :3]).T; t_wl = -R_wl @ poses[f, 3:]
        w = pvo.transform_cloud(R_wl, t_wl, d["src_local"][lo:hi])
        oq, opt, opl = pvo.associate_p2plane(tgt, np.eye(3), np.zeros(3), w, R_wl, t_wl, tol, thr, k, True)
        consts = np.zeros((len(oq), 12)); consts[:, :3] = opt; consts[:, 3:7] = opl; consts[:, 7] = 1.0
        blk = pvo.Blocks(np.full(len(oq), rtype), 0, 1, consts, hub, 1)
        ro, Jo, co = blk.evaluate(np.stack([np.zeros(6), poses[f]]), apply_loss=True) if len(oq) else (np.zeros(0), np.zeros((0, 12)), np.zeros(0))
        exp_sys.append((oq, opl, Jo[:, 6:], ro, co, opt))
    for stage, ctx in ctxs.items():
        ctx.dense_set_target(tgt, cell)
        ctx.dense_set_sources(d["src_local"], d["src_off"])
        prm = ctx.dense_params(tol, thr, k, rtype, 1, hub, 1.0)
        s = ctx.dense_evaluate(poses, prm)
        valid, pt, pl, r, j6 = ctx.dense_get_rows(poses, prm)
        for f in range(nf):
            lo, hi = d["src_off"][f], d["src_off"][f + 1]
            oq, opl, J6, ro, co, opt = exp_sys[f]
            gq = np.nonzero(valid[lo:hi])[0]
            if not np.array_equal(gq, oq):
                print("MISMATCH association set", dict(case=case, stage=stage, frame=f, cell=cell, thr=thr, k=k, tol=tol, n_gpu=len(gq), n_cpu=len(oq)))
                sys.exit(1)
            if len(oq):
                # plane parity where it matters: the offset of the two planes at the query (coefficients of a plane far from the origin are
                # individually ill-conditioned: d changes by |dn| * distance) + the residuals themselves
                R_wl = pvo.aa_to_R(poses[f, :3]).T; t_wl = -R_wl @ poses[f, 3:]
                pw = opt @ R_wl.T + t_wl
                dpl = pl[lo:hi][gq] - opl
                dp = max(np.abs((dpl[:, :3] * pw).sum(1) + dpl[:, 3]).max(), (np.abs(r[lo:hi][gq] - ro) / np.maximum(1e-9, np.abs(ro))).max() * 1e-3)
                worst["plane"] = max(worst["plane"], dp)
                H = J6.T @ J6; gv = J6.T @ ro
                iu = np.triu_indices(6)
                ds = max(np.abs(s[f, :21] - H[iu]).max() / max(1e-30, np.abs(H).max()), np.abs(s[f, 21:27] - gv).max() / max(1e-30, np.abs(gv).max(), 1e-9 * np.abs(H).max()))
                worst["sys"] = max(worst["sys"], ds)
                if dp > 1e-8 or ds > 1e-7 or s[f, 28] != len(oq) or abs(s[f, 27] - co.sum()) > 1e-8 * max(1e-30, co.sum()):
                    print("MISMATCH values", dict(case=case, stage=stage, frame=f, plane=dp, sys=ds, n=(s[f, 28], len(oq)), cost=(s[f, 27], co.sum())))
                    sys.exit(1)
            n_assoc_total += len(oq)
print(json.dumps({"cases": n_cases, "associations_checked": n_assoc_total, "worst_plane_abs": worst["plane"], "worst_system_rel": worst["sys"], "staged_vs_global_tiles": ctxs[1].debug_counters()}))
