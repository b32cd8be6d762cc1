"""Runs the Room-shaped reprojection evaluation (k_reproj_rows, rows mode) and a 64 x 156,250-point undistortion (k_undistort) standalone, for ncu
captures of the two kernels:  ncu --set full --clock-control none --import-source on -k regex:'k_reproj_rows|k_undistort' -c 2 -o ... python tools/profile_new_kernels.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

ctx = panovlm_b200.Context(0)
d = synth.make_ba_problem(n_cams=454, n_points=200_000, track_len=(3, 10), seed=7)
ctx.reproj_set(d["cam"], d["point"], d["bearing"], 454, 200_000, huber=4.0 * np.pi / 180.0)
ctx.reproj_evaluate(d["cams"], d["points"], True, False)
rng = np.random.default_rng(3)
nf, per = 64, 156_250
off = (np.arange(nf + 1) * per).astype(np.int32)
cloud = (rng.normal(size=(nf * per, 4)) * 20).astype(np.float32)
T_wl = np.tile(np.eye(4), (nf, 1, 1))
T_we = T_wl.copy()
for f in range(nf):
    T_we[f][:3, :3] = synth.rotvec_to_R(rng.normal(size=3) * 0.02)
    T_we[f][:3, 3] = rng.normal(size=3) * 0.1
out = ctx.undistort_clouds(cloud, off, T_wl, T_we)
print(json.dumps({"n_observations": int(len(d["cam"])), "n_points_undistorted": int(len(cloud)), "moved": float(np.abs(out[:, :3] - cloud[:, :3]).max())}))
ctx.close()
