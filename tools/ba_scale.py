"""Room-shaped bundle adjustment of the camera-camera term on one GPU (BASELINE.json configs[2] shape: 454 panoramas): timings of the
reprojection residual + Jacobian kernel, of one evaluation with the reduced blocks and of the whole Schur-complement LM, with the CPU oracle's
one-functor-at-a-time evaluation beside it.  Prints one JSON object; used for profiles/ and DESIGN.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

n_cams = int(sys.argv[1]) if len(sys.argv) > 1 else 454
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
d = synth.make_ba_problem(n_cams=n_cams, n_points=n_points, track_len=(3, 10), seed=11)
n_obs = len(d["cam"])
ctx = panovlm_b200.Context(0)
huber = 4.0 * np.pi / 180.0
ctx.reproj_set(d["cam"], d["point"], d["bearing"], n_cams, n_points, huber=huber)
out = {"n_cams": n_cams, "n_points": n_points, "n_observations": n_obs}
for rows, system, key in ((True, False, "rows"), (False, True, "reduced")):
    ts, ks = [], []
    for it in range(8):
        t0 = time.perf_counter()
        ctx.reproj_evaluate(d["cams"], d["points"], rows, system)
        ts.append(time.perf_counter() - t0)
        ks.append(ctx.reproj_kernel_time_ms())
    k_ms = float(np.median(ks[2:]))
    # algorithmic bytes per observation of the residual kernel: bearing 24 + indices 8 + point 24 (+ poses, amortised) in; r + J[9] = 80 out
    # (rows) or the 88-byte work row (reduced)
    b_alg = 24 + 8 + 24 + (80 + 4 if rows else 88)
    out[key] = {"evaluate_ms_end_to_end": float(np.median(ts[2:]) * 1e3), "kernel_ms": k_ms, "evals_per_s": n_obs / (k_ms * 1e-3), "algorithmic_bytes_per_obs": b_alg,
                "achieved_GBs": n_obs * b_alg / (k_ms * 1e-3) / 1e9}
cam_const = np.zeros((n_cams, 6), np.uint8)
cam_const[0] = 1
ctx.reproj_solve_lm(d["cams"], d["points"], cam_const, None, max_iterations=2)     # warm-up (contribution lists, buffers)
launches0 = ctx.kernel_launches
t0 = time.perf_counter()
c, p, s = ctx.reproj_solve_lm(d["cams"], d["points"], cam_const, None, max_iterations=20)
out["lm"] = {"seconds": time.perf_counter() - t0, "summary": s, "kernel_launches": ctx.kernel_launches - launches0, "unknowns": int(6 * (n_cams - 1) + 3 * n_points)}
# the CPU leg of this comparison (the oracle's one-Jet<9>-at-a-time evaluation of the same observations) lives in tests/ba_cpu_baseline_tool.py: only tests/, smoke() and
# bench.py's cpu_baseline leg may call the oracle
print(json.dumps(out))
