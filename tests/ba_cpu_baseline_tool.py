"""TEST INFRASTRUCTURE (not collected by pytest).  CPU leg of tools/ba_scale.py: times the oracle's one-Jet<9>-at-a-time evaluation (what Ceres does with
AutoDiffCostFunction<PanoramaReprojResidual_1Angle, 1, 3, 3, 3>) of a Room-shaped reprojection problem on all host threads.
Usage: python tests/ba_cpu_baseline_tool.py [n_cams n_points]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pvo  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

n_cams = int(sys.argv[1]) if len(sys.argv) > 1 else 454
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
d = synth.make_ba_problem(n_cams=n_cams, n_points=n_points, track_len=(3, 10), seed=11)       # the problem of tools/ba_scale.py
R = pvo.Reproj(d["cam"], d["point"], d["bearing"], huber=4.0 * np.pi / 180.0)
t0 = time.perf_counter()
for _ in range(3):
    R.evaluate(d["cams"], d["points"])
cpu_s = (time.perf_counter() - t0) / 3
print(json.dumps({"n_observations": int(len(d["cam"])), "evaluate_s": cpu_s, "evals_per_s": len(d["cam"]) / cpu_s, "threads": pvo.num_threads()}))
