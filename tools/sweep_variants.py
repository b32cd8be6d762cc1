"""Kernel-variant sweep on one GPU (development tool): generates the bench workload once and times the fused
associate kernel for combinations of the PVB_MINB / PVB_WALK / cell-size knobs (new context per combination)."""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

n_target = int(os.environ.get("SWEEP_TARGET", 10_000_000))
d = synth.make_dense_sweep(n_target=n_target, n_frames=64, pts_per_frame=156_250 * n_target // 10_000_000, seed=20260929, source_seed=20260930)
out = []
for minb, walk, cell in itertools.product((5, 6), (0, 1), (0.0,)):          # "walk" column = PVB_STAGE
    os.environ["PVB_MINB"], os.environ["PVB_STAGE"] = str(minb), str(walk)
    ctx = panovlm_b200.Context(0)
    ctx.dense_set_target(d["target"], cell)
    ctx.dense_set_sources(d["src_local"], d["src_off"])
    prm = ctx.dense_params(0.05, 1.0, 10, panovlm_b200.P2PLANE_METER, 1, 0.2, 1.0)
    ms = []
    for it in range(6):
        s = ctx.dense_evaluate(d["poses_lw_init"], prm)
        ms.append(ctx.dense_kernel_time_ms())
    rec = {"minb": minb, "stage": walk, "staged_vs_global_tiles": ctx.debug_counters(), "cell": cell, "kernel_ms": float(np.median(ms[2:])), "cost": float(s[:, 27].sum()), "n": float(s[:, 28].sum())}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    ctx.close()
best = min(out, key=lambda r: r["kernel_ms"])
print("BEST", json.dumps(best))
