"""Split-variant sweep of the fused associate kernel on one GPU (development tool): the bench workload once, then the kernel time of the fused kernel
(PVB_SPLIT=0) and of the search + tail pair for PVB_SPLIT = 8 / 10 / 12 (resident blocks per SM of the search kernel).  The reduced systems of every
variant must equal the fused kernel's bit for bit."""
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
ref = None
for split in (0, 8, 10, 12):
    os.environ["PVB_SPLIT"] = str(split)
    ctx = panovlm_b200.Context(0)
    ctx.dense_set_target(d["target"], 0.0)
    ctx.dense_set_sources(d["src_local"], d["src_off"])
    prm = ctx.dense_params(0.05, 1.0, 10, panovlm_b200.P2PLANE_METER, 1, 0.2, 1.0)
    ms = []
    for it in range(6):
        s = ctx.dense_evaluate(d["poses_lw_init"], prm)
        ms.append(ctx.dense_kernel_time_ms())
    if ref is None:
        ref = s.copy()
    print(json.dumps({"split": split, "kernel_ms": float(np.median(ms[2:])), "cost": float(s[:, 27].sum()), "n": float(s[:, 28].sum()),
                      "identical_to_fused": bool(np.array_equal(s, ref))}), flush=True)
    ctx.close()
