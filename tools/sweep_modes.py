"""Search-mode sweep on one GPU (development tool): the bench workload (configs[4]) through a 6-step Gauss-Newton run for
PVB_MODE = 1 (pruned two-pass walk, round 1) and 2 (buffered single pass) with / without search-radius hints; prints the device
time of the fused kernel at every step (step 0 after an upload is always a cold search) and checks that every variant
produces the same reduced systems bit for bit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

n_target = int(os.environ.get("SWEEP_TARGET", 10_000_000))
cells = [float(c) for c in os.environ.get("SWEEP_CELLS", "0").split(",")]
variants = os.environ.get("SWEEP_VARIANTS", "1:0,2:0,2:1").split(",")            # mode:hints
d = synth.make_dense_sweep(n_target=n_target, n_frames=64, pts_per_frame=156_250 * n_target // 10_000_000, seed=20260929, source_seed=20260930)
ref = None
for cell in cells:
    for v in variants:
        mode, hints = (int(x) for x in v.split(":"))
        os.environ["PVB_MODE"] = str(mode)
        ctx = panovlm_b200.Context(0)
        ctx.dense_set_hints(bool(hints))
        ctx.dense_set_target(d["target"], cell)
        ctx.dense_set_sources(d["src_local"], d["src_off"])
        prm = ctx.dense_params(0.05, 1.0, 10, panovlm_b200.P2PLANE_METER, 1, 0.2, 1.0)
        poses, ms, systems = d["poses_lw_init"].copy(), [], []
        for it in range(6):
            s = ctx.dense_evaluate(poses, prm)
            ms.append(round(ctx.dense_kernel_time_ms(), 3))
            systems.append(s)
            poses = ctx.dense_gauss_newton_step(s, poses, 1e-6)
        # back to the initial poses: the hints now come from the converged poses (largest displacement of the run)
        s = ctx.dense_evaluate(d["poses_lw_init"], prm)
        ms.append(round(ctx.dense_kernel_time_ms(), 3))
        systems.append(s)
        systems = np.array(systems)
        same = None
        if cell == cells[0]:
            if ref is None:
                ref = systems
            same = bool(np.array_equal(ref, systems))
        print(json.dumps({"mode": mode, "hints": hints, "cell": cell, "kernel_ms_per_step": ms, "identical_to_first_variant": same,
                          "accepted": float(systems[0][:, 28].sum()), "cost_first_last": [float(systems[0][:, 27].sum()), float(systems[5][:, 27].sum())]}), flush=True)
        ctx.close()
