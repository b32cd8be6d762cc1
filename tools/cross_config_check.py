"""The association result of the fused kernel must not depend on the search structure: runs the same dense problem with different
cell sizes / start radii of the pruned walk (fresh Context per configuration: the knobs are read at pvb_create) and compares the
per-query validity flags, planes and residuals.  Differences are listed with their k-th neighbour distances (exact ties only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import synth  # noqa: E402

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = synth.make_dense_sweep(n_target=n_target, n_frames=8, pts_per_frame=50_000, seed=5)
prm = panovlm_b200.Context.dense_params(plane_tolerance=0.05, dist_threshold=1.0, k=10, huber=0.2)
ref = None
for (r0, cell, cap, stage) in [(1, 0.0, 4, 0), (1, 0.0, 4, 1), (1, 0.12, 8, 0), (2, 0.12, 8, 0), (2, 0.10, 8, 0), (2, 0.08, 16, 0), (2, 0.065, 32, 0), (1, 0.3, 4, 0)]:
    os.environ.update(PVB_R0=str(r0), PVB_CELLCAP=str(cap), PVB_STAGE=str(stage))
    ctx = panovlm_b200.Context(0)
    ctx.dense_set_target(d["target"], cell)
    ctx.dense_set_sources(d["src_local"], d["src_off"])
    v, pt, pl, r, J = ctx.dense_get_rows(d["poses_lw_init"], prm)
    S = ctx.dense_evaluate(d["poses_lw_init"], prm)
    if ref is None:
        ref = (v, pl, r, S)
        print(f"reference r0={r0} cell=auto stage={stage}: {v.sum()} of {len(v)} associated, cost {S[:, 27].sum():.12f}")
        continue
    dv = np.nonzero(v != ref[0])[0]
    both = v & ref[0]
    dpl = np.abs(pl[both] - ref[1][both]).max() if both.any() else 0.0
    dr = np.abs(r[both] - ref[2][both]).max() if both.any() else 0.0
    print(f"r0={r0} cell={cell} stage={stage}: valid flags differ at {len(dv)} queries, max |plane diff| {dpl:.3e}, max |residual diff| {dr:.3e}, "
          f"system diff {np.abs(S - ref[3]).max():.3e}, cost {S[:, 27].sum():.12f}")
    del ctx
