"""Device Cholesky (pvb_solver.cuh) alone at pose-graph sizes: factor + substitution time and the achieved FP64 rate."""
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panovlm_b200
ctx = panovlm_b200.Context(0)
rng = np.random.default_rng(2)
sizes = [int(a) for a in sys.argv[1:]] or [2718, 5000, 9552]
for nn in sizes:
    M = rng.normal(size=(nn, 64)); A = M @ M.T + np.eye(nn) * nn
    b = rng.normal(size=nn)
    ctx.cholesky_solve(A, b)
    x, ms = ctx.cholesky_solve(A, b)
    print(nn, ms, "ms", (nn ** 3 / 3) / (ms * 1e-3) / 1e12, "TFLOP/s", np.abs(A @ x - b).max())
