import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import panovlm_b200
ctx = panovlm_b200.Context(0)
rng = np.random.default_rng(2)
for nn in (2718, 5000, 9552):
    M = rng.normal(size=(nn, 64)); A = M @ M.T + np.eye(nn) * nn
    b = rng.normal(size=nn)
    ctx.cholesky_solve(A, b)
    x, ms = ctx.cholesky_solve(A, b)
    print(nn, ms, "ms", (nn ** 3 / 3) / (ms * 1e-3) / 1e12, "TFLOP/s", np.abs(A @ x - b).max())
