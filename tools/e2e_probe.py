"""Where does the end-to-end step go?  (development tool)  Times the pieces of bench.py's e2e step on the bench workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import panovlm_b200
from panovlm_b200 import synth
d = synth.make_dense_sweep(n_target=10_000_000, n_frames=64, pts_per_frame=156_250, seed=20260929, source_seed=20260930)
ctx = panovlm_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.dense_set_target(d["target"])
src = torch.from_numpy(d["src_local"]).pin_memory()
prm = ctx.dense_params(0.05, 1.0, 10, panovlm_b200.P2PLANE_METER, 1, 0.2, 1.0)
poses = d["poses_lw_init"].copy()
def wall(f, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return [round(t, 3) for t in ts]
ctx.dense_set_sources_ptr(src.data_ptr(), d["src_off"]); s = ctx.dense_evaluate(poses, prm)
for it in range(4):
    poses = ctx.dense_gauss_newton_step(s, poses, 1e-6); s = ctx.dense_evaluate(poses, prm)
print("evaluate only (resident, steady)", wall(lambda: ctx.dense_evaluate(poses, prm)))
print("upload only (set_sources + sync)", wall(lambda: ctx.dense_set_sources_ptr(src.data_ptr(), d["src_off"])))
def step():
    ctx.dense_set_sources_ptr(src.data_ptr(), d["src_off"]); ctx.dense_evaluate(poses, prm)
print("upload + evaluate", wall(step))
print("kernel window of the last one (ev0..ev1)", ctx.dense_kernel_time_ms())
os.environ["X"] = "1"
ctx.close()
