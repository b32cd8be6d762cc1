"""Room-scale (BASELINE.json configs[1] shape) run of the pose-graph path on one GPU: N synthetic frames on a loop,
FindNeighbors(6), point-to-plane association of every edge in ONE fused-kernel launch, line-to-line associations,
residual blocks, one LM solve (first frame constant).  Prints timings; used for profiles/ and DESIGN.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import odometry, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 454
t0 = time.time()
frames = synth.make_sequence(n, n_az=1800)
gen_s = time.time() - t0
from scipy.spatial.transform import Rotation  # noqa: E402


def aa_to_R(a):
    return Rotation.from_rotvec(a).as_matrix()


def R_to_aa(R):
    return Rotation.from_matrix(R).as_rotvec()


rng = np.random.default_rng(1)
R0 = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.005, 3) * (i > 0)).as_matrix() for i, f in enumerate(frames)]
t0_ = [f["t_wl"] + rng.normal(0, 0.02, 3) * (i > 0) for i, f in enumerate(frames)]
poses = odometry.pose_blocks_from_world(R0, t0_, R_to_aa)
ctx = panovlm_b200.Context(0)
cfg = odometry.OdometryConfig(line_to_line=False)
out = {"n_frames": n, "synth_s": gen_s}
# association of all edges
R_wl, t_wl = odometry.world_from_pose_blocks(poses, aa_to_R)
nb = panovlm_b200.Context.find_neighbors(np.array(t_wl), None, None, 6)
edges = [(i, j) for i in range(n) for j in nb[i] if 0 <= j < n and j != i]
ref = np.array([e[0] for e in edges], np.int32); nei = np.array([e[1] for e in edges], np.int32)
ctx.frames_set([f["surfLessFlat"] for f in frames], [f["surfFlat"] for f in frames])
for rep in range(3):
    t = time.time()
    e, q, pt, pl = ctx.frames_associate_point2plane(poses, ref, nei, cfg.plane_tolerance, cfg.plane_dis_threshold, 10)
    out["associate_all_edges_s"] = time.time() - t
out.update(n_edges=len(edges), n_queries=int(sum(len(frames[j]["surfFlat"]) for j in nei)), n_point2plane=int(len(e)))
t = time.time()
bl, _ = odometry.build_problem(ctx, frames, poses, cfg, aa_to_R)
out["build_problem_s"] = time.time() - t
v = bl.view()
t = time.time()
ctx.blocks_set(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"], n)
out["blocks_set_s"] = time.time() - t
t = time.time()
for _ in range(5):
    ctx.blocks_evaluate(poses, want_rows=False, want_system=True)
out["evaluate_reduced_s"] = (time.time() - t) / 5
out["k_eval_blocks_ms"] = ctx.blocks_kernel_time_ms()
mask = np.zeros(n, np.uint8); mask[0] = 1
from panovlm_b200 import api  # noqa: E402
for name, kind in (("device", api.SOLVER_DEVICE), ("host", api.SOLVER_HOST)):
    if name == "host" and os.environ.get("PVB_SKIP_HOST_LM"):
        continue
    ctx.blocks_set_linear_solver(kind)
    t = time.time()
    new_poses, s = ctx.blocks_solve_lm(poses, mask, cfg.max_lm_iterations)
    out[f"solve_lm_{name}_s"] = time.time() - t
    s["n_blocks"] = bl.n
    out[f"lm_{name}"] = s
    out[f"pose_err_after_{name}"] = float(np.abs(new_poses - odometry.pose_blocks_from_world([f["R_wl"] for f in frames], [f["t_wl"] for f in frames], R_to_aa)).max())
# one whole RefinePose (association + residual blocks + LM) with the blocks built on the host vs on the device
ctx.blocks_set_linear_solver(api.SOLVER_DEVICE)
for name, dev in (("host_blocks", False), ("device_blocks", True)):
    for rep in range(2):
        t = time.time()
        p2, s2 = odometry.refine_pose(ctx, frames, poses, cfg, aa_to_R, device_blocks=dev)
        out[f"refine_pose_{name}_s"] = time.time() - t
    out[f"refine_pose_{name}_final_cost"] = s2["final_cost"]
# the device Cholesky alone at this size
rng2 = np.random.default_rng(2)
nn = 6 * (n - 1)
M = rng2.normal(size=(nn, 64)); Amat = M @ M.T + np.eye(nn) * nn
x, ms = ctx.cholesky_solve(Amat, rng2.normal(size=nn))
x, ms = ctx.cholesky_solve(Amat, rng2.normal(size=nn))
out["device_cholesky"] = {"n": nn, "factor_and_solve_ms": ms, "gflops": (nn ** 3 / 3) / (ms * 1e-3) / 1e9}
truth = odometry.pose_blocks_from_world([f["R_wl"] for f in frames], [f["t_wl"] for f in frames], R_to_aa)
out["pose_err_before_after"] = [float(np.abs(poses - truth).max()), float(np.abs(new_poses - truth).max())]
print(json.dumps(out))
