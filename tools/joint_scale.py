"""Room-shaped joint camera-LiDAR refinement on one GPU (BASELINE.json configs[2]: 454 frames, equirectangular projection + line association +
the three residual families in one LM).  Prints one JSON object with the timings of one CameraLidarOptimizer::Optimize call."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import joint, synth  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 454
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
n_az = int(sys.argv[3]) if len(sys.argv) > 3 else 900


def aa_to_R(a):
    return Rotation.from_rotvec(a).as_matrix()


t0 = time.time()
d = synth.make_joint_problem(n_frames=n, n_points=n_points, n_az=n_az, clutter=100, track_len=(3, 10))
out = {"n_frames": n, "n_points": n_points, "synth_s": time.time() - t0}
ctx = panovlm_b200.Context(0)
ctx.blocks_set_linear_solver(panovlm_b200.api.SOLVER_DEVICE)
cfg = joint.JointConfig(max_lm_iterations=20)
for rep in range(2):
    t = time.time()
    pairs = joint.associate_lines(ctx, d["frames"], d["image_lines"], d["cams"], d["lidars"], d["rows"], d["cols"], aa_to_R)
    out["associate_lines_s"] = time.time() - t
    t = time.time()
    v, _, counts = joint.build_problem(ctx, d, d["cams"], d["lidars"], d["points"], cfg, aa_to_R)
    out["build_problem_s"] = time.time() - t
    l0 = ctx.kernel_launches
    t = time.time()
    cams, lidars, points, summ, _ = joint.optimize(ctx, d, d["cams"], d["lidars"], d["points"], cfg, aa_to_R)
    out["optimize_s"] = time.time() - t
    out["kernel_launches"] = ctx.kernel_launches - l0
out["summary"] = summ
out["unknowns"] = int(12 * n - 6 + 3 * n_points)
out["lidar_t_err_before_after"] = [float(np.abs(d["lidars"][:, 3:] - d["lidars_gt"][:, 3:]).mean()), float(np.abs(lidars[:, 3:] - d["lidars_gt"][:, 3:]).mean())]
print(json.dumps(out))
