"""Device line tails against host line tails at a few hundred frames: same block count and RefinePose result, and what the build of one outer iteration costs either way.
usage: python tools/line_tails_probe.py [n_frames]  ->  one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import panovlm_b200
    from panovlm_b200 import odometry, synth
    from scipy.spatial.transform import Rotation
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    t0 = time.time()
    frames = synth.make_sequence(n, n_az=600, tilt=0.3)
    synth_s = time.time() - t0
    rng = np.random.default_rng(5)
    R0 = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.004, 3) * (i > 0)).as_matrix() for i, f in enumerate(frames)]
    t_0 = [f["t_wl"] + rng.normal(0, 0.02, 3) * (i > 0) for i, f in enumerate(frames)]
    aa_to_R, R_to_aa = (lambda a: Rotation.from_rotvec(a).as_matrix()), (lambda R: Rotation.from_matrix(R).as_rotvec())   # the mirror takes its Rodrigues pair as arguments
    poses = odometry.pose_blocks_from_world(R0, t_0, R_to_aa)
    ctx = panovlm_b200.Context(0)
    cfg = odometry.OdometryConfig()
    out = {"n_frames": n, "synth_s": synth_s}
    res = {}
    for name, dev in (("host_tails", False), ("device_tails", True), ("host_tails_again", False), ("device_tails_again", True)):
        t1 = time.time()
        p, s = odometry.refine_pose(ctx, frames, poses, cfg, aa_to_R, device_line_blocks=dev)
        res[name] = p
        out[name] = {"total_s": time.time() - t1, "build_s": s["build_s"], "lm_s": s["lm_s"], "n_blocks": s["n_blocks"], "n_edges": s["n_edges"], "iterations": s["iterations"],
                     "final_cost": s["final_cost"]}
    out["max_pose_difference"] = float(np.abs(res["host_tails_again"] - res["device_tails_again"]).max())
    out["same_block_count"] = out["host_tails_again"]["n_blocks"] == out["device_tails_again"]["n_blocks"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
