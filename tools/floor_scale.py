"""Floor-shaped run (BASELINE.json configs[3]: 1593 frames of LiDAR odometry sharded across GPUs, one NCCL allreduce of the normal equations per
evaluation).  Launch with torchrun (one process per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/floor_scale.py [n_frames] [n_az]
or plainly for one GPU.  Every rank generates the same synthetic frames (seeded), associates and builds the residual blocks of its own range of
reference frames, and all ranks take identical LM steps.  Rank 0 prints one JSON object (timings are the max over ranks)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panovlm_b200  # noqa: E402
from panovlm_b200 import odometry, synth  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1593
n_az = int(sys.argv[2]) if len(sys.argv) > 2 else 600
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def aa_to_R(a):
    return Rotation.from_rotvec(a).as_matrix()


def R_to_aa(R):
    return Rotation.from_matrix(R).as_rotvec()


t0 = time.time()
frames = synth.make_sequence(n, n_az=n_az)
synth_s = time.time() - t0
rng = np.random.default_rng(1)
R0 = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.005, 3) * (i > 0)).as_matrix() for i, f in enumerate(frames)]
t0_ = [f["t_wl"] + rng.normal(0, 0.02, 3) * (i > 0) for i, f in enumerate(frames)]
poses = odometry.pose_blocks_from_world(R0, t0_, R_to_aa)
ctx = panovlm_b200.Context(local)
ctx.blocks_set_linear_solver(panovlm_b200.api.SOLVER_DEVICE)
cfg = odometry.OdometryConfig(line_to_line=False)
out = {"n_frames": n, "n_az": n_az, "n_gpus": world, "synth_s": synth_s}
res = None
for rep in range(2):                                   # the first pass warms up buffers and NCCL
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.time()
    if world > 1:
        new_poses, s = odometry.refine_pose_sharded(ctx, frames, poses, cfg, aa_to_R, world, rank)
    else:
        new_poses, s = odometry.refine_pose(ctx, frames, poses, cfg, aa_to_R)
    torch.cuda.synchronize()
    dt = torch.tensor([time.time() - t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res = (float(dt.item()), s)
out["refine_pose_s"] = res[0]
out["summary"] = res[1]
# every rank must hold the same poses
chk = torch.tensor(new_poses, device="cuda")
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["poses_identical_on_all_ranks"] = bool(torch.equal(lo, hi))
out["poses_sha"] = __import__("hashlib").sha256(np.ascontiguousarray(new_poses).tobytes()).hexdigest()[:16]
out["max_pose_update"] = float(np.abs(new_poses - poses).max())
if rank == 0:
    print(json.dumps(out))
ctx.close()
if world > 1:
    dist.destroy_process_group()
