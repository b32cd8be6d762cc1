"""The N>1 path on CPU: world_size-2 gloo processes shard the source frames, exchange the packed normal-equation blocks
with ONE allreduce and take the same Gauss-Newton step as a single process.  (Per-frame systems come from the oracle
here — there is no GPU in this container; on the GPU box bench.py runs the same exchange over NCCL.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _data():
    from panovlm_b200 import synth
    return synth.make_dense_sweep(n_target=60_000, n_frames=5, pts_per_frame=1500, seed=21)


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pvo
    from panovlm_b200 import Context, dist as pd
    d = _data()
    nf = len(d["src_off"]) - 1
    lo, hi = pd.shard_range(nf, world, rank)
    off = d["src_off"][lo:hi + 1] - d["src_off"][lo]
    src = d["src_local"][d["src_off"][lo]:d["src_off"][hi]]
    local, _, _ = pvo.dense_icp_eval(d["target"], src, off, d["poses_lw_init"][lo:hi], 0.05, 1.0, 10, 0.2, 1.0, 1)
    glob = pd.allreduce_systems(local, nf, lo).numpy()
    import ctypes as C
    import panovlm_b200
    L = panovlm_b200.load_library()
    poses = np.ascontiguousarray(d["poses_lw_init"], dtype=np.float64).copy()
    rc = L.pvb_dense_gauss_newton_step(glob.ctypes.data_as(C.c_void_p), C.c_int(nf), C.c_double(1e-6), poses.ctypes.data_as(C.c_void_p))
    assert rc == 0
    out[rank] = (glob, poses)
    dist.destroy_process_group()


def test_two_rank_allreduce_of_normal_equations_matches_single_process():
    from oracle import pvo
    pvo.build(); pvo.lib()
    import panovlm_b200
    if not os.path.exists(panovlm_b200.lib_path()):
        panovlm_b200.build_library()
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    d = _data()
    ref, _, _ = pvo.dense_icp_eval(d["target"], d["src_local"], d["src_off"], d["poses_lw_init"], 0.05, 1.0, 10, 0.2, 1.0, 1)
    g0, p0 = out[0]
    g1, p1 = out[1]
    assert np.array_equal(g0, g1) and np.array_equal(p0, p1)                      # every rank holds the same systems and takes the same step
    assert np.abs(g0 - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(p0 - d["poses_lw_init"]).max() > 1e-5                             # the step moved the poses
    from panovlm_b200 import dist as pd
    assert [pd.shard_range(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [pd.shard_range(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert pd.reduce_single_pose(ref).shape == (29,)
