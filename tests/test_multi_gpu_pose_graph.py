"""Sharded pose graphs (BASELINE.json configs[3]): edges partitioned by reference frame, every rank reduces its residual blocks into
the GLOBAL edge layout, ONE sum-allreduce of the edge systems per evaluation, identical LM steps on every rank.
CPU: the sharding helpers and the additivity of the exchange through 2 gloo ranks (systems from the oracle).
GPU: 2 processes on cuda:0 (gloo stages the device buffer through the host; on the multi-GPU box the same hook reduces with NCCL in place) must
reproduce the single-process LM bit for bit on both ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _problem(nb=10, n=6000, seed=41):
    """consistent pose graph: plane / line residual blocks between neighbouring frames (|ref - nei| <= 3)"""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    truth = np.concatenate([rng.normal(0, 0.3, (nb, 3)), rng.normal(0, 1.0, (nb, 3))], axis=1)
    truth[0] = 0
    typ = rng.integers(0, 4, n).astype(np.int32)
    ref = rng.integers(0, nb, n).astype(np.int32)
    nei = np.clip(ref + rng.choice([-3, -2, -1, 1, 2, 3], n), 0, nb - 1).astype(np.int32)
    nei = np.where(nei == ref, (ref + 1) % nb, nei).astype(np.int32)
    consts = np.zeros((n, 12))
    for i in range(n):
        pw = rng.normal(0, 4, 3)
        Rr, tr = Rotation.from_rotvec(truth[ref[i], :3]).as_matrix(), truth[ref[i], 3:]
        Rn, tn = Rotation.from_rotvec(truth[nei[i], :3]).as_matrix(), truth[nei[i], 3:]
        p_ref, p_nei = Rr @ pw + tr, Rn @ pw + tn
        consts[i, :3] = p_nei + rng.normal(0, 0.01, 3)
        if typ[i] < 2:
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            d = -nrm @ p_ref
            if d < 0:
                nrm, d = -nrm, -d
            consts[i, 3:6] = nrm; consts[i, 6] = d; consts[i, 7] = 1.0
        else:
            dr = rng.normal(size=3); dr /= np.linalg.norm(dr)
            consts[i, 3:6] = p_ref + 0.3 * dr; consts[i, 6:9] = dr; consts[i, 9] = 1.0
    hub = np.where(typ % 2 == 1, 2 * np.pi / 180, 0.2)
    start = truth + np.concatenate([rng.normal(0, 0.01, (nb, 3)), rng.normal(0, 0.03, (nb, 3))], axis=1)
    start[0] = 0
    mask = np.zeros(nb, np.uint8); mask[0] = 1
    return dict(type=typ, ref=ref, nei=nei, consts=consts, huber=hub, start=start, mask=mask, nb=nb)


def _shard(P, world, rank):
    from panovlm_b200 import dist as pd
    weights = np.bincount(P["ref"], minlength=P["nb"])                 # residual count per reference frame
    b = pd.shard_frames_by_weight(weights, world)
    mine = (P["ref"] >= b[rank]) & (P["ref"] < b[rank + 1])
    return mine, b


def test_sharding_helpers():
    from panovlm_b200 import dist as pd
    assert pd.shard_frames_by_weight([1, 1, 1, 1], 2).tolist() == [0, 2, 4]
    assert pd.shard_frames_by_weight([10, 1, 1, 1, 1, 10], 3).tolist() == [0, 1, 5, 6]
    b = pd.shard_frames_by_weight(np.ones(1593), 8)
    assert b[0] == 0 and b[-1] == 1593 and np.all(np.diff(b) >= 199) and np.all(np.diff(b) <= 200)
    assert pd.shard_frames_by_weight([0, 0, 5], 4).tolist() == [0, 2, 3, 3, 3]     # empty shards are allowed
    er, en = pd.global_edge_list([3, 1, 1, 3], [2, 0, 0, 1])
    assert er.tolist() == [1, 3, 3] and en.tolist() == [0, 1, 2]
    P = _problem()
    m0, _ = _shard(P, 2, 0)
    m1, _ = _shard(P, 2, 1)
    assert np.all(m0 ^ m1) and 0.3 < m0.mean() < 0.7


def _cpu_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pvo
    from panovlm_b200 import dist as pd
    P = _problem()
    mine, _ = _shard(P, world, rank)
    blk = pvo.Blocks(P["type"][mine], P["ref"][mine], P["nei"][mine], P["consts"][mine], P["huber"][mine], 1)
    H, g, cost = blk.normal_equations(P["start"])
    packed = np.concatenate([H.ravel(), g, [cost]])
    out[rank] = pd.allreduce_edge_systems(packed.reshape(1, -1)).ravel()
    dist.destroy_process_group()


def test_two_rank_exchange_is_additive_on_cpu(oracle):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_cpu_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    P = _problem()
    H, g, cost = oracle.Blocks(P["type"], P["ref"], P["nei"], P["consts"], P["huber"], 1).normal_equations(P["start"])
    ref = np.concatenate([H.ravel(), g, [cost]])
    assert np.array_equal(out[0], out[1])
    assert np.abs(out[0] - ref).max() <= 1e-12 * np.abs(ref).max()


def _gpu_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import panovlm_b200
    from panovlm_b200 import api, dist as pd
    torch.cuda.set_device(0)
    ctx = panovlm_b200.Context(0)
    P = _problem()
    mine, bounds = _shard(P, world, rank)
    er, en = pd.global_edge_list(P["ref"], P["nei"])
    ctx.blocks_set_edge_list(er, en)
    if rank == 1:
        mine[:] = False                                                  # rank 1 owns no residual at all: it still takes part in every exchange
    else:
        mine[:] = True
    res = {}
    for variant in ("balanced", "one_rank_empty"):
        sel = _shard(P, world, rank)[0] if variant == "balanced" else mine
        ctx.blocks_set(P["type"][sel], P["ref"][sel], P["nei"][sel], P["consts"][sel], P["huber"][sel], 1, P["nb"])
        hook = pd.install_allreduce_hook(ctx)
        for kind in (api.SOLVER_HOST, api.SOLVER_DEVICE):
            ctx.blocks_set_linear_solver(kind)
            poses, summ = ctx.blocks_solve_lm(P["start"], is_const=P["mask"], max_iterations=12)
            res[(variant, kind)] = (poses, summ, hook["calls"]["n"])
        pd.remove_allreduce_hook(ctx)
    out[rank] = res
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_pose_graph_lm_equals_single_process(gpu_ctx):
    from panovlm_b200 import api
    P = _problem()
    single = {}
    try:
        gpu_ctx.blocks_set_edge_list(None)
        gpu_ctx.blocks_set(P["type"], P["ref"], P["nei"], P["consts"], P["huber"], 1, P["nb"])
        for kind in (api.SOLVER_HOST, api.SOLVER_DEVICE):
            gpu_ctx.blocks_set_linear_solver(kind)
            single[kind] = gpu_ctx.blocks_solve_lm(P["start"], is_const=P["mask"], max_iterations=12)
    finally:
        gpu_ctx.blocks_set_linear_solver(api.SOLVER_AUTO)
    assert single[api.SOLVER_HOST][1]["final_cost"] < 0.2 * single[api.SOLVER_HOST][1]["initial_cost"]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gpu_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for key, (p0, s0, calls0) in out[0].items():
        p1, s1, calls1 = out[1][key]
        assert np.array_equal(p0, p1) and s0 == s1 and calls0 == calls1 and calls0 > 0        # both ranks take the same steps
        ps, ss = single[key[1]]
        for k in ("iterations", "successful", "unsuccessful", "termination"):
            assert s0[k] == ss[k]
        # the sum over two ranks associates differently from the single-process sum: equal to rounding, far inside the 1e-4 gate
        assert abs(s0["final_cost"] - ss["final_cost"]) < 1e-9 * ss["final_cost"]
        assert np.abs(p0 - ps).max() < 1e-8 * np.abs(ps - P["start"]).max()
    # with one empty rank the sum has a single non-zero term: bit-identical to the single process
    for kind in (api.SOLVER_HOST, api.SOLVER_DEVICE):
        assert np.array_equal(out[0][("one_rank_empty", kind)][0], single[kind][0])


def _seq_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import panovlm_b200
    from panovlm_b200 import odometry, synth
    from oracle import pvo
    torch.cuda.set_device(0)
    ctx = panovlm_b200.Context(0)
    frames, poses0, cfg = _sequence()
    res = {}
    for dev in (True, False):
        res[dev] = odometry.refine_pose_sharded(ctx, frames, poses0, cfg, pvo.aa_to_R, world, rank, device_blocks=dev)
    out[rank] = res
    ctx.close()
    dist.destroy_process_group()


def _sequence():
    from panovlm_b200 import odometry, synth
    from oracle import pvo
    from scipy.spatial.transform import Rotation
    frames = synth.make_sequence(10, n_az=360, tilt=0.3)
    rng = np.random.default_rng(3)
    R0 = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.005, 3) * (i > 0)).as_matrix() for i, f in enumerate(frames)]
    t0 = [f["t_wl"] + rng.normal(0, 0.02, 3) * (i > 0) for i, f in enumerate(frames)]
    return frames, odometry.pose_blocks_from_world(R0, t0, pvo.R_to_aa), odometry.OdometryConfig(line_to_line=True)


@pytest.mark.gpu
def test_sharded_refine_pose_with_device_built_blocks(gpu_ctx, oracle):
    """RefinePose sharded over 2 ranks (reference frames split by weight): the point-to-plane correspondences of a rank's own edges become residual blocks on the
    device and are filed under the GLOBAL edge list (pvb_frames_point2plane_blocks with pvb_blocks_set_edge_list), the line blocks built on the host are appended.
    Both ranks end at bit-identical poses; they equal the host-built sharded run and the single-process RefinePose to rounding (same rows, different tile splits)."""
    from panovlm_b200 import odometry
    frames, poses0, cfg = _sequence()
    ps, ss = odometry.refine_pose(gpu_ctx, frames, poses0, cfg, oracle.aa_to_R)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_seq_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for dev in (True, False):
        (p0, s0), (p1, s1) = out[0][dev], out[1][dev]
        assert np.array_equal(p0, p1) and s0["allreduces"] == s1["allreduces"] > 0
        assert s0["n_blocks_local"] + s1["n_blocks_local"] == ss["n_blocks"]
        for k in ("iterations", "successful", "unsuccessful", "termination"):
            assert s0[k] == ss[k]
        assert abs(s0["final_cost"] - ss["final_cost"]) < 1e-9 * ss["final_cost"]
        assert np.abs(p0 - ps).max() < 1e-8 * max(1e-12, np.abs(ps - poses0).max())
    # device-built and host-built shards hold the same rows in a different tile split: equal to rounding
    assert np.abs(out[0][True][0] - out[0][False][0]).max() < 1e-8 * np.abs(ps - poses0).max()
