"""The multi-GPU exchange from a C++ host: libpanovlm_b200_nccl.so (ncclAllReduce as the reduce hook of a sharded pose graph) driven by tests/nccl_harness.cpp -
one process, one thread + one context per GPU, ncclCommInitAll.  CPU part: the library and the harness build and export their symbols; GPU part (needs >= 2 GPUs,
`gpurun --gpus 2`): every rank ends with the same poses, and they agree with the single-GPU solve of the same problem."""
import ctypes as C
import os

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)  # noqa: E731


def test_nccl_library_and_cpp_harness_build_and_export():
    import panovlm_b200
    panovlm_b200.load_library()
    L = C.CDLL(os.path.join(ROOT, "panovlm_b200", "libpanovlm_b200_nccl.so"))
    for name in ("pvb_nccl_attach", "pvb_nccl_detach", "pvb_nccl_allreduce", "pvb_nccl_last_result"):
        assert hasattr(L, name), name
    assert L.pvb_nccl_attach(None, None) != 0                     # argument check, no device needed
    from conftest import build_nccl_harness
    H = C.CDLL(build_nccl_harness())
    assert hasattr(H, "nccl_pose_graph_run")


def _problem(seed=21, n=60000, nb=24):
    """plane blocks on a chain-like pose graph (every frame linked to its next three), small pose noise: a well-posed LM problem"""
    rng = np.random.default_rng(seed)
    c = cases.random_blocks(seed, n, nb=nb)
    ref = rng.integers(0, nb, n)
    nei = (ref + rng.integers(1, 4, n)) % nb
    c["ref"], c["nei"] = ref.astype(np.int32), nei.astype(np.int32)
    c["type"] = np.where(rng.random(n) < 0.5, 0, 1).astype(np.int32)
    truth = np.concatenate([rng.normal(0, 0.2, (nb, 3)), rng.normal(0, 1.0, (nb, 3))], axis=1)
    truth[0] = 0
    from scipy.spatial.transform import Rotation
    consts = np.zeros((n, 12))
    for i in range(n):
        # a point p (nei frame) on a plane (ref frame): P = R_r R_n^T (p - t_n) + t_r lies on the plane at the true poses
        Rr, Rn = Rotation.from_rotvec(truth[ref[i], :3]).as_matrix(), Rotation.from_rotvec(truth[nei[i], :3]).as_matrix()
        pl = rng.normal(0, 3, 3)
        P = Rr @ Rn.T @ (pl - truth[nei[i], 3:]) + truth[ref[i], 3:]
        nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
        consts[i, :3] = pl; consts[i, 3:6] = nrm; consts[i, 6] = -nrm @ P + rng.normal(0, 0.01); consts[i, 7] = 1.0
    c["consts"] = consts
    c["huber"] = np.where(c["type"] == 0, 0.2, 2 * np.pi / 180)
    c["normalize"] = np.ones(n, np.int32)
    start = truth.copy()
    start[1:] += np.concatenate([rng.normal(0, 0.01, (nb - 1, 3)), rng.normal(0, 0.03, (nb - 1, 3))], axis=1)
    return c, start, truth


@pytest.mark.gpu
def test_cpp_host_shards_a_pose_graph_over_two_gpus_with_nccl(gpu_ctx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from conftest import build_nccl_harness
    H = C.CDLL(build_nccl_harness())
    c, start, truth = _problem()
    n, nb = len(c["type"]), c["nb"]
    mask = np.zeros(nb, np.uint8); mask[0] = 1
    ng = 2
    poses, summ, nblk = np.zeros((ng, nb, 6)), np.zeros((ng, 6)), np.zeros(ng, np.int32)
    rc = H.nccl_pose_graph_run(C.c_int(ng), C.c_long(n), p(c["type"]), p(c["ref"]), p(c["nei"]), p(c["normalize"]), p(c["huber"].astype(np.float64)), p(c["consts"]), C.c_int(nb),
                               p(start), p(mask), C.c_int(20), p(poses), p(summ), p(nblk))
    assert rc == 0, rc
    assert nblk.sum() == n and nblk.min() > 0
    assert np.array_equal(poses[0], poses[1]) and np.array_equal(summ[0], summ[1])            # identical steps on every rank
    gpu_ctx.blocks_set(c["type"], c["ref"], c["nei"], c["consts"], c["huber"], c["normalize"], nb)
    single, s1 = gpu_ctx.blocks_solve_lm(start, mask, 20)
    assert s1["iterations"] == summ[0][2] and abs(s1["final_cost"] - summ[0][1]) < 1e-9 * s1["final_cost"]
    assert np.abs((poses[0] - start) - (single - start)).max() < 1e-6 * np.abs(single - start).max()
    assert np.abs(poses[0] - truth).max() < 0.2 * np.abs(start - truth).max()
