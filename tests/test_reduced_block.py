"""include/panovlm_b200_reduced.hpp: the per-edge normal equations (H upper | g | cost | n) as ONE 13-residual block for a least-squares
solver (north_star: "Ceres only sees the reduced system").  CPU part: the factorisation alone; GPU part: both Ceres bridges of
include/panovlm_b200_ceres_adapter.hpp driven through the ceres surface give the same normal equations and cost."""
import ctypes as C

import numpy as np
import pytest

import cases

p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)  # noqa: E731


def _system(J, r):
    H, g = J.T @ J, J.T @ r
    S = np.zeros(92)
    S[:78] = H[np.triu_indices(12)]
    S[78:90] = g
    S[90] = 0.5 * float(r @ r)
    S[91] = len(r)
    return S, H, g


@pytest.fixture(scope="module")
def harness_lib():
    from conftest import build_adapter_harness
    return C.CDLL(build_adapter_harness())


@pytest.mark.parametrize("rows", [400, 12, 7, 1, 0])
def test_reduced_block_reproduces_the_normal_equations_and_the_cost(harness_lib, rows):
    rng = np.random.default_rng(rows)
    for trial in range(20):
        J = rng.normal(size=(rows, 12)) * rng.uniform(0.01, 30, 12)            # badly scaled columns like angle / metre blocks
        if trial % 3 == 1 and rows > 0:
            J[:, 3:6] = 0.0                                                     # a parameter block the rows do not depend on
        if trial % 3 == 2 and rows > 1:
            J[:, 7] = J[:, 6] * 2.0                                             # linearly dependent columns
        r = rng.normal(size=rows)
        S, H, g = _system(J, r)
        Jt, rt, rank = np.zeros((12, 12)), np.zeros(13), C.c_int(0)
        harness_lib.reduced_block(p(S), p(Jt), p(rt), C.byref(rank))
        scale = max(1e-300, np.abs(H).max())
        assert np.abs(Jt.T @ Jt - H).max() <= 1e-11 * scale
        assert np.abs(Jt.T @ rt[:12] - g).max() <= 1e-9 * max(1e-300, np.abs(g).max(), np.sqrt(scale))
        assert abs(0.5 * float(rt @ rt) - S[90]) <= 1e-12 * max(1.0, S[90])
        assert rank.value == np.linalg.matrix_rank(J) if rows else rank.value == 0
        assert np.all(Jt[rank.value:] == 0.0)                                    # rows beyond the rank are empty


@pytest.mark.gpu
def test_reduced_bridge_gives_ceres_the_same_normal_equations_as_the_row_bridge(harness_lib):
    """Row bridge (one SizedCostFunction<1,3,3,3,3> + ceres::HuberLoss per correspondence, raw device rows) against the reduced bridge (one
    SizedCostFunction<13,3,3,3,3> per pose-graph edge, loss applied on the device before the reduction): J^T J, J^T r and the robust cost
    as Ceres assembles them - with and without a constant parameter block, and on an edge with fewer rows than unknowns."""
    c = cases.random_blocks(11, 4000, nb=6)
    c["nei"] = np.where(c["nei"] == c["ref"], (c["ref"] + 1) % c["nb"], c["nei"]).astype(np.int32)     # a Ceres block cannot hold one parameter block twice
    n = len(c["type"])
    # make one edge rank deficient: only 3 blocks on (4 -> 5)
    keep = ~((c["ref"] == 4) & (c["nei"] == 5))
    idx = np.nonzero(~keep)[0]
    keep[idx[:3]] = True
    arr = {k: np.ascontiguousarray(c[k][keep]) for k in ("type", "ref", "nei", "normalize", "huber", "consts")}
    n = int(keep.sum())
    D = 6 * c["nb"]
    ints = [np.ascontiguousarray(arr[k], np.int32) for k in ("type", "ref", "nei", "normalize")]
    for const_block in (-1, 0, 4):
        Hr, gr, Hd, gd, costs, counts = np.zeros((D, D)), np.zeros(D), np.zeros((D, D)), np.zeros(D), np.zeros(2), np.zeros(2, np.int64)
        rc = harness_lib.adapter_reduced_run(0, C.c_long(n), p(ints[0]), p(ints[1]), p(ints[2]), p(ints[3]), p(arr["huber"].astype(np.float64)), p(arr["consts"].astype(np.float64)),
                                             C.c_int(c["nb"]), p(np.ascontiguousarray(c["poses"], np.float64)), C.c_int(const_block), p(Hr), p(gr), p(Hd), p(gd), p(costs), p(counts))
        assert rc == 0, rc
        assert counts[0] == n and counts[1] == len(set(zip(arr["ref"].tolist(), arr["nei"].tolist())))
        assert counts[1] < counts[0] / 50
        scale = np.abs(Hr).max()
        assert np.abs(Hr - Hd).max() <= 1e-10 * scale
        assert np.abs(gr - gd).max() <= 1e-10 * np.abs(gr).max()
        assert abs(costs[0] - costs[1]) <= 1e-10 * costs[0]
        if const_block >= 0:
            assert np.all(Hd[6 * const_block:6 * const_block + 6] == 0) and np.all(gd[6 * const_block:6 * const_block + 6] == 0)
