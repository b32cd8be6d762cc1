"""The product's `__host__ __device__` math (panovlm_b200/csrc/pvb_math.cuh, pvb_knn.cuh, pvb_host.hpp) compiled
with g++ (tests/host_harness.cpp) and compared with the oracle on the CPU — catches arithmetic/parity bugs before any
GPU time is spent.  The same functions run inside the CUDA kernels; `-m gpu` tests then check the kernels proper."""
import ctypes as C
import os

import numpy as np
import pytest

import cases

G = os.path.join(os.path.dirname(__file__), "golden")
p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None  # noqa: E731


def _eval(h, c, loss):
    n = len(c["type"])
    r, J, cost = np.zeros(n), np.zeros((n, 12)), np.zeros(n)
    h.pvbh_eval_blocks(C.c_long(n), p(c["type"]), p(c["ref"]), p(c["nei"]), p(c["normalize"]), p(c["huber"]), p(np.ascontiguousarray(c["consts"])),
                       p(np.ascontiguousarray(c["poses"])), C.c_int(c["nb"]), C.c_int(loss), p(r), p(J), p(cost))
    return r, J, cost


@pytest.mark.parametrize("gen", [cases.random_blocks, cases.random_blocks_f6])
def test_analytic_jacobian_matches_jet_oracle(oracle, harness, gen):
    c = gen(1, 6000)
    b = oracle.Blocks(c["type"], c["ref"], c["nei"], c["consts"], c["huber"], c["normalize"])
    for loss in (0, 1):
        r, J, cost = b.evaluate(c["poses"], apply_loss=bool(loss))
        r2, J2, c2 = _eval(harness, c, loss)
        assert (np.abs(r - r2) / np.maximum(1e-9, np.abs(r))).max() < 1e-8           # gate: 1e-5 relative
        assert (np.abs(J - J2).max(1) / np.maximum(1e-12, np.abs(J).max(1))).max() < 1e-7   # gate: 1e-6 relative
        assert np.abs(cost - c2).max() < 1e-10


def test_zero_rows_and_golden_functors(harness):
    c = cases.on_plane_blocks()
    r, J, _ = _eval(harness, c, 0)
    assert np.all(r == 0) and np.all(J == 0)
    for name in ("functors.npz", "functors_f6.npz", "ref_functors.npz"):      # the last one: outputs of the reference's own CostFunction.h (oracle/_ref)
        g = dict(np.load(os.path.join(G, name)))
        g["nb"] = int(g["nb"])
        r, J, _ = _eval(harness, g, 0)
        assert np.abs(r - g["residual"]).max() < 1e-8
        assert (np.abs(J - g["jacobian"]) / np.maximum(1e-9, np.abs(g["jacobian"]).max(1, keepdims=True))).max() < 1e-7


@pytest.mark.parametrize("prune", [4, 3, 1, 2, 0])
def test_grid_knn_and_association_equal_oracle(oracle, harness, prune):
    """prune = 4: the buffered single-pass search over merged super-rows (dense mode default); 3: the buffered single-pass search (frames mode default); 1 / 2: the pruned two-pass walk starting from a 3x3x3 / 5x5x5 block
    (rows / cells beyond the running K-th distance are skipped); 0: the exhaustive block walk."""
    harness.pvbh_set_prune(C.c_int(prune))
    harness.pvbh_set_hints(None, None)
    g = np.load(os.path.join(G, "assoc_pair.npz"))
    refw, neiw = np.ascontiguousarray(g["ref_world"]), np.ascontiguousarray(g["nei_world"])
    R_ref, t_ref, R_nei, t_nei = (np.ascontiguousarray(g[k]) for k in ("R_ref", "t_ref", "R_nei", "t_nei"))
    for (h, thr, tol, K) in [(0.5, 1.0, 0.05, 10), (0.17, 1.0, 0.05, 10), (1.3, 1.0, 0.01, 10), (0.3, 0.3, 0.05, 5), (0.07, 0.3, 0.05, 5)]:
        q, pt, pl = oracle.associate_p2plane(refw, R_ref, t_ref, neiw, R_nei, t_nei, tol, thr, K, True)
        m = len(neiw)
        valid, pl2, pt2 = np.zeros(m, np.uint8), np.zeros((m, 4)), np.zeros((m, 3))
        ni, nd = np.zeros((m, K), np.int32), np.zeros((m, K), np.float32)
        harness.pvbh_associate(p(refw), C.c_int(len(refw)), p(R_ref), p(t_ref), p(neiw), C.c_int(m), p(R_nei), p(t_nei), C.c_double(h), C.c_float(thr),
                               C.c_double(tol), C.c_int(K), p(valid), p(pt2), p(pl2), p(ni), p(nd))
        q2 = np.nonzero(valid)[0]
        assert np.array_equal(q, q2)
        assert np.abs(pl - pl2[q2]).max() < 1e-12 and np.abs(pt - pt2[q2]).max() == 0.0
        idx, d2 = oracle.knn(refw, neiw, K, False)
        full = d2[:, K - 1] <= np.float32(thr) * np.float32(thr)
        assert np.array_equal(d2[full], nd[full]) and np.array_equal(idx[full], ni[full])       # bit-exact float32 distances
        assert np.all(ni[~full][:, K - 1] == -1)                                                  # fewer than K within the threshold


def test_host_lm_matches_oracle_lm(oracle, harness):
    g = np.load(os.path.join(G, "assoc_pair.npz"))
    n = len(g["query"])
    consts = np.zeros((n, 12)); consts[:, :3] = g["point"]; consts[:, 3:7] = g["plane"]; consts[:, 7] = 1.0
    for bt, hub in ((1, 2 * np.pi / 180), (0, 0.2)):
        blk = oracle.Blocks(np.full(n, bt), 0, 1, consts, hub, 1)
        P1, s1 = blk.solve_lm(np.zeros((2, 6)), is_const=[1, 0], max_iter=20)
        P2, summ, mask = np.zeros((2, 6)), np.zeros(6), np.array([1, 0], np.uint8)
        harness.pvbh_solve_lm(C.c_long(n), p(blk.type), p(blk.ref), p(blk.nei), p(blk.normalize), p(blk.huber), p(blk.consts), p(P2), C.c_int(2), p(mask), C.c_int(20), p(summ))
        assert summ[2] == s1["iterations"] and summ[3] == s1["successful"]
        assert np.abs(P1[1] - P2[1]).max() / np.abs(P1[1]).max() < 1e-8                           # gate: 1e-4 relative on pose deltas


def test_float_paths_are_bit_exact(oracle, harness):
    g = np.load(os.path.join(G, "fast_atan2.npz"))
    yf, xf = g["y"].astype(np.float32), g["x"].astype(np.float32)
    out = np.zeros_like(yf)
    harness.pvbh_fast_atan2_f(C.c_long(len(yf)), p(yf), p(xf), p(out))
    assert np.array_equal(out, oracle.fast_atan2(yf, xf))
    outd = np.zeros_like(g["y"])
    harness.pvbh_fast_atan2_d(C.c_long(len(outd)), p(np.ascontiguousarray(g["y"])), p(np.ascontiguousarray(g["x"])), p(outd))
    assert np.array_equal(outd, oracle.fast_atan2(g["y"], g["x"]))
    rng = np.random.default_rng(0)
    cam = rng.normal(0, 5, (5000, 3)).astype(np.float32)
    px = np.zeros((5000, 2), np.float32)
    harness.pvbh_cam_to_image_f(C.c_int(2880), C.c_int(5760), C.c_long(5000), p(cam), p(px))
    assert np.array_equal(px, oracle.cam_to_image(2880, 5760, cam))
    a = np.load(os.path.join(G, "assoc_pair.npz"))
    out = np.empty_like(a["ref_local"])
    harness.pvbh_transform_cloud(p(np.ascontiguousarray(a["R_ref"])), p(np.ascontiguousarray(a["t_ref"])), p(np.ascontiguousarray(a["ref_local"])), C.c_int(len(out)), p(out))
    assert np.array_equal(out, a["ref_world"])


def test_rank_deficient_neighbour_sets_follow_the_qr_basic_solution(oracle, harness):
    """Neighbours with one coordinate exactly 0 in the reference frame (a floor through the origin): Eigen's rank-revealing
    QR returns a basic solution that may still pass the tolerance test; the streaming plane fit must fall back to it."""
    from panovlm_b200 import synth
    harness.pvbh_set_prune(C.c_int(3)); harness.pvbh_set_hints(None, None)
    rng = np.random.default_rng(12)
    n = 60000
    tgt = np.concatenate([synth.sample_floor_plan(n, rng, xlim=(0, 12.0)), np.ones((n, 1))], axis=1).astype(np.float32)   # floor at z == 0, walls at x == 0 / y == 0
    qry = tgt[rng.choice(n, 3000, replace=False)].copy()
    qry[:, :3] += rng.normal(0, 0.01, (3000, 3)).astype(np.float32)
    I, z = np.eye(3), np.zeros(3)
    oq, opt, opl = oracle.associate_p2plane(tgt, I, z, qry, I, z, 0.05, 1.0, 10, True)
    m = len(qry)
    valid, pl2, pt2 = np.zeros(m, np.uint8), np.zeros((m, 4)), np.zeros((m, 3))
    ni, nd = np.zeros((m, 10), np.int32), np.zeros((m, 10), np.float32)
    harness.pvbh_associate(p(tgt), C.c_int(n), p(I.copy()), p(z), p(np.ascontiguousarray(qry)), C.c_int(m), p(I.copy()), p(z), C.c_double(0.3), C.c_float(1.0),
                           C.c_double(0.05), C.c_int(10), p(valid), p(pt2), p(pl2), p(ni), p(nd))
    q2 = np.nonzero(valid)[0]
    assert np.array_equal(oq, q2)
    degenerate = np.abs(opl[:, 2]) < 1e-12                       # bogus vertical planes through floor points (n_z == 0)
    assert degenerate.sum() > 50
    assert np.abs(opl - pl2[q2]).max() < 1e-9


def test_point2line_association_equals_oracle(oracle, harness):
    """AssociatePoint2Line (5-NN + PCA line in the world frame): same accepted queries; end points equal up to the arbitrary sign
    of the eigenvector (a <-> b swap), which leaves the line and therefore the residuals unchanged."""
    from panovlm_b200 import synth
    A, B = synth.make_pair(seed=20260925, n_az=1800)
    I, z = np.eye(3), np.zeros(3)
    refw = oracle.transform_cloud(A["R_wl"], A["t_wl"], A["cornerLessSharp"])
    neiw = oracle.transform_cloud(I, z, B["cornerLessSharp"])
    for thr, h in ((0.3, 0.3), (0.7, 0.25)):
        oq, opt, oa, ob = oracle.associate_p2line(refw, A["R_wl"], A["t_wl"], neiw, I, z, thr, True)
        m = len(neiw)
        valid, pt, pa, pb = np.zeros(m, np.uint8), np.zeros((m, 3)), np.zeros((m, 3)), np.zeros((m, 3))
        harness.pvbh_associate_lines(p(np.ascontiguousarray(refw)), C.c_int(len(refw)), p(np.ascontiguousarray(A["R_wl"])), p(np.ascontiguousarray(A["t_wl"])), p(np.ascontiguousarray(neiw)),
                                     C.c_int(m), p(I.copy()), p(z), C.c_double(h), C.c_float(thr), p(valid), p(pt), p(pa), p(pb))
        q = np.nonzero(valid)[0]
        assert np.array_equal(q, oq) and len(q) > 100
        assert np.abs(pt[q] - opt).max() == 0.0
        same = np.abs(pa[q] - oa).max(1) < 1e-9
        swapped = np.abs(pa[q] - ob).max(1) < 1e-9
        assert np.all(same | swapped)
        assert np.all(np.where(same, np.abs(pb[q] - ob).max(1), np.abs(pb[q] - oa).max(1)) < 1e-9)


def test_pruned_block_walk_is_exact_on_random_surface_clouds(oracle, harness):
    """Fuzz of the pruned walk (knn_select_pruned): planar clouds with duplicates, queries inside / outside the grid, cell sizes from far
    below to far above the k-NN radius; the neighbour lists must equal the brute-force search bit for bit."""
    I, z = np.eye(3), np.zeros(3)
    rng = np.random.default_rng(21)
    harness.pvbh_set_hints(None, None)
    for trial in range(40):
        harness.pvbh_set_prune(C.c_int(1 + trial % 4))
        harness.pvbh_set_static(C.c_int((trial // 4) % 2))                              # mode 4 with / without the static bound of the target
        n = int(rng.integers(400, 3000))
        uv = rng.uniform(-2, 2, (n, 2))
        which = rng.integers(0, 3, n)
        pts = np.zeros((n, 4), np.float32)
        pts[:, 0] = np.where(which == 0, uv[:, 0], np.where(which == 1, 1.5, uv[:, 0]))
        pts[:, 1] = np.where(which == 0, uv[:, 1], np.where(which == 1, uv[:, 0], -0.7))
        pts[:, 2] = np.where(which == 0, 0.3, uv[:, 1])
        pts[:, :3] += rng.normal(0, 0.005, (n, 3)).astype(np.float32)
        pts[: n // 10, :3] = pts[n // 10: 2 * (n // 10), :3][: n // 10]                # exact duplicates => distance ties
        m = 300
        qry = np.zeros((m, 4), np.float32)
        qry[:, :3] = pts[rng.integers(0, n, m), :3] + rng.normal(0, 0.02, (m, 3)).astype(np.float32)
        qry[:20, :3] += rng.normal(0, 3.0, (20, 3)).astype(np.float32)                # far outside the grid: clamped cell coordinates
        K = int(rng.choice([5, 10])); h = float(rng.choice([0.03, 0.08, 0.2, 0.6])); thr = float(rng.choice([0.25, 1.0]))
        valid, pl2, pt2 = np.zeros(m, np.uint8), np.zeros((m, 4)), np.zeros((m, 3))
        ni, nd = np.zeros((m, K), np.int32), np.zeros((m, K), np.float32)
        harness.pvbh_associate(p(pts), C.c_int(n), p(I.copy()), p(z), p(qry), C.c_int(m), p(I.copy()), p(z), C.c_double(h), C.c_float(thr),
                               C.c_double(0.05), C.c_int(K), p(valid), p(pt2), p(pl2), p(ni), p(nd))
        idx, d2 = oracle.knn(pts, qry, K, False)
        full = d2[:, K - 1] <= np.float32(thr) * np.float32(thr)
        assert np.array_equal(d2[full], nd[full]), (trial, h, K)
        # among exact ties the choice of equal-distance points is free; everything strictly inside tau must match
        for r in np.nonzero(full)[0]:
            inside = d2[r] < d2[r, K - 1]
            assert set(idx[r][inside]) <= set(ni[r]), (trial, r)
        assert np.all(ni[~full][:, K - 1] == -1)


def _fuzz_cloud(rng, n):
    uv = rng.uniform(-2, 2, (n, 2))
    which = rng.integers(0, 3, n)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = np.where(which == 0, uv[:, 0], np.where(which == 1, 1.5, uv[:, 0]))
    pts[:, 1] = np.where(which == 0, uv[:, 1], np.where(which == 1, uv[:, 0], -0.7))
    pts[:, 2] = np.where(which == 0, 0.3, uv[:, 1])
    pts[:, :3] += rng.normal(0, 0.005, (n, 3)).astype(np.float32)
    pts[: n // 10, :3] = pts[n // 10: 2 * (n // 10), :3][: n // 10]                # exact duplicates => distance ties
    return pts


def test_search_radius_hints_never_change_the_result(oracle, harness):
    """The buffered search with hints: (a) hints written by a first evaluation and used after the queries moved (small and large moves),
    (b) wrong hints (far too small radii, positions elsewhere) - the neighbour lists must equal the brute-force search bit for bit in
    every case, because a hint only bounds the walk and a search that finds fewer than K below it starts again without it."""
    I, z = np.eye(3), np.zeros(3)
    rng = np.random.default_rng(5)
    for trial in range(16):
        harness.pvbh_set_prune(C.c_int(3 + (trial // 2) % 2))                              # per-row walk / merged super-rows
        harness.pvbh_set_flat(C.c_int(trial % 2))                                    # flattened / nested hinted walk
        n = int(rng.integers(1500, 6000))
        pts = _fuzz_cloud(rng, n)
        m = 400
        qry = np.zeros((m, 4), np.float32)
        qry[:, :3] = pts[rng.integers(0, n, m), :3] + rng.normal(0, 0.02, (m, 3)).astype(np.float32)
        K = int(rng.choice([5, 10])); h = float(rng.choice([0.05, 0.1, 0.2, 0.5])); thr = float(rng.choice([0.25, 1.0]))

        def run(q, hint_in, hint_out):
            valid, pl2, pt2 = np.zeros(m, np.uint8), np.zeros((m, 4)), np.zeros((m, 3))
            ni, nd = np.zeros((m, K), np.int32), np.zeros((m, K), np.float32)
            harness.pvbh_set_hints(p(hint_in), p(hint_out))
            harness.pvbh_associate(p(pts), C.c_int(n), p(I.copy()), p(z), p(q), C.c_int(m), p(I.copy()), p(z), C.c_double(h), C.c_float(thr),
                                   C.c_double(0.05), C.c_int(K), p(valid), p(pt2), p(pl2), p(ni), p(nd))
            harness.pvbh_set_hints(None, None)
            return valid, pl2, ni, nd

        def check(q, ni, nd):
            idx, d2 = oracle.knn(pts, q, K, False)
            full = d2[:, K - 1] <= np.float32(thr) * np.float32(thr)
            assert np.array_equal(d2[full], nd[full]), (trial, h, K)
            for r in np.nonzero(full)[0]:
                inside = d2[r] < d2[r, K - 1]
                assert set(idx[r][inside]) <= set(ni[r]), (trial, r)
            assert np.all(ni[~full][:, K - 1] == -1)

        hints = np.zeros((m, 4), np.float32)
        v0, pl0, ni0, nd0 = run(qry, None, hints)
        check(qry, ni0, nd0)
        full = np.isfinite(hints[:, 3])
        assert np.array_equal(hints[full, 3], nd0[full, K - 1]) and np.array_equal(hints[:, :3], qry[:, :3])
        for move in (0.0, 0.003, 0.05, 0.4):                                        # the same queries after a pose update
            q2 = qry.copy(); q2[:, :3] += rng.normal(0, move, (m, 3)).astype(np.float32) if move > 0 else 0
            h2 = np.zeros((m, 4), np.float32)
            v_h, pl_h, ni_h, nd_h = run(q2, hints, h2)
            v_n, pl_n, ni_n, nd_n = run(q2, None, None)
            check(q2, ni_h, nd_h)
            assert np.array_equal(v_h, v_n) and np.array_equal(pl_h, pl_n) and np.array_equal(ni_h, ni_n) and np.array_equal(nd_h, nd_n)   # hinted == unhinted, bit for bit
        bad = hints.copy()
        bad[:, 3] = rng.uniform(0, 1e-4, m).astype(np.float32)                        # radii far too small
        bad[::3, :3] = qry[::3, :3]                                                   # ... some of them at the right place
        v_b, pl_b, ni_b, nd_b = run(qry, bad, None)
        assert np.array_equal(v_b, v0) and np.array_equal(pl_b, pl0) and np.array_equal(ni_b, ni0) and np.array_equal(nd_b, nd0)


def test_product_fast_atan2_equals_the_reference_code_outputs(harness):
    """The product's fast_atan2_f32 / fast_atan2_f64 (compiled for the host) against the outputs of the reference's own base/Math.h
    (tests/golden/ref_fast_atan2.npz, produced through oracle/_ref): bit for bit."""
    g = np.load(os.path.join(G, "ref_fast_atan2.npz"))
    yf, xf = np.ascontiguousarray(g["yf"]), np.ascontiguousarray(g["xf"])
    out = np.zeros_like(yf)
    harness.pvbh_fast_atan2_f(C.c_long(len(yf)), p(yf), p(xf), p(out))
    assert np.array_equal(out, g["out_f32"])
    y, x = np.ascontiguousarray(g["y"]), np.ascontiguousarray(g["x"])
    outd = np.zeros_like(y)
    harness.pvbh_fast_atan2_d(C.c_long(len(y)), p(y), p(x), p(outd))
    assert np.array_equal(outd, g["out_f64"])
