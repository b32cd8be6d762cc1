"""Host-side builders of the library (pvb_find_neighbors, pvb_build_*_blocks): pure C++ host code, no GPU needed.
Expected values are rebuilt here in Python from the reference's description with the oracle's primitives."""
import numpy as np
import pytest

import panovlm_b200
from panovlm_b200 import BlockList, Context, LineFrame, synth


def py_find_neighbors(t_wl, pose_valid, frame_valid, k):
    """lidar_mapping/LidarFeatureAssociate.cpp:19-111 restated in Python (float32 centres, brute-force k-NN)."""
    n = len(t_wl)
    cf = [i for i in range(n) if pose_valid[i] and frame_valid[i]]
    c = t_wl[cf].astype(np.float32)
    out = []
    for i in range(n):
        if not pose_valid[i]:
            out.append([i - j for j in range(-(k // 2), k // 2 + 1)])
            continue
        q = t_wl[i].astype(np.float32)
        d = ((q - c) ** 2).astype(np.float32)
        d2 = (d[:, 0] + d[:, 1]) + d[:, 2]
        order = sorted(range(len(cf)), key=lambda j: (d2[j], j))
        nb = [cf[j] for j in order[:k]][1:]
        s = set(nb)
        p = i - 1
        while p >= 0 and not pose_valid[p]:
            p -= 1
        if p >= 0 and p not in s:
            nb.append(p)
        p = i + 1
        while p < n and not pose_valid[p]:
            p += 1
        if p < n and p not in s:
            nb.append(p)
        for j in order:
            if not d2[j] < np.float32(400.0):
                break
            cand, same = cf[j], 0
            for it in sorted(s):
                if abs(cand - it) <= 200:
                    same += 1
                if same >= 2:
                    break
            if same < 2 and cand not in s:
                nb.append(cand); s.add(cand)
        out.append(nb)
    return out


def test_find_neighbors_matches_reference_rules():
    rng = np.random.default_rng(0)
    n = 700                                                     # long enough for loop closures (> 200 frames apart)
    a = np.linspace(0, 4 * np.pi, n)
    t = np.stack([30 * np.cos(a), rng.normal(0, 0.05, n), 30 * np.sin(a)], axis=1) + rng.normal(0, 0.2, (n, 3))
    pv = np.ones(n, np.uint8); pv[[5, 6, 300]] = 0
    fv = np.ones(n, np.uint8); fv[[10, 11]] = 0
    got = Context.find_neighbors(t, pv, fv, 6)
    exp = py_find_neighbors(t, pv, fv, 6)
    assert got == exp
    assert any(abs(i - j) > 200 for i, nb in enumerate(got) for j in nb if pv[i])   # loop closures are found
    assert got[5] == [5 - j for j in range(-3, 4)]


def test_point2plane_and_line_blocks(oracle):
    A, B = synth.make_pair(seed=3, n_az=900)
    bl = BlockList(10000)
    pts, pls = np.random.default_rng(1).normal(size=(7, 3)), np.random.default_rng(2).normal(size=(7, 4))
    Context.build_point2plane_blocks(bl, pts, pls, 2, 5, True, True, 0.01)
    Context.build_point2plane_blocks(bl, pts, pls, 2, 5, False, False, 0.1)
    v = bl.view()
    assert np.all(v["type"][:7] == 1) and np.all(v["type"][7:] == 0) and np.all(v["ref"] == 2) and np.all(v["nei"] == 5)
    assert np.allclose(v["huber"][:7], 2 * np.pi / 180) and np.allclose(v["huber"][7:], 0.2)
    assert np.array_equal(v["consts"][:7, :3], pts) and np.array_equal(v["consts"][:7, 3:7], pls) and np.all(v["consts"][7:, 7] == 0.1)
    assert np.all(v["normalize"][:7] == 1) and np.all(v["normalize"][7:] == 0)
    # line-to-line: one block per point of the neighbour segment, point taken from the float32 WORLD cloud back to the sensor frame
    fr = LineFrame(B["cornerLessSharp"], B["p2s_off"], B["p2s_ids"], B["segment_coeffs"], B["end_points"], B["R_wl"], B["t_wl"])
    world = oracle.transform_cloud(B["R_wl"], B["t_wl"], B["cornerLessSharp"])
    a, b = np.array([1.0, 2.0, 3.0]), np.array([1.1, 2.0, 2.5])
    bl2 = BlockList(1000)
    Context.build_line2line_blocks(bl2, fr, world, 4, a, b, 0, 1, True, True, 1.0)
    v = bl2.view()
    members = [i for i in range(len(world)) if 4 in B["p2s_ids"][B["p2s_off"][i]:B["p2s_off"][i + 1]]]
    assert len(v["type"]) == len(members) > 4 and np.all(v["type"] == 3) and np.all(v["huber"] == 0.0)     # angle line residuals: loss == nullptr
    pl = oracle.world2local(B["R_wl"], B["t_wl"], world[members, :3].astype(np.float64))
    assert np.array_equal(v["consts"][:, :3], pl)
    d = (a - b) / np.linalg.norm(a - b)
    assert np.allclose(v["consts"][:, 3:6], a) and np.abs(v["consts"][:, 6:9] - d).max() < 1e-15
    Context.build_line2line_blocks(bl2, fr, world, 4, a, b, 0, 1, False, True, 0.5)
    v = bl2.view()
    assert np.all(v["type"][len(members):] == 2) and np.all(v["huber"][len(members):] == 0.2) and np.all(v["consts"][len(members):, 9] == 0.5)


def test_point2plane_blocks_of_many_edges_equal_the_per_edge_loop():
    """pvb_build_point2plane_blocks_edges == AddLidarPointToPlaneResidual's loop over (i, n) pairs (Optimization.cpp:520-557), one call for all edges."""
    rng = np.random.default_rng(5)
    n_edges = 9
    edge_ref, edge_nei = rng.integers(0, 6, n_edges).astype(np.int32), rng.integers(0, 6, n_edges).astype(np.int32)
    counts = rng.integers(0, 40, n_edges)
    counts[3] = 0                                                   # an edge without correspondences
    edge = np.repeat(np.arange(n_edges), counts).astype(np.int32)
    pts, pls = rng.normal(size=(len(edge), 3)), rng.normal(size=(len(edge), 4))
    for angle, norm, w in ((True, True, 1.0), (False, False, 0.3)):
        one = BlockList(len(edge) + 4)
        Context.build_point2plane_blocks_edges(one, edge, pts, pls, edge_ref, edge_nei, angle, norm, w)
        loop = BlockList(len(edge) + 4)
        lo = 0
        for e in range(n_edges):
            hi = lo + counts[e]
            if hi > lo:
                Context.build_point2plane_blocks(loop, pts[lo:hi], pls[lo:hi], int(edge_ref[e]), int(edge_nei[e]), angle, norm, w)
            lo = hi
        a, b = one.view(), loop.view()
        assert one.n == loop.n == len(edge)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    # an edge index outside the table is refused, and so is a full list
    with pytest.raises(Exception):
        Context.build_point2plane_blocks_edges(BlockList(100), np.array([n_edges], np.int32), pts[:1], pls[:1], edge_ref, edge_nei, True, True, 1.0)
    with pytest.raises(Exception):
        Context.build_point2plane_blocks_edges(BlockList(3), edge, pts, pls, edge_ref, edge_nei, True, True, 1.0)


def test_camera_lidar_blocks_evaluate_like_the_reference_construction(oracle):
    rows, cols = 2880, 5760
    rng = np.random.default_rng(5)
    lines = np.stack([rng.uniform(0, cols, 12), rng.uniform(0, rows, 12), rng.uniform(0, cols, 12), rng.uniform(0, rows, 12)], axis=1).astype(np.float32)
    start, end = rng.normal(0, 3, (12, 3)), rng.normal(0, 3, (12, 3))
    pw = rng.uniform(0.5, 2, 12).astype(np.float32)
    bl = BlockList(100)
    Context.build_camera_lidar_blocks(bl, rows, cols, lines, start, end, pw, 0, 1, 25.0)
    v = bl.view()
    assert len(v["type"]) == 24 and np.all(v["type"][0::2] == 4) and np.all(v["type"][1::2] == 5) and np.allclose(v["huber"], 3 * np.pi / 180)
    p1 = oracle.image_to_cam(rows, cols, lines[:, :2].astype(np.float64))
    p2 = oracle.image_to_cam(rows, cols, lines[:, 2:].astype(np.float64))
    nrm = np.cross(p2 - p1, -p1)                                   # FormPlane(p1, p2, 0)
    d = -(nrm * p1).sum(1)
    nn = np.linalg.norm(nrm, axis=1, keepdims=True)
    c1, c2 = v["consts"][0::2], v["consts"][1::2]
    assert np.abs(c1[:, :3] - nrm / nn).max() < 1e-12 and np.array_equal(c1[:, 3:6], end) and np.array_equal(c1[:, 6:9], start)
    assert np.allclose(c1[:, 9], pw.astype(np.float64) * 25.0)
    assert np.abs(c2[:, :3] - nrm / nn).max() < 1e-12 and np.abs(c2[:, 3] - d / nn[:, 0]).max() < 1e-12
    assert np.allclose(c2[:, 4:7], (end + start) / 2) and np.allclose(c2[:, 7:10], (p1 + p2) / 2)
    ang = np.arccos(np.clip((p1 * p2).sum(1), -1, 1))
    assert np.abs(c2[:, 10] - ang).max() < 1e-12 and np.all(c2[:, 11] == 50.0)
    # the blocks evaluate (oracle functors) to finite residuals with the camera / LiDAR pose blocks
    blk = oracle.Blocks(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"])
    r, J, _ = blk.evaluate(np.concatenate([rng.normal(0, 0.1, (2, 3)), rng.normal(0, 0.5, (2, 3))], axis=1))
    assert np.all(np.isfinite(r)) and np.all(np.isfinite(J))


def _seg_pair(n_az=1800):
    A, B = synth.make_pair(seed=20260925, n_az=n_az)
    fa = LineFrame(A["cornerLessSharp"], A["p2s_off"], A["p2s_ids"], A["segment_coeffs"], A["end_points"], A["R_wl"], A["t_wl"])
    fb = LineFrame(B["cornerLessSharp"], B["p2s_off"], B["p2s_ids"], B["segment_coeffs"], B["end_points"], np.eye(3), np.zeros(3))
    return A, B, fa, fb


def test_segment_knn_tails_match_oracle(oracle):
    """pvb_point2line_segment_knn_tail / pvb_line2line_knn_tail (host tails of LidarFeatureAssociate.cpp:238-317, 385-440) fed with the oracle's own
    5-NN: membership counting, the all-5 / >= 3 rules, World2Local and FindAssociations must reproduce the oracle's associations."""
    A, B, fa, fb = _seg_pair()
    ref_w = oracle.transform_cloud(A["R_wl"], A["t_wl"], A["cornerLessSharp"])
    nei_w = oracle.transform_cloud(np.eye(3), np.zeros(3), B["cornerLessSharp"])
    for thr in (0.3, 0.6):
        idx, d2 = oracle.knn(ref_w, nei_w, 5, True)
        idx = idx.copy(); idx[d2[:, 4] > np.float32(thr) * np.float32(thr)] = -1
        q, ln, pt, a, b = Context.point2line_segment_knn_tail(fa, fb, idx)
        oq, oln, opt, oa, ob = oracle.associate_p2line_segment_knn(ref_w, A["p2s_off"], A["p2s_ids"], A["segment_coeffs"], nei_w, np.eye(3), np.zeros(3), thr)
        assert len(oq) > 20
        assert np.array_equal(q, oq) and np.array_equal(ln, oln) and np.array_equal(pt, opt) and np.array_equal(a, oa) and np.array_equal(b, ob)
        nl, rl, a2, b2 = Context.line2line_knn_tail(fa, fb, idx)
        M = oracle.line2line_knn_votes(ref_w, A["p2s_off"], A["p2s_ids"], len(A["segment_coeffs"]), nei_w, B["p2s_off"], B["p2s_ids"], len(B["segment_coeffs"]), thr)
        ref_lw = oracle.transform_lines(A["R_wl"], A["t_wl"], A["segment_coeffs"]); nei_lw = oracle.transform_lines(np.eye(3), np.zeros(3), B["segment_coeffs"])
        on, orf, oa2, ob2 = oracle.find_associations(A["segment_coeffs"], ref_lw, nei_lw, np.diff(B["seg_off"]), M)
        assert len(on) >= 3 and M.sum() > 50
        assert np.array_equal(nl, on) and np.array_equal(rl, orf) and np.array_equal(a2, oa2) and np.array_equal(b2, ob2)
    # nothing within reach => no association; a frame without segments => nothing (CheckLidarSegment)
    q, *_ = Context.point2line_segment_knn_tail(fa, fb, np.full((len(B["cornerLessSharp"]), 5), -1, np.int32))
    assert len(q) == 0
    empty = LineFrame(B["cornerLessSharp"], np.zeros(len(B["cornerLessSharp"]) + 1, np.int32), np.zeros(0, np.int32), np.zeros((0, 6)), None, np.eye(3), np.zeros(3))
    assert len(Context.line2line_knn_tail(fa, empty, idx)[0]) == 0


def _random_matches(rng, n_frames, n_lines, n_pairs, p_match):
    pa, pb, off, ma, mb = [], [], [0], [], []
    for _ in range(n_pairs):
        a, b = rng.choice(n_frames, 2, replace=False)
        m = {(int(x), int(y)) for x, y in zip(rng.integers(0, n_lines, 6), rng.integers(0, n_lines, 6)) if rng.random() < p_match}
        pa.append(a); pb.append(b)
        for x, y in sorted(m):
            ma.append(x); mb.append(y)
        off.append(len(ma))
    return np.array(pa, np.int32), np.array(pb, np.int32), np.array(off, np.int32), np.array(ma, np.int32), np.array(mb, np.int32)


def test_line_tracks_match_oracle_and_connected_components(oracle):
    """pvb_line_tracks_build / pvb_line_tracks_gate vs the oracle's restatement of TrackBuilder (util/Tracks.cpp:58-186) and vs scipy's connected
    components (independent pin: a track = a component spanning >= min_length distinct frames, numbered by its smallest feature)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    rng = np.random.default_rng(5)
    for trial in range(30):
        nfr, nl = int(rng.integers(3, 12)), int(rng.integers(2, 9))
        pa, pb, off, ma, mb = _random_matches(rng, nfr, nl, int(rng.integers(1, 25)), rng.uniform(0.2, 0.9))
        for min_len, multi in ((3, True), (2, True), (2, False), (1, True)):
            got = Context.line_tracks_build(pa, pb, off, ma, mb, min_len, multi)
            exp = oracle.line_tracks(pa, pb, off, ma, mb, min_len, multi)
            assert len(got) == len(exp) and all(np.array_equal(g, e) for g, e in zip(got, exp))
            # independent: connected components over feature ids frame * nl + line
            if len(ma) == 0:
                assert got == []
                continue
            u = np.concatenate([pa[p] * nl + ma[off[p]:off[p + 1]] for p in range(len(pa))]); v = np.concatenate([pb[p] * nl + mb[off[p]:off[p + 1]] for p in range(len(pa))])
            ncomp, lab = connected_components(coo_matrix((np.ones(len(u)), (u, v)), shape=(nfr * nl, nfr * nl)), directed=False)
            used = np.unique(np.concatenate([u, v]))
            comps = {}
            for f in used:
                comps.setdefault(lab[f], []).append((f // nl, f % nl))
            want = []
            for feats in sorted(comps.values(), key=lambda fs: min(fs)):
                frames = [f for f, _ in feats]
                if len(set(frames)) < min_len or (not multi and len(set(frames)) != len(frames)) or len(feats) < 2:
                    continue
                want.append(np.array(sorted(feats), np.int32))
            assert len(got) == len(want) and all(np.array_equal(g, w) for g, w in zip(got, want))
            # gate: random candidate associations between two frames
            rf, nf_ = int(rng.integers(0, nfr)), int(rng.integers(0, nfr))
            rl, nln = rng.integers(0, nl, 20), rng.integers(0, nl, 20)
            keep = Context.line_tracks_gate(got, rf, nf_, rl, nln)
            assert np.array_equal(keep, oracle.line_track_gate(exp, rf, nf_, rl, nln))
            tr = {(int(f), int(l)): t for t, fs in enumerate(got) for f, l in fs}
            assert np.array_equal(keep, [tr.get((rf, int(a)), -1) == tr.get((nf_, int(b)), -2) for a, b in zip(rl, nln)])
    assert Context.line_tracks_build(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32)) == []


def test_unique_line_pairs_matches_oracle_and_is_one_to_one(oracle):
    """pvb_unique_line_pairs vs the oracle's restatement of UniqueLinePair (CameraLidarLineAssociate.cpp:754-876), incl. score ties (strict
    inequalities: a tie never displaces) and the three-way cases."""
    rng = np.random.default_rng(12)
    for trial in range(200):
        n = int(rng.integers(0, 40))
        il, ll = rng.integers(0, 8, n), rng.integers(0, 8, n)
        keep = np.unique(np.stack([il, ll], 1), axis=0, return_index=True)[1] if n else np.zeros(0, int)
        keep.sort()
        il, ll = il[keep], ll[keep]
        sc = (rng.integers(0, 12, len(il)) * np.float32(0.01)).astype(np.float32)          # coarse scores => many ties
        got = Context.unique_line_pairs(il, ll, sc)
        exp = oracle.unique_line_pairs(il, ll, sc)
        assert all(np.array_equal(g, e) for g, e in zip(got, exp))
        assert len(set(got[0].tolist())) == len(got[0]) and len(set(got[1].tolist())) == len(got[1])
        for i, l, s_ in zip(*got):                                                         # every surviving pair is one of the candidates
            assert np.any((il == i) & (ll == l) & (sc == s_))
    # hand cases from the reference's comment (:812-820): L1-A 0.3, L2-B 0.5, then L1-B
    assert [x.tolist() for x in Context.unique_line_pairs([1, 2, 1], [0, 1, 1], [0.3, 0.5, 0.1])[:2]] == [[1], [1]]            # case 1: replaces both
    assert [x.tolist() for x in Context.unique_line_pairs([1, 2, 1], [0, 1, 1], [0.3, 0.5, 0.4])[:2]] == [[1], [0]]            # case 2: L2-B dropped
    assert [x.tolist() for x in Context.unique_line_pairs([1, 2, 1], [0, 1, 1], [0.5, 0.3, 0.4])[:2]] == [[2], [1]]            # case 3: L1-A dropped
    assert [x.tolist() for x in Context.unique_line_pairs([1, 2, 1], [0, 1, 1], [0.3, 0.5, 0.9])[:2]] == [[1, 2], [0, 1]]      # case 4: unchanged


def test_neighbor_each_frame_and_lidar_mask_match_restatement():
    """NeighborEachFrame / LidarMaskByTrack (joint_optimization/CameraLidarOptimizer.cpp:551-642) against a Python restatement."""
    rng = np.random.default_rng(8)
    # temporal windows (:556-567)
    for n_frames, n_lidars, size in ((10, 10, 1), (10, 10, 4), (7, 5, 3), (3, 2, 6), (0, 0, 2)):
        got = Context.neighbor_each_frame(n_frames, n_lidars, size, True)
        for f in range(n_frames):
            lo = max(0, f - size // 2); hi = min(n_lidars, lo + size); lo = max(0, hi - size)
            assert got[f] == list(range(lo, hi))
    # spatial: float32 k-NN over the usable LiDAR centres + previous / next index (:570-606)
    n = 40
    t_wl = np.cumsum(rng.normal(0, 0.5, (n, 3)), axis=0)
    t_wc = t_wl + rng.normal(0, 0.05, (n, 3))
    lpv, lv, fpv = (rng.random(n) > 0.15).astype(np.uint8), (rng.random(n) > 0.1).astype(np.uint8), (rng.random(n) > 0.1).astype(np.uint8)
    got = Context.neighbor_each_frame(n, n, 5, False, t_wc, fpv, t_wl, lpv, lv)
    usable = [i for i in range(n) if lpv[i] and lv[i]]
    C32 = t_wl[usable].astype(np.float32)
    for f in range(n):
        if not fpv[f]:
            assert got[f] == []
            continue
        q = t_wc[f].astype(np.float32)
        d = (C32 - q) ** 2
        d2 = (d[:, 0] + d[:, 1]) + d[:, 2]
        order = np.lexsort((np.arange(len(usable)), d2))[:5]
        exp = [usable[j] for j in order]
        have = set(exp)
        if f - 1 >= 0 and f - 1 not in have:
            exp.append(f - 1)
        if f + 1 < n and f + 1 not in have:
            exp.append(f + 1)
        assert got[f] == exp
    # LidarMaskByTrack: lines that belong to a track are switched on
    tracks = [np.array([[0, 2], [1, 0], [3, 1]]), np.array([[1, 3], [2, 0]])]
    masks = Context.lidar_mask_by_track(tracks, [4, 5, 1, 2])
    assert [m.tolist() for m in masks] == [[False, False, True, False], [True, False, False, True, False], [True], [False, True]]
    assert [m.tolist() for m in Context.lidar_mask_by_track([], [2, 0])] == [[False, False], []]
    with pytest.raises(Exception):
        Context.lidar_mask_by_track([np.array([[0, 9]])], [4])


def test_calibration_blocks_match_oracle(oracle):
    """pvb_build_calibration_blocks == the block construction of the calibration-mode Optimize (CameraLidarOptimizer.cpp:32-64)."""
    rng = np.random.default_rng(9)
    n = 40
    lines = np.stack([rng.uniform(0, 5760, n), rng.uniform(200, 2600, n), rng.uniform(0, 5760, n), rng.uniform(200, 2600, n)], axis=1).astype(np.float32)
    start, end = rng.normal(0, 3, (n, 3)), rng.normal(0, 3, (n, 3))
    typ, hub, consts = oracle.build_calibration_blocks(2880, 5760, lines, start, end)
    bl = BlockList(2 * n + 2)
    Context.build_calibration_blocks(bl, 2880, 5760, lines, start, end, 0)
    v = bl.view()
    assert bl.n == 2 * n
    assert np.array_equal(v["type"], typ) and np.array_equal(v["huber"], hub) and np.all(v["ref"] == 0) and np.all(v["nei"] == 0)
    assert np.array_equal(v["consts"], consts)
    # geometric meaning: the plane normal is perpendicular to both image rays, the half arc is half the angle between them
    p = oracle.image_to_cam(2880, 5760, lines.reshape(-1, 2).astype(np.float64)).reshape(n, 2, 3) if hasattr(oracle, "image_to_cam") else None
    if p is not None:
        assert np.abs(np.sum(consts[0::2, :3] * p[:, 0], 1)).max() < 1e-5 and np.abs(np.sum(consts[0::2, :3] * p[:, 1], 1)).max() < 1e-5
    assert np.all(consts[0::2, 9] == 1.0) and np.all(consts[1::2, 11] == 2.0) and np.all(consts[1::2, 3] == 0.0)
    assert np.array_equal(consts[0::2, 3:6], end) and np.array_equal(consts[0::2, 6:9], start)
    with pytest.raises(Exception):
        Context.build_calibration_blocks(BlockList(3), 2880, 5760, lines, start, end, 0)


def test_refine_pose_holds_the_first_valid_frame_constant():
    """LidarOdometry.cpp:59-66: the constant pose blocks are those of the first frame with a valid pose AND valid data, not of frame 0."""
    from panovlm_b200 import odometry
    assert odometry.first_valid_frame([{}, {}, {}]) == 0
    assert odometry.first_valid_frame([{"valid": False}, {"pose_valid": False}, {"valid": True}, {}]) == 2
    assert odometry.first_valid_frame([{"valid": False}]) == 0
