"""pvb_pixel_fit_line (host): FitLineRANSAC + the end-point tail of the pixel-space Associate (joint_optimization/CameraLidarLineAssociate.cpp:105-144, 717-752).

PARITY UNPINNED for the sample-consensus part: the reference calls pcl::SACSegmentation and PCL is neither in the reference tree nor installed here, so there is no
reference output to compare with.  What is checked: (1) the C++ against an independently written numpy float32 restatement of the same published algorithm
(PCL 1.10 RANSAC over SACMODEL_LINE with boost's mt19937 / uniform_int(0, INT_MAX); numpy's legacy RandomState uses the same init_genrand seeding),
bit for bit; (2) properties: every inlier within the threshold of the sampled model, the planted line recovered, determinism, the `false` cases
(< 3 inliers, < 2 points, all points identical); (3) the reference's own tail - the inlier POSITIONS indexing the candidate list (:136-137) and
ProjectPoint2Line3D in double (base/Geometry.hpp:151-161)."""
import numpy as np
import pytest

from panovlm_b200 import Context

f32 = np.float32


def _sqdist_f32(a, d, p):
    v = (a - p).astype(f32)
    c = np.array([f32(v[1] * d[2]) - f32(v[2] * d[1]), f32(v[2] * d[0]) - f32(v[0] * d[2]), f32(v[0] * d[1]) - f32(v[1] * d[0])], f32)
    return f32(f32(f32(c[0] * c[0]) + f32(c[1] * c[1])) + f32(c[2] * c[2]))


def _normalize_f32(v):
    z = f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2]))
    return (v / np.sqrt(z)).astype(f32) if z > 0 else v


def twin_ransac(P, thr=0.1, max_it=50, prob=0.99):
    """numpy restatement of RandomSampleConsensus::computeModel + selectWithinDistance for a line model; returns (model6 float32 | None, inlier indices)."""
    n = len(P)
    if n < 2:
        return None, []
    rs = np.random.RandomState(12345)
    rnd = lambda: int(rs.randint(0, 2 ** 32, dtype=np.uint32)) >> 1  # noqa: E731
    sh = list(range(n))
    k, it, skipped, best, model_best = np.inf, 0, 0, -1, None
    while it < k and skipped < max_it * 10:
        good = False
        for _ in range(1000):
            for i in range(2):
                j = i + rnd() % (n - i)
                sh[i], sh[j] = sh[j], sh[i]
            if np.any(P[sh[0]] != P[sh[1]]):
                good = True
                break
        if not good:
            break
        a, b = P[sh[0]], P[sh[1]]
        if np.all(np.abs(a - b) <= np.finfo(f32).eps):
            skipped += 1
            continue
        d = _normalize_f32((b - a).astype(f32))
        d2 = _normalize_f32(d)
        cnt = sum(float(_sqdist_f32(a, d2, p)) < thr * thr for p in P)
        if cnt > best:
            best, model_best = cnt, np.concatenate([a, d]).astype(f32)
            w = best / n
            pno = min(1 - np.finfo(float).eps, max(np.finfo(float).eps, 1.0 - w * w))
            k = np.log(1 - prob) / np.log(pno)
        it += 1
        if it > max_it:
            break
    if model_best is None:
        return None, []
    d2 = _normalize_f32(model_best[3:])
    return model_best, [i for i, p in enumerate(P) if float(_sqdist_f32(model_best[:3], d2, p)) < thr * thr]


def twin_tail(P, inl):
    """FitLineRANSAC :731-749 + Associate :117-137 on a given inlier list (float32 accumulation in index order; float64 eigen solve for the direction)."""
    Q = P[inl]
    cen = np.zeros(3, f32)
    for q in Q:
        cen = (cen + q).astype(f32)
    cen = (cen / f32(len(Q))).astype(f32)
    D = (Q - cen).astype(np.float64)
    w, V = np.linalg.eigh(D.T @ D)
    best, s, e = -1.0, 0, 0
    for i in range(len(Q)):
        for j in range(i + 1, len(Q)):
            dd = (Q[i] - Q[j]).astype(f32)
            d = f32(f32(f32(dd[0] * dd[0]) + f32(dd[1] * dd[1])) + f32(dd[2] * dd[2]))
            if d > best:
                best, s, e = d, i, j
    return cen, V[:, 2], s, e


def _cloud(seed, n_line=40, n_out=12, noise=0.01, dup=True):
    rng = np.random.default_rng(seed)
    a, d = rng.uniform(-3, 3, 3), rng.normal(size=3)
    d /= np.linalg.norm(d)
    t = rng.uniform(-2, 2, n_line)
    P = np.concatenate([a + t[:, None] * d + rng.normal(0, noise, (n_line, 3)), rng.uniform(-4, 6, (n_out, 3))]).astype(f32)
    if dup:                                    # the candidate lists of the first stage hold a point once per neighbouring sub-line: duplicates are normal
        P = np.concatenate([P, P[rng.integers(0, len(P), 9)]])
        P = P[rng.permutation(len(P))]
    return P, a, d


@pytest.mark.parametrize("seed", range(6))
def test_fit_line_equals_the_numpy_restatement(seed):
    P, a, d = _cloud(seed)
    coeff, inl, start, end = Context.pixel_fit_line(P)
    model, inl_t = twin_ransac(P)
    assert np.array_equal(inl, inl_t) and len(inl) >= 40
    cen, dir64, s, e = twin_tail(P, inl_t)
    assert np.array_equal(coeff[:3], cen)                                                    # float32 centroid, same accumulation order
    assert abs(abs(np.dot(coeff[3:].astype(np.float64), dir64)) - 1) < 1e-5                  # closed-form float32 eigenvector vs LAPACK in double
    assert abs(abs(np.dot(dir64, d)) - 1) < 1e-3                                             # the planted line
    c = coeff.astype(np.float64)
    for got, pos in ((start, s), (end, e)):
        p = P[pos].astype(np.float64)                                                        # POSITION in the inlier list used as the index (:136-137)
        k = np.dot(c[3:], p - c[:3]) / np.dot(c[3:], c[3:])
        assert np.abs(got - (c[:3] + k * c[3:])).max() < 1e-12
    again = Context.pixel_fit_line(P)
    assert all(np.array_equal(x, y) for x, y in zip((coeff, inl, start, end), again))        # fixed seed: deterministic


def test_inliers_are_within_the_threshold_of_a_sampled_pair():
    P, _, _ = _cloud(11, dup=False)
    model, inl_t = twin_ransac(P, thr=0.05)
    _, inl, _, _ = Context.pixel_fit_line(P, dist_threshold=0.05)
    assert np.array_equal(inl, inl_t)
    a, d = model[:3].astype(np.float64), model[3:].astype(np.float64)
    dist = np.linalg.norm(np.cross(P[inl].astype(np.float64) - a, d), axis=1)
    assert dist.max() < 0.05 + 1e-6
    assert any(np.array_equal(model[:3], p) for p in P)                                      # the model point is one of the input points


def test_no_line_cases():
    assert Context.pixel_fit_line(np.zeros((0, 3), f32)) is None
    assert Context.pixel_fit_line(np.ones((1, 3), f32)) is None
    assert Context.pixel_fit_line(np.ones((8, 3), f32)) is None                              # no good sample: all points identical
    two = np.array([[0, 0, 0], [1, 0, 0], [0, 5, 0], [0, 0, 7]], f32)                        # every pair explains exactly two points
    assert Context.pixel_fit_line(two) is None
    P = np.zeros((5, 4), f32)                                                                # PointXYZI-like rows (stride 4)
    P[:, 0] = np.arange(5)
    P[:, 3] = 99.0
    coeff, inl, s, e = Context.pixel_fit_line(P)
    assert np.array_equal(inl, np.arange(5)) and np.allclose(np.abs(coeff[3:]), [1, 0, 0], atol=1e-6) and np.allclose(coeff[:3], [2, 0, 0])
    assert np.allclose(s, [0, 0, 0]) and np.allclose(e, [4, 0, 0])


def test_stride_of_a_pcl_point_gives_the_same_fit():
    """pcl::PointXYZI is 8 floats wide (x y z 1 | intensity + padding): the INTEGRATION.md binding passes stride 8."""
    P, _, _ = _cloud(3)
    wide = np.zeros((len(P), 8), f32)
    wide[:, :3], wide[:, 3], wide[:, 4] = P, 1.0, 7.0
    a, b = Context.pixel_fit_line(P), Context.pixel_fit_line(wide)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_batched_fit_equals_the_per_list_calls_and_checks_its_arguments():
    from panovlm_b200 import PvbError
    clouds = [_cloud(s)[0] for s in range(4)]
    cam = np.zeros((sum(len(c) for c in clouds), 4), f32)
    cam[:, :3] = np.concatenate(clouds)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(cam))                                                         # lists address the cloud through an index array
    inv = np.argsort(perm)
    off = np.concatenate([[0], np.cumsum([len(c) for c in clouds]), [sum(len(c) for c in clouds)]]).astype(np.int32)      # last list empty
    n_in, s, e = Context.pixel_fit_lines(cam[perm], off, inv.astype(np.int32))
    assert n_in[-1] == 0
    for l, c in enumerate(clouds):
        _, inl, s1, e1 = Context.pixel_fit_line(c)
        assert n_in[l] == len(inl) and np.array_equal(s[l], s1) and np.array_equal(e[l], e1)
    with pytest.raises(PvbError):
        Context.pixel_fit_lines(cam, np.array([0, 3], np.int32), np.array([0, 1, len(cam)], np.int32))                   # index out of range
    with pytest.raises(PvbError):
        Context.pixel_fit_lines(cam, np.array([0, 3, 2], np.int32), np.array([0, 1, 2], np.int32))                       # offsets not ascending
