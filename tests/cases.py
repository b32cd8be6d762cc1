"""Seeded test inputs shared by the oracle-pinning tests, the GPU parity tests and tests/make_golden.py."""
import numpy as np


def random_blocks(seed, n, nb=8):
    """Random residual blocks of all six functor types with poses that hit the small-angle and large-angle branches."""
    rng = np.random.default_rng(seed)
    types = rng.integers(0, 6, n).astype(np.int32)
    consts = np.zeros((n, 12))
    for i, bt in enumerate(types):
        c = consts[i]
        if bt in (0, 1):
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            c[:3] = rng.normal(0, 3, 3); c[3:6] = nrm; c[6] = abs(rng.normal(2, 1)); c[7] = 0.7
        elif bt in (2, 3):
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            c[:3] = rng.normal(0, 3, 3); c[3:6] = rng.normal(0, 3, 3); c[6:9] = d; c[9] = 0.7
        elif bt == 4:
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            c[:3] = nrm; c[3:6] = rng.normal(0, 3, 3); c[6:9] = rng.normal(0, 3, 3); c[9] = 0.5
        else:
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            m = rng.normal(size=3)
            c[:3] = nrm; c[3] = 0.0; c[4:7] = rng.normal(0, 3, 3); c[7:10] = m / np.linalg.norm(m); c[10] = rng.uniform(0, 2); c[11] = 2.0
    poses = np.concatenate([rng.normal(0, 0.4, (nb, 3)), rng.normal(0, 1, (nb, 3))], axis=1)
    poses[0] = 0.0                      # identity: the theta^2 <= eps branch of Ceres' rotation helpers
    poses[1, :3] = 1e-10                # still the first-order branch
    poses[2, :3] = [3.0, 0.5, 0.2]      # close to pi
    ref = rng.integers(0, nb, n).astype(np.int32)
    nei = rng.integers(0, nb, n).astype(np.int32)
    normalize = rng.integers(0, 2, n).astype(np.int32)
    huber = np.where(rng.random(n) < 0.5, 0.05, 0.0)
    return dict(type=types, ref=ref, nei=nei, consts=consts, huber=huber, normalize=normalize, poses=poses, nb=nb)


def random_blocks_f6(seed, n, nb=8):
    """Calibration-mode functors (types 6, 7, 8: Plane2Plane_Relative, PlaneRelativeIOUResidual, Line2Line_Angle)."""
    rng = np.random.default_rng(seed)
    c = random_blocks(seed + 1, n, nb)
    types = rng.integers(6, 9, n).astype(np.int32)
    consts = np.zeros((n, 12))
    for i, bt in enumerate(types):
        k = consts[i]
        nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
        if bt == 6:
            k[:3] = nrm; k[3:6] = rng.normal(0, 3, 3); k[6:9] = rng.normal(0, 3, 3); k[9] = 0.5
        elif bt == 7:
            m = rng.normal(size=3)
            k[:3] = nrm; k[3] = 0.0; k[4:7] = rng.normal(0, 3, 3); k[7:10] = m / np.linalg.norm(m); k[10] = rng.uniform(0, 2); k[11] = 2.0
        else:
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            k[:3] = nrm; k[3:6] = d
    c["type"], c["consts"] = types, consts
    same = rng.random(n) < 0.1                 # Line2Line_Angle with identical rotations and directions: the < 1e-3 => 0 branch
    for i in np.nonzero(same & (types == 8))[0]:
        c["nei"][i] = c["ref"][i]; consts[i, 3:6] = consts[i, :3] * (1 if i % 2 else -1)
    return c


def on_plane_blocks():
    """Closed-form cases: point on the plane => r = 0 and a zero Jacobian row (angle types)."""
    consts = np.zeros((2, 12))
    consts[0, :3] = [1.0, 2.0, 3.0]; consts[0, 3:7] = [0, 0, 1, -3.0]; consts[0, 7] = 1.0          # on plane z = 3
    consts[1, :3] = [1.0, 2.0, 3.0]; consts[1, 3:6] = [0, 0, 3.0]; consts[1, 6:9] = [1, 0, 0]; consts[1, 9] = 1.0
    consts[1, 1] = 0.0                                                                              # on the line y=0,z=3
    return dict(type=np.array([1, 3], np.int32), ref=np.array([0, 0], np.int32), nei=np.array([1, 1], np.int32), consts=consts,
                huber=np.zeros(2), normalize=np.ones(2, np.int32), poses=np.zeros((2, 6)), nb=2)


def ref_functor_cases(seed, n, nb=8):
    """CONSTRUCTOR arguments (raw, 16 doubles per row, declaration order of base/CostFunction.h) for all nine functor types of the path, derived
    from random_blocks / random_blocks_f6 so that the poses hit the same branches.  Un-normalised planes / directions on purpose: the constructors'
    own normalisations are part of what the reference-compiled fixture pins."""
    rng = np.random.default_rng(seed)
    a, b = random_blocks(seed, n, nb), random_blocks_f6(seed + 7, n, nb)
    pick = rng.random(n) < 0.6
    c = {k: (np.where(pick, a[k], b[k]) if k in ("type", "ref", "nei", "normalize") else a[k]) for k in a}
    consts = np.where(pick[:, None], a["consts"], b["consts"])
    c["huber"] = np.zeros(n)
    raw = np.zeros((n, 16))
    for i, bt in enumerate(c["type"]):
        k, r = consts[i], raw[i]
        if bt in (0, 1):
            r[:8] = k[:8]                                             # point, unit plane (the functor expects it normalised), weight
        elif bt in (2, 3):
            r[:3] = k[:3]; r[3:6] = k[3:6]; r[6:9] = k[3:6] - k[6:9] * rng.uniform(0.2, 5); r[9] = k[9]      # point, a, b = a - L d
        elif bt in (4, 6):
            r[:3] = k[:3] * rng.uniform(0.1, 7); r[3:9] = k[3:9]; r[9] = k[9]                               # scaled normal
        elif bt == 5:
            s = rng.uniform(0.1, 7)
            r[:3] = k[:3] * s; r[3] = rng.normal() * s; r[4:10] = k[4:10]; r[10] = k[10]; r[11] = k[11]
        elif bt == 7:
            s = rng.uniform(0.1, 7)
            st, en = rng.normal(size=3), rng.normal(size=3)
            r[:3] = k[:3] * s; r[3] = 0.0; r[4:7] = k[4:7]; r[7:10] = (5 * st / np.linalg.norm(st)).astype(np.float32); r[10:13] = (5 * en / np.linalg.norm(en)).astype(np.float32); r[13] = k[11]
        else:
            r[:3] = k[:3] * rng.uniform(0.1, 7); r[3:6] = k[3:6] * rng.uniform(0.1, 7); r[6] = 1.0
    c["raw"] = raw
    # parameter blocks in the functor's call order
    P = c["poses"]
    params = np.zeros((n, 12))
    for i, bt in enumerate(c["type"]):
        if bt in (6, 7):
            params[i, :6] = P[c["ref"][i]]
        elif bt == 8:
            params[i, :3] = P[c["ref"][i], :3]; params[i, 3:6] = P[c["nei"][i], :3]
        else:
            params[i, :6] = P[c["ref"][i]]; params[i, 6:] = P[c["nei"][i]]
    c["params"] = params
    return c


def ref_jacobian_to_block_layout(type, J):
    """Reference Jacobians (blocks in call order) -> the 1x12 row of the block list: (aa_ref, t_ref, aa_nei, t_nei)."""
    J = np.array(J)
    l2l = np.asarray(type) == 8
    J[l2l, 6:9] = J[l2l, 3:6]; J[l2l, 3:6] = 0
    return J
