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
