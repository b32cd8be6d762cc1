"""Independent float64 twin of the residual functors (torch autograd + closed-form Rodrigues), used
ONLY to pin the C++ oracle (tests/test_oracle_*.py) and to generate tests/golden/*.npz.

It deliberately does NOT follow the reference's angle-axis detour (R_rn -> angle-axis -> rotate): it
evaluates P = R_r R_n^T (p - t_n) + t_r directly, so agreement with the oracle checks both the detour's
value and the Jet derivatives against an independent derivation (torch reverse-mode autograd).
"""
import math

import torch

torch.set_default_dtype(torch.float64)


def hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def exp_so3(a):
    th2 = (a * a).sum()
    K = hat(a)
    if th2.item() > 2.220446049250313e-16:
        th = torch.sqrt(th2)
        return torch.eye(3) + torch.sin(th) / th * K + (1 - torch.cos(th)) / th2 * (K @ K)
    return torch.eye(3) + K


def transform(aa_r, t_r, aa_n, t_n, p):
    return exp_so3(aa_r) @ (exp_so3(aa_n).T @ (p - t_n)) + t_r


def vector_angle(v1, v2):
    c = (v1 * v2).sum() / (v1.norm() * v2.norm())
    if c.item() >= 1.0:
        return c * 0.0
    if c.item() <= -1.0:
        return c * 0.0 + math.pi
    return torch.acos(c)


def plane_angle(n1, n2):
    c = (n1 * n2).sum().abs() / (n1.norm() * n2.norm())
    if c.item() >= 1.0:
        return c * 0.0
    return torch.acos(c)


def angle_tail(P, Pp, normalize):
    if normalize:
        nrm = Pp.norm()
        ratio = (nrm - 1.0) / nrm
        c = ratio * Pp
        return vector_angle(Pp - c, P - c)
    return vector_angle(P, Pp)


def residual(btype, c, normalize, aa_r, t_r, aa_n, t_n):
    """btype/c follow oracle/pvo_solver.hpp's Block layout."""
    c = torch.as_tensor(c)
    if btype in (0, 1):
        P = transform(aa_r, t_r, aa_n, t_n, c[0:3])
        n, d = c[3:6], c[6]
        s = (n * P).sum() + d
        dis = s.abs()
        if btype == 0:
            return c[7] * dis
        if dis.item() < 1e-3:
            return dis * 0.0
        Pp = P - dis * n
        if abs(((n * Pp).sum() + d).item()) > 1e-4:
            Pp = P + dis * n
        return angle_tail(P, Pp, normalize)
    if btype in (2, 3):
        P = transform(aa_r, t_r, aa_n, t_n, c[0:3])
        a, dvec = c[3:6], c[6:9]
        k = (dvec * (P - a)).sum()
        if btype == 2:
            k = k / (dvec * dvec).sum()
        Pp = a + k * dvec
        dis = (P - Pp).norm()
        if btype == 2:
            return c[9] * dis
        if dis.item() < 1e-3:
            return dis * 0.0
        return angle_tail(P, Pp, normalize)
    if btype == 4:
        A = transform(aa_r, t_r, aa_n, t_n, c[3:6])
        B = transform(aa_r, t_r, aa_n, t_n, c[6:9])
        return c[9] * plane_angle(c[0:3], torch.linalg.cross(A, B))
    if btype == 5:
        M = transform(aa_r, t_r, aa_n, t_n, c[4:7])
        n, d = c[0:3], c[3]
        dis = ((n * M).sum() + d).abs()
        Pp = M - dis * n
        if abs(((n * Pp).sum() + d).item()) > 1e-4:
            Pp = M + dis * n
        ang = vector_angle(Pp, c[7:10])
        if ang.item() < c[10].item():
            return ang * 0.0
        return c[11] * (ang - c[10])
    if btype == 6:      # Plane2Plane_Relative: one relative pose (the ref block), degrees
        R = exp_so3(aa_r)
        A, B = R @ c[3:6] + t_r, R @ c[6:9] + t_r
        return c[9] * plane_angle(c[0:3], torch.linalg.cross(A, B)) * 180.0 / math.pi
    if btype == 7:      # PlaneRelativeIOUResidual
        M = exp_so3(aa_r) @ c[4:7] + t_r
        n, d = c[0:3], c[3]
        dis = ((n * M).sum() + d).abs()
        Pp = M - dis * n
        if abs(((n * Pp).sum() + d).item()) > 1e-4:
            Pp = M + dis * n
        ang = vector_angle(Pp, c[7:10])
        if ang.item() < c[10].item():
            return ang * 0.0
        return c[11] * (ang - c[10])
    if btype == 8:      # Line2Line_Angle: rotations only, unit directions, no norm division
        dr = exp_so3(aa_r) @ (exp_so3(aa_n).T @ c[3:6])
        cs = (dr * c[0:3]).sum().abs()
        if cs.item() >= 1.0:
            return cs * 0.0
        ang = torch.acos(cs)
        if ang.item() < 1e-3:
            return ang * 0.0
        return ang
    raise ValueError(btype)


def residual_and_jacobian(btype, c, normalize, pose_r, pose_n):
    xs = [torch.tensor(list(pose_r[0:3]), requires_grad=True), torch.tensor(list(pose_r[3:6]), requires_grad=True),
          torch.tensor(list(pose_n[0:3]), requires_grad=True), torch.tensor(list(pose_n[3:6]), requires_grad=True)]
    r = residual(btype, c, normalize, *xs)
    grads = torch.autograd.grad(r, xs, allow_unused=True)
    J = torch.cat([g if g is not None else torch.zeros(3) for g in grads])
    return r.item(), J.numpy()
