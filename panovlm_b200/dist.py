"""Multi-GPU plumbing of the hot path: one process per GPU (torch.distributed), units = frames / pose-graph edges.

The path shards by frame: every rank owns a contiguous range of source frames (association + residuals + per-frame
normal equations need only the frame's own cloud and the replicated target), and the only exchange per Gauss-Newton /
LM evaluation is ONE allreduce of the packed normal-equation blocks (per frame: 6x6 upper (21) + gradient (6) + cost +
count = 29 doubles; every rank contributes zeros outside its slice, so the reduction doubles as the all-gather).
"""
import numpy as np

SYS = 29


def shard_range(n_units, world, rank):
    """Contiguous balanced split: the first (n % world) ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_systems(local_sys, n_total, lo, group=None, device=None):
    """local_sys: (n_local, 29) array or tensor of this rank's frames [lo, lo + n_local).  Returns the (n_total, 29)
    systems of all frames on every rank (torch tensor on `device`)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(local_sys, dtype=torch.float64, device=device)
    buf = torch.zeros((n_total, SYS), dtype=torch.float64, device=t.device)
    buf[lo:lo + t.shape[0]] = t
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, group=group)
    return buf


def reduce_single_pose(systems):
    """Sum per-frame 6x6 systems into ONE 6x6 system (2-frame / rigid-map case: north_star's '6x6 / 6x1 blocks')."""
    return np.asarray(systems).sum(axis=0)
