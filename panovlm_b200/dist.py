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


# ---- pose graphs (BASELINE.json configs[3]: Floor, 1593 frames across 8 GPUs) ---------------------------------------------------
EDGE_SYS = 92   # per edge: 12x12 upper (78) | gradient (12) | cost | residual count


def shard_frames_by_weight(weights, world):
    """Contiguous ranges of reference frames, balanced by `weights` (e.g. the residual count of each frame's edges in the previous outer
    iteration).  Returns world + 1 boundaries."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(w)])
    total = c[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(c, target, side="left"))
        if i > 0 and target - c[i - 1] < c[min(i, len(w))] - target:     # the boundary whose cumulative weight is nearest to the target
            i -= 1
        bounds.append(min(i, len(w)))
    bounds.append(len(w))
    return np.maximum.accumulate(np.array(bounds))


def global_edge_list(ref, nei):
    """The pose graph's edges sorted by (ref, nei), unique: the layout every rank reduces into (pvb_blocks_set_edge_list)."""
    e = np.unique(np.stack([np.asarray(ref, np.int64), np.asarray(nei, np.int64)], axis=1), axis=0)
    return e[:, 0].astype(np.int32), e[:, 1].astype(np.int32)


def allreduce_edge_systems(local_sys, group=None):
    """Host form of the exchange (tests / gloo): local_sys is (n_global_edges, 92) with zeros for the edges of other ranks."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(local_sys), dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t.numpy()


class _DevicePtr:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def install_allreduce_hook(ctx, group=None):
    """Registers the ONE exchange step of a sharded pose-graph evaluation: a sum-allreduce of the device buffer of edge systems, in place, enqueued on
    the context's stream (the hook receives it and makes it torch's current stream for the duration of the call).  NCCL reduces the aliased
    device memory directly; with a gloo group (single-GPU tests) the buffer is staged through the host.  Returns a handle holding the callback
    (keep it alive while the hook is installed) and a call counter."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p)
    calls = {"n": 0}

    def _hook(user, dev_ptr, n_doubles, stream):
        calls["n"] += 1
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        # everything below is enqueued on the library's stream (the kernels that wrote the buffer precede it there, the readers follow)
        with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
            t = torch.as_tensor(_DevicePtr(dev_ptr, n_doubles), device="cuda")
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(t, group=group)
            else:
                h = t.cpu()
                dist.all_reduce(h, group=group)
                t.copy_(h)

    cb = HOOK(_hook)
    ctx._ck(ctx._L.pvb_blocks_set_reduce_hook(ctx._h, cb, None))
    return {"callback": cb, "calls": calls}


def remove_allreduce_hook(ctx):
    ctx._ck(ctx._L.pvb_blocks_set_reduce_hook(ctx._h, None, None))
