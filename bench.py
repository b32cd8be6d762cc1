#!/usr/bin/env python
"""bench.py — residual-evaluations/sec of the fused correspondence-and-residual hot path (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[4], the dense ICP sweep the metric's 1/2/4/8-GPU clause and the
HBM-roofline target are quoted on: a 10,000,000-point target cloud (float32, 200x50x4 m multi-room floor plan) and
64 source frames x 156,250 points per GPU; one *step* = one Gauss-Newton iteration = for every source point: SE(3)
transform (float32 store) -> exact 10-NN on the cell-sorted target -> class test -> LSQ plane fit + collinearity test
-> Point2Plane_Meter residual + analytic Jacobian + Huber -> per-frame 6x6 normal equations -> (N>1: one NCCL
allreduce of the packed 6x6/6x1 blocks) -> 64 tiny solves + pose update on the host.
One residual evaluation = one source point at one pose estimate (points x iterations).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (torchrun for N>1)
  python bench.py --impl reference ...                           the reference algorithm on the host cores (oracle port)

Multi-GPU (SURVEY.md 8e / BASELINE configs[4]): the 64 source frames are sharded across the ranks (64 / N frames each: STRONG scaling, the
default), the target is replicated, the only exchange is one NCCL allreduce of the packed 6x6/6x1 blocks of all 64 frames per step.
`--scaling weak` gives every rank its own 64 frames instead (round-1 behaviour).  Synthetic data, seeds fixed.
The line carries `parity`: the oracle's reduced systems of the first step (all 64 frames against the 10 M-point target, the same
run that is timed as `cpu_baseline`) against the GPU's, and `extra.*`: the other BASELINE configs measured in the same run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_PER_QUERY = {10: 176.0, 5: 96.0}   # bytes: query 16 + k*16 target records, reduced-output mode
HBM_FALLBACK_GBS = 6650.0                 # /opt/skills/guides/B200_PROFILING.md


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-target", type=int, default=10_000_000)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--pts-per-frame", type=int, default=156_250)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--radius", type=float, default=1.0)
    ap.add_argument("--cpu-sample-frames", type=int, default=64, help="source frames in the bounded CPU-baseline sample (64 = the whole step, ~4 s on 64 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--cell", type=float, default=0.0, help="target grid cell size in metres (0: library default)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: shard the 64 frames (strong, SURVEY 8e) or 64 frames per rank (weak)")
    ap.add_argument("--no-configs", action="store_true", help="skip extra.pair_icp / room_refine_pose / room_joint / floor_refine_pose_sharded (BASELINE configs[0,1,2,3])")
    ap.add_argument("--floor-frames", type=int, default=1593)
    ap.add_argument("--room-frames", type=int, default=454)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (recipe's clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_data(args, rank, world=1):
    from panovlm_b200 import synth
    # the target is identical on every rank (same seed).  weak: each rank owns its own 64 source frames; strong: the same 64 frames
    # everywhere, rank r keeps frames [r * 64 / N, (r + 1) * 64 / N)
    weak = args.scaling == "weak"
    d = synth.make_dense_sweep(n_target=args.n_target, n_frames=args.frames, pts_per_frame=args.pts_per_frame, seed=20260929,
                               source_seed=20260930 + (1000 * rank if weak else 0))
    d["frame_lo"], d["frame_hi"] = 0, args.frames
    if not weak and world > 1:
        lo, hi = args.frames * rank // world, args.frames * (rank + 1) // world
        o0, o1 = int(d["src_off"][lo]), int(d["src_off"][hi])
        d["all_src_local"], d["all_src_off"], d["all_poses_lw_init"] = d["src_local"], d["src_off"], d["poses_lw_init"]
        d["src_local"] = np.ascontiguousarray(d["src_local"][o0:o1])
        d["src_off"] = (d["src_off"][lo:hi + 1] - o0).astype(np.int32)
        d["poses_lw_init"] = d["poses_lw_init"][lo:hi].copy()
        d["poses_lw_true"] = d["poses_lw_true"][lo:hi].copy()
        d["frame_lo"], d["frame_hi"] = lo, hi
    return d


def use_physical_cores():
    """The kd-tree search is memory-latency bound: one thread per physical core is faster than all hyper-threads."""
    from oracle import pvo
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            pvo.set_num_threads(n)
    except Exception:
        pass
    return pvo.num_threads()


def cpu_baseline(args, d, steps=1, mode=1, tree=None, build_s=0.0):
    """The oracle port of the reference algorithm on a bounded sample of the same workload (same target, first
    `cpu_sample_frames` source frames), kd-tree prebuilt (the GPU path also builds its grid once, outside the step)."""
    from oracle import pvo
    nf = min(args.cpu_sample_frames, args.frames)
    off = d["src_off"][: nf + 1]
    if tree is None:
        t0 = time.time()
        tree = pvo.KdTreeHandle(d["target"])
        build_s = time.time() - t0
    times, n_assoc, systems = [], 0, None
    for _ in range(steps):
        t0 = time.time()
        systems, tt, n_assoc = pvo.dense_icp_eval(d["target"], d["src_local"][: off[-1]], off, d["poses_lw_init"][:nf], 0.05, args.radius, args.k, 0.2, 1.0, mode, tree)
        times.append(time.time() - t0)
    evals = int(off[-1])
    return {"evals": evals, "seconds": float(np.median(times)), "per_step": times, "kdtree_build_s": build_s, "n_assoc": int(n_assoc),
            "threads": pvo.num_threads(), "frames": nf, "tree": tree, "systems": systems}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d = make_data(args, 0)
    from oracle import pvo
    use_physical_cores()
    nf = min(args.cpu_sample_frames, args.frames)
    off = d["src_off"][: nf + 1]
    tree = pvo.KdTreeHandle(d["target"])
    per = []
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        pvo.dense_icp_eval(d["target"], d["src_local"][: off[-1]], off, d["poses_lw_init"][:nf], 0.05, args.radius, args.k, 0.2, 1.0, 1, tree)
        if it >= args.warmup:
            per.append(time.time() - t0)
    total = sum(per)
    value = int(off[-1]) * args.steps / total
    sample = f"{nf} of {args.frames} source frames ({int(off[-1])} points) per step against the full {args.n_target}-point target, kd-tree prebuilt"
    line = {"impl": "reference", "metric": "residual_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": pvo.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bench_blocks(ctx, peak_gbs, n=1_400_000, nb=454, seed=5):
    """Room-shaped residual-block list: every frame linked to 7 neighbours, 440 plane correspondences per edge."""
    rng = np.random.default_rng(seed)
    ref = np.repeat(np.arange(nb), 7)
    nei = (ref + np.tile(np.array([-3, -2, -1, 1, 2, 3, 40]), nb)) % nb
    per = n // len(ref)
    ref, nei = np.repeat(ref, per).astype(np.int32), np.repeat(nei, per).astype(np.int32)
    n = len(ref)
    consts = np.zeros((n, 12))
    consts[:, :3] = rng.normal(0, 4, (n, 3))
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    consts[:, 3:6] = nrm; consts[:, 6] = np.abs(rng.normal(2, 1, n)); consts[:, 7] = 1.0
    poses = np.concatenate([rng.normal(0, 0.2, (nb, 3)), rng.normal(0, 1.0, (nb, 3))], axis=1)
    ctx.blocks_set(np.full(n, 1, np.int32), ref, nei, consts, 2 * np.pi / 180, 1, nb)
    out = {"n_blocks": n, "n_pose_blocks": nb, "n_edges": int(len(set(zip(ref.tolist(), nei.tolist()))))}
    for name, rows, sysm, balg in (("reduced", False, True, 64.0), ("rows", True, False, 168.0)):
        ms = []
        for _ in range(6):
            ctx.blocks_evaluate(poses, want_rows=rows, want_system=sysm)
            ms.append(ctx.blocks_kernel_time_ms())
        k = float(np.median(ms[2:]))
        out[name] = {"kernel_ms": k, "evals_per_s": n / (k * 1e-3), "algorithmic_bytes_per_row": balg, "achieved_GBs": n * balg / (k * 1e-3) / 1e9,
                     "frac_of_hbm_peak": n * balg / (k * 1e-3) / 1e9 / peak_gbs}
    return out


def bench_reproj(ctx, peak_gbs, n_cams=454, n_points=200_000, seed=7):
    """Room-shaped camera-camera term (configs[2]): 454 panoramas, 200 k structure points, ~1.3 M observations of PanoramaReprojResidual_1Angle."""
    from panovlm_b200 import synth
    d = synth.make_ba_problem(n_cams=n_cams, n_points=n_points, track_len=(3, 10), seed=seed)
    n = len(d["cam"])
    ctx.reproj_set(d["cam"], d["point"], d["bearing"], n_cams, n_points, huber=4.0 * np.pi / 180.0)
    out = {"n_observations": n, "n_cams": n_cams, "n_points": n_points}
    for name, rows, sysm, balg in (("rows", True, False, 140.0), ("reduced", False, True, 144.0)):
        ms = []
        for _ in range(6):
            ctx.reproj_evaluate(d["cams"], d["points"], rows, sysm)
            ms.append(ctx.reproj_kernel_time_ms())
        k = float(np.median(ms[2:]))
        out[name] = {"kernel_ms": k, "evals_per_s": n / (k * 1e-3), "algorithmic_bytes_per_obs": balg, "achieved_GBs": n * balg / (k * 1e-3) / 1e9,
                     "frac_of_hbm_peak": n * balg / (k * 1e-3) / 1e9 / peak_gbs}
    return out


def config_dict(args, world=1):
    weak = args.scaling == "weak"
    per = args.frames if weak else f"{args.frames} / {world}"
    return {"workload": f"configs[4]: dense ICP sweep, {args.n_target}-pt target, {args.frames} source frames x {args.pts_per_frame} pts "
                        f"{'per GPU' if weak else 'in total'}, k={args.k}, radius={args.radius} m, plane_tol=0.05, Point2Plane_Meter + Huber(0.2), per-frame 6x6 reduce",
            "cell_m": args.cell, "n_target": args.n_target, "frames_total": args.frames * (world if weak else 1), "frames_per_gpu": per, "pts_per_frame": args.pts_per_frame,
            "k": args.k, "radius_m": args.radius,
            "l2": "inputs (160 MB target records + 160 MB queries + cell table) exceed the 126 MB L2; no explicit flush",
            "search": "exact 10-NN, one buffered pass over the merged super-rows of the static target (queries ordered by target cell, re-ordered after large pose updates: "
                      "the re-ordering is inside kernel_ms of the steps that need it); per-point search radius bounded by the previous step's 10th distance + displacement, "
                      "else by the target's own 10-NN radius at the nearest target point (static bound), else by list compaction - every bound is exact; warm-up steps "
                      "provide the bounds of the first timed step; roofline.cold_search_kernel_ms = the same launch without them",
            "parallelism": f"frames sharded across GPUs ({args.scaling}), target replicated, one NCCL allreduce of the packed 6x6/6x1 blocks of all frames per step"}


def kernel_sources_hash():
    """sha256 (16 hex digits) of the sources of the fused kernel: profiles/traffic.json is only used when it was captured from these."""
    import hashlib
    h = hashlib.sha256()
    for f in ("pvb_kernels.cuh", "pvb_knn.cuh", "pvb_math.cuh"):
        with open(os.path.join(ROOT, "panovlm_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_traffic():
    """(dram bytes per launch, warp instructions per launch, source) of the fused kernel from the committed ncu capture, or Nones when the
    capture does not belong to the kernel sources of this tree (tools/capture_traffic.py rewrites it)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tp))
    except Exception:
        return None, None, "no profiles/traffic.json"
    if t.get("kernel_sources_sha256_16") != kernel_sources_hash():
        return None, None, f"profiles/traffic.json was captured from other kernel sources ({t.get('kernel_sources_sha256_16')}): ignored"
    return t.get("k_associate_dram_bytes_per_launch"), t.get("k_associate_warp_instructions_per_launch"), t.get("capture", "profiles/traffic.json")


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import panovlm_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    weak = args.scaling == "weak"
    d = make_data(args, rank, world)
    nq = int(d["src_off"][-1])                                       # this rank's source points
    nf = len(d["src_off"]) - 1                                       # this rank's frames
    frames_total = args.frames * world if weak else args.frames
    nq_total = nq * world if weak else args.frames * args.pts_per_frame
    slot0 = rank * nf if weak else d["frame_lo"]                     # this rank's rows in the packed buffer of all frames

    ctx = panovlm_b200.Context(local)
    stream = torch.cuda.Stream()                # a real (non-legacy) stream: library work, torch events, copies and NCCL all on it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.dense_set_target(d["target"], args.cell)
    src_pinned = torch.from_numpy(d["src_local"]).pin_memory()
    ctx.dense_set_sources_ptr(src_pinned.data_ptr(), d["src_off"])
    prm = ctx.dense_params(0.05, args.radius, args.k, panovlm_b200.P2PLANE_METER, 1, 0.2, 1.0)
    sys_all = torch.zeros((frames_total, 29), dtype=torch.float64, device="cuda")     # packed normal-equation blocks of ALL frames
    my_ptr = sys_all.data_ptr() + slot0 * 29 * 8
    sys_host = torch.zeros((frames_total, 29), dtype=torch.float64).pin_memory()

    def gn_step(poses):
        """one Gauss-Newton iteration, inputs resident in HBM"""
        if world > 1:
            sys_all.zero_()
        ctx.dense_evaluate_device(poses, prm, my_ptr)
        if world > 1:
            dist.all_reduce(sys_all)                  # the single exchange step: 6x6/6x1 blocks of every frame
        sys_host.copy_(sys_all, non_blocking=True)
        stream.synchronize()
        mine = sys_host[slot0:slot0 + nf].numpy()
        return ctx.dense_gauss_newton_step(mine, poses, 1e-6), mine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident value: W warm-up + K timed GN iterations
    poses = d["poses_lw_init"].copy()
    for _ in range(args.warmup):
        poses, _ = gn_step(poses)
    poses = d["poses_lw_init"].copy()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, costs, first_sys = [], [], None
    e0.record()
    for it in range(args.steps):
        poses, mine = gn_step(poses)
        kernel_ms.append(ctx.dense_kernel_time_ms())
        costs.append(float(mine[:, 27].sum()))
        if it == 0:
            first_sys = sys_host.numpy().copy()                     # all frames' systems at the initial poses (after the allreduce)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - l0
    if rank == 0:
        sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = nq_total * args.steps / (ms_max * 1e-3)

    # ---- e2e: the same iteration through the C ABI with HOST buffers (source clouds + poses H2D, systems D2H)
    e2e = None
    if not args.no_e2e:
        poses_e = d["poses_lw_init"].copy()
        # the copy rate of this box's pinned host memory -> HBM path (the e2e step moves nq * 16 bytes over it)
        h2d_gbs = None
        try:
            dst = torch.empty(src_pinned.numel(), dtype=src_pinned.dtype, device="cuda").view(src_pinned.shape)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dst.copy_(src_pinned, non_blocking=True); torch.cuda.synchronize()
            g0.record(); dst.copy_(src_pinned, non_blocking=True); g1.record(); torch.cuda.synchronize()
            h2d_gbs = src_pinned.numel() * 4 / (g0.elapsed_time(g1) * 1e-3) / 1e9
            del dst
        except Exception:
            pass

        def e2e_step(p):
            ctx.dense_set_sources_ptr(src_pinned.data_ptr(), d["src_off"])       # H2D of this step's source clouds + re-ordering
            return gn_step(p)
        for _ in range(max(1, args.warmup // 2)):
            poses_e, _ = e2e_step(poses_e)
        poses_e = d["poses_lw_init"].copy()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            poses_e, _ = e2e_step(poses_e)
        f1.record()
        barrier()
        te = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": nq_total * args.steps / (float(te.item()) * 1e-3), "unit": "evals/s",
               "h2d_bytes_per_step": int(nq * 16 + nf * (21 + 12) * 8 * 1 + (nf + 1) * 4), "d2h_bytes_per_step": int(frames_total * 29 * 8),
               "ms_per_step": float(te.item()) / args.steps,
               "h2d_gbs_measured": h2d_gbs, "h2d_floor_ms_per_step": (nq * 16 / (h2d_gbs * 1e9) * 1e3) if h2d_gbs else None,
               "note": "per rank and step: its source clouds (pinned host) -> device + ordering by target cell (chunked, overlapped with the copies), poses H2D, fused kernel per chunk, "
                       "allreduce, reduced systems of all frames D2H; target map resident.  h2d_floor_ms_per_step = this step's input bytes / the pinned-memory copy rate measured in this run"}

    # ---- the same launch without search-radius bounds (cold search), for transparency
    ctx.dense_reset_hints()
    ctx.dense_evaluate_device(d["poses_lw_init"], prm, my_ptr)
    cold_ms = ctx.dense_kernel_time_ms()

    # ---- BASELINE configs[3] under the same launch: Floor RefinePose sharded over the ranks (every rank takes part)
    floor = None
    if not args.no_configs and not args.no_extra:
        try:
            from tools import bench_configs
            floor = bench_configs.floor_refine_pose(ctx, world, rank, n_frames=args.floor_frames)
        except Exception as e:                                   # never lose the headline line
            floor = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused associate+residual), CUDA-event time measured inside the library
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = HBM_FALLBACK_GBS, "fallback"
    if os.path.exists(peaks_path):
        try:
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    k_ms = float(np.mean(kernel_ms))
    alg_bytes = nq * B_ALG_PER_QUERY[args.k]
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, wi, traffic_src = load_traffic()
    if world > 1 and not weak:
        traffic = wi = None                                      # the capture is a full 64-frame launch
    # instruction-issue view of the same launch (the kernel is issue-bound): executed warp instructions per
    # launch from the committed ncu capture vs 4 schedulers x 148 SMs x SM clock
    issue = None
    if wi:
        issue = {"warp_instructions": wi, "achieved_ginst_s": wi / (k_ms * 1e-3) / 1e9, "peak_ginst_s": 4 * 148 * 1.965, "frac": wi / (k_ms * 1e-3) / 1e9 / (4 * 148 * 1.965)}
    roofline = {"bound": "hbm", "kernel": "k_associate<10,true,...,MODE 4> (merged super-rows, warp-synchronous single pass)", "issue": issue, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": k_ms, "kernel_ms_per_step": [round(x, 4) for x in kernel_ms], "cold_search_kernel_ms": cold_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_share_of_step": k_ms * args.steps / ms}

    line = {"metric": "residual_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world), "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "gn_cost_first_last": [costs[0], costs[-1]], "ms_per_gauss_newton_iter": ms_max / args.steps}
    line["extra"] = {}
    if floor is not None:
        line["extra"]["floor_refine_pose_sharded"] = floor

    # ---- secondary kernel: K3/K4 over a Room-shaped correspondence list (configs[1] sizes: 454 pose blocks, ~3.2 k edges,
    #      1.4 M residual blocks = Point2Plane_Angle + Huber 2 deg), device time of k_eval_blocks per evaluation
    if world == 1 and not args.no_extra:
        try:
            line["extra"]["k_eval_blocks"] = bench_blocks(ctx, peak)
        except Exception as e:                                   # never lose the headline line
            line["extra"]["k_eval_blocks"] = {"error": str(e)}
        try:
            line["extra"]["k_reproj_rows"] = bench_reproj(ctx, peak)
        except Exception as e:
            line["extra"]["k_reproj_rows"] = {"error": str(e)}
        if not args.no_configs:
            from tools import bench_configs
            for name, fn in (("pair_icp", bench_configs.pair_icp), ("room_refine_pose", lambda c: bench_configs.room_refine_pose(c, n_frames=args.room_frames)),
                             ("room_joint", lambda c: bench_configs.room_joint(c, n_frames=args.room_frames))):
                try:
                    line["extra"][name] = fn(ctx)
                except Exception as e:
                    line["extra"][name] = {"error": repr(e)}

    # ---- CPU baseline: the oracle port timed on this box's host cores (bounded sample), N=1 only; its systems are the parity check of the GPU's first step
    if world == 1 and not args.no_cpu_baseline:
        from oracle import pvo
        use_physical_cores()
        cb = cpu_baseline(args, d, steps=1, mode=1)
        v = cb["evals"] / cb["seconds"]
        sample = (f"{cb['frames']} of {args.frames} source frames ({cb['evals']} points, {cb['n_assoc']} accepted) against the full {args.n_target}-point target, "
                  f"kd-tree prebuilt ({cb['kdtree_build_s']:.1f} s, not counted), association + Jet<12> residuals on all threads")
        line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": cb["threads"], "kind": "port", "sample": sample, "seconds": cb["seconds"]}
        s_cpu, s_gpu = np.asarray(cb["systems"]), first_sys[:cb["frames"]]
        scale = float(np.abs(s_cpu).max())
        line["parity"] = {"against": "oracle (cpu_baseline run), reduced systems of the first timed step, all source frames of the sample",
                          "counts_equal": bool(np.array_equal(s_cpu[:, 28], s_gpu[:, 28])), "n_accepted": int(s_gpu[:, 28].sum()), "n_accepted_oracle": int(s_cpu[:, 28].sum()),
                          "sys_max_rel": float(np.abs(s_cpu - s_gpu).max() / scale), "gate": 1e-8, "frames": int(cb["frames"])}
        line["parity"]["ok"] = bool(line["parity"]["counts_equal"] and line["parity"]["sys_max_rel"] < 1e-8)
        cb0 = cpu_baseline(argparse.Namespace(**{**vars(args), "cpu_sample_frames": 1}), d, steps=1, mode=0, tree=cb["tree"])
        line["cpu_baseline"]["reference_faithful"] = {"value": cb0["evals"] / cb0["seconds"], "unit": "evals/s",
                                                      "note": "association serial on 1 core (util/Optimization.cpp:506-562 has no omp), evaluation on all threads; 1 frame"}
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
