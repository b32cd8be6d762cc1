"""Host loop the hot path drops into: a Python mirror of LidarOdometry::RefinePose / EstimatePose
(lidar_mapping/LidarOdometry.cpp:15-114, 116-187) driving the C ABI.  Every numeric step runs in the library:
FindNeighbors (host C++), point-to-plane association of all pose-graph edges (fused kNN kernel), line-to-line vote
matrices (kernel) + tails (host C++), LiDAR line tracks and their gate (LidarLineMatch::GenerateTracks with neighbour size 4 /
track length 3 as LidarOdometry.cpp:47-50, util/Optimization.cpp:383-400), residual-block builders (host C++), Levenberg-Marquardt
with device evaluation.
"""
import time

import numpy as np

from .api import BlockList, Context, LineFrame


# the line-to-line tails and blocks on the device (pvb_frames_line2line_blocks_device) instead of the host cores; same block constants either way
# (tests/test_device_line_blocks.py).  The sharded mirror keeps the host tails: its device variant has not been run on more than one GPU.
DEVICE_LINE_BLOCKS = True


class OdometryConfig:
    def __init__(self, point_to_plane=True, line_to_line=True, point_to_line=False, use_segment=True, angle_residual=True, normalize_distance=True, plane_dis_threshold=1.0,
                 line_dis_threshold=0.3, plane_tolerance=0.05, lidar_weight=0.01, neighbor_size=6, max_lm_iterations=20, line_tracks=True,
                 track_neighbor_size=4, min_track_length=3):
        self.point_to_plane, self.line_to_line = point_to_plane, line_to_line
        self.point_to_line, self.use_segment = point_to_line, use_segment                           # config point_to_line_residual (off in Room.txt:72)
        self.angle_residual, self.normalize_distance = angle_residual, normalize_distance          # config/Room.txt:67-74
        self.plane_dis_threshold, self.line_dis_threshold, self.plane_tolerance = plane_dis_threshold, line_dis_threshold, plane_tolerance
        self.lidar_weight, self.neighbor_size, self.max_lm_iterations = lidar_weight, neighbor_size, max_lm_iterations
        self.line_tracks, self.track_neighbor_size, self.min_track_length = line_tracks, track_neighbor_size, min_track_length   # LidarOdometry.cpp:47-50
        # weight handed to the point-to-plane / point-to-line builders.  RefinePose passes none (default 1.0, LidarOdometry.cpp:38-57: lidar_weight is NOT used there);
        # the joint stage passes config.lidar_weight to the point-to-plane builder (CameraLidarOptimizer.cpp:455-458) - see joint.joint_lidar_config.  The angle
        # functors ignore their weight (base/CostFunction.h:714-717, 916-920), so this only matters with angle_residual = False.
        self.plane_weight, self.point_line_weight = 1.0, 1.0


def pose_blocks_from_world(R_wl, t_wl, R_to_aa):
    """T_wl -> (aa_lw, t_lw) blocks (LidarOdometry.cpp:25-33)."""
    out = np.zeros((len(R_wl), 6))
    for i, (R, t) in enumerate(zip(R_wl, t_wl)):
        R_lw = R.T
        out[i, :3] = R_to_aa(R_lw)
        out[i, 3:] = -R_lw @ t
    return out


def world_from_pose_blocks(poses, aa_to_R):
    if hasattr(aa_to_R, "batch"):                 # optional vectorised form of the same map: aa_to_R.batch(aa[n, 3]) -> R[n, 3, 3]
        poses = np.asarray(poses, dtype=np.float64)
        R = np.transpose(np.asarray(aa_to_R.batch(poses[:, :3])), (0, 2, 1))
        t = -np.einsum("nij,nj->ni", R, poses[:, 3:])
        return list(R), list(t)
    R_wl, t_wl = [], []
    for p in poses:
        R = aa_to_R(p[:3]).T
        R_wl.append(R); t_wl.append(-R @ p[3:])
    return R_wl, t_wl


def first_valid_frame(frames):
    """RefinePose holds the pose blocks of the FIRST VALID frame constant (LidarOdometry.cpp:59-76), not blindly those of frame 0: `if(!lidars[i].IsPoseValid() || !lidars[i].valid) continue;`.  A frame dict may carry valid = False / pose_valid = False."""
    for i, f in enumerate(frames):
        if f.get("valid", True) and f.get("pose_valid", True):
            return i
    return 0


def pose_graph_edges(poses, cfg: OdometryConfig, aa_to_R):
    """FindNeighbors (LidarFeatureAssociate.cpp:19-111) -> the (reference frame, neighbour frame) pairs of one outer iteration."""
    n = len(poses)
    _, t_wl = world_from_pose_blocks(poses, aa_to_R)
    neighbors = Context.find_neighbors(np.array(t_wl), None, None, cfg.neighbor_size)
    return [(i, j) for i in range(n) for j in neighbors[i] if 0 <= j < n and j != i]


def build_problem(ctx: Context, frames, poses, cfg: OdometryConfig, aa_to_R, frame_range=None, host_point2plane=True, per_edge_line_calls=False, all_edges=None,
                  device_line_blocks=False):
    """One outer iteration's residual blocks (the Add*Residual calls of RefinePose); returns a BlockList and the GLOBAL edge list.
    frame_range = (lo, hi): only the edges whose reference frame lies in [lo, hi) are associated and turned into blocks (one rank's shard of a
    pose graph split across GPUs); the clouds of all frames stay available as halo.
    host_point2plane = False leaves the point-to-plane family to Context.frames_point2plane_blocks (association and blocks without a host round trip):
    the BlockList then only holds the other families and the edge list of the shard is returned as third value.
    device_line_blocks (with host_point2plane = False): the line-to-line tails and blocks stay on the device as well (Context.frames_line2line_blocks_device);
    the blocks wait in HBM for the frames_point2plane_blocks call of the caller."""
    n = len(frames)
    R_wl, t_wl = world_from_pose_blocks(poses, aa_to_R)
    if all_edges is None:
        all_edges = pose_graph_edges(poses, cfg, aa_to_R)
    edges = all_edges if frame_range is None else [(i, j) for (i, j) in all_edges if frame_range[0] <= i < frame_range[1]]
    cap = sum(len(frames[j]["surfFlat"]) for _, j in edges) + sum(len(frames[j]["cornerLessSharp"]) for _, j in edges) * 6 + 16
    bl = BlockList(cap)
    lf = None
    if cfg.line_to_line or cfg.point_to_line:
        lf = [LineFrame(f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["end_points"], R_wl[i], t_wl[i]) for i, f in enumerate(frames)]
    if cfg.point_to_line:                                           # AddLidarPointToLineResidual (Optimization.cpp:443-504): consecutive frames only (:475)
        near = [(i, j) for (i, j) in edges if abs(i - j) <= 1]
        if cfg.use_segment:                                         # AssociatePoint2LineSegmentKNN (:482)
            for (i, j) in near:
                _, _, pt, a, b = ctx.point2line_segment_knn_associate(lf[i], lf[j], cfg.line_dis_threshold)
                Context.build_point2line_blocks(bl, pt, a, b, i, j, cfg.angle_residual, cfg.normalize_distance, cfg.point_line_weight)
        elif near:                                                  # AssociatePoint2Line (:484)
            ctx.frames_set_corners([f["cornerLessSharp"] for f in frames])
            e, _, pt, a, b = ctx.frames_associate_point2line(poses, np.array([x[0] for x in near], np.int32), np.array([x[1] for x in near], np.int32), cfg.line_dis_threshold)
            bounds = np.searchsorted(e, np.arange(len(near) + 1))
            for ei, (i, j) in enumerate(near):
                lo, hi = bounds[ei], bounds[ei + 1]
                if hi > lo:
                    Context.build_point2line_blocks(bl, pt[lo:hi], a[lo:hi], b[lo:hi], i, j, cfg.angle_residual, cfg.normalize_distance, cfg.point_line_weight)
    if cfg.line_to_line:                                            # AddLidarLineToLineResidual2 (Optimization.cpp:329-441)
        tracks = None
        if cfg.line_tracks:                                         # LidarLineMatch::GenerateTracks (LidarLineMatch.cpp:36-86), threshold hard-coded 0.3
            tracks = ctx.generate_line_tracks(lf, Context.find_neighbors(np.array(t_wl), None, None, cfg.track_neighbor_size), None, 0.3, cfg.min_track_length)
        if per_edge_line_calls:                                     # one library call per edge (the round-1 path; kept for the equivalence test)
            world = [ctx.transform_cloud(f["cornerLessSharp"], R_wl[i], t_wl[i]) for i, f in enumerate(frames)]
            for (i, j) in edges:
                nl, rl, a, b = ctx.line2line_associate(lf[i], lf[j], cfg.line_dis_threshold)
                keep = Context.line_tracks_gate(tracks, i, j, rl, nl) if tracks is not None else np.ones(len(nl), bool)   # Optimization.cpp:383-400
                for k in np.nonzero(keep)[0]:
                    Context.build_line2line_blocks(bl, lf[j], world[j], nl[k], a[k], b[k], i, j, cfg.angle_residual, cfg.normalize_distance, 1.0)
        elif edges and device_line_blocks and not host_point2plane and cfg.point_to_plane:   # votes, tails and blocks on the device
            ctx.frames_line2line_blocks_device(lf, edges, cfg.line_dis_threshold, tracks, cfg.angle_residual, cfg.normalize_distance, 1.0)
        elif edges:                                                 # all edges in one call: batched device votes, tails and blocks on the host cores
            ctx.frames_line2line_blocks(bl, lf, edges, cfg.line_dis_threshold, tracks, cfg.angle_residual, cfg.normalize_distance, 1.0)
    if cfg.point_to_plane:                                          # AddLidarPointToPlaneResidual (Optimization.cpp:506-562)
        ctx.frames_set([f["surfLessFlat"] for f in frames], [f["surfFlat"] for f in frames])
    if not host_point2plane:
        return bl, all_edges, edges
    if cfg.point_to_plane and edges:
        ref = np.array([e[0] for e in edges], np.int32)
        nei = np.array([e[1] for e in edges], np.int32)
        e, q, pt, pl = ctx.frames_associate_point2plane(poses, ref, nei, cfg.plane_tolerance, cfg.plane_dis_threshold, 10)
        # correspondences come back edge-major with their edge index: all blocks of the outer iteration in one builder call
        Context.build_point2plane_blocks_edges(bl, e, pt, pl, ref, nei, cfg.angle_residual, cfg.normalize_distance, cfg.plane_weight)
    return bl, all_edges


def refine_pose(ctx: Context, frames, poses, cfg: OdometryConfig, aa_to_R, device_blocks=True, device_line_blocks=DEVICE_LINE_BLOCKS):
    """RefinePose: build the problem at `poses`, fix the first frame, solve (LidarOdometry.cpp:15-114).  device_blocks: the point-to-plane
    correspondences become residual blocks on the device (no download / rebuild / upload); the other families are built on the host and appended."""
    t0 = time.time()
    if device_blocks and cfg.point_to_plane:
        bl, edges, mine = build_problem(ctx, frames, poses, cfg, aa_to_R, host_point2plane=False, device_line_blocks=device_line_blocks)
        ref, nei = np.array([e[0] for e in mine], np.int32), np.array([e[1] for e in mine], np.int32)
        n_total = ctx.frames_point2plane_blocks(poses, ref, nei, cfg.plane_tolerance, cfg.plane_dis_threshold, cfg.angle_residual, cfg.normalize_distance, cfg.plane_weight,
                                                len(frames), extra=bl.view())
        bl.n = n_total
    else:
        bl, edges = build_problem(ctx, frames, poses, cfg, aa_to_R)
        v = bl.view()
        ctx.blocks_set(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"], len(frames))
    mask = np.zeros(len(frames), np.uint8); mask[first_valid_frame(frames)] = 1
    t_lm = time.time()
    new_poses, summary = ctx.blocks_solve_lm(poses, mask, cfg.max_lm_iterations)
    summary["lm_s"] = time.time() - t_lm
    summary["build_s"] = t_lm - t0
    summary["n_blocks"], summary["n_edges"] = bl.n, len(edges)
    return new_poses, summary


def refine_pose_sharded(ctx: Context, frames, poses, cfg: OdometryConfig, aa_to_R, world, rank, group=None, device_blocks=True, device_line_blocks=False):
    """RefinePose on one rank of a pose graph sharded across GPUs (BASELINE.json configs[3]): contiguous ranges of reference frames
    balanced by the query count of their edges, association + residual blocks of the rank's own edges only, the GLOBAL edge list as reduction layout
    and ONE allreduce of the edge systems per evaluation (panovlm_b200.dist.install_allreduce_hook).  Every rank returns the same poses."""
    from . import dist as pd
    t0 = time.time()
    all_edges = pose_graph_edges(poses, cfg, aa_to_R)
    t_edges = time.time() - t0
    weights = np.zeros(len(frames))
    for (i, j) in all_edges:
        weights[i] += len(frames[j]["surfFlat"]) + 4 * len(frames[j]["cornerLessSharp"])
    bounds = pd.shard_frames_by_weight(weights, world)
    fr = (int(bounds[rank]), int(bounds[rank + 1]))
    er, en = pd.global_edge_list([e[0] for e in all_edges], [e[1] for e in all_edges])
    ctx.blocks_set_edge_list(er, en)
    hook = pd.install_allreduce_hook(ctx, group)
    try:
        if device_blocks and cfg.point_to_plane:
            # the rank's own point-to-plane correspondences become residual blocks on the device, filed under the GLOBAL edge list (no download / rebuild / upload)
            bl, _, mine = build_problem(ctx, frames, poses, cfg, aa_to_R, frame_range=fr, host_point2plane=False, all_edges=all_edges, device_line_blocks=device_line_blocks)
            ref, nei = np.array([e[0] for e in mine], np.int32), np.array([e[1] for e in mine], np.int32)
            bl.n = ctx.frames_point2plane_blocks(poses, ref, nei, cfg.plane_tolerance, cfg.plane_dis_threshold, cfg.angle_residual, cfg.normalize_distance, cfg.plane_weight,
                                                 len(frames), extra=bl.view())
        else:
            bl, _ = build_problem(ctx, frames, poses, cfg, aa_to_R, frame_range=fr, all_edges=all_edges)
            v = bl.view()
            ctx.blocks_set(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"], len(frames))
        mask = np.zeros(len(frames), np.uint8); mask[first_valid_frame(frames)] = 1
        t_lm = time.time()
        new_poses, summary = ctx.blocks_solve_lm(poses, mask, cfg.max_lm_iterations)
        summary["lm_s"] = time.time() - t_lm
        summary["edges_s"], summary["build_s"] = t_edges, t_lm - t0 - t_edges
    finally:
        pd.remove_allreduce_hook(ctx)
        ctx.blocks_set_edge_list(None)
    summary["n_blocks_local"], summary["n_edges"], summary["allreduces"] = bl.n, len(all_edges), hook["calls"]["n"]
    summary["frame_range"] = [int(bounds[rank]), int(bounds[rank + 1])]
    return new_poses, summary


def estimate_pose(ctx: Context, frames, poses, cfg: OdometryConfig, aa_to_R, max_iteration=7, refine_fn=None):
    """EstimatePose's outer loop with its early exits (LidarOdometry.cpp:166-183): RefinePose up to max_iteration times; stop when the cost changed by
    less than 1 % of the PREVIOUS cost (the first comparison divides by last_cost = 0 and never fires) or when two consecutive iterations took fewer
    than 5 successful LM steps (last_step starts at INT16_MAX)."""
    refine_fn = refine_fn or refine_pose
    last_cost, last_step, log = 0.0, 2 ** 15 - 1, []
    for it in range(max_iteration):
        poses, s = refine_fn(ctx, frames, poses, cfg, aa_to_R)
        log.append(s)
        curr_cost, curr_step = s["final_cost"], s["successful"]
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(np.float64(curr_cost) - last_cost) / np.float64(last_cost)
        if rel < 0.01:
            break
        if curr_step < 5 and last_step < 5:
            break
        last_cost, last_step = curr_cost, curr_step
    return poses, log
