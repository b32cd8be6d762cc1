"""Host loop of the joint camera-LiDAR refinement: a Python mirror of CameraLidarOptimizer::Optimize / AssociateLineMulti
(joint_optimization/CameraLidarOptimizer.cpp:330-548) driving the C ABI.  Pose blocks are laid out [cameras 0..n) | LiDARs n..2n):
camera-camera reprojection observations (pvb_reproj_set), camera-LiDAR line pairs (AssociateByAngle on the device + builders) and the LiDAR-LiDAR
blocks of RefinePose (panovlm_b200.odometry.build_problem) meet in ONE trust-region problem solved by pvb_joint_solve_lm."""
import numpy as np

from . import odometry
from .api import BlockList, Context, LineFrame


class JointConfig:
    def __init__(self, camera_weight=1.0, camera_lidar_weight=1.0, lidar=None, refine_camera_rotation=True, refine_camera_trans=True, refine_lidar_rotation=True,
                 refine_lidar_trans=True, refine_structure=True, max_lm_iterations=50):      # SetOptionsSfM leaves Ceres' default max_num_iterations = 50 (util/Optimization.cpp:611-636)
        self.camera_weight, self.camera_lidar_weight = camera_weight, camera_lidar_weight        # config/Room.txt:81-83
        self.lidar = lidar or odometry.OdometryConfig(line_to_line=False)
        self.refine_camera_rotation, self.refine_camera_trans = refine_camera_rotation, refine_camera_trans
        self.refine_lidar_rotation, self.refine_lidar_trans, self.refine_structure = refine_lidar_rotation, refine_lidar_trans, refine_structure
        self.max_lm_iterations = max_lm_iterations


def joint_lidar_config(lidar: odometry.OdometryConfig):
    """The LiDAR-LiDAR builders as CameraLidarOptimizer::Optimize calls them (CameraLidarOptimizer.cpp:436-458) - which is NOT how RefinePose calls them:
    * AddLidarPointToLineResidual is called WITHOUT its `use_segment` argument (:441-443), so every later argument lands one slot early: use_segment <- angle_residual,
      angle_residual <- normalize_distance, normalized_distance <- (lidar_weight != 0), weight <- default 1.0;
    * AddLidarLineToLineResidual2 runs whenever line_to_line_residual is set (no `use_segment` gate) and gets no weight (1.0);
    * AddLidarPointToPlaneResidual gets weight = lidar_weight.
    Returns (config for the point-to-line family, config for the other two).  Pinned against the reference's own Optimize (tests/test_reference_pinning.py)."""
    import copy
    main = copy.copy(lidar)
    main.plane_weight, main.point_to_line = lidar.lidar_weight, False
    p2l = copy.copy(lidar)
    p2l.point_to_plane, p2l.line_to_line = False, False
    p2l.use_segment, p2l.angle_residual, p2l.normalize_distance, p2l.point_line_weight = bool(lidar.angle_residual), bool(lidar.normalize_distance), lidar.lidar_weight != 0, 1.0
    return p2l, main


def _lidar_blocks(ctx, frames, lidars, lidar_cfg, aa_to_R, host_point2plane=True):
    """The LiDAR-LiDAR blocks of the joint stage in the reference's registration order: point-to-line (if enabled), then line-to-line, then point-to-plane."""
    p2l, main = joint_lidar_config(lidar_cfg)
    res = odometry.build_problem(ctx, frames, lidars, main, aa_to_R, host_point2plane=host_point2plane)
    if lidar_cfg.point_to_line:
        pre, _ = odometry.build_problem(ctx, frames, lidars, p2l, aa_to_R)
        a, b = pre.view(), res[0].view()
        merged = BlockList(pre.n + res[0].n + 1)
        merged.extend({k: np.concatenate([a[k], b[k]]) for k in a})
        res = (merged,) + tuple(res[1:])
    return res, main


def _T_from_block(block, aa_to_R):
    T = np.eye(4)
    T[:3, :3] = aa_to_R(block[:3])
    T[:3, 3] = block[3:]
    return T


def associate_lines(ctx: Context, frames, image_lines, cams, lidars, rows, cols, aa_to_R, neighbors=None, lidar_masks=None, image_masks=None):
    """AssociateLineMulti (CameraLidarOptimizer.cpp:330-384): image i against every LiDAR of neighbors[i] at T_cl = T_cw T_wl (:347-353), optionally
    restricted to the LiDAR lines / image lines that belong to tracks (lidar_masks from Context.lidar_mask_by_track, :339-342).  neighbors = None:
    neighbor_size_joint = 1 in temporal mode, i.e. LiDAR i (Context.neighbor_each_frame gives the general lists)."""
    pairs = {}
    n = len(frames)
    if neighbors is None:
        neighbors = Context.neighbor_each_frame(n, n, 1, True)
    for i in range(n):
        T_cw = _T_from_block(cams[i], aa_to_R)
        for li in neighbors[i]:
            f = frames[li]
            T_cl = T_cw @ np.linalg.inv(_T_from_block(lidars[li], aa_to_R))
            if len(f["segment_coeffs"]) == 0:                            # no LiDAR segments: the pixel-space Associate(lines, cornerLessSharp, T_cl) (:366-367)
                il, s, e, ang = ctx.pixel_associate(rows, cols, image_lines[i], f["cornerLessSharp"], T_cl)
                pairs[(i, li)] = (il, np.full(len(il), -1, np.int32), s, e, ang)
                continue
            lf = LineFrame(f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["end_points"], np.eye(3), np.zeros(3))
            il, ll, s, e, ang = ctx.camera_lidar_associate(rows, cols, image_lines[i], lf, T_cl, True, True, None if image_masks is None else image_masks[i],
                                                           None if lidar_masks is None else lidar_masks[li])
            pairs[(i, li)] = (il, ll, s, e, ang)
    return pairs


def build_problem(ctx: Context, data, cams, lidars, points, cfg: JointConfig, aa_to_R):
    """Steps 2-4 of Optimize (:418-460): returns the residual-block arrays over the pose blocks [cameras | LiDARs] and the line pairs."""
    frames, n = data["frames"], len(data["frames"])
    pairs = associate_lines(ctx, frames, data["image_lines"], cams, lidars, data["rows"], data["cols"], aa_to_R)
    n_pairs = sum(len(p[0]) for p in pairs.values())
    bl_cl = BlockList(2 * n_pairs + 16)
    for (ci, li), (il, ll, s, e, ang) in pairs.items():                                    # AddCameraLidarResidual (util/Optimization.cpp:564-607)
        if len(il):
            Context.build_camera_lidar_blocks(bl_cl, data["rows"], data["cols"], data["image_lines"][ci][il], s, e, np.ones(len(il), np.float32), ci, n + li,
                                              cfg.camera_lidar_weight)
    (bl_ll, _), _ = _lidar_blocks(ctx, frames, lidars, cfg.lidar, aa_to_R)                # step 4: the LiDAR-LiDAR blocks, called as Optimize calls the builders
    a, b = bl_cl.view(), bl_ll.view()
    out = {k: np.concatenate([a[k], b[k] + n if k in ("ref", "nei") else b[k]]) for k in ("type", "ref", "nei", "normalize", "huber", "consts")}
    return out, pairs, (bl_cl.n, bl_ll.n)


def optimize(ctx: Context, data, cams, lidars, points, cfg: JointConfig, aa_to_R, device_blocks=True):
    """One call of CameraLidarOptimizer::Optimize: build the three residual families at the current estimate and solve.
    device_blocks: the LiDAR point-to-plane correspondences become residual blocks on the device (Context.frames_point2plane_blocks); only the
    camera-LiDAR blocks and the other LiDAR families are built on the host and appended."""
    n = len(data["frames"])
    poses = np.concatenate([cams, lidars])
    if device_blocks and cfg.lidar.point_to_plane:
        frames = data["frames"]
        pairs = associate_lines(ctx, frames, data["image_lines"], cams, lidars, data["rows"], data["cols"], aa_to_R)
        bl_cl = BlockList(2 * sum(len(p[0]) for p in pairs.values()) + 16)
        for (ci, li), (il, ll, s, e, ang) in pairs.items():
            if len(il):
                Context.build_camera_lidar_blocks(bl_cl, data["rows"], data["cols"], data["image_lines"][ci][il], s, e, np.ones(len(il), np.float32), ci, n + li,
                                                  cfg.camera_lidar_weight)
        (bl_ll, _, mine), main_cfg = _lidar_blocks(ctx, frames, lidars, cfg.lidar, aa_to_R, host_point2plane=False)
        a, b = bl_cl.view(), bl_ll.view()
        extra = {k: np.concatenate([a[k], b[k] + n if k in ("ref", "nei") else b[k]]) for k in ("type", "ref", "nei", "normalize", "huber", "consts")}
        ref, nei = np.array([e[0] for e in mine], np.int32), np.array([e[1] for e in mine], np.int32)
        n_total = ctx.frames_point2plane_blocks(lidars, ref, nei, cfg.lidar.plane_tolerance, cfg.lidar.plane_dis_threshold, cfg.lidar.angle_residual,
                                                cfg.lidar.normalize_distance, main_cfg.plane_weight, 2 * n, block_offset=n, extra=extra)
        v, counts = None, (bl_cl.n, n_total - bl_cl.n)
    else:
        v, pairs, counts = build_problem(ctx, data, cams, lidars, points, cfg, aa_to_R)
        ctx.blocks_set(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"], 2 * n)
    ctx.reproj_set(data["cam"], data["point"], data["bearing"], n, len(points), weight=cfg.camera_weight, huber=4.0 * np.pi / 180.0)   # :431-432
    const = np.zeros((2 * n, 6), np.uint8)
    const[:n, :3] = 0 if cfg.refine_camera_rotation else 1                               # :466-476
    const[:n, 3:] = 0 if cfg.refine_camera_trans else 1
    const[n:, :3] = 0 if cfg.refine_lidar_rotation else 1                                # :478-488
    const[n:, 3:] = 0 if cfg.refine_lidar_trans else 1
    const[0] = 1                                                                         # :490-491 camera 0 constant
    pt_const = None if cfg.refine_structure else np.ones(len(points), np.uint8)          # :462-465
    if const.all() and pt_const is None:
        # Optimize(refine_* = false, refine_structure = true) is a structure-only bundle adjustment in the reference: every pose block constant, the points free
        _, new_points, summary = ctx.reproj_solve_lm(cams, points, np.ones((n, 6), np.uint8), None, cfg.max_lm_iterations)
        new_poses = poses.copy()
    else:
        new_poses, new_points, summary = ctx.joint_solve_lm(poses, points, const, pt_const, cfg.max_lm_iterations)
    summary.update(n_camera_lidar_blocks=counts[0], n_lidar_blocks=counts[1], n_reproj=len(data["cam"]), n_line_pairs=sum(len(p[0]) for p in pairs.values()))
    return new_poses[:n], new_poses[n:], new_points, summary, (v, const, pt_const)


def calibrate(ctx: Context, frames, image_lines, rows, cols, T_cl_init, aa_to_R, R_to_aa, max_iterations=35, associate_fn=None, solve_fn=None):
    """Calibration mode of CameraLidarOptimizer::JointOptimize (CameraLidarOptimizer.cpp:195-233): one relative pose T_cl for all (image i, LiDAR i)
    pairs.  Per iteration: AssociateLineSingle at the current T_cl (:301-317, AssociateByAngle with its defaults: one-to-one pairs), the residual blocks of
    Optimize(line_pairs, T_cl) (:32-64) on the single pose block, LM with max_num_iterations = 50 (:68), stop when the rotation changed by less than
    0.1 deg and the translation by less than 0.01 (:229).
    associate_fn(T_cl) -> per frame (image line ids, LiDAR line ids, start, end, score) and solve_fn(blocks, pose) -> (new pose, summary) replace the device association
    and the device LM (used by the CPU test that runs this loop next to the reference's own JointOptimize with a scripted solver)."""
    T = np.array(T_cl_init, dtype=np.float64)
    if associate_fn is None:
        lfs = [LineFrame(f["cornerLessSharp"], f["p2s_off"], f["p2s_ids"], f["segment_coeffs"], f["end_points"], np.eye(3), np.zeros(3)) if len(f["segment_coeffs"]) else None
               for f in frames]

        def associate(T_cl):                                                   # AssociateLineSingle (:301-317); frames without LiDAR segments take the pixel-space Associate (:313-314)
            out = []
            for i, f in enumerate(frames):
                if len(f["segment_coeffs"]) == 0:
                    il, s, e, ang = ctx.pixel_associate(rows, cols, image_lines[i], f["cornerLessSharp"], T_cl)
                    out.append((il, np.full(len(il), -1, np.int32), s, e, ang))
                else:
                    out.append(ctx.camera_lidar_associate(rows, cols, image_lines[i], lfs[i], T_cl, True, False))
            return out
    else:
        associate = associate_fn

    pairs = associate(T)
    log = []
    for it in range(max_iterations):
        bl = BlockList(2 * sum(len(p[0]) for p in pairs) + 4)
        for i, (il, ll, s, e, ang) in enumerate(pairs):
            if len(il):
                Context.build_calibration_blocks(bl, rows, cols, image_lines[i][il], s, e, 0)
        v = bl.view()
        pose = np.concatenate([R_to_aa(T[:3, :3]), T[:3, 3]])[None, :]
        if solve_fn is None:
            ctx.blocks_set(v["type"], v["ref"], v["nei"], v["consts"], v["huber"], v["normalize"], 1)
            new_pose, summary = ctx.blocks_solve_lm(pose, None, 50)
        else:
            new_pose, summary = solve_fn(v, pose)
        T_new = np.eye(4)
        T_new[:3, :3] = aa_to_R(new_pose[0, :3])
        T_new[:3, 3] = new_pose[0, 3:]
        rot_change = np.float32(np.arccos(np.clip((np.trace(T[:3, :3].T @ T_new[:3, :3]) - 1) / 2.0, -1.0, 1.0))) * np.float32(180.0 / np.pi)
        trans_change = np.float32(np.linalg.norm(T[:3, 3] - T_new[:3, 3]))
        T = T_new
        pairs = associate(T)
        summary.update(n_pairs=bl.n // 2, rotation_change_deg=float(rot_change), translation_change=float(trans_change))
        log.append(summary)
        if rot_change < 0.1 and trans_change < 0.01:
            break
    return T, log


def joint_optimize(ctx: Context, data, cams, lidars, points, cfg: JointConfig, aa_to_R, num_iteration_joint=5, optimize_fn=None):
    """Mapping mode of CameraLidarOptimizer::JointOptimize (CameraLidarOptimizer.cpp:236-287): Optimize (which re-associates the lines at the current
    estimate) repeated up to num_iteration_joint times with the reference's two early exits - the cost changed by less than 1 % (:271, compared with the
    previous iteration; the first comparison divides by last_cost = 0 and never fires) or fewer than 5 successful LM steps in two consecutive
    iterations (:276; last_step starts at INT32_MAX)."""
    optimize_fn = optimize_fn or optimize
    last_cost, last_step = 0.0, 2 ** 31 - 1
    log = []
    for it in range(num_iteration_joint):
        cams, lidars, points, summary, _ = optimize_fn(ctx, data, cams, lidars, points, cfg, aa_to_R)
        log.append(summary)
        curr_cost, curr_step = summary["final_cost"], summary["successful"]
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(np.float64(curr_cost) - last_cost) / np.float64(last_cost)
        if rel < 0.01:
            break
        if curr_step < 5 and last_step < 5:
            break
        last_cost, last_step = curr_cost, curr_step
    return cams, lidars, points, log
