"""Synthetic inputs for the hot path (numpy only): the workloads of BASELINE.json's configs.

There is no network for the Room/Floor datasets, so the benchmark and the parity tests run on seeded
synthetic data of the same shape: VLP-16 ray casts (16 rings x 1800 azimuth steps, sensor axes
X-right / Y-down / Z-forward as after sensors/Velodyne.cpp:125-132) inside a box room with pillars, LOAM-style
feature subsets (surfFlat <= 16 rings x 6 sectors x 4, surfLessFlat = 0.2 m voxel grid of the rest, corner
segments on vertical edges), and a dense multi-room floor plan for the HBM-roofline stress sweep.
"""
import numpy as np

POINT_NORMAL = 1.0   # sensors/Velodyne.h point classes carried in PointXYZI.intensity
POINT_GROUND = 16.0


def rotvec_to_R(a):
    a = np.asarray(a, dtype=np.float64)
    th = np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


# ------------------------------------------------------------------------------------------------
# box room + pillars, VLP-16 ray cast
# ------------------------------------------------------------------------------------------------
ROOM_MIN = np.array([-4.0, -1.5, -3.0])   # x, y (down: ceiling at -1.5, floor at +1.5), z
ROOM_MAX = np.array([4.0, 1.5, 3.0])
PILLARS = [(-2.0, -1.2), (2.2, -1.0), (-1.8, 1.4), (1.9, 1.5)]   # (x, z) centres, 0.4 m square, full height
PILLAR_HALF = 0.2


def _ray_box(o, d, bmin, bmax):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0 = (bmin - o) * inv
        t1 = (bmax - o) * inv
    tmin = np.minimum(t0, t1).max(axis=1)
    tmax = np.maximum(t0, t1).min(axis=1)
    return tmin, tmax


def vlp16_dirs(n_az=1800):
    elev = np.deg2rad(np.arange(-15, 16, 2.0))            # 16 rings
    az = np.arange(n_az) * (2 * np.pi / n_az)
    E, A = np.meshgrid(elev, az, indexing="ij")           # ring-major
    d = np.stack([np.cos(E) * np.sin(A), -np.sin(E), np.cos(E) * np.cos(A)], axis=-1)
    return d.reshape(-1, 3), np.repeat(np.arange(16), n_az), np.tile(np.arange(n_az), 16)


def raycast_room(R_wl, t_wl, rng, n_az=1800, noise=0.01):
    """Returns local-frame points (n,3 float64), ring ids, azimuth ids of one VLP-16 sweep."""
    dl, ring, azi = vlp16_dirs(n_az)
    dw = dl @ R_wl.T
    o = np.broadcast_to(t_wl, dw.shape)
    _, t_room = _ray_box(o, dw, ROOM_MIN, ROOM_MAX)       # inside the box: exit distance
    t_hit = t_room.copy()
    for (px, pz) in PILLARS:
        bmin = np.array([px - PILLAR_HALF, ROOM_MIN[1], pz - PILLAR_HALF])
        bmax = np.array([px + PILLAR_HALF, ROOM_MAX[1], pz + PILLAR_HALF])
        tin, tout = _ray_box(o, dw, bmin, bmax)
        hit = (tin < tout) & (tin > 0)
        t_hit = np.where(hit & (tin < t_hit), tin, t_hit)
    rng_m = t_hit + rng.normal(0.0, noise, size=t_hit.shape)
    keep = rng_m > 0.5                                    # Velodyne.cpp:123 drops range < 0.5 m
    pts = dl * rng_m[:, None]
    return pts[keep], ring[keep], azi[keep]


def _voxel_downsample(p, leaf=0.2):
    key = np.floor(p / leaf).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.ravel()
    cnt = np.bincount(inv)
    out = np.stack([np.bincount(inv, weights=p[:, k]) / cnt for k in range(3)], axis=1)
    return out


def _edge_segments(R_wl, t_wl, rng, noise=0.01):
    """Corner points on the vertical pillar / wall edges, one point per ring that crosses the edge."""
    edges = []
    for (px, pz) in PILLARS:
        for sx in (-1, 1):
            for sz in (-1, 1):
                edges.append((px + sx * PILLAR_HALF, pz + sz * PILLAR_HALF))
    for x in (ROOM_MIN[0], ROOM_MAX[0]):
        for z in (ROOM_MIN[2], ROOM_MAX[2]):
            edges.append((x, z))
    elev = np.deg2rad(np.arange(-15, 16, 2.0))
    segs = []
    for (ex, ez) in edges:
        # vertical line (ex, y, ez) in world; find y where the ray from the sensor at each ring elevation meets it
        # (assumes small sensor tilt: solve in the sensor frame exactly)
        pts = []
        for e in elev:
            # param: world point q(y) = (ex, y, ez); local = R^T (q - t); want -local_y / hypot(local_x, local_z) = tan(e)
            ys = np.linspace(ROOM_MIN[1], ROOM_MAX[1], 601)
            q = np.stack([np.full_like(ys, ex), ys, np.full_like(ys, ez)], axis=1)
            l = (q - t_wl) @ R_wl
            f = -l[:, 1] / np.hypot(l[:, 0], l[:, 2]) - np.tan(e)
            s = np.where(np.sign(f[:-1]) != np.sign(f[1:]))[0]
            if len(s) == 0:
                continue
            i = s[0]
            w = f[i] / (f[i] - f[i + 1])
            pts.append(l[i] * (1 - w) + l[i + 1] * w)
        if len(pts) >= 5:
            P = np.array(pts) + rng.normal(0, noise, size=(len(pts), 3))
            if np.linalg.norm(P.max(0) - P.min(0)) > 0.3:
                segs.append(P)
    return segs


def fit_segment(P):
    """(centroid, unit PCA direction) + extreme points projected on the line (Velodyne.cpp:1275-1282)."""
    c = P.mean(0)
    w, v = np.linalg.eigh((P - c).T @ (P - c))
    d = v[:, 2] / np.linalg.norm(v[:, 2])
    s = (P - c) @ d
    return np.concatenate([c, d]), np.stack([c + s.min() * d, c + s.max() * d])


def make_frame(R_wl, t_wl, seed, n_az=1800, ground_class=False):
    """One synthetic LiDAR frame in the SENSOR frame with the feature clouds the hot path consumes."""
    rng = np.random.default_rng(seed)
    pts, ring, azi = raycast_room(R_wl, t_wl, rng, n_az)
    # LOAM curvature on each ring (Velodyne.cpp ExtractFeatures): |sum_{j=-5..5} (p_j - p_0)|^2
    curv = np.full(len(pts), np.inf)
    for r in range(16):
        idx = np.where(ring == r)[0]
        if len(idx) < 11:
            continue
        P = pts[idx]
        k = np.zeros_like(P)
        for j in range(-5, 6):
            if j:
                k += np.roll(P, -j, axis=0) - P
        c = (k ** 2).sum(1)
        c[:5] = np.inf
        c[-5:] = np.inf
        curv[idx] = c
    sector = (azi * 6) // n_az
    flat_idx = []
    for r in range(16):
        for s in range(6):
            idx = np.where((ring == r) & (sector == s))[0]
            if len(idx) == 0:
                continue
            order = idx[np.argsort(curv[idx])][:4]
            flat_idx.extend(order[np.isfinite(curv[order])].tolist())
    flat_idx = np.array(sorted(flat_idx), dtype=np.int64)
    mask = np.ones(len(pts), dtype=bool)
    mask[flat_idx] = False
    less_flat = _voxel_downsample(pts[mask], 0.2)

    def cls(p):
        if not ground_class:
            return np.full(len(p), POINT_NORMAL)
        pw = p @ R_wl.T + t_wl
        return np.where(pw[:, 1] > ROOM_MAX[1] - 0.1, POINT_GROUND, POINT_NORMAL)

    segs = _edge_segments(R_wl, t_wl, rng)
    corner, p2s_off, p2s_ids, coeffs, ends, seg_off = [], [0], [], [], [], [0]
    for si, P in enumerate(segs):
        co, en = fit_segment(P)
        coeffs.append(co)
        ends.append(en)
        for p in P:
            corner.append(p)
            p2s_ids.append(si)
            p2s_off.append(len(p2s_ids))
        seg_off.append(len(corner))
    n_clutter = 40                                         # unsegmented corner points (empty point_to_segment sets)
    clutter = pts[rng.choice(len(pts), n_clutter, replace=False)]
    for p in clutter:
        corner.append(p)
        p2s_off.append(len(p2s_ids))
    corner = np.array(corner).reshape(-1, 3)

    def xyzi(p, c):
        return np.concatenate([p, np.asarray(c).reshape(-1, 1)], axis=1).astype(np.float32)

    S = len(segs)
    return dict(
        R_wl=R_wl.copy(), t_wl=np.asarray(t_wl, dtype=np.float64).copy(),
        cloud=xyzi(pts, np.ones(len(pts))),
        surfFlat=xyzi(pts[flat_idx], cls(pts[flat_idx])),
        surfLessFlat=xyzi(less_flat, cls(less_flat)),
        cornerLessSharp=xyzi(corner, np.ones(len(corner))),
        p2s_off=np.array(p2s_off, dtype=np.int32), p2s_ids=np.array(p2s_ids, dtype=np.int32),
        segment_coeffs=np.array(coeffs).reshape(S, 6), end_points=np.array(ends).reshape(S, 2, 3),
        seg_off=np.array(seg_off, dtype=np.int32),          # edge_segmented[s] = cornerLessSharp[seg_off[s]:seg_off[s+1]]
    )


def make_pair(seed=20260925, n_az=1800, ground_class=False):
    """configs[0]: frame A at the origin, frame B = A o (rotvec ~ N(0,0.02^2), t ~ N(0,0.05^2))."""
    rng = np.random.default_rng(seed)
    RA, tA = np.eye(3), np.zeros(3)
    RB, tB = rotvec_to_R(rng.normal(0, 0.02, 3)), rng.normal(0, 0.05, 3)
    return make_frame(RA, tA, seed + 1, n_az, ground_class), make_frame(RB, tB, seed + 2, n_az, ground_class)


def make_sequence(n_frames, seed=20260926, n_az=1800, radius=1.2, tilt=0.0):
    """configs[1]-like: frames along a closed loop inside the room.
    tilt (rad): the sensor is additionally pitched / rolled by up to `tilt` along the loop (a hand-held or backpack rig).  With tilt = 0 the
    +-15 degree rings of a level VLP-16 at the room centre only ever see the walls and pillars - never the floor or the ceiling - so the
    vertical translation of a frame is not observable from point-to-plane / line-to-line constraints (the optimiser is free to slide along it)."""
    rng = np.random.default_rng(seed)
    frames = []
    for i in range(n_frames):
        a = 2 * np.pi * i / n_frames
        t = np.array([radius * np.cos(a), rng.normal(0, 0.02), 0.6 * radius * np.sin(a)])
        R = rotvec_to_R(np.array([rng.normal(0, 0.01), a * 0.5 + rng.normal(0, 0.01), rng.normal(0, 0.01)]))
        if tilt != 0.0:
            R = R @ rotvec_to_R(np.array([tilt * np.cos(3 * a), 0.0, tilt * np.sin(2 * a + 0.7)]))
        frames.append(make_frame(R, t, seed + 10 + i, n_az))
    return frames


# ------------------------------------------------------------------------------------------------
# dense multi-room floor plan (configs[4]: HBM-roofline stress)
# ------------------------------------------------------------------------------------------------
def _floor_plan_rects(L=200.0, W=50.0, H=4.0, room=10.0):
    """List of axis-aligned rectangles (origin, edge_u, edge_v) covering floor, ceiling, outer and inner walls."""
    rects = [(np.array([0, 0, 0.0]), np.array([L, 0, 0]), np.array([0, W, 0])),
             (np.array([0, 0, H]), np.array([L, 0, 0]), np.array([0, W, 0]))]
    xs = np.arange(0, L + 1e-9, room)
    ys = np.arange(0, W + 1e-9, room)
    for x in xs:
        rects.append((np.array([x, 0, 0.0]), np.array([0, W, 0]), np.array([0, 0, H])))
    for y in ys:
        rects.append((np.array([0, y, 0.0]), np.array([L, 0, 0]), np.array([0, 0, H])))
    return rects


def sample_floor_plan(n, rng, xlim=None, noise=0.0, L=200.0, W=50.0, H=4.0):
    """n points on the floor-plan surfaces (uniform by area), optionally restricted to xlim=(x0,x1)."""
    rects = _floor_plan_rects(L, W, H)
    if xlim is not None:
        clipped = []
        for (o, u, v) in rects:
            lo, hi = o[0], o[0] + u[0] + v[0]
            a, b = max(lo, xlim[0]), min(hi, xlim[1])
            if u[0] + v[0] == 0:                      # rectangle at constant x
                if xlim[0] <= o[0] <= xlim[1]:
                    clipped.append((o, u, v))
                continue
            if b <= a:
                continue
            o2 = o.copy(); o2[0] = a
            u2 = u.copy()
            v2 = v.copy()
            if u[0] > 0:
                u2[0] = b - a
            else:
                v2[0] = b - a
            clipped.append((o2, u2, v2))
        rects = clipped
    area = np.array([np.linalg.norm(np.cross(u, v)) for (_, u, v) in rects])
    counts = rng.multinomial(n, area / area.sum())
    out = np.empty((n, 3), dtype=np.float64)
    pos = 0
    for (o, u, v), c in zip(rects, counts):
        if c == 0:
            continue
        a = rng.random((c, 1))
        b = rng.random((c, 1))
        out[pos:pos + c] = o + a * u + b * v
        pos += c
    if noise > 0:
        out += rng.normal(0, noise, size=out.shape)
    return out


def make_dense_sweep(n_target=10_000_000, n_frames=64, pts_per_frame=156_250, seed=20260929, L=200.0, source_seed=None):
    """configs[4]: target cloud (world == target frame) + n_frames source frames, each a local scan of a
    slab of the building (like a LiDAR at that position), expressed in its own sensor frame whose true pose is
    a small perturbation; the initial guess handed to the optimiser is the unperturbed slab pose.
    Returns dict(target[n,4] f32, src_local[m,4] f32, src_off[n_frames+1] i32, poses_lw_init[n_frames,6],
    poses_lw_true[n_frames,6])."""
    rng = np.random.default_rng(seed)                    # target cloud
    scale = n_target / 10_000_000.0
    Lx = max(10.0, L * scale)                          # keep the surface density (~400 pts/m^2) when scaled down
    # world origin inside the building, above the floor (a plane through the origin is degenerate for the reference's
    # 'A x = -1' plane parameterisation, Geometry.hpp:345-373; real sensor-frame data never has it)
    origin = np.array([0.37 * Lx, 23.3, 1.6])
    tgt = sample_floor_plan(n_target, rng, xlim=(0, Lx), L=L) - origin
    target = np.concatenate([tgt, np.ones((n_target, 1))], axis=1).astype(np.float32)
    rng = np.random.default_rng(seed + 1 if source_seed is None else source_seed)   # source frames (per-rank seed in bench.py)
    src, off, init, true = [], [0], [], []
    from scipy.spatial.transform import Rotation
    slab = Lx / n_frames
    for f in range(n_frames):
        x0, x1 = f * slab, (f + 1) * slab
        pw = sample_floor_plan(pts_per_frame, rng, xlim=(x0, x1), noise=0.01, L=L) - origin
        # sensor pose: centre of the slab, true pose = nominal o small perturbation
        c = np.array([(x0 + x1) / 2, 25.0, 2.0]) - origin
        R_nom = Rotation.from_rotvec([0, 0, 0.3 * np.sin(f)]).as_matrix()
        dR = Rotation.from_rotvec(rng.normal(0, 0.004, 3)).as_matrix()
        dt = rng.normal(0, 0.02, 3)
        R_true, t_true = R_nom @ dR, c + dt
        pl = (pw - t_true) @ R_true                    # R^T (p - t)
        src.append(np.concatenate([pl, np.ones((len(pl), 1))], axis=1).astype(np.float32))
        off.append(off[-1] + len(pl))
        for (R, t, dst) in ((R_nom, c, init), (R_true, t_true, true)):
            R_lw = R.T
            dst.append(np.concatenate([Rotation.from_matrix(R_lw).as_rotvec(), -R_lw @ t]))
    return dict(target=target, src_local=np.concatenate(src), src_off=np.array(off, dtype=np.int32),
                poses_lw_init=np.array(init), poses_lw_true=np.array(true), extent=(Lx, 50.0, 4.0))


# ------------------------------------------------------------------------------------------------
# panoramic structure-from-motion problem (the camera-camera term of the joint optimisation)
# ------------------------------------------------------------------------------------------------
def make_ba_problem(n_cams=12, n_points=300, track_len=(2, 8), pixel_noise_rad=2e-3, pose_noise=(0.01, 0.03), point_noise=0.03, seed=20261001, rows=2880, cols=5760):
    """Cameras on a loop inside the box room, points on its walls, every point seen by `track_len` cameras (panoramas see all directions).
    Returns the ground truth, perturbed initial values (camera blocks (aa_cw, t_cw), points) and the observation list with unit-sphere
    bearings + the same key points as pixels (for the observation builder)."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 2 * np.pi, n_cams, endpoint=False)
    cams_gt = np.zeros((n_cams, 6))
    Rs = []
    for c in range(n_cams):
        R_wc = rotvec_to_R(np.array([0.0, ang[c] * 0.3, 0.0]) + rng.normal(size=3) * 0.05)
        t_wc = np.array([1.5 * np.cos(ang[c]), rng.normal() * 0.05, 1.0 * np.sin(ang[c])])
        R_cw = R_wc.T
        # angle-axis of R_cw through the matrix logarithm
        th = np.arccos(np.clip((np.trace(R_cw) - 1) / 2, -1, 1))
        ax = np.array([R_cw[2, 1] - R_cw[1, 2], R_cw[0, 2] - R_cw[2, 0], R_cw[1, 0] - R_cw[0, 1]])
        aa = ax / (2 * np.sin(th)) * th if th > 1e-9 else ax / 2
        cams_gt[c, :3], cams_gt[c, 3:] = aa, -R_cw @ t_wc
        Rs.append(rotvec_to_R(aa))
    # points on the six faces of the room
    face = rng.integers(0, 6, n_points)
    pts = rng.uniform(ROOM_MIN, ROOM_MAX, size=(n_points, 3))
    for a in range(3):
        pts[face == 2 * a, a] = ROOM_MIN[a]
        pts[face == 2 * a + 1, a] = ROOM_MAX[a]
    cam, point, bearing, pix = [], [], [], []
    for p in range(n_points):
        L = int(rng.integers(track_len[0], min(track_len[1], n_cams) + 1))
        for c in np.sort(rng.choice(n_cams, size=L, replace=False)):
            Pc = Rs[c] @ pts[p] + cams_gt[c, 3:]
            d = Pc / np.linalg.norm(Pc) + rng.normal(size=3) * pixel_noise_rad
            d /= np.linalg.norm(d)
            cam.append(c); point.append(p); bearing.append(d)
            lon, lat = np.arctan2(d[0], d[2]), -np.arcsin(d[1])
            pix.append([cols * (0.5 + lon / (2 * np.pi)), rows * (0.5 - lat / np.pi)])
    cams0 = cams_gt.copy()
    cams0[1:, :3] += rng.normal(size=(n_cams - 1, 3)) * pose_noise[0]
    cams0[1:, 3:] += rng.normal(size=(n_cams - 1, 3)) * pose_noise[1]
    pts0 = pts + rng.normal(size=pts.shape) * point_noise
    return {"cams_gt": cams_gt, "points_gt": pts, "cams": cams0, "points": pts0, "cam": np.array(cam, np.int32), "point": np.array(point, np.int32),
            "bearing": np.array(bearing), "pixels": np.array(pix, np.float32), "rows": rows, "cols": cols}


# ------------------------------------------------------------------------------------------------
# joint camera-LiDAR problem (configs[2]: Room joint refinement)
# ------------------------------------------------------------------------------------------------
def _R_to_aa(R):
    th = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return ax / (2 * np.sin(th)) * th if th > 1e-9 else ax / 2


def pixel_of(cam_pts, rows, cols):
    """Equirectangular projection (sensors/Equirectangular.h:41-96, exact atan2): camera-frame points -> pixels."""
    p = np.asarray(cam_pts, dtype=np.float64).reshape(-1, 3)
    lon = np.arctan2(p[:, 0], p[:, 2])
    lat = -np.arctan2(p[:, 1], np.hypot(p[:, 0], p[:, 2]))
    return np.stack([cols * (0.5 + lon / (2 * np.pi)), rows * (0.5 - lat / np.pi)], axis=1)


def make_joint_problem(n_frames=8, n_points=400, n_az=600, seed=20261005, rows=2880, cols=5760, clutter=20, track_len=(2, 6), pixel_noise=2.0, bearing_noise=2e-3,
                       pose_noise=(0.004, 0.02), point_noise=0.03):
    """LiDAR frames on a loop (make_sequence) with a panoramic camera rigidly mounted on the sensor (T_cl), image lines = projections of the
    LiDAR line segments (+ pixel noise) plus clutter, structure points on the room's faces seen by a few cameras each.  Returns ground truth and
    perturbed initial pose blocks: cameras (aa_cw, t_cw) and LiDARs (aa_lw, t_lw)."""
    rng = np.random.default_rng(seed)
    frames = make_sequence(n_frames, seed=seed + 1, n_az=n_az)
    T_cl = np.eye(4)
    T_cl[:3, :3] = rotvec_to_R(np.array([0.01, 0.02, -0.01]))
    T_cl[:3, 3] = [0.03, -0.05, 0.02]
    T_lc = np.linalg.inv(T_cl)
    cams_gt, lidars_gt, image_lines = np.zeros((n_frames, 6)), np.zeros((n_frames, 6)), []
    Rcw, tcw = [], []
    for i, f in enumerate(frames):
        T_wl = np.eye(4); T_wl[:3, :3] = f["R_wl"]; T_wl[:3, 3] = f["t_wl"]
        T_lw = np.linalg.inv(T_wl)
        T_cw = T_cl @ T_lw
        lidars_gt[i, :3], lidars_gt[i, 3:] = _R_to_aa(T_lw[:3, :3]), T_lw[:3, 3]
        cams_gt[i, :3], cams_gt[i, 3:] = _R_to_aa(T_cw[:3, :3]), T_cw[:3, 3]
        Rcw.append(rotvec_to_R(cams_gt[i, :3])); tcw.append(cams_gt[i, 3:].copy())
        ends_cam = f["end_points"].reshape(-1, 3) @ T_cl[:3, :3].T + T_cl[:3, 3]
        px = pixel_of(ends_cam, rows, cols).reshape(-1, 4) + rng.normal(0, pixel_noise, (len(f["end_points"]), 4))
        cl = np.stack([rng.uniform(0, cols, clutter), rng.uniform(0, rows, clutter), rng.uniform(0, cols, clutter), rng.uniform(0, rows, clutter)], axis=1)
        image_lines.append(np.concatenate([px, cl]).astype(np.float32))
    face = rng.integers(0, 6, n_points)
    pts = rng.uniform(ROOM_MIN, ROOM_MAX, size=(n_points, 3))
    for a in range(3):
        pts[face == 2 * a, a] = ROOM_MIN[a]
        pts[face == 2 * a + 1, a] = ROOM_MAX[a]
    cam, point, bearing = [], [], []
    for p in range(n_points):
        L = int(rng.integers(track_len[0], min(track_len[1], n_frames) + 1))
        for c in np.sort(rng.choice(n_frames, size=L, replace=False)):
            Pc = Rcw[c] @ pts[p] + tcw[c]
            d = Pc / np.linalg.norm(Pc) + rng.normal(size=3) * bearing_noise
            cam.append(c); point.append(p); bearing.append(d / np.linalg.norm(d))
    cams0, lidars0 = cams_gt.copy(), lidars_gt.copy()
    for arr in (cams0, lidars0):
        arr[1:, :3] += rng.normal(size=(n_frames - 1, 3)) * pose_noise[0]
        arr[1:, 3:] += rng.normal(size=(n_frames - 1, 3)) * pose_noise[1]
    lidars0[0, :3] += rng.normal(size=3) * pose_noise[0]
    lidars0[0, 3:] += rng.normal(size=3) * pose_noise[1]
    return {"frames": frames, "T_cl": T_cl, "T_lc": T_lc, "image_lines": image_lines, "rows": rows, "cols": cols, "cams_gt": cams_gt, "lidars_gt": lidars_gt,
            "cams": cams0, "lidars": lidars0, "points_gt": pts, "points": pts + rng.normal(size=pts.shape) * point_noise,
            "cam": np.array(cam, np.int32), "point": np.array(point, np.int32), "bearing": np.array(bearing)}
