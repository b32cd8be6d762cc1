"""examples/refine_pose_driver.cpp: a C++ host program that runs LidarOdometry::RefinePose (point-to-plane, and line-to-line gated by line tracks) through the C ABI alone -
FindNeighbors, GenerateTracks, the frames' clouds, association + residual blocks, the LM solve - with no Python in between.  CPU: it builds against include/panovlm_b200.h and links the library.
GPU: its result equals the Python mirror's (panovlm_b200.odometry.refine_pose), which the other tests compare with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver():
    out = os.path.join(ROOT, "examples", "refine_pose_driver")
    src = os.path.join(ROOT, "examples", "refine_pose_driver.cpp")
    deps = [src, os.path.join(ROOT, "include", "panovlm_b200.h"), os.path.join(ROOT, "panovlm_b200", "libpanovlm_b200.so")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", out, src, "-L", os.path.join(ROOT, "panovlm_b200"), "-lpanovlm_b200",
                               "-Wl,-rpath,$ORIGIN/../panovlm_b200"])
    return out


def test_cpp_refine_pose_driver_builds_against_the_public_header():
    import panovlm_b200
    panovlm_b200.load_library()                                            # the library the driver links
    exe = build_driver()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("lines", [False, True], ids=["point2plane", "point2plane+line2line"])
def test_cpp_refine_pose_driver_equals_the_python_mirror(gpu_ctx, oracle, tmp_path, lines):
    from panovlm_b200 import odometry, synth
    from scipy.spatial.transform import Rotation
    frames = synth.make_sequence(8, n_az=600, tilt=0.3)
    rng = np.random.default_rng(17)
    R0 = [f["R_wl"] @ Rotation.from_rotvec(rng.normal(0, 0.005, 3) * (i > 0)).as_matrix() for i, f in enumerate(frames)]
    t0 = [f["t_wl"] + rng.normal(0, 0.02, 3) * (i > 0) for i, f in enumerate(frames)]
    poses0 = odometry.pose_blocks_from_world(R0, t0, oracle.R_to_aa)
    cfg = odometry.OdometryConfig(line_to_line=lines)
    exp_poses, exp = odometry.refine_pose(gpu_ctx, frames, poses0, cfg, oracle.aa_to_R)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<i", len(frames)))
        for fr in frames:
            f.write(struct.pack("<ii", len(fr["surfLessFlat"]), len(fr["surfFlat"])))
        f.write(np.ascontiguousarray(poses0, np.float64).tobytes())
        for fr in frames:
            f.write(np.ascontiguousarray(fr["surfLessFlat"], np.float32).tobytes()); f.write(np.ascontiguousarray(fr["surfFlat"], np.float32).tobytes())
        f.write(struct.pack("<ddii", cfg.plane_tolerance, cfg.plane_dis_threshold, int(cfg.angle_residual), int(cfg.normalize_distance)))
        f.write(struct.pack("<i", int(lines)))
        if lines:
            for fr in frames:
                corner = np.ascontiguousarray(fr["cornerLessSharp"], np.float32).reshape(-1, 4)
                off, ids = np.ascontiguousarray(fr["p2s_off"], np.int32), np.ascontiguousarray(fr["p2s_ids"], np.int32)
                coeffs, ends = np.ascontiguousarray(fr["segment_coeffs"], np.float64).reshape(-1, 6), np.ascontiguousarray(fr["end_points"], np.float64).reshape(-1, 6)
                assert len(off) == len(corner) + 1 and len(ends) == len(coeffs)
                f.write(struct.pack("<iii", len(corner), len(ids), len(coeffs)))
                for a in (corner, off, ids, coeffs, ends):
                    f.write(a.tobytes())
            f.write(struct.pack("<diii", cfg.line_dis_threshold, int(cfg.line_tracks), cfg.track_neighbor_size, cfg.min_track_length))
    r = subprocess.run([build_driver(), str(fin), str(fout), str(cfg.max_lm_iterations)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    n = len(frames)
    poses = np.frombuffer(raw, np.float64, 6 * n).reshape(n, 6)
    summ = np.frombuffer(raw, np.float64, 6, 48 * n)
    n_blocks, = struct.unpack_from("<q", raw, 48 * n + 48); n_edges, = struct.unpack_from("<i", raw, 48 * n + 56)
    assert n_blocks == exp["n_blocks"] and n_edges == exp["n_edges"]
    assert (summ[2], summ[3], summ[5]) == (exp["iterations"], exp["successful"], exp["termination"])
    assert abs(summ[1] - exp["final_cost"]) < 1e-9 * exp["final_cost"] and summ[1] < 0.5 * summ[0]
    assert np.abs(poses - exp_poses).max() < (1e-8 if lines else 1e-9)      # the line blocks' constants carry the host's R_wl (two Rodrigues implementations, last bits)
    line_blocks = int(r.stdout.split("(")[1].split()[0])
    assert (line_blocks > 0) == lines
    assert np.array_equal(poses[0], poses0[0])                              # the first frame is held constant
