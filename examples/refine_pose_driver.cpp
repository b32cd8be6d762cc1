// A C++ host for the pose-graph path: LidarOdometry::RefinePose (lidar_mapping/LidarOdometry.cpp:15-114) with the residual families of the shipped configs
// (point-to-plane + line-to-line gated by line tracks, config/Room.txt:67-76), written against the C ABI only (include/panovlm_b200.h) the way a PanoVLM maintainer
// would call it without Ceres - no Python anywhere:
//   poses (aa_lw, t_lw) -> frame centres -> pvb_find_neighbors (FindNeighbors, :35) -> pose-graph edges (i, n) for n in N(i)
//   -> [lines] pvb_generate_line_tracks (LidarLineMatch::GenerateTracks, :47-50) -> pvb_frames_line2line_blocks (AddLidarLineToLineResidual2, :51-52: batched votes on the
//      device, FindAssociations tails + track gate + one Point2Line block per point of a kept segment on the host cores)
//   -> pvb_frames_set (the frames' surfLessFlat / surfFlat clouds, sensor frame) -> pvb_frames_point2plane_blocks (AddLidarPointToPlaneResidual, :53-56:
//      Transform2LidarWorld, AssociatePoint2Plane of every edge, residual blocks - all on the device; the line blocks are appended) -> first valid frame constant (:59-66)
//   -> pvb_blocks_solve_lm (SetOptionsLidar + ceres::Solve, :78-80; the linear solver is chosen by size like SetOptionsLidar does).
//
//   refine_pose_driver <in.bin> <out.bin> [max_lm_iterations = 20]
// in.bin : int32 n_frames | per frame: int32 n_target, int32 n_query | float64 poses[n_frames][6] | per frame: float32 target[n_target][4], float32 query[n_query][4]
//          | float64 plane_tolerance, float64 dist_threshold, int32 angle_residual, int32 normalize_distance | int32 line_to_line
//          | if line_to_line: per frame: int32 n_corner, n_ids, n_segments, float32 corner[n_corner][4], int32 p2s_off[n_corner + 1], int32 p2s_ids[n_ids],
//            float64 segment_coeffs[n_segments][6], float64 end_points[n_segments][6]; then float64 line_dis_threshold, int32 line_tracks, track_neighbor_size, min_track_length
// out.bin: float64 poses[n_frames][6] | float64 summary[6] (initial cost, final cost, iterations, successful, unsuccessful, termination) | int64 n_blocks | int32 n_edges
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "panovlm_b200.h"

static void die(const char* what, pvb_ctx* ctx = nullptr) { std::fprintf(stderr, "refine_pose_driver: %s%s%s\n", what, ctx ? ": " : "", ctx ? pvb_last_error(ctx) : ""); std::exit(1); }

template <typename T> static void rd(std::FILE* f, T* dst, size_t n) { if (n && std::fread(dst, sizeof(T), n, f) != n) die("short read"); }

// angle-axis -> rotation matrix (Rodrigues; the same map as ceres::AngleAxisToRotationMatrix), row-major
static void aa_to_R(const double* a, double* R) {
  const double th2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  if (th2 > 1e-30) {
    const double th = std::sqrt(th2), wx = a[0] / th, wy = a[1] / th, wz = a[2] / th, c = std::cos(th), s = std::sin(th), v = 1.0 - c;
    R[0] = c + wx * wx * v;      R[1] = wx * wy * v - wz * s; R[2] = wy * s + wx * wz * v;
    R[3] = wz * s + wx * wy * v; R[4] = c + wy * wy * v;      R[5] = -wx * s + wy * wz * v;
    R[6] = -wy * s + wx * wz * v; R[7] = wx * s + wy * wz * v; R[8] = c + wz * wz * v;
  } else {
    R[0] = 1; R[1] = -a[2]; R[2] = a[1]; R[3] = a[2]; R[4] = 1; R[5] = -a[0]; R[6] = -a[1]; R[7] = a[0]; R[8] = 1;
  }
}

int main(int argc, char** argv) {
  if (argc < 3) die("usage: refine_pose_driver <in.bin> <out.bin> [max_lm_iterations]");
  const int max_it = argc > 3 ? std::atoi(argv[3]) : 20;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) die("cannot open the input");
  int32_t n = 0;
  rd(f, &n, 1);
  std::vector<int32_t> cnt(2 * (size_t)n);
  rd(f, cnt.data(), cnt.size());
  std::vector<double> poses(6 * (size_t)n);
  rd(f, poses.data(), poses.size());
  std::vector<std::vector<float>> tgt(n), qry(n);
  for (int i = 0; i < n; ++i) { tgt[i].resize(4 * (size_t)cnt[2 * i]); qry[i].resize(4 * (size_t)cnt[2 * i + 1]); rd(f, tgt[i].data(), tgt[i].size()); rd(f, qry[i].data(), qry[i].size()); }
  double plane_tol = 0, dist_thr = 0; int32_t angle_residual = 1, normalize_distance = 1;
  rd(f, &plane_tol, 1); rd(f, &dist_thr, 1); rd(f, &angle_residual, 1); rd(f, &normalize_distance, 1);
  int32_t line_to_line = 0, line_tracks = 0, track_nb = 4, min_track = 3; double line_thr = 0.3;
  rd(f, &line_to_line, 1);
  struct Lines { std::vector<float> corner; std::vector<int32_t> p2s_off, p2s_ids; std::vector<double> coeffs, ends; int n_seg = 0; };
  std::vector<Lines> L(n);
  if (line_to_line) {
    for (int i = 0; i < n; ++i) {
      int32_t c3[3]; rd(f, c3, 3);
      L[i].corner.resize(4 * (size_t)c3[0]); L[i].p2s_off.resize((size_t)c3[0] + 1); L[i].p2s_ids.resize((size_t)c3[1]); L[i].n_seg = c3[2];
      L[i].coeffs.resize(6 * (size_t)c3[2]); L[i].ends.resize(6 * (size_t)c3[2]);
      rd(f, L[i].corner.data(), L[i].corner.size()); rd(f, L[i].p2s_off.data(), L[i].p2s_off.size()); rd(f, L[i].p2s_ids.data(), L[i].p2s_ids.size());
      rd(f, L[i].coeffs.data(), L[i].coeffs.size()); rd(f, L[i].ends.data(), L[i].ends.size());
    }
    rd(f, &line_thr, 1); rd(f, &line_tracks, 1); rd(f, &track_nb, 1); rd(f, &min_track, 1);
  }
  std::fclose(f);

  // frame centres t_wl = -R_lw^T t_lw (LidarOdometry.cpp:23-33 keeps T_lw as parameter blocks)
  std::vector<double> t_wl(3 * (size_t)n), R_wl(9 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double R[9]; aa_to_R(&poses[6 * i], R);
    const double* t = &poses[6 * i + 3];
    for (int r = 0; r < 3; ++r) {
      t_wl[3 * i + r] = -(R[0 + r] * t[0] + R[3 + r] * t[1] + R[6 + r] * t[2]);
      for (int c = 0; c < 3; ++c) R_wl[9 * i + 3 * r + c] = R[3 * c + r];          // R_wl = R_lw^T
    }
  }
  // FindNeighbors(lidars, 6) -> edges
  std::vector<int> off(n + 1), nb((size_t)n * 70);
  if (pvb_find_neighbors(n, t_wl.data(), nullptr, nullptr, 6, off.data(), nb.data(), (int)nb.size()) < 0) die("pvb_find_neighbors failed");
  std::vector<int> ref, nei;
  for (int i = 0; i < n; ++i)
    for (int k = off[i]; k < off[i + 1]; ++k)
      if (nb[k] >= 0 && nb[k] < n && nb[k] != i) { ref.push_back(i); nei.push_back(nb[k]); }

  pvb_ctx* ctx = nullptr;
  if (pvb_create(0, &ctx) != PVB_OK) die("pvb_create failed (no CUDA device?)");
  // ---- line-to-line family: tracks, then one Point2Line block per point of every kept neighbour segment (host arrays, appended below)
  std::vector<int> x_type, x_ref, x_nei, x_norm; std::vector<double> x_huber, x_consts;
  long n_extra = 0;
  if (line_to_line) {
    std::vector<pvb_line_frame> lf(n);
    long cap_b = 16; int seg_sum = 0;
    for (int i = 0; i < n; ++i) {
      lf[i].corner_local = L[i].corner.data(); lf[i].n_corner = (int)(L[i].corner.size() / 4); lf[i].p2s_off = L[i].p2s_off.data();
      lf[i].p2s_ids = L[i].p2s_ids.empty() ? nullptr : L[i].p2s_ids.data(); lf[i].n_segments = L[i].n_seg;
      lf[i].segment_coeffs = L[i].n_seg ? L[i].coeffs.data() : nullptr; lf[i].end_points = L[i].n_seg ? L[i].ends.data() : nullptr;
      lf[i].R_wl = &R_wl[9 * i]; lf[i].t_wl = &t_wl[3 * i];
      seg_sum += L[i].n_seg > 0 ? L[i].n_seg : 1;
    }
    for (size_t e = 0; e < ref.size(); ++e) cap_b += 6l * lf[nei[e]].n_corner;
    int n_tracks = -1;
    std::vector<int> track_off, feat_frame, feat_line;
    if (line_tracks) {                                            // GenerateTracks(neighbor size 4, minimum track length 3), threshold hard-coded 0.3 (LidarLineMatch.cpp:84)
      std::vector<int> off4(n + 1), nb4((size_t)n * 70);
      if (pvb_find_neighbors(n, t_wl.data(), nullptr, nullptr, track_nb, off4.data(), nb4.data(), (int)nb4.size()) < 0) die("pvb_find_neighbors (tracks) failed");
      int widest = 1;
      for (int i = 0; i < n; ++i) widest = off4[i + 1] - off4[i] > widest ? off4[i + 1] - off4[i] : widest;
      const int cap_t = 2 * seg_sum * widest + 1;
      track_off.resize((size_t)cap_t + 1); feat_frame.resize(cap_t); feat_line.resize(cap_t);
      if (pvb_generate_line_tracks(ctx, n, lf.data(), nullptr, off4.data(), nb4.data(), 0.3, min_track, cap_t, &n_tracks, track_off.data(), feat_frame.data(), feat_line.data()) != PVB_OK)
        die("pvb_generate_line_tracks", ctx);
    }
    x_type.resize(cap_b); x_ref.resize(cap_b); x_nei.resize(cap_b); x_norm.resize(cap_b); x_huber.resize(cap_b); x_consts.resize(12 * (size_t)cap_b);
    n_extra = pvb_frames_line2line_blocks(ctx, n, lf.data(), (int)ref.size(), ref.data(), nei.data(), line_thr, n_tracks, n_tracks >= 0 ? track_off.data() : nullptr,
                                          n_tracks >= 0 ? feat_frame.data() : nullptr, n_tracks >= 0 ? feat_line.data() : nullptr, angle_residual, normalize_distance, 1.0, 0, cap_b,
                                          x_type.data(), x_ref.data(), x_nei.data(), x_norm.data(), x_huber.data(), x_consts.data());
    if (n_extra < 0) die("pvb_frames_line2line_blocks", ctx);
  }
  std::vector<pvb_frame> fr(n);
  for (int i = 0; i < n; ++i) { fr[i].surf_target = tgt[i].data(); fr[i].n_target = cnt[2 * i]; fr[i].surf_query = qry[i].data(); fr[i].n_query = cnt[2 * i + 1]; }
  if (pvb_frames_set(ctx, n, fr.data()) != PVB_OK) die("pvb_frames_set", ctx);
  pvb_assoc_params ap; ap.plane_tolerance = plane_tol; ap.dist_threshold = (float)dist_thr; ap.k = 10; ap.cell_size = 0.0;
  long n_blocks = 0;
  if (pvb_frames_point2plane_blocks(ctx, poses.data(), (int)ref.size(), ref.data(), nei.data(), &ap, angle_residual, normalize_distance, 1.0, 0, n, n_extra, n_extra ? x_type.data() : nullptr,
                                    n_extra ? x_ref.data() : nullptr, n_extra ? x_nei.data() : nullptr, n_extra ? x_norm.data() : nullptr, n_extra ? x_huber.data() : nullptr,
                                    n_extra ? x_consts.data() : nullptr, &n_blocks) != PVB_OK) die("pvb_frames_point2plane_blocks", ctx);
  std::vector<unsigned char> is_const(n, 0);
  is_const[0] = 1;                                                // every frame of the input is valid: the first one is held constant
  double summary[6] = {0, 0, 0, 0, 0, 0};
  if (pvb_blocks_solve_lm(ctx, poses.data(), is_const.data(), max_it, summary) != PVB_OK) die("pvb_blocks_solve_lm", ctx);
  pvb_destroy(ctx);

  f = std::fopen(argv[2], "wb");
  if (!f) die("cannot open the output");
  const int64_t nb64 = n_blocks; const int32_t ne = (int32_t)ref.size();
  std::fwrite(poses.data(), sizeof(double), poses.size(), f); std::fwrite(summary, sizeof(double), 6, f); std::fwrite(&nb64, 8, 1, f); std::fwrite(&ne, 4, 1, f);
  std::fclose(f);
  std::printf("refine_pose_driver: %d frames, %d edges, %ld residual blocks (%ld line blocks), cost %.6g -> %.6g in %d LM iterations\n", n, ne, n_blocks, n_extra, summary[0], summary[1], (int)summary[2]);
  return 0;
}
