// A C++ host for the pose-graph path: LidarOdometry::RefinePose (lidar_mapping/LidarOdometry.cpp:15-114) for the point-to-plane family, written against the C ABI
// only (include/panovlm_b200.h) the way a PanoVLM maintainer would call it without Ceres - no Python anywhere:
//   poses (aa_lw, t_lw) -> frame centres -> pvb_find_neighbors (FindNeighbors, :35) -> pose-graph edges (i, n) for n in N(i)
//   -> pvb_frames_set (the frames' surfLessFlat / surfFlat clouds, sensor frame) -> pvb_frames_point2plane_blocks (AddLidarPointToPlaneResidual, :53-56:
//      Transform2LidarWorld, AssociatePoint2Plane of every edge, residual blocks - all on the device) -> first valid frame constant (:59-66)
//   -> pvb_blocks_solve_lm (SetOptionsLidar + ceres::Solve, :78-80; the linear solver is chosen by size like SetOptionsLidar does).
//
//   refine_pose_driver <in.bin> <out.bin> [max_lm_iterations = 20]
// in.bin : int32 n_frames | per frame: int32 n_target, int32 n_query | float64 poses[n_frames][6] | per frame: float32 target[n_target][4], float32 query[n_query][4]
//          | float64 plane_tolerance, float64 dist_threshold, int32 angle_residual, int32 normalize_distance
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
  std::fclose(f);

  // frame centres t_wl = -R_lw^T t_lw (LidarOdometry.cpp:23-33 keeps T_lw as parameter blocks)
  std::vector<double> t_wl(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double R[9]; aa_to_R(&poses[6 * i], R);
    const double* t = &poses[6 * i + 3];
    for (int r = 0; r < 3; ++r) t_wl[3 * i + r] = -(R[0 + r] * t[0] + R[3 + r] * t[1] + R[6 + r] * t[2]);
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
  std::vector<pvb_frame> fr(n);
  for (int i = 0; i < n; ++i) { fr[i].surf_target = tgt[i].data(); fr[i].n_target = cnt[2 * i]; fr[i].surf_query = qry[i].data(); fr[i].n_query = cnt[2 * i + 1]; }
  if (pvb_frames_set(ctx, n, fr.data()) != PVB_OK) die("pvb_frames_set", ctx);
  pvb_assoc_params ap; ap.plane_tolerance = plane_tol; ap.dist_threshold = (float)dist_thr; ap.k = 10; ap.cell_size = 0.0;
  long n_blocks = 0;
  if (pvb_frames_point2plane_blocks(ctx, poses.data(), (int)ref.size(), ref.data(), nei.data(), &ap, angle_residual, normalize_distance, 1.0, 0, n, 0, nullptr, nullptr, nullptr, nullptr,
                                    nullptr, nullptr, &n_blocks) != PVB_OK) die("pvb_frames_point2plane_blocks", ctx);
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
  std::printf("refine_pose_driver: %d frames, %d edges, %ld residual blocks, cost %.6g -> %.6g in %d LM iterations\n", n, ne, n_blocks, summary[0], summary[1], (int)summary[2]);
  return 0;
}
