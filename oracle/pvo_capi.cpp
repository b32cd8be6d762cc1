// ORACLE — TEST INFRASTRUCTURE ONLY (see pvo_math.hpp header).  C entry points for ctypes
// (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).  Nothing in
// panovlm_b200/ may link, import or call this library.
#include <chrono>
#include <cstdio>
#include <thread>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "pvo_assoc.hpp"
#include "pvo_solver.hpp"
#include "pvo_tracks.hpp"
#include "pvo_undistort.hpp"
#include "pvo_ba.hpp"

using namespace pvo;

extern "C" {

int pvo_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void pvo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

// ---- rotation / geometry primitives (for pinning against scipy / numpy) --------------------------------
void pvo_aa_to_R(const double* aa, double* R_colmajor) { AngleAxisToRotationMatrix(aa, R_colmajor); }
void pvo_R_to_aa(const double* R_colmajor, double* aa) { RotationMatrixToAngleAxis(R_colmajor, aa); }
void pvo_aa_rotate(const double* aa, const double* p, double* out) { AngleAxisRotatePoint(aa, p, out); }
void pvo_form_plane(int m, const double* pts, double tol, double* out4) { FormPlaneLSQ(m, pts, tol, out4); }
int pvo_form_line(int m, const double* pts, double tol, double dis_thr, double* out6) { return FormLinePCA(m, pts, tol, dis_thr, out6) ? 1 : 0; }
void pvo_sym_eig3(const double* A, double* eval, double* evec) { double v[3][3]; SymEig3(A, eval, v); for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) evec[k * 3 + r] = v[k][r]; }
void pvo_fast_atan2_f(long n, const float* y, const float* x, float* out) { for (long i = 0; i < n; ++i) out[i] = FastAtan2(y[i], x[i]); }
void pvo_fast_atan2_d(long n, const double* y, const double* x, double* out) { for (long i = 0; i < n; ++i) out[i] = FastAtan2(y[i], x[i]); }
void pvo_image_to_cam_d(int rows, int cols, long n, const double* px, double* cam) { Equirect eq{rows, cols}; for (long i = 0; i < n; ++i) eq.ImageToCam(px + 2 * i, 1.0, cam + 3 * i); }
void pvo_cam_to_image_d(int rows, int cols, long n, const double* cam, double* px) { Equirect eq{rows, cols}; for (long i = 0; i < n; ++i) eq.CamToImage(cam + 3 * i, px + 2 * i); }
void pvo_cam_to_image_f(int rows, int cols, long n, const float* cam, float* px) { Equirect eq{rows, cols}; for (long i = 0; i < n; ++i) eq.CamToImage(cam + 3 * i, px + 2 * i); }
void pvo_image_to_cam_f(int rows, int cols, long n, const float* px, float r, float* cam) { Equirect eq{rows, cols}; for (long i = 0; i < n; ++i) eq.ImageToCam(px + 2 * i, r, cam + 3 * i); }
int pvo_break_to_segments(int rows, int cols, const float* line4, float seg_length, int cap, float* out2) {
  Equirect eq{rows, cols};
  const auto s = eq.BreakToSegments(line4, line4 + 2, seg_length);
  if ((int)s.size() > cap) return -1;
  for (size_t i = 0; i < s.size(); ++i) { out2[2 * i] = s[i].first; out2[2 * i + 1] = s[i].second; }
  return (int)s.size();
}
void pvo_form_plane3(const double* p1, const double* p2, const double* p3, double* out4) { FormPlane3(p1, p2, p3, out4); }
// out5 = PointToLineDistance3D(point, line6), PointToPlaneDistance(plane4, point, false), the same with the plane normalised by |n| and normalized = true,
// VectorAngle3D(point, line6[0:3]), PlaneAngle(point, line6[0:3]); proj6 = ProjectPointToPlane(point, plane4, false / true as above)
void pvo_geometry_helpers(const double* point, const double* line6, const double* plane4, double* out5, double* proj6) {
  double pn[4]; const double nn = std::sqrt(plane4[0] * plane4[0] + plane4[1] * plane4[1] + plane4[2] * plane4[2]);
  for (int i = 0; i < 4; ++i) pn[i] = i < 3 ? plane4[i] / nn : plane4[i];
  out5[0] = PointToLineDistance3D(point, line6);
  out5[1] = PointToPlaneDistance(plane4, point, false);
  out5[2] = PointToPlaneDistance(pn, point, true);
  out5[3] = VectorAngle3D(point, line6, false);
  out5[4] = PlaneAngle(point, line6, false);
  ProjectPointToPlane(point, plane4, proj6, false);
  ProjectPointToPlane(point, pn, proj6 + 3, true);
}

// ---- residual blocks ------------------------------------------------------------------------------------
// blocks are passed as parallel arrays: type/ref/nei/normalize (int32), huber (f64), consts (n x 12 f64).
static inline Block MakeBlock(long i, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts) {
  Block b; b.type = type[i]; b.ref = ref[i]; b.nei = nei[i]; b.normalize = normalize[i]; b.huber = huber[i];
  std::memcpy(b.c, consts + 12 * i, 96); return b;
}

// Per-block residual / 1x12 Jacobian / cost, one autodiff functor at a time (OpenMP over blocks like
// Ceres' num_threads).  apply_loss: Huber corrector applied.  out_J may be null.
void pvo_eval_blocks(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber,
                     const double* consts, const double* poses, int apply_loss, double* out_r, double* out_J, double* out_cost) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    Block b = MakeBlock(i, type, ref, nei, normalize, huber, consts);
    double r, J[12];
    const double c = EvalBlock(b, poses, &r, out_J ? J : nullptr, apply_loss != 0);
    out_r[i] = r; if (out_cost) out_cost[i] = c;
    if (out_J) std::memcpy(out_J + 12 * i, J, 96);
  }
}

// Dense normal equations (6nb x 6nb) + gradient + cost.
double pvo_normal_equations(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber,
                            const double* consts, const double* poses, int nb, double* H, double* g) {
  std::vector<Block> blocks(n);
  for (long i = 0; i < n; ++i) blocks[i] = MakeBlock(i, type, ref, nei, normalize, huber, consts);
  return NormalEquations(blocks.data(), n, poses, nb, H, g);
}

// Levenberg-Marquardt over the block list; poses (nb x 6: aa, t) updated in place.  summary[6] =
// {initial_cost, final_cost, iterations, successful, unsuccessful, termination}.
void pvo_solve_lm(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber,
                  const double* consts, double* poses, int nb, const unsigned char* is_const, int max_iter, double* summary) {
  std::vector<Block> blocks(n);
  for (long i = 0; i < n; ++i) blocks[i] = MakeBlock(i, type, ref, nei, normalize, huber, consts);
  auto eval = [&](const double* x, double* H, double* g) { return NormalEquations(blocks.data(), n, x, nb, H, g); };
  LMSummary S = SolveLM(eval, poses, nb, is_const, max_iter);
  summary[0] = S.initial_cost; summary[1] = S.final_cost; summary[2] = S.iterations; summary[3] = S.successful; summary[4] = S.unsuccessful; summary[5] = S.termination;
}

// ---- transforms / kNN / associations ---------------------------------------------------------------------
void pvo_transform_cloud(const double* R_rowmajor, const double* t, const float* in, int n, float* out) { TransformCloud(R_rowmajor, t, in, n, out, 4); }
void pvo_world2local(const double* R, const double* t, long n, const double* pw, double* out) { for (long i = 0; i < n; ++i) World2Local(R, t, pw + 3 * i, out + 3 * i); }

void pvo_knn(const float* pts, int n, const float* queries, int nq, int k, int use_kdtree, int* out_idx, float* out_d2) {
  KdTree tree; if (use_kdtree) tree.Build(pts, n, 4);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; ++i) {
    KnnResult r;
    if (use_kdtree) tree.Knn(queries + (size_t)i * 4, k, r); else KnnBrute(pts, n, 4, queries + (size_t)i * 4, k, r);
    for (int j = 0; j < k; ++j) { out_idx[(size_t)i * k + j] = j < (int)r.idx.size() ? r.idx[j] : -1; out_d2[(size_t)i * k + j] = j < (int)r.d2.size() ? r.d2[j] : INFINITY; }
  }
}

// returns the number of associations; out arrays sized for n_nei entries.
int pvo_associate_p2plane(const float* ref_world, int n_ref, const double* R_ref, const double* t_ref,
                          const float* nei_world, int n_nei, const double* R_nei, const double* t_nei,
                          double plane_tol, float dist_thr, int k, int use_kdtree,
                          int* out_query, double* out_point, double* out_plane) {
  std::vector<P2PlaneAssoc> v;
  AssociatePoint2Plane(ref_world, n_ref, R_ref, t_ref, nei_world, n_nei, R_nei, t_nei, plane_tol, dist_thr, k, use_kdtree != 0, v);
  for (size_t i = 0; i < v.size(); ++i) { out_query[i] = v[i].query_idx; std::memcpy(out_point + 3 * i, v[i].point, 24); std::memcpy(out_plane + 4 * i, v[i].plane, 32); }
  return (int)v.size();
}

int pvo_associate_p2line(const float* ref_world, int n_ref, const double* R_ref, const double* t_ref, const float* nei_world, int n_nei,
                          const double* R_nei, const double* t_nei, float dist_thr, int use_kdtree, int* out_query, double* out_point, double* out_a, double* out_b) {
  std::vector<P2LineAssoc> v;
  AssociatePoint2Line(ref_world, n_ref, R_ref, t_ref, nei_world, n_nei, R_nei, t_nei, dist_thr, use_kdtree != 0, v);
  for (size_t i = 0; i < v.size(); ++i) { out_query[i] = v[i].query_idx; std::memcpy(out_point + 3 * i, v[i].point, 24); std::memcpy(out_a + 3 * i, v[i].a, 24); std::memcpy(out_b + 3 * i, v[i].b, 24); }
  return (int)v.size();
}

void pvo_transform_lines(const double* R, const double* t, int S, const double* in, double* out) { for (int s = 0; s < S; ++s) TransformLine(R, t, in + 6 * s, out + 6 * s); }

void pvo_line_votes(const double* ref_lines_world, int S_ref, const float* nei_corner_world, int n_pts, const int* p2s_off, const int* p2s_ids,
                    int S_nei, double dist_thr, int* M) {
  std::vector<int> m; Line2LineVotes(ref_lines_world, S_ref, nei_corner_world, n_pts, p2s_off, p2s_ids, S_nei, dist_thr, m);
  std::copy(m.begin(), m.end(), M);
}

int pvo_find_associations(const double* ref_coeffs_local, const double* ref_lines_world, int S_ref, const double* nei_lines_world, int S_nei,
                          const int* seg_sizes_nei, const int* M, int* out_nei, int* out_ref, double* out_a, double* out_b) {
  std::vector<int> m(M, M + (size_t)S_nei * S_ref); std::vector<L2LAssoc> v;
  FindAssociations(ref_coeffs_local, ref_lines_world, S_ref, nei_lines_world, S_nei, seg_sizes_nei, m, v);
  for (size_t i = 0; i < v.size(); ++i) { out_nei[i] = v[i].nei_line; out_ref[i] = v[i].ref_line; std::memcpy(out_a + 3 * i, v[i].a, 24); std::memcpy(out_b + 3 * i, v[i].b, 24); }
  return (int)v.size();
}

int pvo_associate_p2line_segment_knn(const float* ref_world, int n_ref, const int* ref_p2s_off, const int* ref_p2s_ids, const double* ref_coeffs_local,
                                     const float* nei_world, int n_nei, const double* R_nei, const double* t_nei, float dist_threshold, int use_kdtree,
                                     int* out_query, int* out_line, double* out_point, double* out_a, double* out_b) {
  std::vector<P2SegAssoc> v;
  AssociatePoint2LineSegmentKNN(ref_world, n_ref, ref_p2s_off, ref_p2s_ids, ref_coeffs_local, nei_world, n_nei, R_nei, t_nei, dist_threshold, use_kdtree != 0, v);
  for (size_t i = 0; i < v.size(); ++i) { out_query[i] = v[i].query_idx; out_line[i] = v[i].ref_line; std::memcpy(out_point + 3 * i, v[i].point, 24); std::memcpy(out_a + 3 * i, v[i].a, 24); std::memcpy(out_b + 3 * i, v[i].b, 24); }
  return (int)v.size();
}

int pvo_associate_p2line_segment(const double* ref_lines_world, const double* ref_coeffs_local, int S_ref, const float* nei_world, int n_nei,
                                 const double* R_nei, const double* t_nei, float dist_threshold, int* out_query, int* out_line, double* out_point, double* out_a, double* out_b) {
  std::vector<P2SegAssoc> v;
  AssociatePoint2LineSegment(ref_lines_world, ref_coeffs_local, S_ref, nei_world, n_nei, R_nei, t_nei, dist_threshold, v);
  for (size_t i = 0; i < v.size(); ++i) { out_query[i] = v[i].query_idx; out_line[i] = v[i].ref_line; std::memcpy(out_point + 3 * i, v[i].point, 24); std::memcpy(out_a + 3 * i, v[i].a, 24); std::memcpy(out_b + 3 * i, v[i].b, 24); }
  return (int)v.size();
}

void pvo_line2line_knn_votes(const float* ref_world, int n_ref, const int* ref_p2s_off, const int* ref_p2s_ids, int S_ref, const float* nei_world, int n_nei,
                             const int* nei_p2s_off, const int* nei_p2s_ids, int S_nei, float dist_threshold, int use_kdtree, int* M) {
  std::vector<int> m;
  Line2LineKnnVotes(ref_world, n_ref, ref_p2s_off, ref_p2s_ids, S_ref, nei_world, n_nei, nei_p2s_off, nei_p2s_ids, S_nei, dist_threshold, use_kdtree != 0, m);
  std::copy(m.begin(), m.end(), M);
}

// line tracks: CSR in (pairs, matches), CSR out (track -> features), returns the number of tracks; keep[] = gate result per association
int pvo_line_tracks(int n_pairs, const int* pair_a, const int* pair_b, const int* match_off, const int* match_a, const int* match_b, int min_length, int allow_multiple_map,
                    int* track_off, int* feat_frame, int* feat_line) {
  std::vector<std::pair<size_t, size_t>> pairs(n_pairs); std::vector<std::set<Feature>> matches(n_pairs);
  for (int p = 0; p < n_pairs; ++p) {
    pairs[p] = {(size_t)pair_a[p], (size_t)pair_b[p]};
    for (int e = match_off[p]; e < match_off[p + 1]; ++e) matches[p].insert(Feature((uint32_t)match_a[e], (uint32_t)match_b[e]));
  }
  LineTracks T; BuildLineTracks(pairs, matches, (uint32_t)min_length, allow_multiple_map != 0, T);
  int at = 0;
  for (size_t t = 0; t < T.tracks.size(); ++t) {
    track_off[t] = at;
    for (const Feature& f : T.tracks[t]) { feat_frame[at] = (int)f.first; feat_line[at] = (int)f.second; ++at; }
  }
  track_off[T.tracks.size()] = at;
  return (int)T.tracks.size();
}

void pvo_line_track_gate(int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int ref_frame, int nei_frame, int n, const int* ref_line,
                         const int* nei_line, unsigned char* keep) {
  LineTracks T; T.tracks.resize(n_tracks);
  for (int t = 0; t < n_tracks; ++t) for (int e = track_off[t]; e < track_off[t + 1]; ++e) T.tracks[t].insert(Feature((uint32_t)feat_frame[e], (uint32_t)feat_line[e]));
  const auto l2t = LinesToTrack(T);
  for (int i = 0; i < n; ++i) keep[i] = TrackGate(T, l2t, Feature((uint32_t)ref_frame, (uint32_t)ref_line[i]), Feature((uint32_t)nei_frame, (uint32_t)nei_line[i])) ? 1 : 0;
}

void pvo_angle_votes(int rows, int cols, const float* lines, int L, const float* cloud_local, int P, const int* p2s_off, const int* p2s_ids, int S,
                     const double* T_cl, int* counts) {
  std::vector<int> c; AngleVotes(Equirect{rows, cols}, lines, L, cloud_local, P, p2s_off, p2s_ids, S, T_cl, c);
  std::copy(c.begin(), c.end(), counts);
}

int pvo_associate_by_angle(int rows, int cols, const float* lines, int L, const float* cloud_local, int P, const int* p2s_off, const int* p2s_ids, int S,
                           const int* seg_sizes, const double* end_points, const double* T_cl, int filter_by_length, int max_out,
                           int* out_img, int* out_lidar, double* out_start, double* out_end, float* out_angle, int multiple_association,
                           const unsigned char* image_mask, const unsigned char* lidar_mask) {
  std::vector<CamLidarPair> v;
  AssociateByAngle(Equirect{rows, cols}, lines, L, cloud_local, P, p2s_off, p2s_ids, S, seg_sizes, end_points, T_cl, filter_by_length != 0, v, multiple_association != 0,
                   image_mask, lidar_mask);
  const int n = std::min<int>((int)v.size(), max_out);
  for (int i = 0; i < n; ++i) { out_img[i] = v[i].image_line; out_lidar[i] = v[i].lidar_line; std::memcpy(out_start + 3 * i, v[i].start, 24); std::memcpy(out_end + 3 * i, v[i].end, 24); out_angle[i] = v[i].angle; }
  return (int)v.size();
}

int pvo_unique_line_pairs(int n, const int* image_line, const int* lidar_line, const float* score, int* out_img, int* out_lidar, float* out_score) {
  std::vector<CamLidarPair> v(n);
  for (int i = 0; i < n; ++i) { v[i] = CamLidarPair{}; v[i].image_line = image_line[i]; v[i].lidar_line = lidar_line[i]; v[i].angle = score[i]; }
  UniqueLinePair(v);
  for (size_t i = 0; i < v.size(); ++i) { out_img[i] = v[i].image_line; out_lidar[i] = v[i].lidar_line; out_score[i] = v[i].angle; }
  return (int)v.size();
}

void pvo_project_depth(const float* cloud, int n, int rows, int cols, const double* T_cl, int size, uint16_t* img, float* uvd) {
  ProjectLidar2PanoramaDepth(cloud, n, 4, rows, cols, T_cl, size, img, uvd);
}

// ---- dense ICP sweep (BASELINE.json configs[4]) — the timed CPU baseline ----------------------------------
// One Gauss-Newton evaluation of the reference algorithm for `n_frames` source frames against one target
// cloud (world == target frame, T = identity):  per source point  pcl::transformPointCloud -> kd-tree
// 10-NN (float32) -> class check -> FormPlane / FormLine -> Point2Plane_Meter autodiff (Jet<12>) -> Huber
// corrector -> per-frame 6x6 normal equations of the source ("nei") pose.
//   src_local: concatenated n x 4 float32; frame f owns [src_off[f], src_off[f+1]).
//   poses_lw: per frame (aa_lw[3], t_lw[3]) i.e. world->lidar, like the optimiser's parameter blocks.
//   out_sys: per frame 29 doubles = H upper (21, row-major) + g (6) + cost + n_residuals.
//   mode: 0 = reference-faithful (association serial on 1 thread, evaluation on all threads),
//         1 = best-effort (everything on all threads).
// Returns seconds: out_times[0] = kd-tree build, [1] = association, [2] = residual/Jacobian/reduce.
void pvo_dense_icp_eval(const float* target_world, int n_target, const float* src_local, const int* src_off, int n_frames,
                        const double* poses_lw, double plane_tol, float dist_thr, int k, double huber, double weight, int mode,
                        const void* prebuilt_tree, double* out_sys, double* out_times, long* out_n_assoc) {
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  KdTree local; const KdTree* tree = (const KdTree*)prebuilt_tree;
  if (!tree) { local.Build(target_world, n_target, 4); tree = &local; }
  auto t1 = clk::now();
  const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z3[3] = {0, 0, 0};
  std::vector<std::vector<P2PlaneAssoc>> assoc(n_frames);
  std::vector<std::vector<float>> world(n_frames);
  // association
  auto do_frame = [&](int f) {
    const int n = src_off[f + 1] - src_off[f];
    double Rlw[9], R_wl[9], t_wl[3];
    AngleAxisToRotationMatrix(poses_lw + 6 * f, Rlw);  // column-major R_lw == row-major R_wl
    for (int i = 0; i < 9; ++i) R_wl[i] = Rlw[i];
    const double* tl = poses_lw + 6 * f + 3;
    for (int r = 0; r < 3; ++r) t_wl[r] = -(R_wl[r * 3] * tl[0] + R_wl[r * 3 + 1] * tl[1] + R_wl[r * 3 + 2] * tl[2]);
    world[f].resize((size_t)n * 4);
    TransformCloud(R_wl, t_wl, src_local + (size_t)src_off[f] * 4, n, world[f].data(), 4);
    AssociatePoint2Plane(target_world, n_target, I9, Z3, world[f].data(), n, R_wl, t_wl, plane_tol, dist_thr, k, true, assoc[f], tree);
  };
  if (mode == 0) { for (int f = 0; f < n_frames; ++f) do_frame(f); }
  else {
    // best effort: split each frame's queries over threads (order of results restored by query index)
    for (int f = 0; f < n_frames; ++f) {
      const int n = src_off[f + 1] - src_off[f];
      double Rlw[9], R_wl[9], t_wl[3];
      AngleAxisToRotationMatrix(poses_lw + 6 * f, Rlw);
      for (int i = 0; i < 9; ++i) R_wl[i] = Rlw[i];
      const double* tl = poses_lw + 6 * f + 3;
      for (int r = 0; r < 3; ++r) t_wl[r] = -(R_wl[r * 3] * tl[0] + R_wl[r * 3 + 1] * tl[1] + R_wl[r * 3 + 2] * tl[2]);
      world[f].resize((size_t)n * 4);
      TransformCloud(R_wl, t_wl, src_local + (size_t)src_off[f] * 4, n, world[f].data(), 4);
      const int chunk = 4096, nchunks = (n + chunk - 1) / chunk;
      std::vector<std::vector<P2PlaneAssoc>> parts(nchunks);
#pragma omp parallel for schedule(dynamic, 1)
      for (int c = 0; c < nchunks; ++c) {
        const int lo = c * chunk, hi = std::min(n, lo + chunk);
        AssociatePoint2Plane(target_world, n_target, I9, Z3, world[f].data() + (size_t)lo * 4, hi - lo, R_wl, t_wl, plane_tol, dist_thr, k, true, parts[c], tree);
        for (auto& a : parts[c]) a.query_idx += lo;
      }
      for (auto& p : parts) assoc[f].insert(assoc[f].end(), p.begin(), p.end());
    }
  }
  auto t2 = clk::now();
  // residual + Jacobian one functor at a time, reduce per frame.  ref pose = identity block, nei = frame.
  long total = 0;
  for (int f = 0; f < n_frames; ++f) {
    const auto& A = assoc[f];
    const long m = (long)A.size(); total += m;
    double acc[29]; for (double& v : acc) v = 0;
#pragma omp parallel
    {
      double loc[29]; for (double& v : loc) v = 0;
#pragma omp for schedule(static) nowait
      for (long i = 0; i < m; ++i) {
        Point2Plane_Meter fn; std::memcpy(fn.p, A[i].point, 24); std::memcpy(fn.plane, A[i].plane, 32); fn.weight = weight;
        const double zero6[6] = {0, 0, 0, 0, 0, 0};
        double r, J[12], c;
        EvaluateAutoDiff4(fn, zero6, zero6 + 3, poses_lw + 6 * f, poses_lw + 6 * f + 3, &r, J);
        HuberCorrect(huber, &r, J, 12, &c);
        const double* j = J + 6; int o = 0;
        for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) loc[o++] += j[a] * j[b];
        for (int a = 0; a < 6; ++a) loc[21 + a] += j[a] * r;
        loc[27] += c; loc[28] += 1.0;
      }
#pragma omp critical
      for (int q = 0; q < 29; ++q) acc[q] += loc[q];
    }
    std::memcpy(out_sys + 29 * f, acc, sizeof(acc));
  }
  auto t3 = clk::now();
  out_times[0] = std::chrono::duration<double>(t1 - t0).count();
  out_times[1] = std::chrono::duration<double>(t2 - t1).count();
  out_times[2] = std::chrono::duration<double>(t3 - t2).count();
  *out_n_assoc = total;
}

// ---- pose interpolation / motion undistortion ------------------------------------------------
void pvo_slerp_pose(const double* pose_w1, const double* pose_w2, double ratio, double* out16) { SlerpPose(pose_w1, pose_w2, ratio, out16); }
void pvo_undistort_cloud(const double* R_wl, const double* t_wl, const double* R_we, const double* t_we, const float* in, long n, float* out) {
  UndistortCloud(R_wl, t_wl, R_we, t_we, in, (size_t)n, out);
}
void pvo_undistort_end_poses(int n, const double* poses, const unsigned char* pose_valid, const unsigned char* frame_valid, float gap_time, double* out_pose,
                             unsigned char* has) {
  UndistortEndPoses(n, poses, pose_valid, frame_valid, gap_time, out_pose, has);
}

// ---- camera reprojection residuals / bundle adjustment ---------------------------------------------
static std::vector<ReprojObs> MakeObs(long n, const int* cam, const int* point, const double* bearing) {
  std::vector<ReprojObs> v(n);
  for (long i = 0; i < n; ++i) { v[i].cam = cam[i]; v[i].point = point[i]; std::memcpy(v[i].bearing, bearing + 3 * i, 24); }
  return v;
}
void pvo_reproj_eval(long n, const int* cam, const int* point, const double* bearing, double weight, double huber, const double* cams, const double* points,
                     int apply_loss, double* out_r, double* out_J9, double* out_cost) {
  std::vector<ReprojObs> obs = MakeObs(n, cam, point, bearing);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    double r, J[9];
    const double c = EvalReproj(obs[i], weight, apply_loss ? huber : 0.0, cams, points, &r, out_J9 ? J : nullptr);
    out_r[i] = r; if (out_cost) out_cost[i] = c;
    if (out_J9) std::memcpy(out_J9 + 9 * i, J, 72);
  }
}
double pvo_reproj_normal_equations(long n, const int* cam, const int* point, const double* bearing, double weight, double huber, const double* x, int nc, long np,
                                   double* H, double* g) {
  std::vector<ReprojObs> obs = MakeObs(n, cam, point, bearing);
  return ReprojNormalEquations(obs.data(), n, weight, huber, x, nc, np, H, g);
}
// x = [cams (6 nc) | points (3 np)] updated in place; param_const per parameter
void pvo_reproj_solve_lm(long n, const int* cam, const int* point, const double* bearing, double weight, double huber, double* x, int nc, long np,
                         const unsigned char* param_const, int max_iter, double* summary) {
  std::vector<ReprojObs> obs = MakeObs(n, cam, point, bearing);
  auto eval = [&](const double* xx, double* H, double* g) { return ReprojNormalEquations(obs.data(), n, weight, huber, xx, nc, np, H, g); };
  LMSummary S = SolveLMParams(eval, x, (int)(6L * nc + 3L * np), param_const, max_iter);
  summary[0] = S.initial_cost; summary[1] = S.final_cost; summary[2] = S.iterations; summary[3] = S.successful; summary[4] = S.unsuccessful; summary[5] = S.termination;
}

// joint problem of CameraLidarOptimizer::Optimize: pose blocks [cameras | LiDARs] (6 each) + structure points; the residual-block list
// (LiDAR-LiDAR, camera-LiDAR) and the reprojection observations (camera index = pose block index) share the pose array.
// x = [poses (6 nb) | points (3 np)] updated in place.
void pvo_joint_solve_lm(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts,
                        long n_obs, const int* cam, const int* point, const double* bearing, double weight, double obs_huber,
                        double* x, int nb, long np, const unsigned char* param_const, int max_iter, double* summary) {
  std::vector<Block> blocks(n);
  for (long i = 0; i < n; ++i) blocks[i] = MakeBlock(i, type, ref, nei, normalize, huber, consts);
  std::vector<ReprojObs> obs = MakeObs(n_obs, cam, point, bearing);
  const long D = 6L * nb + 3L * np, Dp = 6L * nb;
  std::vector<double> Hp((size_t)Dp * Dp), gp(Dp);
  auto eval = [&](const double* xx, double* H, double* g) {
    double cost = ReprojNormalEquations(obs.data(), n_obs, weight, obs_huber, xx, nb, np, H, g);
    cost += NormalEquations(blocks.data(), n, xx, nb, H ? Hp.data() : nullptr, g ? gp.data() : nullptr);
    if (H) for (long i = 0; i < Dp; ++i) for (long j = 0; j < Dp; ++j) H[(size_t)i * D + j] += Hp[(size_t)i * Dp + j];
    if (g) for (long i = 0; i < Dp; ++i) g[i] += gp[i];
    return cost;
  };
  LMSummary S = SolveLMParams(eval, x, (int)D, param_const, max_iter);
  summary[0] = S.initial_cost; summary[1] = S.final_cost; summary[2] = S.iterations; summary[3] = S.successful; summary[4] = S.unsuccessful; summary[5] = S.termination;
}

void pvo_pixel_line_neighbors(int rows, int cols, const float* lines, int L, const float* cloud_local, int P, const double* T_cl, int* line3, float* d2_3, float* pixel2) {
  Equirect eq{rows, cols};
  PixelLineNeighbors(eq, lines, L, cloud_local, P, T_cl, line3, d2_3, pixel2);
}
int pvo_pixel_sub_lines(int rows, int cols, const float* lines, int L, int cap, float* mid2, int* sub_to_line) {
  Equirect eq{rows, cols};
  std::vector<float> mid; std::vector<int> s2l;
  PixelSubLines(eq, lines, L, mid, s2l);
  if ((int)s2l.size() > cap) return -1;
  std::memcpy(mid2, mid.data(), mid.size() * 4); std::memcpy(sub_to_line, s2l.data(), s2l.size() * 4);
  return (int)s2l.size();
}

void pvo_filter_line_pairs(int rows, int cols, int n, const float* image_line4, const double* start3, const double* end3, int by_angle, int by_length, unsigned char* keep,
                           float* angle) {
  Equirect eq{rows, cols};
  FilterLinePairs(eq, n, image_line4, start3, end3, by_angle != 0, by_length != 0, keep, angle);
}

void pvo_build_calibration_blocks(int rows, int cols, int n, const float* image_line4, const double* start3, const double* end3, int* type, double* huber, double* consts) {
  Equirect eq{rows, cols};
  BuildCalibrationBlocks(eq, n, image_line4, start3, end3, type, huber, consts);
}

void* pvo_kdtree_build(const float* pts, int n) { KdTree* t = new KdTree(); t->Build(pts, n, 4); return t; }
void pvo_kdtree_free(void* t) { delete (KdTree*)t; }

}  // extern "C"
