// ORACLE/_ref — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE'S OWN lidar_mapping/LidarFeatureAssociate.cpp (FindNeighbors,
// AssociatePoint2Plane, AssociateLine2Line + FindAssociations + TransformLines, AssociatePoint2Line, AssociatePoint2LineSegmentKNN,
// AssociatePoint2LineSegment, AssociateLine2LineKNN), compiled from the file where it lies under /root/reference (never copied) together with the
// headers it includes (LidarFeatureAssociate.h, sensors/Velodyne.h and its sub-headers, base/Geometry.hpp ...).  PCL / Eigen / OpenCV / glog / Boost are
// absent from this image: oracle/shim/ provides stand-ins - an Eigen-like matrix, and pcl::KdTreeFLANN as an exact float32 search (shim/pvo_shim_pcl.hpp).
// sensors/Velodyne.cpp is compiled too (its feature-extraction members call into translation units that are not: those entry points abort, see below), so
// the Velodyne class - poses, World2Local, Transform2LidarWorld / Transform2Local, UndistortCloud - is the reference's own.
// What this pins: the control flow, thresholds, class test, vote rules, conflict resolution and output order of the reference's association functions.
// Built by `make -C oracle ref` into oracle/_ref/libpvo_ref_assoc.so; used by tests/test_reference_pinning.py and tests/make_golden.py only.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <omp.h>
#define private public          // the wrapper reads / fills private members (Velodyne::world, Frame::keypoints_all, CameraLidarOptimizer's private stages)
#define REF(path) REF_STR(REFERENCE_ROOT/path)
#define REF_STR(x) REF_STR2(x)
#define REF_STR2(x) #x
#include REF(base/common.cpp)                                  // SplitString, IterateFiles (the file system stand-in lists nothing)
#include REF(base/ProcessBar.cpp)
#include REF(sensors/Velodyne.cpp)                             // the class itself: poses, Transform2LidarWorld / Transform2Local, World2Local, UndistortCloud ...
#include REF(sensors/Frame.cpp)
#include REF(sensors/Equirectangular.cpp)
#include REF(lidar_mapping/LidarFeatureAssociate.cpp)
#include REF(util/Tracks.cpp)                                  // TrackBuilder (union-find over (frame, line) features), Filter, ExportTracks
#include REF(lidar_mapping/LidarLineMatch.cpp)                 // GenerateTracks = FindNeighbors + AssociateLine2Line(nei, i, 0.3) + TrackBuilder
#include REF(util/Optimization.cpp)                            // the residual-block builders; ceres::Problem = the shim's recorder
#include REF(util/FileIO.cpp)                                  // ReadPoseT / ExportPoseT (pose text files)
#include REF(sfm/Structure.cpp)
#include REF(lidar_mapping/LidarOdometry.cpp)                  // RefinePose, UndistortLidars
#include REF(joint_optimization/CameraLidarLineAssociate.cpp)
#include REF(joint_optimization/CameraLidarOptimizer.cpp)      // NeighborEachFrame, LidarMaskByTrack, AssociateLineMulti, Optimize
#undef private

// ---- defined by the reference in translation units that are NOT compiled here (ground segmentation, LiDAR line / plane extraction, drawing) and reached only from
// Velodyne's feature-extraction members, which no entry point of this library calls: bodies that abort loudly, so that nothing can silently depend on them ----
static void not_compiled(const char* what) { std::fprintf(stderr, "oracle/_ref: %s belongs to a reference file that is not part of this build\n", what); std::abort(); }
cv::Vec3b Gray2Color(uchar) { not_compiled("Gray2Color (util/Visualization.cpp)"); return cv::Vec3b(); }
std::vector<std::vector<int>> PlaneSegmentation2(const pcl::PointCloud<pcl::PointXYZI>::Ptr, const Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic>&,
                                                 const std::vector<std::pair<size_t, size_t>>&, const std::vector<std::vector<int>>&, int, int) {
  not_compiled("PlaneSegmentation2 (sensors/LidarPlaneExtraction.cpp)"); return {};
}
void ExtractLineFeatures(const pcl::PointCloud<pcl::PointXYZI>&, const std::vector<std::pair<size_t, size_t>>&, std::vector<pcl::PointCloud<pcl::PointXYZI>>&, eigen_vector<Vector6d>&) {
  not_compiled("ExtractLineFeatures (sensors/LidarLineExtraction.cpp)");
}
GroundSegmentation::GroundSegmentation(const GroundSegmentationParams& p) : params_(p) { not_compiled("GroundSegmentation (sensors/ground_segmentation.cpp)"); }
void GroundSegmentation::segment(const PointCloud&, std::vector<int>&) { not_compiled("GroundSegmentation::segment"); }

// ---- class PanoramaLine (util/PanoramaLine.cpp is image line DETECTION on real OpenCV algorithms and is not compiled): the three trivial members the joint stage needs
// to carry already-detected lines around.  Everything else of that file, and the other non-compiled files, are abort placeholders in ref_unresolved_stubs.c ----
PanoramaLine::PanoramaLine() : id(-1), rows(0), cols(0) {}
PanoramaLine::~PanoramaLine() {}
void PanoramaLine::SetName(const std::string& _name) { name = _name; }
const std::vector<cv::Vec4f>& PanoramaLine::GetLines() const { return lines; }

namespace {
void fill_cloud(pcl::PointCloud<PointType>& c, const float* xyzi, int n) {
  c.clear();
  for (int i = 0; i < n; ++i) { PointType p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; c.push_back(p); }
}
}  // namespace

extern "C" {
// A frame as the association code sees it: pose (R row-major, t; pose_valid = 0 leaves the constructor's "no pose" state), the three feature clouds (float32
// x, y, z, intensity), the point -> segment sets of cornerLessSharp (CSR), the segment coefficients in the SENSOR frame.  clouds_in_world = 0: the clouds are
// given in the SENSOR frame and the reference's own Velodyne::Transform2LidarWorld() moves them (the state RefinePose starts from, LidarOdometry.cpp:17-21);
// clouds_in_world = 1: they are already world-frame and only the flag is set; 2: sensor-frame clouds, left there (RefinePose / LidarMaskByTrack move them themselves).
void* ref_frame_create(int id, int valid, int pose_valid, const double* R_wl, const double* t_wl, const float* corner, int n_corner, const int* p2s_off,
                       const int* p2s_ids, int S, const double* coeffs_local, const int* seg_sizes, const float* surf_flat, int n_flat,
                       const float* surf_less_flat, int n_less, int clouds_in_world, const double* end_points) {
  Velodyne* v = new Velodyne();
  v->id = id; v->valid = valid != 0;
  if (pose_valid) {
    Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wl[3 * i + j];
    v->SetPose(R, Eigen::Vector3d(t_wl[0], t_wl[1], t_wl[2]));
  }
  fill_cloud(v->cornerLessSharp, corner, n_corner);
  fill_cloud(v->surfFlat, surf_flat, n_flat);
  fill_cloud(v->surfLessFlat, surf_less_flat, n_less);
  v->point_to_segment.resize(n_corner);
  if (p2s_off) for (int i = 0; i < n_corner; ++i) for (int k = p2s_off[i]; k < p2s_off[i + 1]; ++k) v->point_to_segment[i].insert(p2s_ids[k]);
  v->edge_segmented.resize(S);
  for (int s = 0; s < S; ++s) {
    Vector6d c; for (int k = 0; k < 6; ++k) c[k] = coeffs_local[6 * s + k];
    v->segment_coeffs.push_back(c);
  }
  if (end_points) for (int i = 0; i < 2 * S; ++i) v->end_points.push_back(Eigen::Vector3d(end_points[3 * i], end_points[3 * i + 1], end_points[3 * i + 2]));   // sensor frame, projected on the line
  // the points of a segment = the cornerLessSharp points whose set holds it (Velodyne::EdgeToLine fills both from the same lists); when the caller
  // gives explicit sizes they must agree
  for (int i = 0; i < n_corner; ++i) for (int s : v->point_to_segment[i]) v->edge_segmented[s].push_back(v->cornerLessSharp.points[i]);
  if (seg_sizes) for (int s = 0; s < S; ++s) if ((int)v->edge_segmented[s].size() != seg_sizes[s]) { delete v; return nullptr; }
  if (clouds_in_world == 1) v->world = 1;
  else if (clouds_in_world == 0 && pose_valid) v->Transform2LidarWorld();       // 2: sensor-frame clouds stay where they are (the caller's pipeline moves them)
  return v;
}
// which: 0 cornerLessSharp, 1 surfFlat, 2 surfLessFlat; out: n x 4 float.  Returns the point count.
int ref_frame_get_cloud(const void* f, int which, int cap, float* out) {
  const Velodyne* v = static_cast<const Velodyne*>(f);
  const pcl::PointCloud<PointType>& c = which == 0 ? v->cornerLessSharp : (which == 1 ? v->surfFlat : v->surfLessFlat);
  if ((int)c.size() > cap) return -1;
  for (size_t i = 0; i < c.size(); ++i) { out[4 * i] = c.points[i].x; out[4 * i + 1] = c.points[i].y; out[4 * i + 2] = c.points[i].z; out[4 * i + 3] = c.points[i].intensity; }
  return (int)c.size();
}
// Velodyne::Transform2Local() after Transform2LidarWorld(): the round trip the joint stage makes every outer iteration (CameraLidarOptimizer.cpp:416, 535-536)
void ref_frame_to_local(void* f) { static_cast<Velodyne*>(f)->Transform2Local(); }
void ref_frame_to_world(void* f) { static_cast<Velodyne*>(f)->Transform2LidarWorld(); }
// Velodyne::UndistortCloud(R_we, t_we) (sensors/Velodyne.cpp:1642-1674) on a raw sweep: cloud n x 4 float in, undistorted cloud out.  Returns 1 / 0 like the member.
int ref_undistort_cloud(const double* R_wl, const double* t_wl, const double* R_we, const double* t_we, const float* cloud, long n, float* out) {
  Velodyne v;
  Eigen::Matrix3d Rl, Re;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Rl(i, j) = R_wl[3 * i + j]; Re(i, j) = R_we[3 * i + j]; }
  v.SetPose(Rl, Eigen::Vector3d(t_wl[0], t_wl[1], t_wl[2]));
  fill_cloud(v.cloud, cloud, (int)n);
  const bool ok = v.UndistortCloud(Re, Eigen::Vector3d(t_we[0], t_we[1], t_we[2]));
  for (long i = 0; i < n; ++i) { out[4 * i] = v.cloud.points[i].x; out[4 * i + 1] = v.cloud.points[i].y; out[4 * i + 2] = v.cloud.points[i].z; out[4 * i + 3] = v.cloud.points[i].intensity; }
  return ok ? 1 : 0;
}
void ref_frame_destroy(void* f) { delete static_cast<Velodyne*>(f); }

int ref_associate_point2plane(const void* ref, const void* nei, double plane_tolerance, float dist_threshold, int cap, double* point3, double* plane4) {
  const std::vector<Point2Plane> a = AssociatePoint2Plane(*static_cast<const Velodyne*>(ref), *static_cast<const Velodyne*>(nei), plane_tolerance, dist_threshold, false);
  if ((int)a.size() > cap) return -1;
  for (size_t i = 0; i < a.size(); ++i) { for (int k = 0; k < 3; ++k) point3[3 * i + k] = a[i].point[k]; for (int k = 0; k < 4; ++k) plane4[4 * i + k] = a[i].plane_coeff[k]; }
  return (int)a.size();
}
static int put_p2l(const std::vector<Point2Line>& a, int cap, double* point3, double* a3, double* b3) {
  if ((int)a.size() > cap) return -1;
  for (size_t i = 0; i < a.size(); ++i) for (int k = 0; k < 3; ++k) { point3[3 * i + k] = a[i].point[k]; a3[3 * i + k] = a[i].line_point1[k]; b3[3 * i + k] = a[i].line_point2[k]; }
  return (int)a.size();
}
static int put_l2l(const std::vector<Line2Line>& a, int cap, int* nei_idx, int* ref_idx, double* a3, double* b3) {
  if ((int)a.size() > cap) return -1;
  for (size_t i = 0; i < a.size(); ++i) { nei_idx[i] = a[i].neighbor_line_idx; ref_idx[i] = a[i].ref_line_idx; for (int k = 0; k < 3; ++k) { a3[3 * i + k] = a[i].line_point1[k]; b3[3 * i + k] = a[i].line_point2[k]; } }
  return (int)a.size();
}
int ref_associate_point2line(const void* ref, const void* nei, float dist_threshold, int cap, double* point3, double* a3, double* b3) {
  return put_p2l(AssociatePoint2Line(*static_cast<const Velodyne*>(ref), *static_cast<const Velodyne*>(nei), dist_threshold, false), cap, point3, a3, b3);
}
int ref_associate_point2line_segment_knn(const void* ref, const void* nei, float dist_threshold, int cap, double* point3, double* a3, double* b3) {
  return put_p2l(AssociatePoint2LineSegmentKNN(*static_cast<const Velodyne*>(ref), *static_cast<const Velodyne*>(nei), dist_threshold, false), cap, point3, a3, b3);
}
int ref_associate_point2line_segment(const void* ref, const void* nei, float dist_threshold, int cap, double* point3, double* a3, double* b3) {
  return put_p2l(AssociatePoint2LineSegment(*static_cast<const Velodyne*>(ref), *static_cast<const Velodyne*>(nei), dist_threshold, false), cap, point3, a3, b3);
}
int ref_associate_line2line(const void* ref, const void* nei, float dist_threshold, int cap, int* nei_idx, int* ref_idx, double* a3, double* b3) {
  return put_l2l(AssociateLine2Line(*static_cast<const Velodyne*>(ref), *static_cast<const Velodyne*>(nei), dist_threshold, false), cap, nei_idx, ref_idx, a3, b3);
}
int ref_associate_line2line_knn(const void* ref, const void* nei, float dist_threshold, int cap, int* nei_idx, int* ref_idx, double* a3, double* b3) {
  return put_l2l(AssociateLine2LineKNN(*static_cast<const Velodyne*>(ref), *static_cast<const Velodyne*>(nei), dist_threshold, false), cap, nei_idx, ref_idx, a3, b3);
}
void ref_transform_lines(const double* T_rowmajor16, int n, const double* lines6, double* out6) {
  Eigen::Matrix4d T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T_rowmajor16[4 * i + j];
  eigen_vector<Vector6d> in;
  for (int i = 0; i < n; ++i) { Vector6d c; for (int k = 0; k < 6; ++k) c[k] = lines6[6 * i + k]; in.push_back(c); }
  const eigen_vector<Vector6d> out = TransformLines(in, T);
  for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) out6[6 * i + k] = out[i][k];
}
// LidarLineMatch::GenerateTracks over frames made by ref_frame_create (copied into a std::vector<Velodyne>).  Output: track t = features
// [off[t], off[t + 1]) as (frame, line) pairs in the set's order.  Returns the number of tracks, or -1 when cap is too small.
int ref_generate_line_tracks(int n, void* const* frames, int neighbor_size, int min_track_length, int cap, int* off, int* feat_frame, int* feat_line) {
  std::vector<Velodyne> lidars;
  for (int i = 0; i < n; ++i) lidars.push_back(*static_cast<const Velodyne*>(frames[i]));
  LidarLineMatch m(lidars);
  m.SetNeighborSize(neighbor_size);
  m.SetMinTrackLength(min_track_length);
  m.GenerateTracks();
  const std::vector<LineTrack>& tr = m.GetTracks();
  int total = 0; off[0] = 0;
  for (size_t t = 0; t < tr.size(); ++t) {
    for (const auto& f : tr[t].feature_pairs) { if (total >= cap) return -1; feat_frame[total] = (int)f.first; feat_line[total] = (int)f.second; ++total; }
    off[t + 1] = total;
  }
  return (int)tr.size();
}

// ---- what ceres::Solve sees: a snapshot taken by the stand-in's solve hook (the solver itself is not reproduced) ----
namespace {
struct Snapshot {
  std::vector<int> ref, nei; std::vector<double> huber, residual, jac; std::vector<const double*> constant; std::vector<double> poses;
  int n_frames = 0, first_valid = 0, max_num_iterations = 0, linear_solver = -1; long status = 0;
};
Snapshot* g_snap = nullptr;
// RefinePose's problem: 4-block residuals over (aa_lw[i], t_lw[i]); the first valid frame's two blocks are constant (LidarOdometry.cpp:58-64), which gives the
// base addresses of the two pose lists (contiguous eigen_vector<Vector3d>)
void lidar_hook(const ceres::Solver::Options& o, ceres::Problem* p, ceres::Solver::Summary* s) {
  Snapshot& S = *g_snap;
  S.max_num_iterations = o.max_num_iterations; S.linear_solver = (int)o.linear_solver_type;
  S.constant = p->constant_blocks;
  if (p->constant_blocks.size() != 2) { S.status = -3; return; }
  const Eigen::Vector3d* aa = (const Eigen::Vector3d*)p->constant_blocks[0] - S.first_valid;
  const Eigen::Vector3d* tt = (const Eigen::Vector3d*)p->constant_blocks[1] - S.first_valid;
  S.poses.resize(6 * S.n_frames);
  for (int i = 0; i < S.n_frames; ++i) for (int k = 0; k < 3; ++k) { S.poses[6 * i + k] = aa[i][k]; S.poses[6 * i + 3 + k] = tt[i][k]; }
  for (const ceres::Problem::Block& blk : p->blocks) {
    if (blk.params.size() != 4) { S.status = -2; return; }
    const int r = (int)((const Eigen::Vector3d*)blk.params[0] - aa), n = (int)((const Eigen::Vector3d*)blk.params[2] - aa);
    if (r < 0 || r >= S.n_frames || n < 0 || n >= S.n_frames || blk.params[1] != (const double*)(tt + r) || blk.params[3] != (const double*)(tt + n)) { S.status = -2; return; }
    S.ref.push_back(r); S.nei.push_back(n);
    const ceres::HuberLoss* h = dynamic_cast<const ceres::HuberLoss*>(blk.loss);
    S.huber.push_back(h ? h->a() : 0.0);
    double res, jb[4][3]; double* jp[4] = {jb[0], jb[1], jb[2], jb[3]};
    if (!blk.cost->Evaluate(blk.params.data(), &res, jp)) { S.status = -2; return; }
    S.residual.push_back(res); S.jac.insert(S.jac.end(), &jb[0][0], &jb[0][0] + 12);
  }
  s->usable = true; s->final_cost = 0; s->num_successful_steps = 0;      // keeps RefinePose off its failure branch (which writes debug files)
}
Config make_config(int point_to_plane, int line_to_line, int point_to_line, int angle_residual, int normalize_distance, float plane_dis_threshold, float line_dis_threshold,
                   float plane_tolerance) {
  Config c;                                             // base/Config.h defaults; the fields below are the ones RefinePose reads (config/*.txt: lines 67-76)
  c.point_to_plane_residual = point_to_plane != 0; c.line_to_line_residual = line_to_line != 0; c.point_to_line_residual = point_to_line != 0;
  c.angle_residual = angle_residual != 0; c.normalize_distance = normalize_distance != 0;
  c.point_to_plane_dis_threshold = plane_dis_threshold; c.point_to_line_dis_threshold = line_dis_threshold; c.lidar_plane_tolerance = plane_tolerance;
  c.num_threads = 1;
  return c;
}
}  // namespace

// The residual blocks of one LidarOdometry::RefinePose call: the reference's OWN RefinePose (lidar_mapping/LidarOdometry.cpp:15-114) runs - pose blocks, FindNeighbors(6),
// GenerateTracks(4, 3), the builders, SetParameterBlockConstant of the first valid frame, SetOptionsLidar - and the stand-in's ceres::Solve hook records the problem it
// is handed, evaluating every block once through ceres::CostFunction::Evaluate.  The thresholds travel as `float`, as they do in base/Config.h.
// Per block: reference / neighbour frame, Huber parameter a (0 = loss nullptr), raw residual, raw 1x12 Jacobian.  poses_out: n x 6 pose blocks; info4 = {max_num_iterations,
// linear_solver_type, number of constant blocks, first valid frame}.  Returns the number of blocks, -1 when cap is too small, -2 / -3 on an inconsistency.
long ref_refine_pose_blocks(int n, void* const* frames, int point_to_plane, int line_to_line, int point_to_line, int use_segment, int angle_residual, int normalize_distance,
                            double plane_dis_threshold, double line_dis_threshold, double plane_tolerance, long cap, int* ref_frame, int* nei_frame, double* huber_a,
                            double* residual, double* jac12, double* poses_out, int* info4) {
  std::vector<Velodyne> lidars;
  for (int i = 0; i < n; ++i) lidars.push_back(*static_cast<const Velodyne*>(frames[i]));
  for (Velodyne& l : lidars) if (l.IsInWorldCoordinate()) l.Transform2Local();                 // RefinePose itself moves them to the world frame (:17-21)
  const Config config = make_config(point_to_plane, line_to_line, point_to_line, angle_residual, normalize_distance, (float)plane_dis_threshold, (float)line_dis_threshold,
                                    (float)plane_tolerance);
  LidarOdometry odo(lidars, config);
  Snapshot S; S.n_frames = n;
  for (S.first_valid = 0; S.first_valid < n; ++S.first_valid) if (lidars[S.first_valid].IsPoseValid() && lidars[S.first_valid].valid) break;
  g_snap = &S; ceres::solve_hook() = lidar_hook;
  double cost = 0; int steps = 0;
  odo.RefinePose(cost, steps, use_segment != 0);
  ceres::solve_hook() = nullptr; g_snap = nullptr;
  if (S.status < 0) return S.status;
  if ((long)S.residual.size() > cap) return -1;
  for (size_t b = 0; b < S.residual.size(); ++b) { ref_frame[b] = S.ref[b]; nei_frame[b] = S.nei[b]; huber_a[b] = S.huber[b]; residual[b] = S.residual[b]; }
  if (!S.jac.empty()) std::memcpy(jac12, S.jac.data(), S.jac.size() * sizeof(double));
  if (!S.poses.empty()) std::memcpy(poses_out, S.poses.data(), S.poses.size() * sizeof(double));
  info4[0] = S.max_num_iterations; info4[1] = S.linear_solver; info4[2] = (int)S.constant.size(); info4[3] = S.first_valid;
  return (long)S.residual.size();
}

// LidarOdometry::UndistortLidars (lidar_mapping/LidarOdometry.cpp:189-263): sweep-end pose selection (next / previous frame with a pose, SlerpPose with the
// duration ratio) + Velodyne::UndistortCloud per frame.  n frames with pose (R row-major 9, t 3), pose_valid / valid flags, raw sweeps concatenated (off[n + 1],
// n_points x 4 float).  out: the sweeps as the reference leaves them in `lidars[i].cloud`.
void ref_undistort_lidars(int n, const double* R_wl, const double* t_wl, const unsigned char* pose_valid, const unsigned char* valid, const int* off, const float* clouds,
                          float gap_time, float* out) {
  std::vector<Velodyne> lidars(n);
  for (int f = 0; f < n; ++f) {
    lidars[f].id = f; lidars[f].valid = valid[f] != 0;
    if (pose_valid[f]) { Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wl[9 * f + 3 * i + j]; lidars[f].SetPose(R, Eigen::Vector3d(t_wl[3 * f], t_wl[3 * f + 1], t_wl[3 * f + 2])); }
    fill_cloud(lidars[f].cloud, clouds + 4 * (size_t)off[f], off[f + 1] - off[f]);
  }
  Config config; config.num_threads = 1;
  LidarOdometry odo(lidars, config);
  odo.UndistortLidars(gap_time);
  const std::vector<Velodyne>& res = odo.GetLidarData();
  for (int f = 0; f < n; ++f) for (int i = 0; i < off[f + 1] - off[f]; ++i) {
    const PointType& p = res[f].cloud.points[i]; float* o = out + 4 * ((size_t)off[f] + i);
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.intensity;
  }
}

// ExportPoseT / ReadPoseT (util/FileIO.cpp:11-73, 168-191): pose text files.  names may be null (no name column).
void ref_export_pose_t(const char* path, int n, const double* R9, const double* t3, const char* const* names) {
  eigen_vector<Eigen::Matrix3d> Rl; eigen_vector<Eigen::Vector3d> tl; std::vector<std::string> nl;
  for (int f = 0; f < n; ++f) {
    Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R9[9 * f + 3 * i + j];
    Rl.push_back(R); tl.push_back(Eigen::Vector3d(t3[3 * f], t3[3 * f + 1], t3[3 * f + 2]));
    if (names) nl.push_back(names[f]);
  }
  ExportPoseT(path, Rl, tl, nl);
}
int ref_read_pose_t(const char* path, int with_invalid, int cap, double* R9, double* t3, char* names_256) {
  eigen_vector<Eigen::Matrix3d> Rl; eigen_vector<Eigen::Vector3d> tl; std::vector<std::string> nl;
  if (!ReadPoseT(path, with_invalid != 0, Rl, tl, nl)) return -1;
  if ((int)Rl.size() > cap) return -2;
  for (size_t f = 0; f < Rl.size(); ++f) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[9 * f + 3 * i + j] = Rl[f](i, j);
    for (int k = 0; k < 3; ++k) t3[3 * f + k] = tl[f][k];
    std::snprintf(names_256 + 256 * f, 256, "%s", nl[f].c_str());
  }
  return (int)Rl.size();
}

// The problem that the reference's OWN mapping-mode CameraLidarOptimizer::Optimize (joint_optimization/CameraLidarOptimizer.cpp:387-548) hands to ceres::Solve, with the line
// pairs coming from its own AssociateLineMulti(neighbor_size_joint, temporal = true, no track masks) (:330-384).  Inputs: n camera frames (pose T_wc, image lines CSR,
// key points CSR), the LiDAR frames (ref_frame_create with clouds_in_world = 2 and end points), structure tracks (CSR of (frame, key point)) with their 3-D points,
// T_cl_init, the weights / switches of base/Config.h.  Per recorded block: n_params (4: pose-pose, 3: reprojection), a / b = pose-block indices in the layout
// [cameras 0..n) | LiDARs n..n+m) (4-block) or (camera, track) (3-block), Huber a, raw residual, raw Jacobian (12 or 9 entries of 12).  const_part: for every constant
// parameter block an id: pose block b rotation -> 2 b, translation -> 2 b + 1, track t -> 2 (n + m) + t.  poses_out: (n + m) x 6.
namespace {
struct JointSnapshot {
  int n_cams = 0, n_lidars = 0; const std::vector<PointTrack>* structure = nullptr;
  std::vector<int> n_params, a, b, const_part; std::vector<double> huber, residual, jac, poses; long status = 0;
};
JointSnapshot* g_joint = nullptr;
void joint_hook(const ceres::Solver::Options&, ceres::Problem* p, ceres::Solver::Summary* s) {
  JointSnapshot& S = *g_joint;
  typedef const Eigen::Vector3d* V;
  const size_t nc = p->constant_blocks.size();
  if (nc < 2) { S.status = -3; return; }
  const V caa = (V)p->constant_blocks[nc - 2], ctt = (V)p->constant_blocks[nc - 1];            // :490-491: camera 0 is made constant last
  auto in = [](V base, int n, const double* q) { return (V)q >= base && (V)q < base + n; };
  V laa = nullptr, ltt = nullptr;                                                               // the LiDAR lists: lowest addresses seen outside the camera lists
  for (const ceres::Problem::Block& blk : p->blocks) if (blk.params.size() == 4) for (int k = 0; k < 4; ++k) {
    const double* q = blk.params[k];
    if (in(caa, S.n_cams, q) || in(ctt, S.n_cams, q)) continue;
    V& base = (k % 2 == 0) ? laa : ltt;
    if (!base || (V)q < base) base = (V)q;
  }
  auto pose_index = [&](const double* q, int part) -> int {                                       // part 0: rotation list, 1: translation list
    if (in(part ? ctt : caa, S.n_cams, q)) return (int)((V)q - (part ? ctt : caa));
    V base = part ? ltt : laa;
    if (base && in(base, S.n_lidars, q)) return S.n_cams + (int)((V)q - base);
    return -1;
  };
  for (const ceres::Problem::Block& blk : p->blocks) {
    const int np = (int)blk.params.size();
    int ia = -1, ib = -1;
    if (np == 4) {
      ia = pose_index(blk.params[0], 0); ib = pose_index(blk.params[2], 0);
      if (ia < 0 || ib < 0 || pose_index(blk.params[1], 1) != ia || pose_index(blk.params[3], 1) != ib) { S.status = -2; return; }
    } else if (np == 3) {
      ia = pose_index(blk.params[0], 0);
      for (size_t t = 0; t < S.structure->size(); ++t) if (blk.params[2] == (*S.structure)[t].point_3d.data()) { ib = (int)t; break; }
      if (ia < 0 || ib < 0 || pose_index(blk.params[1], 1) != ia) { S.status = -2; return; }
    } else { S.status = -2; return; }
    const ceres::HuberLoss* h = dynamic_cast<const ceres::HuberLoss*>(blk.loss);
    double res, jb[4][3] = {{0}}; double* jp[4] = {jb[0], jb[1], jb[2], jb[3]};
    if (!blk.cost->Evaluate(blk.params.data(), &res, jp)) { S.status = -2; return; }
    S.n_params.push_back(np); S.a.push_back(ia); S.b.push_back(ib); S.huber.push_back(h ? h->a() : 0.0); S.residual.push_back(res);
    S.jac.insert(S.jac.end(), &jb[0][0], &jb[0][0] + 12);
  }
  for (const double* q : p->constant_blocks) {
    int id = -1, k;
    if ((k = pose_index(q, 0)) >= 0) id = 2 * k;
    else if ((k = pose_index(q, 1)) >= 0) id = 2 * k + 1;
    else for (size_t t = 0; t < S.structure->size(); ++t) if (q == (*S.structure)[t].point_3d.data()) { id = 2 * (S.n_cams + S.n_lidars) + (int)t; break; }
    S.const_part.push_back(id);
  }
  S.poses.assign(6 * (S.n_cams + S.n_lidars), 0.0);
  for (int i = 0; i < S.n_cams; ++i) for (int k = 0; k < 3; ++k) { S.poses[6 * i + k] = caa[i][k]; S.poses[6 * i + 3 + k] = ctt[i][k]; }
  if (laa && ltt) for (int i = 0; i < S.n_lidars; ++i) for (int k = 0; k < 3; ++k) { S.poses[6 * (S.n_cams + i) + k] = laa[i][k]; S.poses[6 * (S.n_cams + i) + 3 + k] = ltt[i][k]; }
  s->usable = true; s->final_cost = 0;
}
}  // namespace

long ref_joint_optimize_blocks(int rows, int cols, int n_cams, const double* R_wc, const double* t_wc, const int* line_off, const float* lines4, const int* kp_off, const float* kp_xy,
                               int n_lidars, void* const* lidar_frames, int n_tracks, const int* track_off, const int* feat_frame, const int* feat_index, const double* points3,
                               const double* T_cl_init16, int neighbor_size_joint, double camera_weight, double lidar_weight, double camera_lidar_weight, int point_to_plane,
                               int line_to_line, int point_to_line, int angle_residual, int normalize_distance, double plane_dis_threshold, double line_dis_threshold,
                               double plane_tolerance, int refine_camera_rotation, int refine_camera_trans, int refine_lidar_rotation, int refine_lidar_trans, int refine_structure,
                               long cap, int* n_params, int* a, int* b, double* huber_a, double* residual, double* jac12, int const_cap, int* const_part, int* n_const,
                               double* poses_out, int* n_line_pairs) {
  std::vector<Frame> frames; std::vector<Velodyne> lidars; std::vector<PanoramaLine> image_lines(n_cams);
  for (int f = 0; f < n_cams; ++f) {
    frames.push_back(Frame(rows, cols, f, "frame"));
    Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wc[9 * f + 3 * i + j];
    frames[f].SetPose(R, Eigen::Vector3d(t_wc[3 * f], t_wc[3 * f + 1], t_wc[3 * f + 2]));
    for (int k = kp_off[f]; k < kp_off[f + 1]; ++k) { cv::KeyPoint kp; kp.pt = cv::Point2f(kp_xy[2 * k], kp_xy[2 * k + 1]); frames[f].keypoints_all.push_back(kp); }
    image_lines[f].id = f; image_lines[f].rows = rows; image_lines[f].cols = cols;
    for (int k = line_off[f]; k < line_off[f + 1]; ++k) image_lines[f].lines.push_back(cv::Vec4f(lines4[4 * k], lines4[4 * k + 1], lines4[4 * k + 2], lines4[4 * k + 3]));
  }
  for (int i = 0; i < n_lidars; ++i) lidars.push_back(*static_cast<const Velodyne*>(lidar_frames[i]));
  std::vector<PointTrack> structure;
  structure.reserve(n_tracks);
  for (int t = 0; t < n_tracks; ++t) {
    std::set<std::pair<uint32_t, uint32_t>> fp;
    for (int k = track_off[t]; k < track_off[t + 1]; ++k) fp.insert(std::make_pair((uint32_t)feat_frame[k], (uint32_t)feat_index[k]));
    structure.push_back(PointTrack(t, fp, Eigen::Vector3d(points3[3 * t], points3[3 * t + 1], points3[3 * t + 2])));
  }
  Config config = make_config(point_to_plane, line_to_line, point_to_line, angle_residual, normalize_distance, (float)plane_dis_threshold, (float)line_dis_threshold, (float)plane_tolerance);
  config.camera_weight = camera_weight; config.lidar_weight = lidar_weight; config.camera_lidar_weight = camera_lidar_weight; config.neighbor_size_joint = neighbor_size_joint;
  Eigen::Matrix4d T_cl; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_cl(i, j) = T_cl_init16[4 * i + j];
  CameraLidarOptimizer opt(T_cl, lidars, frames, config);
  opt.image_lines_all = image_lines;
  const eigen_map<std::pair<size_t, size_t>, std::vector<CameraLidarLinePair>> line_pairs = opt.AssociateLineMulti(neighbor_size_joint, true, false, false);
  *n_line_pairs = 0; for (const auto& kv : line_pairs) *n_line_pairs += (int)kv.second.size();
  JointSnapshot S; S.n_cams = n_cams; S.n_lidars = n_lidars; S.structure = &structure;
  g_joint = &S; ceres::solve_hook() = joint_hook;
  double cost = 0; int steps = 0;
  opt.Optimize(line_pairs, structure, refine_camera_rotation != 0, refine_camera_trans != 0, refine_lidar_rotation != 0, refine_lidar_trans != 0, refine_structure != 0, cost, steps);
  ceres::solve_hook() = nullptr; g_joint = nullptr;
  if (S.status < 0) return S.status;
  if ((long)S.residual.size() > cap || (int)S.const_part.size() > const_cap) return -1;
  for (size_t k = 0; k < S.residual.size(); ++k) { n_params[k] = S.n_params[k]; a[k] = S.a[k]; b[k] = S.b[k]; huber_a[k] = S.huber[k]; residual[k] = S.residual[k]; }
  if (!S.jac.empty()) std::memcpy(jac12, S.jac.data(), S.jac.size() * sizeof(double));
  for (size_t k = 0; k < S.const_part.size(); ++k) const_part[k] = S.const_part[k];
  *n_const = (int)S.const_part.size();
  std::memcpy(poses_out, S.poses.data(), S.poses.size() * sizeof(double));
  return (long)S.residual.size();
}

// The OUTER loop of the mapping mode: the reference's own CameraLidarOptimizer::JointOptimize (:236-287) - AssociateLineMulti, then up to num_iteration_joint times Optimize +
// re-association, with its two early exits - run with a SCRIPTED solver: the k-th ceres::Solve call reports final_cost = script_cost[k] and num_successful_steps =
// script_steps[k] and leaves the parameters alone.  Returns how many times Optimize reached the solver (or < 0).  Same inputs as ref_joint_optimize_blocks.
namespace {
struct LoopScript { const double* cost; const int* steps; int n, calls; };
LoopScript* g_loop = nullptr;
void loop_hook(const ceres::Solver::Options&, ceres::Problem*, ceres::Solver::Summary* s) {
  LoopScript& L = *g_loop;
  const int k = std::min(L.calls, L.n - 1);
  s->usable = true; s->final_cost = L.cost[k]; s->num_successful_steps = L.steps[k];
  ++L.calls;
}
}  // namespace
int ref_joint_optimize_loop(int rows, int cols, int n_cams, const double* R_wc, const double* t_wc, const int* line_off, const float* lines4, int n_lidars, void* const* lidar_frames,
                            const double* T_cl_init16, int num_iteration_joint, int n_script, const double* script_cost, const int* script_steps) {
  std::vector<Frame> frames; std::vector<Velodyne> lidars; std::vector<PanoramaLine> image_lines(n_cams);
  for (int f = 0; f < n_cams; ++f) {
    frames.push_back(Frame(rows, cols, f, "frame"));
    Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wc[9 * f + 3 * i + j];
    frames[f].SetPose(R, Eigen::Vector3d(t_wc[3 * f], t_wc[3 * f + 1], t_wc[3 * f + 2]));
    image_lines[f].id = f; image_lines[f].rows = rows; image_lines[f].cols = cols;
    for (int k = line_off[f]; k < line_off[f + 1]; ++k) image_lines[f].lines.push_back(cv::Vec4f(lines4[4 * k], lines4[4 * k + 1], lines4[4 * k + 2], lines4[4 * k + 3]));
  }
  for (int i = 0; i < n_lidars; ++i) lidars.push_back(*static_cast<const Velodyne*>(lidar_frames[i]));
  Config config = make_config(1, 0, 0, 1, 1, 1.0f, 0.3f, 0.05f);
  config.num_iteration_joint = num_iteration_joint; config.neighbor_size_joint = 1;
  Eigen::Matrix4d T_cl; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_cl(i, j) = T_cl_init16[4 * i + j];
  CameraLidarOptimizer opt(T_cl, lidars, frames, config);
  opt.image_lines_all = image_lines;
  opt.SetOptimizationMode(MAPPING);
  LoopScript L{script_cost, script_steps, n_script, 0};
  g_loop = &L; ceres::solve_hook() = loop_hook;
  const bool ok = opt.JointOptimize(false);
  ceres::solve_hook() = nullptr; g_loop = nullptr;
  return ok ? L.calls : -1;
}

// Calibration mode: the reference's own AssociateLineSingle(T_cl) (joint_optimization/CameraLidarOptimizer.cpp:300-328: image i against LiDAR i, AssociateByAngle with its
// defaults = one-to-one pairs) followed by Optimize(line_pairs, T_cl) (:32-96: Plane2Plane_Relative with HuberLoss(2 deg) + PlaneRelativeIOUResidual without a loss on ONE
// relative pose block), recorded at ceres::Solve.  Per block: Huber a, raw residual, raw 1x6 Jacobian (aa_cl, t_cl); pose6 = the pose block; info3 = {number of line
// pairs, max_num_iterations, linear_solver_type}.  Returns the number of blocks or < 0.
namespace {
struct CalibSnapshot { std::vector<double> huber, residual, jac, pose; int max_it = 0, solver = -1; long status = 0; };
CalibSnapshot* g_calib = nullptr;
void calib_hook(const ceres::Solver::Options& o, ceres::Problem* p, ceres::Solver::Summary* s) {
  CalibSnapshot& S = *g_calib;
  S.max_it = o.max_num_iterations; S.solver = (int)o.linear_solver_type;
  for (const ceres::Problem::Block& blk : p->blocks) {
    if (blk.params.size() != 2 || blk.params[0] != p->blocks[0].params[0] || blk.params[1] != p->blocks[0].params[1]) { S.status = -2; return; }
    const ceres::HuberLoss* h = dynamic_cast<const ceres::HuberLoss*>(blk.loss);
    double res, jb[2][3]; double* jp[2] = {jb[0], jb[1]};
    if (!blk.cost->Evaluate(blk.params.data(), &res, jp)) { S.status = -2; return; }
    S.huber.push_back(h ? h->a() : 0.0); S.residual.push_back(res); S.jac.insert(S.jac.end(), &jb[0][0], &jb[0][0] + 6);
  }
  if (!p->blocks.empty()) { S.pose.assign(p->blocks[0].params[0], p->blocks[0].params[0] + 3); S.pose.insert(S.pose.end(), p->blocks[0].params[1], p->blocks[0].params[1] + 3); }
  s->usable = true;
}
}  // namespace
// Scripts the answers of the pcl::SACSegmentation stand-in inside THIS library (oracle/shim/pvo_shim_pcl.hpp: the k-th segment() call reports inlier list k), so that the
// pixel-space Associate() behind AssociateLineSingle / AssociateLineMulti (frames without LiDAR segments) runs to its end.  n_lists < 0 clears the script; returns the
// number of segment() calls answered since the script was set.
int ref_set_sac_script(int n_lists, const int* off, const int* idx) {
  static std::pair<std::vector<std::vector<int>>, size_t> script;
  const int answered = (int)script.second;
  script.first.clear(); script.second = 0;
  if (n_lists < 0) { pcl::sac_script() = nullptr; return answered; }
  for (int k = 0; k < n_lists; ++k) script.first.push_back(std::vector<int>(idx + off[k], idx + off[k + 1]));
  pcl::sac_script() = &script;
  return answered;
}
long ref_calibration_blocks(int rows, int cols, int n, const int* line_off, const float* lines4, void* const* lidar_frames, const double* T_cl16, long cap, double* huber_a,
                            double* residual, double* jac6, double* pose6, int* info3) {
  std::vector<Frame> frames; std::vector<Velodyne> lidars; std::vector<PanoramaLine> image_lines(n);
  for (int f = 0; f < n; ++f) {
    frames.push_back(Frame(rows, cols, f, "frame"));
    image_lines[f].id = f; image_lines[f].rows = rows; image_lines[f].cols = cols;
    for (int k = line_off[f]; k < line_off[f + 1]; ++k) image_lines[f].lines.push_back(cv::Vec4f(lines4[4 * k], lines4[4 * k + 1], lines4[4 * k + 2], lines4[4 * k + 3]));
    lidars.push_back(*static_cast<const Velodyne*>(lidar_frames[f]));
  }
  Config config; config.num_threads = 1;
  Eigen::Matrix4d T_cl; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_cl(i, j) = T_cl16[4 * i + j];
  CameraLidarOptimizer opt(T_cl, lidars, frames, config);
  opt.image_lines_all = image_lines;
  const eigen_map<std::pair<size_t, size_t>, std::vector<CameraLidarLinePair>> line_pairs = opt.AssociateLineSingle(T_cl);
  info3[0] = 0; for (const auto& kv : line_pairs) info3[0] += (int)kv.second.size();
  CalibSnapshot S; g_calib = &S; ceres::solve_hook() = calib_hook;
  opt.Optimize(line_pairs, T_cl);
  ceres::solve_hook() = nullptr; g_calib = nullptr;
  if (S.status < 0) return S.status;
  if ((long)S.residual.size() > cap) return -1;
  for (size_t k = 0; k < S.residual.size(); ++k) { huber_a[k] = S.huber[k]; residual[k] = S.residual[k]; }
  if (!S.jac.empty()) std::memcpy(jac6, S.jac.data(), S.jac.size() * sizeof(double));
  if (S.pose.size() == 6) std::memcpy(pose6, S.pose.data(), 48);
  info3[1] = S.max_it; info3[2] = S.solver;
  return (long)S.residual.size();
}

// The OUTER loop of the calibration mode: the reference's own JointOptimize (:195-233: AssociateLineSingle, up to 35 x [Optimize(line_pairs, T_cl), rotation / translation
// change, re-association], exit when the rotation changed by < 0.1 deg AND the translation by < 0.01) with a SCRIPTED solver: the k-th ceres::Solve adds delta6[k] to the
// relative pose block (aa_cl, t_cl); after the script ends it adds nothing.  Returns the number of solver calls; T_out16 = GetResult().
namespace {
struct CalibScript { const double* delta6; int n, calls; };
CalibScript* g_cscript = nullptr;
void calib_loop_hook(const ceres::Solver::Options&, ceres::Problem* p, ceres::Solver::Summary* s) {
  CalibScript& C = *g_cscript;
  if (C.calls < C.n && !p->blocks.empty()) for (int k = 0; k < 3; ++k) { p->blocks[0].params[0][k] += C.delta6[6 * C.calls + k]; p->blocks[0].params[1][k] += C.delta6[6 * C.calls + 3 + k]; }
  ++C.calls; s->usable = true;
}
}  // namespace
int ref_calibration_loop(int rows, int cols, int n, const int* line_off, const float* lines4, void* const* lidar_frames, const double* T_cl16, int n_script, const double* delta6,
                         double* T_out16) {
  std::vector<Frame> frames; std::vector<Velodyne> lidars; std::vector<PanoramaLine> image_lines(n);
  for (int f = 0; f < n; ++f) {
    frames.push_back(Frame(rows, cols, f, "frame"));
    image_lines[f].id = f; image_lines[f].rows = rows; image_lines[f].cols = cols;
    for (int k = line_off[f]; k < line_off[f + 1]; ++k) image_lines[f].lines.push_back(cv::Vec4f(lines4[4 * k], lines4[4 * k + 1], lines4[4 * k + 2], lines4[4 * k + 3]));
    lidars.push_back(*static_cast<const Velodyne*>(lidar_frames[f]));
  }
  Config config; config.num_threads = 1;
  Eigen::Matrix4d T_cl; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_cl(i, j) = T_cl16[4 * i + j];
  CameraLidarOptimizer opt(T_cl, lidars, frames, config);
  opt.image_lines_all = image_lines;
  opt.SetOptimizationMode(CALIBRATION);
  CalibScript C{delta6, n_script, 0};
  g_cscript = &C; ceres::solve_hook() = calib_loop_hook;
  const bool ok = opt.JointOptimize(false);
  ceres::solve_hook() = nullptr; g_cscript = nullptr;
  const Eigen::Matrix4d T = opt.GetResult();
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_out16[4 * i + j] = T(i, j);
  return ok ? C.calls : -1;
}

// CameraLidarOptimizer::NeighborEachFrame (joint_optimization/CameraLidarOptimizer.cpp:551-607) and LidarMaskByTrack (:609-642), called on an optimizer object
// built from n_frames camera poses and the LiDAR frames.  CSR outputs.
int ref_neighbor_each_frame(int n_frames, const double* R_wc, const double* t_wc, const unsigned char* frame_pose_valid, int n_lidars, const double* R_wl, const double* t_wl,
                            const unsigned char* lidar_pose_valid, const unsigned char* lidar_valid, int neighbor_size, int temporal, int cap, int* off, int* ids) {
  std::vector<Frame> frames; std::vector<Velodyne> lidars(n_lidars);
  for (int f = 0; f < n_frames; ++f) {
    frames.push_back(Frame(2880, 5760, f, "frame"));
    if (frame_pose_valid[f]) { Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wc[9 * f + 3 * i + j]; frames[f].SetPose(R, Eigen::Vector3d(t_wc[3 * f], t_wc[3 * f + 1], t_wc[3 * f + 2])); }
  }
  for (int f = 0; f < n_lidars; ++f) {
    lidars[f].id = f; lidars[f].valid = lidar_valid[f] != 0;
    if (lidar_pose_valid[f]) { Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wl[9 * f + 3 * i + j]; lidars[f].SetPose(R, Eigen::Vector3d(t_wl[3 * f], t_wl[3 * f + 1], t_wl[3 * f + 2])); }
  }
  Config config; config.num_threads = 1;
  CameraLidarOptimizer opt(Eigen::Matrix4d::Identity(), lidars, frames, config);
  const std::vector<std::vector<int>> nb = opt.NeighborEachFrame(neighbor_size, temporal != 0);
  int total = 0; off[0] = 0;
  for (int f = 0; f < n_frames; ++f) { for (int v : nb[f]) { if (total >= cap) return -1; ids[total++] = v; } off[f + 1] = total; }
  return total;
}
int ref_lidar_mask_by_track(int n, void* const* lidar_frames, int min_track_length, int neighbor_size, int cap, int* off, unsigned char* mask) {
  std::vector<Velodyne> lidars; std::vector<Frame> frames;
  for (int i = 0; i < n; ++i) lidars.push_back(*static_cast<const Velodyne*>(lidar_frames[i]));
  Config config; config.num_threads = 1;
  CameraLidarOptimizer opt(Eigen::Matrix4d::Identity(), lidars, frames, config);
  const std::vector<std::vector<bool>> m = opt.LidarMaskByTrack(min_track_length, neighbor_size);
  int total = 0; off[0] = 0;
  for (int f = 0; f < n; ++f) { for (bool v : m[f]) { if (total >= cap) return -1; mask[total++] = v ? 1 : 0; } off[f + 1] = total; }
  return total;
}

// AddCameraLidarResidual (util/Optimization.cpp:564-607) for ONE (image, LiDAR) pair of frames: n line pairs (image line in pixels, LiDAR segment start / end
// in the LiDAR frame, pair weight), camera pose T_wc and LiDAR pose T_wl (rotation row-major), global weight.  The pose blocks are built as
// CameraLidarOptimizer::Optimize does (inverse pose + RotationMatrixToAngleAxis).  Two blocks per pair, in registration order: raw residual, raw 1x12
// Jacobian (camera block = reference, LiDAR block = neighbour).  poses_out: 2 x 6 = (aa_cw, t_cw), (aa_lw, t_lw).  Returns the block count or < 0.
long ref_camera_lidar_blocks(int rows, int cols, int n, const float* image_line4, const double* start3, const double* end3, const float* pair_weight, const double* R_wc,
                             const double* t_wc, const double* R_wl, const double* t_wl, double weight, long cap, double* residual, double* jac12, double* poses_out) {
  std::vector<Frame> frames; frames.push_back(Frame(rows, cols, 0, "frame"));
  std::vector<Velodyne> lidars(1); lidars[0].id = 0;
  Eigen::Matrix3d Rc, Rl;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Rc(i, j) = R_wc[3 * i + j]; Rl(i, j) = R_wl[3 * i + j]; }
  frames[0].SetPose(Rc, Eigen::Vector3d(t_wc[0], t_wc[1], t_wc[2]));
  lidars[0].SetPose(Rl, Eigen::Vector3d(t_wl[0], t_wl[1], t_wl[2]));
  eigen_vector<Eigen::Vector3d> aa_cw(1), t_cw(1), aa_lw(1), t_lw(1);
  { Eigen::Matrix4d T = frames[0].GetPose().inverse(); Eigen::Matrix3d R = T.block<3, 3>(0, 0); ceres::RotationMatrixToAngleAxis(R.data(), aa_cw[0].data()); t_cw[0] = T.block<3, 1>(0, 3); }
  { Eigen::Matrix4d T = lidars[0].GetPose().inverse(); Eigen::Matrix3d R = T.block<3, 3>(0, 0); ceres::RotationMatrixToAngleAxis(R.data(), aa_lw[0].data()); t_lw[0] = T.block<3, 1>(0, 3); }
  eigen_map<std::pair<size_t, size_t>, std::vector<CameraLidarLinePair>> line_pairs;
  std::vector<CameraLidarLinePair>& v = line_pairs[std::pair<size_t, size_t>(0, 0)];
  for (int i = 0; i < n; ++i) {
    CameraLidarLinePair lp;
    lp.image_id = 0; lp.lidar_id = 0;
    lp.image_line = cv::Vec4f(image_line4[4 * i], image_line4[4 * i + 1], image_line4[4 * i + 2], image_line4[4 * i + 3]);
    lp.lidar_line_start = Eigen::Vector3d(start3[3 * i], start3[3 * i + 1], start3[3 * i + 2]);
    lp.lidar_line_end = Eigen::Vector3d(end3[3 * i], end3[3 * i + 1], end3[3 * i + 2]);
    lp.weight = pair_weight ? pair_weight[i] : 1.f;
    v.push_back(lp);
  }
  ceres::Problem problem;
  ceres::LossFunction* loss = new ceres::HuberLoss(3.0 * M_PI / 180.0);
  AddCameraLidarResidual(frames, lidars, aa_cw, t_cw, aa_lw, t_lw, line_pairs, loss, problem, weight);
  for (int k = 0; k < 3; ++k) { poses_out[k] = aa_cw[0][k]; poses_out[3 + k] = t_cw[0][k]; poses_out[6 + k] = aa_lw[0][k]; poses_out[9 + k] = t_lw[0][k]; }
  if ((long)problem.blocks.size() > cap) return -1;
  for (size_t b = 0; b < problem.blocks.size(); ++b) {
    const ceres::Problem::Block& blk = problem.blocks[b];
    if (blk.params.size() != 4 || blk.params[0] != aa_cw[0].data() || blk.params[1] != t_cw[0].data() || blk.params[2] != aa_lw[0].data() || blk.params[3] != t_lw[0].data()) return -2;
    double jb[4][3]; double* jp[4] = {jb[0], jb[1], jb[2], jb[3]};
    if (!blk.cost->Evaluate(blk.params.data(), residual + b, jp)) return -2;
    std::memcpy(jac12 + 12 * b, jb, sizeof(jb));
  }
  delete loss;
  return (long)problem.blocks.size();
}

// AddCameraResidual (util/Optimization.cpp:172-222) with residual_type = ANGLE_RESIDUAL_1 (PanoramaReprojResidual_1Angle, the type SfMGlobalBA and the joint
// stage use): n_frames cameras (pose T_wc, rotation row-major; pose_valid = 0 leaves the constructor's "no pose" state) with their key points (CSR: kp_off,
// kp_xy float pixels), tracks (CSR: track_off, (feat_frame, feat_index)) with their 3-D points, weight.  Pose blocks as SfMGlobalBA builds them (:13-30:
// inverse pose + RotationMatrixToAngleAxis).  Per registered block, in registration order: camera, track, raw residual, raw 1x9 Jacobian (aa, t, X).
long ref_camera_residual_blocks(int rows, int cols, int n_frames, const double* R_wc, const double* t_wc, const unsigned char* pose_valid, const int* kp_off, const float* kp_xy,
                                int n_tracks, const int* track_off, const int* feat_frame, const int* feat_index, const double* points3, double weight, long cap,
                                int* cam, int* track, double* residual, double* jac9, double* cams_out) {
  std::vector<Frame> frames;
  for (int f = 0; f < n_frames; ++f) {
    frames.push_back(Frame(rows, cols, f, "frame"));
    if (pose_valid[f]) { Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wc[9 * f + 3 * i + j]; frames[f].SetPose(R, Eigen::Vector3d(t_wc[3 * f], t_wc[3 * f + 1], t_wc[3 * f + 2])); }
    for (int k = kp_off[f]; k < kp_off[f + 1]; ++k) { cv::KeyPoint kp; kp.pt = cv::Point2f(kp_xy[2 * k], kp_xy[2 * k + 1]); frames[f].keypoints_all.push_back(kp); }
  }
  eigen_vector<Eigen::Vector3d> aa(n_frames, Eigen::Vector3d::Zero()), tt(n_frames, Eigen::Vector3d::Zero());
  for (int f = 0; f < n_frames; ++f) {
    if (!frames[f].IsPoseValid()) continue;
    Eigen::Matrix4d T = frames[f].GetPose().inverse(); Eigen::Matrix3d R = T.block<3, 3>(0, 0);
    ceres::RotationMatrixToAngleAxis(R.data(), aa[f].data()); tt[f] = T.block<3, 1>(0, 3);
  }
  std::vector<PointTrack> structure;
  for (int t = 0; t < n_tracks; ++t) {
    std::set<std::pair<uint32_t, uint32_t>> fp;
    for (int k = track_off[t]; k < track_off[t + 1]; ++k) fp.insert(std::make_pair((uint32_t)feat_frame[k], (uint32_t)feat_index[k]));
    structure.push_back(PointTrack(t, fp, Eigen::Vector3d(points3[3 * t], points3[3 * t + 1], points3[3 * t + 2])));
  }
  ceres::Problem problem;
  AddCameraResidual(frames, aa, tt, structure, problem, ANGLE_RESIDUAL_1, weight);
  for (int f = 0; f < n_frames; ++f) for (int k = 0; k < 3; ++k) { cams_out[6 * f + k] = aa[f][k]; cams_out[6 * f + 3 + k] = tt[f][k]; }
  if ((long)problem.blocks.size() > cap) return -1;
  for (size_t b = 0; b < problem.blocks.size(); ++b) {
    const ceres::Problem::Block& blk = problem.blocks[b];
    if (blk.params.size() != 3) return -2;
    cam[b] = (int)((Eigen::Vector3d*)blk.params[0] - &aa[0]);
    track[b] = -1;
    for (int t = 0; t < n_tracks; ++t) if (blk.params[2] == structure[t].point_3d.data()) { track[b] = t; break; }
    double jb[3][3]; double* jp[3] = {jb[0], jb[1], jb[2]};
    if (!blk.cost->Evaluate(blk.params.data(), residual + b, jp)) return -2;
    std::memcpy(jac9 + 9 * b, jb, sizeof(jb));
  }
  return (long)problem.blocks.size();
}

// FindNeighbors over n frames given by pose (R row-major 9, t 3), pose_valid and valid flags; CSR output (off[n + 1], ids[cap]); returns total or -1
int ref_find_neighbors(int n, const double* R_wl, const double* t_wl, const unsigned char* pose_valid, const unsigned char* valid, int neighbor_size, int cap, int* off, int* ids) {
  std::vector<Velodyne> lidars(n);
  for (int f = 0; f < n; ++f) {
    lidars[f].id = f; lidars[f].valid = valid[f] != 0;
    if (pose_valid[f]) {
      Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wl[9 * f + 3 * i + j];
      lidars[f].SetPose(R, Eigen::Vector3d(t_wl[3 * f], t_wl[3 * f + 1], t_wl[3 * f + 2]));
    }
  }
  const std::vector<std::vector<int>> nb = FindNeighbors(lidars, neighbor_size);
  int total = 0; off[0] = 0;
  for (int f = 0; f < n; ++f) { for (int v : nb[f]) { if (total >= cap) return -1; ids[total++] = v; } off[f + 1] = total; }
  return total;
}
}
