// ORACLE/_ref — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE'S OWN lidar_mapping/LidarFeatureAssociate.cpp (FindNeighbors,
// AssociatePoint2Plane, AssociateLine2Line + FindAssociations + TransformLines, AssociatePoint2Line, AssociatePoint2LineSegmentKNN,
// AssociatePoint2LineSegment, AssociateLine2LineKNN), compiled from the file where it lies under /root/reference (never copied) together with the
// headers it includes (LidarFeatureAssociate.h, sensors/Velodyne.h and its sub-headers, base/Geometry.hpp ...).  PCL / Eigen / OpenCV / glog / Boost are
// absent from this image: oracle/shim/ provides stand-ins - an Eigen-like matrix, and pcl::KdTreeFLANN as an exact float32 search (shim/pvo_shim_pcl.hpp).
// sensors/Velodyne.cpp (1900 lines of feature extraction on real PCL algorithms) cannot be compiled that way; the EIGHT small member functions of
// class Velodyne that the association code calls are therefore defined below, each a restatement of the cited lines of sensors/Velodyne.cpp.
// What this pins: the control flow, thresholds, class test, vote rules, conflict resolution and output order of the reference's association functions.
// Built by `make -C oracle ref` into oracle/_ref/libpvo_ref_assoc.so; used by tests/test_reference_pinning.py and tests/make_golden.py only.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <vector>
#include <atomic>
#include <chrono>
#define private public          // the wrapper sets Velodyne::world (set by Transform2LidarWorld, sensors/Velodyne.cpp:1807) after filling world-frame clouds
#include REF_LIDAR_FEATURE_ASSOCIATE_CPP
#include REF_TRACKS_CPP                   // util/Tracks.cpp: TrackBuilder (union-find over (frame, line) features), Filter, ExportTracks
#include REF_LIDAR_LINE_MATCH_CPP         // lidar_mapping/LidarLineMatch.cpp: GenerateTracks = FindNeighbors + AssociateLine2Line(nei, i, 0.3) + TrackBuilder
#undef private

// ---- class Velodyne: the members the association code needs (sensors/Velodyne.cpp, restated) ----
Velodyne::Velodyne() : world(false), scanPeriod(0.1), valid(true), N_SCANS(0), id(-1) {                                   // :61-79
  R_wl = Eigen::Matrix3d::Zero();
  t_wl = std::numeric_limits<double>::infinity() * Eigen::Vector3d::Ones();
  T_wc_wl = Eigen::Matrix4d::Identity();
  cloudCurvature = NULL; cloudSortInd = NULL; cloudState = NULL; left_neighbor = NULL; right_neighbor = NULL;
}
Velodyne::~Velodyne() {}                                                                                                    // :81-89 (clears the clouds)
const Eigen::Vector3d Velodyne::World2Local(Eigen::Vector3d point_w) const { return R_wl.transpose() * point_w - R_wl.transpose() * t_wl; }   // :1850-1853
const Eigen::Vector3d Velodyne::Local2World(Eigen::Vector3d point_local) const { return R_wl * point_local + t_wl; }                          // :1856-1859
void Velodyne::SetPose(const Eigen::Matrix3d _R_wl, const Eigen::Vector3d _t_wl) { R_wl = _R_wl; t_wl = _t_wl; }                              // :1867-1871
const Eigen::Matrix4d Velodyne::GetPose() const {                                                                                            // :1886-1892
  Eigen::Matrix4d T_wl = Eigen::Matrix4d::Identity();
  T_wl.block<3, 3>(0, 0) = R_wl;
  T_wl.block<3, 1>(0, 3) = t_wl;
  return T_wl;
}
const bool Velodyne::IsPoseValid() const {                                                                                                   // :1894-1899
  if (!std::isinf(t_wl(0)) && !std::isnan(t_wl(0)) && !std::isinf(t_wl(1)) && !std::isnan(t_wl(1)) && !std::isinf(t_wl(2)) && !std::isnan(t_wl(2)) && !R_wl.isZero())
    return true;
  return false;
}
const bool Velodyne::IsInWorldCoordinate() const { return world; }                                                                           // :1901-1904

namespace {
void fill_cloud(pcl::PointCloud<PointType>& c, const float* xyzi, int n) {
  c.clear();
  for (int i = 0; i < n; ++i) { PointType p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; c.push_back(p); }
}
}  // namespace

extern "C" {
// A frame as the association code sees it: pose (R row-major, t; pose_valid = 0 leaves the constructor's "no pose" state), the three feature clouds
// ALREADY in the world frame (float32 x, y, z, intensity = what Transform2LidarWorld leaves), the point -> segment sets of cornerLessSharp (CSR),
// the segment coefficients in the SENSOR frame and the number of points of every segment.
void* ref_frame_create(int id, int valid, int pose_valid, const double* R_wl, const double* t_wl, const float* corner_world, int n_corner, const int* p2s_off,
                       const int* p2s_ids, int S, const double* coeffs_local, const int* seg_sizes, const float* surf_flat_world, int n_flat,
                       const float* surf_less_flat_world, int n_less) {
  Velodyne* v = new Velodyne();
  v->id = id; v->valid = valid != 0;
  if (pose_valid) {
    Eigen::Matrix3d R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_wl[3 * i + j];
    v->SetPose(R, Eigen::Vector3d(t_wl[0], t_wl[1], t_wl[2]));
  }
  fill_cloud(v->cornerLessSharp, corner_world, n_corner);
  fill_cloud(v->surfFlat, surf_flat_world, n_flat);
  fill_cloud(v->surfLessFlat, surf_less_flat_world, n_less);
  v->point_to_segment.resize(n_corner);
  if (p2s_off) for (int i = 0; i < n_corner; ++i) for (int k = p2s_off[i]; k < p2s_off[i + 1]; ++k) v->point_to_segment[i].insert(p2s_ids[k]);
  v->edge_segmented.resize(S);
  for (int s = 0; s < S; ++s) {
    Vector6d c; for (int k = 0; k < 6; ++k) c[k] = coeffs_local[6 * s + k];
    v->segment_coeffs.push_back(c);
  }
  // the points of a segment = the cornerLessSharp points whose set holds it (Velodyne::EdgeToLine fills both from the same lists); when the caller
  // gives explicit sizes they must agree
  for (int i = 0; i < n_corner; ++i) for (int s : v->point_to_segment[i]) v->edge_segmented[s].push_back(v->cornerLessSharp.points[i]);
  if (seg_sizes) for (int s = 0; s < S; ++s) if ((int)v->edge_segmented[s].size() != seg_sizes[s]) { delete v; return nullptr; }
  v->world = 1;
  return v;
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
