// panovlm_b200 — host-side builders around the kernels (C++ because the reference's host is C++).
// Mirrors the serial loops of util/Optimization.cpp:329-607 and the tails of the association functions that stay on
// the host (small integer / per-segment work): lidar_mapping/LidarFeatureAssociate.cpp:19-111, 120-197 and
// joint_optimization/CameraLidarLineAssociate.cpp:415-475, 628-715.  Only public pvb_* entry points are used for the
// device work (vote matrices, cloud transforms).
#include <limits>
#include <algorithm>
#include <omp.h>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/panovlm_b200.h"
#include "pvb_math.cuh"

using namespace pvb;

namespace {

// base/Geometry.hpp helpers in plain double (host).  The build uses -ffp-contract=off like the reference's distro libs.
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double sq(double a) { return a * a; }
inline double vector_angle(const double* a, const double* b, bool normalized = false) {   // :450-466
  double c = dot3(a, b);
  if (!normalized) c /= (std::sqrt(sq(a[0]) + sq(a[1]) + sq(a[2])) * std::sqrt(sq(b[0]) + sq(b[1]) + sq(b[2])));
  if (c >= 1.0) return 0.0;
  if (c <= -1.0) return M_PI;
  return std::acos(c);
}
inline double plane_angle(const double* a, const double* b, bool normalized = false) {   // :471-485
  double c = std::fabs(dot3(a, b));
  if (!normalized) c /= (std::sqrt(sq(a[0]) + sq(a[1]) + sq(a[2])) * std::sqrt(sq(b[0]) + sq(b[1]) + sq(b[2])));
  if (c >= 1.0) return 0.0;
  return std::acos(c);
}
inline double point_to_line(const double* p, const double* l) {   // :198-211
  const double k = (l[3] * (p[0] - l[0]) + l[4] * (p[1] - l[1]) + l[5] * (p[2] - l[2])) / (sq(l[3]) + sq(l[4]) + sq(l[5]));
  const double q[3] = {k * l[3] + l[0], k * l[4] + l[1], k * l[5] + l[2]};
  return std::sqrt(sq(q[0] - p[0]) + sq(q[1] - p[1]) + sq(q[2] - p[2]));
}
inline void project_to_plane(const double* p, const double* pl, double* out) {   // :301-316, normalized == true
  const double t = std::fabs(pl[0] * p[0] + pl[1] * p[1] + pl[2] * p[2] + pl[3]);
  for (int k = 0; k < 3; ++k) out[k] = p[k] - t * pl[k];
  if (std::fabs(pl[0] * out[0] + pl[1] * out[1] + pl[2] * out[2] + pl[3]) > 1e-4)
    for (int k = 0; k < 3; ++k) out[k] = p[k] + t * pl[k];
}
inline void plane_through_origin(const double* p1, const double* p2, double* pl) {   // FormPlane(p1, p2, 0) :328-336, then 4-vector normalize
  const double z[3] = {0, 0, 0};
  pl[0] = (p2[1] - p1[1]) * (z[2] - p1[2]) - (p2[2] - p1[2]) * (z[1] - p1[1]);
  pl[1] = (p2[2] - p1[2]) * (z[0] - p1[0]) - (p2[0] - p1[0]) * (z[2] - p1[2]);
  pl[2] = (p2[0] - p1[0]) * (z[1] - p1[1]) - (p2[1] - p1[1]) * (z[0] - p1[0]);
  pl[3] = -(pl[0] * p1[0] + pl[1] * p1[1] + pl[2] * p1[2]);
}
inline void normalize4(double* pl) { const double n = std::sqrt(sq(pl[0]) + sq(pl[1]) + sq(pl[2]) + sq(pl[3])); for (int k = 0; k < 4; ++k) pl[k] /= n; }
inline void transform4(const double* T, const double* p, double* out) {   // (T * p.homogeneous()).hnormalized() for a rigid T
  for (int r = 0; r < 3; ++r) out[r] = T[r * 4] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3];
}
inline void transform_line(const double* R, const double* t, const double* in, double* out) {   // TransformLines :219-236
  for (int r = 0; r < 3; ++r) {
    out[r] = R[r * 3] * in[0] + R[r * 3 + 1] * in[1] + R[r * 3 + 2] * in[2] + t[r];
    out[3 + r] = R[r * 3] * in[3] + R[r * 3 + 1] * in[4] + R[r * 3 + 2] * in[5];
  }
}

// float image -> camera ray of radius r (Equirectangular.h:98-146 with T = float)
inline void image_to_cam_f32(float px, float py, int rows, int cols, float r, float* cam) {
  const float lon = (2 * px / cols - 1) * M_PI;
  const float lat = (0.5 - py / rows) * M_PI;
  const float cy = std::cos(lat);
  cam[0] = r * cy * std::sin(lon);
  cam[1] = -r * std::sin(lat);
  cam[2] = r * cy * std::cos(lon);
}

// Equirectangular::BreakToSegments (sensors/Equirectangular.cpp:20-58)
std::vector<std::pair<float, float>> break_to_segments(int rows, int cols, const float* start, const float* end, float seg_length) {
  float p1[3], p2[3];
  image_to_cam_f32(start[0], start[1], rows, cols, 5.0f, p1);
  image_to_cam_f32(end[0], end[1], rows, cols, 5.0f, p2);
  const float sl[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const float length = std::sqrt((start[0] - end[0]) * (start[0] - end[0]) + (start[1] - end[1]) * (start[1] - end[1]));
  const int count = length / seg_length + 1;
  std::vector<std::pair<float, float>> seg = {{start[0], start[1]}};
  for (int i = 1; i < count; ++i) {
    const float f = i * 1.f / count;
    float u, v;
    cam_to_image_f32(p1[0] + f * sl[0], p1[1] + f * sl[1], p1[2] + f * sl[2], rows, cols, u, v);
    if (std::abs(u - seg.back().first) > 0.8 * cols) {
      const float g = p1[0] / (p1[0] - p2[0]);
      float lu, lv;
      cam_to_image_f32(p1[0] + g * sl[0], p1[1] + g * sl[1], p1[2] + g * sl[2], rows, cols, lu, lv);
      const std::pair<float, float> L{0.f, lv}, R{float(cols - 1), lv};
      if (u > seg.back().first) { seg.push_back(L); seg.push_back(R); }
      else { seg.push_back(R); seg.push_back(L); }
    }
    seg.push_back({u, v});
  }
  seg.push_back({end[0], end[1]});
  return seg;
}

struct BlockOut { long at, cap; int* type; int* ref; int* nei; int* normalize; double* huber; double* consts; };
inline bool push_block(BlockOut& o, int type, int ref, int nei, int normalize, double huber, const double c[12]) {
  if (o.at >= o.cap) return false;
  o.type[o.at] = type; o.ref[o.at] = ref; o.nei[o.at] = nei; o.normalize[o.at] = normalize; o.huber[o.at] = huber;
  std::memcpy(o.consts + 12 * o.at, c, 96);
  ++o.at;
  return true;
}

// FindAssociations (:120-197) on a vote matrix; shared by AssociateLine2Line and AssociateLine2LineKNN.
int find_associations(const pvb_line_frame* ref, const pvb_line_frame* nei, const int* M, int* n_out, int* nei_line, int* ref_line,
                             double* point_a3, double* point_b3) {
  const int Sr = ref->n_segments, Sn = nei->n_segments;
  std::vector<double> ref_w((size_t)Sr * 6), nei_w((size_t)Sn * 6);
  for (int s = 0; s < Sr; ++s) transform_line(ref->R_wl, ref->t_wl, ref->segment_coeffs + 6 * s, &ref_w[6 * s]);
  for (int s = 0; s < Sn; ++s) transform_line(nei->R_wl, nei->t_wl, nei->segment_coeffs + 6 * s, &nei_w[6 * s]);
  std::vector<int> seg_size(Sn, 0);                                            // edge_segmented[s].size()
  for (int i = 0; i < nei->n_corner; ++i) for (int e = nei->p2s_off[i]; e < nei->p2s_off[i + 1]; ++e) seg_size[nei->p2s_ids[e]]++;
  struct A { int nl, rl; double a[3], b[3]; };
  std::map<int, A> assoc;                                                      // eigen_map<int, Line2Line> keyed by the reference line
  for (int s = 0; s < Sn; ++s) {
    int max_col = 0, max_count = M[(size_t)s * Sr];
    for (int c = 1; c < Sr; ++c) if (M[(size_t)s * Sr + c] > max_count) { max_count = M[(size_t)s * Sr + c]; max_col = c; }
    if ((size_t)max_count < (size_t)seg_size[s] / 2) continue;                 // :132
    if (plane_angle(&ref_w[6 * max_col + 3], &nei_w[6 * s + 3]) * 180.0 / M_PI > 7) continue;   // :138
    const double* cl = ref->segment_coeffs + 6 * max_col;
    A a; a.nl = s; a.rl = max_col;
    for (int k = 0; k < 3; ++k) { a.a[k] = 0.1 * cl[3 + k] + cl[k]; a.b[k] = -0.1 * cl[3 + k] + cl[k]; }   // :144-145
    auto it = assoc.find(max_col);
    if (it == assoc.end()) assoc.insert({max_col, a});
    else {
      const double d1 = point_to_line(&nei_w[6 * it->second.nl], &ref_w[6 * max_col]);
      const double d2 = point_to_line(&nei_w[6 * s], &ref_w[6 * max_col]);
      if (d2 < d1) it->second = a;
    }
  }
  int n = 0;
  for (auto& kv : assoc) {
    nei_line[n] = kv.second.nl; ref_line[n] = kv.second.rl;
    std::memcpy(point_a3 + 3 * n, kv.second.a, 24); std::memcpy(point_b3 + 3 * n, kv.second.b, 24);
    ++n;
  }
  *n_out = n;
  return PVB_OK;
}

// std::map<size_t,size_t> seg_count of one query (:264-270, :419-425): ascending segment id, count of its 5 neighbours on it
void segment_counts(const pvb_line_frame* ref, const int* nn5, std::map<int, int>& cnt) {
  cnt.clear();
  for (int j = 0; j < 5; ++j) for (int e = ref->p2s_off[nn5[j]]; e < ref->p2s_off[nn5[j] + 1]; ++e) cnt[ref->p2s_ids[e]]++;
}

void push_segment_assoc(const pvb_line_frame* ref, const pvb_line_frame* nei, int q, int seg, long at, int* query, int* ref_line, double* point3, double* a3, double* b3) {
  const double* cl = ref->segment_coeffs + 6 * seg;
  for (int k = 0; k < 3; ++k) { a3[3 * at + k] = 0.1 * cl[3 + k] + cl[k]; b3[3 * at + k] = -0.1 * cl[3 + k] + cl[k]; }   // :283-284, :349-350
  float wx, wy, wz;                                                            // the neighbour's world cloud is float32 (Transform2LidarWorld)
  const float* pl = nei->corner_local + 4 * (size_t)q;
  transform_point_f32(nei->R_wl, nei->t_wl, pl[0], pl[1], pl[2], wx, wy, wz);
  const double pw[3] = {(double)wx, (double)wy, (double)wz};
  world2local(nei->R_wl, nei->t_wl, pw, point3 + 3 * at);                      // :286, :351
  query[at] = q; ref_line[at] = seg;
}

}  // namespace

extern "C" {

static inline uint32_t bits_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float float_of(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int pvb_find_neighbors(int n, const double* t_wl, const unsigned char* pose_valid, const unsigned char* frame_valid, int neighbor_size,
                       int* out_offsets, int* out_neighbors, int cap) {
  if (n <= 0 || !t_wl || !out_offsets || !out_neighbors || neighbor_size < 1) return PVB_ERR_ARG;
  auto pv = [&](int i) { return !pose_valid || pose_valid[i]; };
  auto fv = [&](int i) { return !frame_valid || frame_valid[i]; };
  std::vector<int> centre_frame;             // lidar_center cloud: frames with a valid pose AND valid data (:25-38)
  std::vector<float> centre;
  for (int i = 0; i < n; ++i)
    if (pv(i) && fv(i)) { centre_frame.push_back(i); for (int k = 0; k < 3; ++k) centre.push_back((float)t_wl[3 * i + k]); }
  const int m = (int)centre_frame.size();
  // every frame's list is independent (a sort of all centre distances per frame): frames in parallel on the host cores, lists concatenated in frame order
  std::vector<std::vector<int>> lists(n);
  // (an explicit thread count: launchers such as torchrun export OMP_NUM_THREADS=1, which would make this loop serial on every rank)
  const int n_threads = std::max(1, std::min(8, omp_get_num_procs()));
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
  for (int i = 0; i < n; ++i) {
    std::vector<int>& neighbors = lists[i];
    if (pv(i)) {
      const float q[3] = {(float)t_wl[3 * i], (float)t_wl[3 * i + 1], (float)t_wl[3 * i + 2]};
      // (squared distance, index) packed into one 64-bit key: non-negative floats order like their bit patterns, so sorting the keys is the
      // lexicographic sort of the pairs at half the cost
      std::vector<unsigned long long> keys(m);
      for (int j = 0; j < m; ++j) keys[j] = ((unsigned long long)bits_of(sqdist_f32(q[0], q[1], q[2], centre[3 * j], centre[3 * j + 1], centre[3 * j + 2])) << 32) | (unsigned)j;
      std::sort(keys.begin(), keys.end());
      struct DI { float first; int second; };
      std::vector<DI> d(m);
      for (int j = 0; j < m; ++j) { d[j].first = float_of((uint32_t)(keys[j] >> 32)); d[j].second = (int)(keys[j] & 0xFFFFFFFFu); }
      const int k = std::min(neighbor_size, m);
      for (int j = 0; j < k; ++j) neighbors.push_back(d[j].second);
      if (!neighbors.empty()) neighbors.erase(neighbors.begin());              // the first is the frame itself (:52)
      for (int& v : neighbors) v = centre_frame[v];
      std::set<int> nset(neighbors.begin(), neighbors.end());
      int p = i - 1;
      while (p >= 0 && !pv(p)) --p;
      if (p >= 0 && nset.count(p) == 0) neighbors.push_back(p);
      p = i + 1;
      while (p < n && !pv(p)) ++p;
      if (p < n && nset.count(p) == 0) neighbors.push_back(p);
      const float r2 = 20.0f * 20.0f;                                          // loop candidates (:79-98), sorted by distance
      for (int j = 0; j < m && d[j].first < r2; ++j) {
        const int cand = centre_frame[d[j].second];
        int same_loop = 0;
        for (int it : nset) { if (std::abs(cand - it) <= 200) ++same_loop; if (same_loop >= 2) break; }
        if (same_loop < 2 && nset.count(cand) == 0) { neighbors.push_back(cand); nset.insert(cand); }
      }
    } else {
      for (int j = -neighbor_size / 2; j <= neighbor_size / 2; ++j) neighbors.push_back(i - j);   // :103-106 (may be out of range; callers filter)
    }
  }
  int total = 0;
  out_offsets[0] = 0;
  for (int i = 0; i < n; ++i) {
    for (int v : lists[i]) { if (total >= cap) return PVB_ERR_ARG; out_neighbors[total++] = v; }
    out_offsets[i + 1] = total;
  }
  return total;
}

// CameraLidarOptimizer::NeighborEachFrame (joint_optimization/CameraLidarOptimizer.cpp:551-610): the LiDAR frames every image is associated with.
// temporal: a window of neighbor_size LiDAR indices around the image index, shifted to stay inside [0, n_lidars) (:556-567);
// otherwise the neighbor_size LiDARs nearest to the camera centre (float32 k-NN over the LiDARs with a valid pose and valid data) plus the previous /
// next LiDAR index when they are not among them (:591-596); images without a valid pose get no neighbour.
int pvb_neighbor_each_frame(int n_frames, int n_lidars, int neighbor_size, int temporal, const double* t_wc, const unsigned char* frame_pose_valid, const double* t_wl,
                            const unsigned char* lidar_pose_valid, const unsigned char* lidar_valid, int* out_offsets, int* out_neighbors, int cap) {
  if (n_frames < 0 || n_lidars < 0 || neighbor_size < 0 || !out_offsets || (cap > 0 && !out_neighbors) || (!temporal && n_frames > 0 && (!t_wc || (n_lidars > 0 && !t_wl)))) return PVB_ERR_ARG;
  int total = 0;
  out_offsets[0] = 0;
  std::vector<int> centre_lidar; std::vector<float> centre;
  if (!temporal)
    for (int i = 0; i < n_lidars; ++i)
      if ((!lidar_pose_valid || lidar_pose_valid[i]) && (!lidar_valid || lidar_valid[i])) { centre_lidar.push_back(i); for (int k = 0; k < 3; ++k) centre.push_back((float)t_wl[3 * i + k]); }
  const int m = (int)centre_lidar.size();
  for (int f = 0; f < n_frames; ++f) {
    std::vector<int> nb;
    if (temporal) {
      int lo = std::max(0, f - neighbor_size / 2);
      const int hi = std::min(n_lidars, lo + neighbor_size);
      lo = std::max(0, hi - neighbor_size);
      for (int l = lo; l < hi; ++l) nb.push_back(l);
    } else if (!frame_pose_valid || frame_pose_valid[f]) {
      const float q[3] = {(float)t_wc[3 * f], (float)t_wc[3 * f + 1], (float)t_wc[3 * f + 2]};
      std::vector<std::pair<float, int>> d(m);
      for (int j = 0; j < m; ++j) d[j] = {sqdist_f32(q[0], q[1], q[2], centre[3 * j], centre[3 * j + 1], centre[3 * j + 2]), j};
      std::sort(d.begin(), d.end());
      for (int j = 0; j < std::min(neighbor_size, m); ++j) nb.push_back(centre_lidar[d[j].second]);
      const std::set<int> have(nb.begin(), nb.end());
      if (have.count(f - 1) == 0 && f - 1 >= 0) nb.push_back(f - 1);
      if (have.count(f + 1) == 0 && f + 1 < n_lidars) nb.push_back(f + 1);
    }
    for (int v : nb) { if (total >= cap) return PVB_ERR_NOMEM; out_neighbors[total++] = v; }
    out_offsets[f + 1] = total;
  }
  return total;
}

// CameraLidarOptimizer::LidarMaskByTrack (:612-642), the part after GenerateTracks: a LiDAR line takes part in the camera-LiDAR association only when it
// belongs to a line track.  seg_off: n_lidars + 1 offsets of every frame's segments in `mask`.
int pvb_lidar_mask_by_track(int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int n_lidars, const int* seg_off, unsigned char* mask) {
  if (n_tracks < 0 || n_lidars < 0 || !seg_off || (n_tracks > 0 && (!track_off || !feat_frame || !feat_line)) || (n_lidars > 0 && seg_off[n_lidars] > 0 && !mask)) return PVB_ERR_ARG;
  std::fill(mask, mask + seg_off[n_lidars], (unsigned char)0);
  for (int t = 0; t < n_tracks; ++t)
    for (int k = track_off[t]; k < track_off[t + 1]; ++k) {
      const int fr = feat_frame[k], ln = feat_line[k];
      if (fr < 0 || fr >= n_lidars || ln < 0 || seg_off[fr] + ln >= seg_off[fr + 1]) return PVB_ERR_ARG;
      mask[seg_off[fr] + ln] = 1;
    }
  return PVB_OK;
}

int pvb_line2line_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, double dist_threshold, int* n_out, int* nei_line, int* ref_line,
                            double* point_a3, double* point_b3) {
  if (!ctx || !ref || !nei || !n_out) return PVB_ERR_ARG;
  *n_out = 0;
  const int Sr = ref->n_segments, Sn = nei->n_segments;
  if (Sr == 0 || Sn == 0) return PVB_OK;                                       // CheckLidarSegment (:208-216)
  std::vector<double> ref_w((size_t)Sr * 6);
  for (int s = 0; s < Sr; ++s) transform_line(ref->R_wl, ref->t_wl, ref->segment_coeffs + 6 * s, &ref_w[6 * s]);
  std::vector<float> world((size_t)std::max(1, nei->n_corner) * 4);
  int rc = pvb_transform_cloud(ctx, nei->corner_local, nei->n_corner, nei->R_wl, nei->t_wl, world.data());
  if (rc) return rc;
  std::vector<int> M((size_t)Sn * Sr, 0);
  rc = pvb_line_votes(ctx, ref_w.data(), Sr, world.data(), nei->n_corner, nei->p2s_off, nei->p2s_ids, Sn, dist_threshold, M.data());
  if (rc) return rc;
  return find_associations(ref, nei, M.data(), n_out, nei_line, ref_line, point_a3, point_b3);
}

// ---- segment-based variants (LidarFeatureAssociate.cpp:238-440): 5-NN / nearest line on the device, membership counting here ---------
int pvb_point2line_segment_knn_tail(const pvb_line_frame* ref, const pvb_line_frame* nei, const int* idx5, long cap, long* n_out, int* query, int* ref_line,
                                    double* point3, double* a3, double* b3) {
  if (!ref || !nei || !n_out || (nei->n_corner > 0 && !idx5)) return PVB_ERR_ARG;
  *n_out = 0;
  if (ref->n_segments == 0 || nei->n_segments == 0) return PVB_OK;             // CheckLidarSegment (:243-247)
  std::map<int, int> cnt;
  long n = 0;
  for (int q = 0; q < nei->n_corner; ++q) {
    if (idx5[5 * (size_t)q] < 0) continue;                                     // :261
    segment_counts(ref, idx5 + 5 * (size_t)q, cnt);
    for (auto& kv : cnt) {
      if (kv.second < 5) continue;                                             // :276 all five neighbours on the segment
      if (n >= cap) return PVB_ERR_NOMEM;
      push_segment_assoc(ref, nei, q, kv.first, n, query, ref_line, point3, a3, b3);
      ++n;
    }
  }
  *n_out = n;
  return PVB_OK;
}

int pvb_line2line_knn_tail(const pvb_line_frame* ref, const pvb_line_frame* nei, const int* idx5, int* n_out, int* nei_line, int* ref_line, double* point_a3,
                           double* point_b3) {
  if (!ref || !nei || !n_out || (nei->n_corner > 0 && !idx5)) return PVB_ERR_ARG;
  *n_out = 0;
  const int Sr = ref->n_segments, Sn = nei->n_segments;
  if (Sr == 0 || Sn == 0) return PVB_OK;
  std::vector<int> M((size_t)Sn * Sr, 0);
  std::map<int, int> cnt;
  for (int q = 0; q < nei->n_corner; ++q) {
    if (idx5[5 * (size_t)q] < 0) continue;                                     // :416
    segment_counts(ref, idx5 + 5 * (size_t)q, cnt);
    for (auto& kv : cnt) {
      if (kv.second < 3) continue;                                             // :428 k_search_size - 2
      for (int e = nei->p2s_off[q]; e < nei->p2s_off[q + 1]; ++e) M[(size_t)nei->p2s_ids[e] * Sr + kv.first] += 1;   // :432-433
    }
  }
  return find_associations(ref, nei, M.data(), n_out, nei_line, ref_line, point_a3, point_b3);
}

int pvb_point2line_segment_knn_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, float dist_threshold, long cap, long* n_out, int* query,
                                         int* ref_line, double* point3, double* a3, double* b3) {
  if (!ctx || !ref || !nei || !n_out) return PVB_ERR_ARG;
  *n_out = 0;
  if (ref->n_segments == 0 || nei->n_segments == 0 || nei->n_corner == 0) return PVB_OK;
  std::vector<int> idx((size_t)nei->n_corner * 5);
  int rc = pvb_pair_knn5(ctx, ref->corner_local, ref->n_corner, ref->R_wl, ref->t_wl, nei->corner_local, nei->n_corner, nei->R_wl, nei->t_wl, dist_threshold, 0.0, idx.data());
  if (rc) return rc;
  return pvb_point2line_segment_knn_tail(ref, nei, idx.data(), cap, n_out, query, ref_line, point3, a3, b3);
}

int pvb_line2line_knn_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, float dist_threshold, int* n_out, int* nei_line, int* ref_line,
                                double* point_a3, double* point_b3) {
  if (!ctx || !ref || !nei || !n_out) return PVB_ERR_ARG;
  *n_out = 0;
  if (ref->n_segments == 0 || nei->n_segments == 0) return PVB_OK;
  std::vector<int> idx((size_t)std::max(1, nei->n_corner) * 5);
  int rc = pvb_pair_knn5(ctx, ref->corner_local, ref->n_corner, ref->R_wl, ref->t_wl, nei->corner_local, nei->n_corner, nei->R_wl, nei->t_wl, dist_threshold, 0.0, idx.data());
  if (rc) return rc;
  return pvb_line2line_knn_tail(ref, nei, idx.data(), n_out, nei_line, ref_line, point_a3, point_b3);
}

int pvb_point2line_segment_associate(pvb_ctx* ctx, const pvb_line_frame* ref, const pvb_line_frame* nei, float dist_threshold, long cap, long* n_out, int* query,
                                     int* ref_line, double* point3, double* a3, double* b3) {
  if (!ctx || !ref || !nei || !n_out) return PVB_ERR_ARG;
  *n_out = 0;
  const int Sr = ref->n_segments;
  if (Sr == 0 || nei->n_segments == 0 || nei->n_corner == 0) return PVB_OK;
  std::vector<double> ref_w((size_t)Sr * 6);
  for (int s = 0; s < Sr; ++s) transform_line(ref->R_wl, ref->t_wl, ref->segment_coeffs + 6 * s, &ref_w[6 * s]);
  std::vector<float> world((size_t)nei->n_corner * 4);
  int rc = pvb_transform_cloud(ctx, nei->corner_local, nei->n_corner, nei->R_wl, nei->t_wl, world.data());
  if (rc) return rc;
  std::vector<int> line(nei->n_corner); std::vector<double> dist(nei->n_corner);
  rc = pvb_nearest_line(ctx, ref_w.data(), Sr, world.data(), nei->n_corner, line.data(), dist.data());
  if (rc) return rc;
  long n = 0;
  for (int q = 0; q < nei->n_corner; ++q) {
    if (!(dist[q] <= (double)dist_threshold)) continue;                        // :343
    if (n >= cap) return PVB_ERR_NOMEM;
    push_segment_assoc(ref, nei, q, line[q], n, query, ref_line, point3, a3, b3);
    ++n;
  }
  *n_out = n;
  return PVB_OK;
}

// ---- LiDAR line tracks (lidar_mapping/LidarLineMatch.cpp:36-86, util/Tracks.cpp:58-186) and the track gate of
//      AddLidarLineToLineResidual2 (util/Optimization.cpp:343-400) ------------------------------------------------------------------
// A track is a connected component of the line-match graph whose lines come from >= min_track_length distinct frames (and, unless
// allow_multiple_map, from pairwise distinct frames).  Tracks are numbered by their smallest (frame, line) feature and list their
// features in ascending order — the order TrackBuilder::ExportTracks produces from its sorted feature index.
int pvb_line_tracks_build(int n_pairs, const int* pair_a, const int* pair_b, const int* match_off, const int* match_a, const int* match_b, int min_track_length,
                          int allow_multiple_map, int cap_features, int* n_tracks, int* track_off, int* feat_frame, int* feat_line) {
  if (n_pairs < 0 || !n_tracks || (n_pairs > 0 && (!pair_a || !pair_b || !match_off))) return PVB_ERR_ARG;
  *n_tracks = 0;
  if (track_off) track_off[0] = 0;
  const int n_matches = n_pairs ? match_off[n_pairs] : 0;
  if (n_matches == 0) return PVB_OK;
  if (!match_a || !match_b || !track_off || !feat_frame || !feat_line) return PVB_ERR_ARG;
  typedef std::pair<int, int> F;
  std::vector<F> feats; feats.reserve(2 * (size_t)n_matches);
  for (int p = 0; p < n_pairs; ++p)
    for (int e = match_off[p]; e < match_off[p + 1]; ++e) { feats.push_back(F(pair_a[p], match_a[e])); feats.push_back(F(pair_b[p], match_b[e])); }
  std::sort(feats.begin(), feats.end());
  feats.erase(std::unique(feats.begin(), feats.end()), feats.end());
  const int nf = (int)feats.size();
  auto id_of = [&](int frame, int line) { return (int)(std::lower_bound(feats.begin(), feats.end(), F(frame, line)) - feats.begin()); };
  std::vector<int> parent(nf);
  for (int i = 0; i < nf; ++i) parent[i] = i;
  auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
  for (int p = 0; p < n_pairs; ++p)
    for (int e = match_off[p]; e < match_off[p + 1]; ++e) {
      const int a = find(id_of(pair_a[p], match_a[e])), b = find(id_of(pair_b[p], match_b[e]));
      if (a != b) parent[std::max(a, b)] = std::min(a, b);          // the root is the smallest feature of the component
    }
  // per component: number of distinct frames and whether a frame repeats (features are sorted by frame within a component)
  std::vector<int> n_frames(nf, 0), last_frame(nf, -1), n_feat(nf, 0); std::vector<char> repeat(nf, 0);
  for (int i = 0; i < nf; ++i) {
    const int r = find(i);
    n_feat[r]++;
    if (last_frame[r] == feats[i].first && n_feat[r] > 1) repeat[r] = 1; else { n_frames[r]++; last_frame[r] = feats[i].first; }
  }
  std::vector<int> track_of(nf, -1);
  int nt = 0, total = 0;
  for (int i = 0; i < nf; ++i) {
    if (parent[i] != i) continue;
    if (n_frames[i] < min_track_length || (!allow_multiple_map && repeat[i]) || n_feat[i] < 2) continue;
    track_of[i] = nt++; total += n_feat[i];
  }
  if (total > cap_features) return PVB_ERR_NOMEM;
  std::vector<int> fill(nt + 1, 0);
  for (int i = 0; i < nf; ++i) if (parent[i] == i && track_of[i] >= 0) fill[track_of[i] + 1] = n_feat[i];
  for (int t = 0; t < nt; ++t) fill[t + 1] += fill[t];
  for (int t = 0; t <= nt; ++t) track_off[t] = fill[t];
  for (int i = 0; i < nf; ++i) {
    const int t = track_of[find(i)];
    if (t < 0) continue;
    feat_frame[fill[t]] = feats[i].first; feat_line[fill[t]] = feats[i].second; fill[t]++;
  }
  *n_tracks = nt;
  return PVB_OK;
}

// keep[i] = 1 iff the reference line (ref_frame, ref_line[i]) and the neighbour line (nei_frame, nei_line[i]) are in one track
int pvb_line_tracks_gate(int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int ref_frame, int nei_frame, int n, const int* ref_line,
                         const int* nei_line, unsigned char* keep) {
  if (n < 0 || n_tracks < 0 || (n > 0 && (!ref_line || !nei_line || !keep)) || (n_tracks > 0 && (!track_off || !feat_frame || !feat_line))) return PVB_ERR_ARG;
  std::map<std::pair<int, int>, int> track_of;                                 // a line belongs to at most one track (components are disjoint)
  for (int t = 0; t < n_tracks; ++t) for (int e = track_off[t]; e < track_off[t + 1]; ++e) track_of[{feat_frame[e], feat_line[e]}] = t;
  for (int i = 0; i < n; ++i) {
    auto a = track_of.find({ref_frame, ref_line[i]});
    auto b = track_of.find({nei_frame, nei_line[i]});
    keep[i] = (a != track_of.end() && b != track_of.end() && a->second == b->second) ? 1 : 0;
  }
  return PVB_OK;
}

// world-frame segment lines of every frame (TransformLines, LidarFeatureAssociate.cpp:219-236), concatenated in frame order
static void all_lines_world(int n_frames, const pvb_line_frame* frames, std::vector<double>& lines, std::vector<int>& seg_off) {
  seg_off.assign(n_frames + 1, 0);
  for (int f = 0; f < n_frames; ++f) seg_off[f + 1] = seg_off[f] + frames[f].n_segments;
  lines.assign((size_t)std::max(1, seg_off[n_frames]) * 6, 0.0);
  for (int f = 0; f < n_frames; ++f)
    for (int s = 0; s < frames[f].n_segments; ++s) transform_line(frames[f].R_wl, frames[f].t_wl, frames[f].segment_coeffs + 6 * s, &lines[(size_t)(seg_off[f] + s) * 6]);
}

// LidarLineMatch::GenerateTracks (:36-86): AssociateLine2Line(lidars[nei], lidars[i], 0.3) over the frame graph, then the track builder.  All vote matrices
// come from ONE batched device pass (pvb_line_votes_batch); the FindAssociations tails run on the host cores, one pair per task.
int pvb_generate_line_tracks(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, const unsigned char* pose_valid, const int* nbr_off, const int* nbr_ids,
                             double dist_threshold, int min_track_length, int cap_features, int* n_tracks, int* track_off, int* feat_frame, int* feat_line) {
  if (!ctx || n_frames < 0 || !n_tracks || (n_frames > 0 && (!frames || !nbr_off))) return PVB_ERR_ARG;
  std::vector<int> pa, pb;                                                     // pair p: frame i = pa[p] (neighbour role), its graph neighbour n = pb[p] (reference role), :66
  for (int i = 0; i < n_frames; ++i) {
    if (pose_valid && !pose_valid[i]) continue;                                // :62
    for (int e = nbr_off[i]; e < nbr_off[i + 1]; ++e) {
      const int n = nbr_ids[e];
      if (n < 0 || n >= n_frames) return PVB_ERR_ARG;
      pa.push_back(i); pb.push_back(n);
    }
  }
  const int np = (int)pa.size();
  std::vector<long long> m_off(np + 1, 0);
  for (int p = 0; p < np; ++p) m_off[p + 1] = m_off[p] + (long long)frames[pa[p]].n_segments * frames[pb[p]].n_segments;
  std::vector<int> M((size_t)std::max<long long>(1, m_off[np]));
  std::vector<double> lines; std::vector<int> seg_off;
  all_lines_world(n_frames, frames, lines, seg_off);
  int rc = pvb_line_votes_batch(ctx, n_frames, frames, lines.data(), np, pb.data(), pa.data(), m_off.data(), m_off[np], dist_threshold, M.data(), nullptr);
  if (rc) return rc;
  std::vector<std::vector<std::pair<int, int>>> matches(np);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int p = 0; p < np; ++p) {
    const pvb_line_frame* rf = &frames[pb[p]]; const pvb_line_frame* nf = &frames[pa[p]];
    if (rf->n_segments == 0 || nf->n_segments == 0) continue;
    const int S = nf->n_segments;
    std::vector<int> nl(S), rl(S); std::vector<double> a3(3 * (size_t)S), b3(3 * (size_t)S);
    int m = 0;
    if (find_associations(rf, nf, M.data() + m_off[p], &m, nl.data(), rl.data(), a3.data(), b3.data()) != PVB_OK) {
#pragma omp atomic write
      bad = 1;
      continue;
    }
    std::set<std::pair<int, int>> uniq;                                        // set<pair<neighbor_line_idx, ref_line_idx>> (:67-69)
    for (int k = 0; k < m; ++k) uniq.insert({nl[k], rl[k]});
    matches[p].assign(uniq.begin(), uniq.end());
  }
  if (bad) return PVB_ERR_ARG;
  std::vector<int> off(1, 0), ma, mb;
  for (int p = 0; p < np; ++p) { for (auto& f : matches[p]) { ma.push_back(f.first); mb.push_back(f.second); } off.push_back((int)ma.size()); }
  return pvb_line_tracks_build(np, pa.data(), pb.data(), off.data(), ma.data(), mb.data(), min_track_length, 1, cap_features, n_tracks, track_off, feat_frame, feat_line);
}

// AddLidarLineToLineResidual2 (util/Optimization.cpp:329-441) for a whole pose graph in one call: AssociateLine2Line of every edge (ref[e], nei[e]) from one batched
// device pass (world clouds + vote matrices), FindAssociations tails, the line-track gate (:383-400; n_tracks < 0: no gate) and one Point2Line block per point of every
// kept neighbour segment (:402-435), appended edge by edge in the reference's order.  Returns the new block count or a negative error.
int pvb_frames_line2line_blocks(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, int n_edges, const int* ref, const int* nei, double dist_threshold,
                                int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int angle_residual, int normalize_distance, double weight,
                                long at, long cap, int* type, int* ref_out, int* nei_out, int* normalize, double* huber, double* consts) {
  if (!ctx || n_frames < 0 || n_edges < 0 || (n_frames > 0 && !frames) || (n_edges > 0 && (!ref || !nei)) || !type) return PVB_ERR_ARG;
  if (n_tracks > 0 && (!track_off || !feat_frame || !feat_line)) return PVB_ERR_ARG;
  std::vector<long long> m_off(n_edges + 1, 0);
  for (int e = 0; e < n_edges; ++e) {
    if (ref[e] < 0 || ref[e] >= n_frames || nei[e] < 0 || nei[e] >= n_frames) return PVB_ERR_ARG;
    m_off[e + 1] = m_off[e] + (long long)frames[ref[e]].n_segments * frames[nei[e]].n_segments;
  }
  std::vector<int> M((size_t)std::max<long long>(1, m_off[n_edges]));
  std::vector<double> lines; std::vector<int> seg_off;
  all_lines_world(n_frames, frames, lines, seg_off);
  std::vector<int> coff(n_frames + 1, 0);
  for (int f = 0; f < n_frames; ++f) coff[f + 1] = coff[f] + frames[f].n_corner;
  std::vector<float> world((size_t)std::max(1, coff[n_frames]) * 4);
  int rc = pvb_line_votes_batch(ctx, n_frames, frames, lines.data(), n_edges, ref, nei, m_off.data(), m_off[n_edges], dist_threshold, M.data(), world.data());
  if (rc) return rc;
  std::vector<int> track_of;                                                   // (frame, line) -> track id, -1: none (a line belongs to at most one track)
  if (n_tracks >= 0) {
    track_of.assign((size_t)std::max(1, seg_off[n_frames]), -1);
    for (int t = 0; t < n_tracks; ++t)
      for (int k = track_off[t]; k < track_off[t + 1]; ++k) {
        const int fr = feat_frame[k], ln = feat_line[k];
        if (fr < 0 || fr >= n_frames || ln < 0 || ln >= frames[fr].n_segments) return PVB_ERR_ARG;
        track_of[(size_t)seg_off[fr] + ln] = t;
      }
  }
  struct Blk { double c[12]; };
  std::vector<std::vector<Blk>> per_edge(n_edges);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (int e = 0; e < n_edges; ++e) {
    const pvb_line_frame* rf = &frames[ref[e]]; const pvb_line_frame* nf = &frames[nei[e]];
    if (rf->n_segments == 0 || nf->n_segments == 0) continue;                  // CheckLidarSegment (:208-216)
    const int S = nf->n_segments;
    std::vector<int> nl(S), rl(S); std::vector<double> a3(3 * (size_t)S), b3(3 * (size_t)S);
    int m = 0;
    if (find_associations(rf, nf, M.data() + m_off[e], &m, nl.data(), rl.data(), a3.data(), b3.data()) != PVB_OK) {
#pragma omp atomic write
      bad = 1;
      continue;
    }
    const float* w = world.data() + (size_t)coff[nei[e]] * 4;
    for (int k = 0; k < m; ++k) {
      if (n_tracks >= 0) {                                                     // Optimization.cpp:383-400: both lines in one track
        const int ta = track_of[(size_t)seg_off[ref[e]] + rl[k]], tb = track_of[(size_t)seg_off[nei[e]] + nl[k]];
        if (ta < 0 || ta != tb) continue;
      }
      const double* a = &a3[3 * (size_t)k]; const double* b = &b3[3 * (size_t)k];
      double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};                   // Point2Line_*: line_direction = (a - b).normalized()
      const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int q = 0; q < 3; ++q) d[q] /= nn;
      for (int i = 0; i < nf->n_corner; ++i) {
        bool member = false;
        for (int q = nf->p2s_off[i]; q < nf->p2s_off[i + 1]; ++q) member = member || nf->p2s_ids[q] == nl[k];
        if (!member) continue;
        const double pw[3] = {w[4 * i], w[4 * i + 1], w[4 * i + 2]};
        Blk blk;
        world2local(nf->R_wl, nf->t_wl, pw, blk.c);                            // Optimization.cpp:407 / :422
        blk.c[3] = a[0]; blk.c[4] = a[1]; blk.c[5] = a[2]; blk.c[6] = d[0]; blk.c[7] = d[1]; blk.c[8] = d[2]; blk.c[9] = weight; blk.c[10] = blk.c[11] = 0.0;
        per_edge[e].push_back(blk);
      }
    }
  }
  if (bad) return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref_out, nei_out, normalize, huber, consts};
  for (int e = 0; e < n_edges; ++e)
    for (const Blk& blk : per_edge[e])                                         // angle residuals are added with loss == nullptr (:417), metre residuals with HuberLoss(0.2) (:430)
      if (!push_block(o, angle_residual ? PVB_P2LINE_ANGLE : PVB_P2LINE_METER, ref[e], nei[e], normalize_distance, angle_residual ? 0.0 : 0.2, blk.c)) return PVB_ERR_NOMEM;
  return (int)o.at;
}

// CameraLidarLineAssociate::UniqueLinePair (:754-876): reduce (image line, LiDAR line, score) candidates, processed in input order, to a
// one-to-one matching in which a pair displaces an existing one only with a strictly smaller score; output ascending by image line.
int pvb_unique_line_pairs(int n, const int* image_line, const int* lidar_line, const float* score, int* n_out, int* out_image, int* out_lidar, float* out_score) {
  if (n < 0 || !n_out || (n > 0 && (!image_line || !lidar_line || !score || !out_image || !out_lidar || !out_score))) return PVB_ERR_ARG;
  int max_i = -1, max_l = -1;
  for (int k = 0; k < n; ++k) { if (image_line[k] < 0 || lidar_line[k] < 0) return PVB_ERR_ARG; max_i = std::max(max_i, image_line[k]); max_l = std::max(max_l, lidar_line[k]); }
  std::vector<int> lidar_of(max_i + 1, -1), image_of(max_l + 1, -1);          // current partner (or -1) of every image / LiDAR line
  std::vector<float> s_img(max_i + 1, 0.f);                                   // score of the pair an image line currently holds
  auto drop_image = [&](int i) { image_of[lidar_of[i]] = -1; lidar_of[i] = -1; };
  auto take = [&](int i, int l, float sc) { lidar_of[i] = l; image_of[l] = i; s_img[i] = sc; };
  for (int k = 0; k < n; ++k) {
    const int i = image_line[k], l = lidar_line[k]; const float sc = score[k];
    const bool hi = lidar_of[i] >= 0, hl = image_of[l] >= 0;
    if (!hi && !hl) take(i, l, sc);
    else if (hi && !hl) { if (sc < s_img[i]) { drop_image(i); take(i, l, sc); } }
    else if (!hi && hl) { if (sc < s_img[image_of[l]]) { drop_image(image_of[l]); take(i, l, sc); } }
    else {
      const float si = s_img[i], sl = s_img[image_of[l]];
      if (sc < std::min(si, sl)) { drop_image(image_of[l]); drop_image(i); take(i, l, sc); }   // beats both: replaces both
      else if (sc > si && sc < sl) drop_image(image_of[l]);                     // would displace l's pair but loses to i's: l's pair goes
      else if (sc < si && sc > sl) drop_image(i);                               // symmetric
    }
  }
  int m = 0;
  for (int i = 0; i <= max_i; ++i) if (lidar_of[i] >= 0) { out_image[m] = i; out_lidar[m] = lidar_of[i]; out_score[m] = s_img[i]; ++m; }
  *n_out = m;
  return PVB_OK;
}

int pvb_camera_lidar_associate(pvb_ctx* ctx, int rows, int cols, const float* lines4, int L, const pvb_line_frame* lidar, const double* T, int filter_by_length,
                               int multiple_association, const unsigned char* image_line_mask, const unsigned char* lidar_line_mask,
                               int cap, int* n_out, int* image_line, int* lidar_line, double* start3, double* end3, float* angle_out) {
  if (!ctx || !lidar || !T || !n_out || !lidar->end_points) return PVB_ERR_ARG;
  *n_out = 0;
  const int S = lidar->n_segments;
  if (L == 0 || S == 0) return PVB_OK;
  std::vector<int> counts((size_t)L * S, 0);
  int rc = pvb_angle_votes(ctx, rows, cols, lines4, L, lidar->corner_local, lidar->n_corner, lidar->p2s_off, lidar->p2s_ids, S, T, counts.data());
  if (rc) return rc;
  std::vector<int> seg_size(S, 0);
  for (int i = 0; i < lidar->n_corner; ++i) for (int e = lidar->p2s_off[i]; e < lidar->p2s_off[i + 1]; ++e) seg_size[lidar->p2s_ids[e]]++;
  std::vector<double> ep((size_t)S * 6), lplane((size_t)S * 4);
  for (int s = 0; s < S; ++s) {                                                // :373-386
    transform4(T, lidar->end_points + 6 * s, &ep[6 * s]);
    transform4(T, lidar->end_points + 6 * s + 3, &ep[6 * s + 3]);
    plane_through_origin(&ep[6 * s], &ep[6 * s + 3], &lplane[4 * s]);
    normalize4(&lplane[4 * s]);
  }
  const double thr = 3.0 / 180.0 * M_PI;
  struct P { int il, ll; double s[3], e[3]; float ang; };
  std::vector<P> pairs;
  for (int l = 0; l < L; ++l) {
    if (image_line_mask && !image_line_mask[l]) continue;                      // :396
    double p1[3], p2[3], plane[4];
    image_to_cam_f64(lines4[l * 4], lines4[l * 4 + 1], rows, cols, p1);
    image_to_cam_f64(lines4[l * 4 + 2], lines4[l * 4 + 3], rows, cols, p2);
    plane_through_origin(p1, p2, plane);
    normalize4(plane);
    const double p4[3] = {(p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0, (p1[2] + p2[2]) / 2.0};
    const double scope = vector_angle(p1, p4);
    for (int s = 0; s < S; ++s) {                                              // std::map order = ascending segment id (:415)
      const int cnt = counts[(size_t)l * S + s];
      if (cnt == 0) continue;
      if ((size_t)cnt < (size_t)seg_size[s] / 2) continue;                     // :417
      if (lidar_line_mask && !lidar_line_mask[s]) continue;                    // :420
      const double ang = plane_angle(plane, &lplane[4 * s], true);
      if (ang > thr) continue;                                                 // :422-424
      double mid[3], midp[3];
      for (int k = 0; k < 3; ++k) mid[k] = (ep[6 * s + k] + ep[6 * s + 3 + k]) / 2.f;
      project_to_plane(mid, plane, midp);
      if (vector_angle(midp, p4) > scope) continue;                            // :433
      const float ang2 = vector_angle(mid, midp);
      if (ang2 > thr / 2.0) continue;                                          // :436
      P pr; pr.il = l; pr.ll = s; pr.ang = ang + ang2;
      std::memcpy(pr.s, &ep[6 * s], 24); std::memcpy(pr.e, &ep[6 * s + 3], 24);
      pairs.push_back(pr);
    }
  }
  double Tlc[16] = {0};                                                         // T_cl^-1 (rigid), :469
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Tlc[r * 4 + c] = T[c * 4 + r];
  for (int r = 0; r < 3; ++r) Tlc[r * 4 + 3] = -(Tlc[r * 4] * T[3] + Tlc[r * 4 + 1] * T[7] + Tlc[r * 4 + 2] * T[11]);
  Tlc[15] = 1;
  std::vector<P> kept;
  for (const P& p : pairs) {
    if (filter_by_length) {                                                    // Filter(false, true): :676-692
      float ua, va, ub, vb;
      cam_to_image_f32((float)p.s[0], (float)p.s[1], (float)p.s[2], rows, cols, ua, va);
      cam_to_image_f32((float)p.e[0], (float)p.e[1], (float)p.e[2], rows, cols, ub, vb);
      const float a[2] = {ua, va}, b[2] = {ub, vb};
      const auto seg = break_to_segments(rows, cols, a, b, 100);
      float len = 0;
      for (size_t i = 0; i + 1 < seg.size(); ++i) {
        if (std::abs(seg[i].first - seg[i + 1].first) > 0.8 * cols) continue;
        const float dx = seg[i].first - seg[i + 1].first, dy = seg[i].second - seg[i + 1].second;
        len += std::sqrt(dx * dx + dy * dy);
      }
      if (len < 100.f || len > 2000.f) continue;
    }
    kept.push_back(p);
  }
  if (!multiple_association && !kept.empty()) {                                // :465-466 UniqueLinePair; pairs rebuilt from the camera-frame end points (:866-875)
    const int m = (int)kept.size();
    std::vector<int> il(m), ll(m), oi(m), ol(m); std::vector<float> sc(m), os(m);
    for (int k = 0; k < m; ++k) { il[k] = kept[k].il; ll[k] = kept[k].ll; sc[k] = kept[k].ang; }
    int mu = 0;
    const int rc2 = pvb_unique_line_pairs(m, il.data(), ll.data(), sc.data(), &mu, oi.data(), ol.data(), os.data());
    if (rc2) return rc2;
    kept.resize(mu);
    for (int k = 0; k < mu; ++k) { kept[k].il = oi[k]; kept[k].ll = ol[k]; kept[k].ang = os[k]; std::memcpy(kept[k].s, &ep[6 * ol[k]], 24); std::memcpy(kept[k].e, &ep[6 * ol[k] + 3], 24); }
  }
  int n = 0;
  for (const P& p : kept) {
    if (n >= cap) return PVB_ERR_ARG;
    image_line[n] = p.il; lidar_line[n] = p.ll; angle_out[n] = p.ang;
    transform4(Tlc, p.s, start3 + 3 * n); transform4(Tlc, p.e, end3 + 3 * n);
    ++n;
  }
  *n_out = n;
  return PVB_OK;
}

// ---- CameraLidarLineAssociate::Filter (joint_optimization/CameraLidarLineAssociate.cpp:628-715) on pairs in the camera frame -----------------
namespace {
// projected length (px) of a camera-frame LiDAR line: CamToImage of both ends, BreakToSegments(100), seam pieces skipped (:676-687)
float projected_length(int rows, int cols, const double* s3, const double* e3) {
  float ua, va, ub, vb;
  cam_to_image_f32((float)s3[0], (float)s3[1], (float)s3[2], rows, cols, ua, va);
  cam_to_image_f32((float)e3[0], (float)e3[1], (float)e3[2], rows, cols, ub, vb);
  const float a[2] = {ua, va}, b[2] = {ub, vb};
  const auto seg = break_to_segments(rows, cols, a, b, 100);
  float len = 0;
  for (size_t i = 0; i + 1 < seg.size(); ++i) {
    if (std::abs(seg[i].first - seg[i + 1].first) > 0.8 * cols) continue;
    const float dx = seg[i].first - seg[i + 1].first, dy = seg[i].second - seg[i + 1].second;
    len += std::sqrt(dx * dx + dy * dy);
  }
  return len;
}
}  // namespace

int pvb_filter_line_pairs(int rows, int cols, int n, const float* image_line4, const double* start3, const double* end3, int filter_by_angle, int filter_by_length,
                          unsigned char* keep, float* angle) {
  if (rows <= 0 || cols <= 0 || n < 0 || (n > 0 && (!image_line4 || !start3 || !end3 || !keep)) || (filter_by_angle && n > 0 && !angle)) return PVB_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    const double* ls = start3 + 3 * (size_t)i; const double* le = end3 + 3 * (size_t)i;
    bool ok = true;
    if (filter_by_angle) {
      // the two great-circle planes: LiDAR line and image line, each through the sphere centre, normalised as 4-vectors (:641-649)
      double pl_lidar[4], pl_img[4], p1[3], p2[3];
      plane_through_origin(ls, le, pl_lidar); normalize4(pl_lidar);
      image_to_cam_f64(image_line4[4 * i], image_line4[4 * i + 1], rows, cols, p1);
      image_to_cam_f64(image_line4[4 * i + 2], image_line4[4 * i + 3], rows, cols, p2);
      plane_through_origin(p1, p2, pl_img); normalize4(pl_img);
      const double deg = plane_angle(pl_lidar, pl_img, true) * 180.0 / M_PI;
      ok = !(deg > 5);                                                                                // :651
      if (ok) {
        angle[i] = (float)deg;
        const double half_arc = vector_angle(p1, p2) / 2.0;
        const double mid[3] = {(p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0, (p1[2] + p2[2]) / 2.0};
        double sp[3], ep[3];
        project_to_plane(ls, pl_img, sp); project_to_plane(le, pl_img, ep);
        ok = !(vector_angle(sp, mid) > half_arc) && !(vector_angle(ep, mid) > half_arc);              // :660-663
      }
      if (ok) {                                                                                       // both ends pushed to radius 5 must stay within 0.4 of the image plane (:665-670)
        const double ns = std::sqrt(sq(ls[0]) + sq(ls[1]) + sq(ls[2])), ne = std::sqrt(sq(le[0]) + sq(le[1]) + sq(le[2]));
        const double a5[3] = {ls[0] / ns * 5, ls[1] / ns * 5, ls[2] / ns * 5}, b5[3] = {le[0] / ne * 5, le[1] / ne * 5, le[2] / ne * 5};
        const double da = std::fabs(pl_img[0] * a5[0] + pl_img[1] * a5[1] + pl_img[2] * a5[2] + pl_img[3]);
        const double db = std::fabs(pl_img[0] * b5[0] + pl_img[1] * b5[1] + pl_img[2] * b5[2] + pl_img[3]);
        ok = !((float)std::min(da, db) > 0.4);
      }
    }
    if (ok && filter_by_length) { const float len = projected_length(rows, cols, ls, le); ok = !(len < 100.f || len > 2000.f); }   // :688-692
    keep[i] = ok ? 1 : 0;
  }
  return PVB_OK;
}

// ---- pixel-space Associate, first stage (joint_optimization/CameraLidarLineAssociate.cpp:22-102) -----------------------------------------
// image lines -> sub-line mid points: BreakToSegments(line, 70), pieces across the +-pi seam skipped (:38-54)
int pvb_pixel_sub_lines(int rows, int cols, const float* lines4, int n_lines, int cap, float* mid2, int* sub_to_line) {
  if (rows <= 0 || cols <= 0 || n_lines < 0 || (n_lines > 0 && !lines4) || cap < 0 || (cap > 0 && (!mid2 || !sub_to_line))) return PVB_ERR_ARG;
  int m = 0;
  for (int l = 0; l < n_lines; ++l) {
    const auto seg = break_to_segments(rows, cols, lines4 + 4 * l, lines4 + 4 * l + 2, 70);
    for (size_t i = 0; i + 1 < seg.size(); ++i) {
      if (std::abs(seg[i].first - seg[i + 1].first) > 0.8 * cols) continue;
      if (m >= cap) return PVB_ERR_NOMEM;
      mid2[2 * m] = (float)((seg[i + 1].first + seg[i].first) / 2.0);
      mid2[2 * m + 1] = (float)((seg[i + 1].second + seg[i].second) / 2.0);
      sub_to_line[m++] = l;
    }
  }
  return m;
}

static int sub_line_cap(const float* lines4, int n_lines) {
  double c = 64;
  for (int l = 0; l < n_lines; ++l) c += std::hypot((double)lines4[4 * l] - lines4[4 * l + 2], (double)lines4[4 * l + 1] - lines4[4 * l + 3]) / 70.0 + 4.0;
  return (int)c;
}

int pvb_pixel_line_neighbors(pvb_ctx* ctx, int rows, int cols, const float* lines4, int n_lines, const float* cloud_local, int n_points, const double* T_cl16,
                             int* line3, float* d2_3, float* pixel2) {
  if (!ctx || n_lines < 0 || n_points < 0 || (n_lines > 0 && !lines4) || (n_points > 0 && (!cloud_local || !line3)) || !T_cl16) return PVB_ERR_ARG;
  const int cap = sub_line_cap(lines4, n_lines);
  std::vector<float> mid((size_t)cap * 2); std::vector<int> s2l(cap);
  const int M = pvb_pixel_sub_lines(rows, cols, lines4, n_lines, cap, mid.data(), s2l.data());
  if (M < 0) return M;
  std::vector<float> d2((size_t)std::max(n_points, 1) * 3);
  const int rc = pvb_pixel_knn3(ctx, rows, cols, mid.data(), M, cloud_local, n_points, T_cl16, line3, d2.data(), pixel2);
  if (rc) return rc;
  for (long i = 0; i < (long)n_points * 3; ++i) {
    const int m = line3[i];
    line3[i] = (m >= 0 && !(d2[i] > 60 * 60)) ? s2l[m] : -1;                      // :81 vecDist > 60 * 60 is skipped
    if (d2_3) d2_3[i] = d2[i];
  }
  return PVB_OK;
}

// line -> LiDAR point lists (`line_lidar`, :83): ascending point index, one entry per neighbouring sub-line (a point may appear up to three times
// under the same line); lines with fewer than min_points entries (6, :92) are emptied
int pvb_pixel_line_candidates(int n_lines, int n_points, const int* line3, int min_points, int cap, int* line_off, int* lidar_idx) {
  if (n_lines < 0 || n_points < 0 || (n_points > 0 && !line3) || !line_off || (cap > 0 && !lidar_idx)) return PVB_ERR_ARG;
  std::vector<int> count(n_lines + 1, 0);
  for (long i = 0; i < (long)n_points * 3; ++i) if (line3[i] >= 0) { if (line3[i] >= n_lines) return PVB_ERR_ARG; count[line3[i]]++; }
  line_off[0] = 0;
  for (int l = 0; l < n_lines; ++l) line_off[l + 1] = line_off[l] + (count[l] >= min_points ? count[l] : 0);
  if (line_off[n_lines] > cap) return PVB_ERR_NOMEM;
  std::vector<int> fill(line_off, line_off + n_lines);
  for (int i = 0; i < n_points; ++i)
    for (int k = 0; k < 3; ++k) { const int l = line3[3 * i + k]; if (l >= 0 && count[l] >= min_points) lidar_idx[fill[l]++] = i; }
  return line_off[n_lines];
}

// ---- FitLineRANSAC + the end-point tail of the pixel-space Associate (joint_optimization/CameraLidarLineAssociate.cpp:105-144, 717-752) ------------
// PARITY UNPINNED for the sample-consensus part: the reference calls pcl::SACSegmentation (SACMODEL_LINE, SAC_RANSAC, threshold 0.1, PCL's defaults
// max_iterations = 50, probability = 0.99, random = false); PCL is a system dependency (libpcl-dev of the reference's Dockerfile, not in /root/reference, not
// installed here), so what follows restates PCL's published sequential algorithm (sample_consensus/ransac.hpp, sac_model.h, sac_model_line.hpp, common/centroid.hpp,
// common/eigen.hpp as of PCL 1.10 - 1.12; the Dockerfile does not pin a version) without an answer to compare with.  Version note: isSampleGood is written as an exact
// comparison in some releases and as an epsilon comparison in others; both reject a duplicated point, and a pair closer than float epsilon in all three coordinates is
// rejected by computeModelCoefficients below at the cost of the same two generator draws, so the sample sequence is the same.  Everything of the reference's OWN code around it (:117-137) is restated as written.
namespace {
// boost::mt19937 seeded with 12345 + boost::uniform_int<>(0, INT_MAX): the bucket method divides the 32-bit output by 2 (SampleConsensusModel::rnd())
struct Mt19937 {
  uint32_t mt[624]; int at;
  explicit Mt19937(uint32_t seed) { mt[0] = seed; for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i; at = 624; }
  uint32_t next() {
    if (at >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      at = 0;
    }
    uint32_t y = mt[at++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
};

// squared float32 distance of point p to the line (a, unit d): |(a - p) x d|^2 (Vector4f::cross3, sac_model_line.hpp countWithinDistance / selectWithinDistance)
inline float line_sqdist_f32(const float* a, const float* d, const float* p) {
  const float vx = a[0] - p[0], vy = a[1] - p[1], vz = a[2] - p[2];
  const float cx = vy * d[2] - vz * d[1], cy = vz * d[0] - vx * d[2], cz = vx * d[1] - vy * d[0];
  return cx * cx + cy * cy + cz * cz;
}
inline void normalize3_f32(float* v) {                                   // Eigen: if (squaredNorm > 0) v /= sqrt(squaredNorm)
  const float z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (z > 0.f) { const float n = std::sqrt(z); v[0] /= n; v[1] /= n; v[2] /= n; }
}

// pcl::computeRoots2 / computeRoots / eigen33(mat, evals) / computeCorrespondingEigenVector in float32 (common/impl/eigen.hpp)
void roots2_f32(float b, float c, float* r) {
  r[0] = 0.f;
  float d = b * b - 4.0f * c;
  if (d < 0.f) d = 0.f;
  const float sd = std::sqrt(d);
  r[2] = 0.5f * (b + sd); r[1] = 0.5f * (b - sd);
}
void roots3_f32(const float m[3][3], float* r) {
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) { roots2_f32(c2, c1, r); return; }
  const float inv3 = 1.0f / 3.0f, sqrt3 = std::sqrt(3.0f);
  const float c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3;
  if (a_3 > 0.f) a_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_3 * (2.f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3;
  if (q > 0.f) q = 0.f;
  const float rho = std::sqrt(-a_3), theta = std::atan2(std::sqrt(-q), half_b) * inv3, ct = std::cos(theta), st = std::sin(theta);
  r[0] = c2_3 + 2.f * rho * ct;
  r[1] = c2_3 - rho * (ct + sqrt3 * st);
  r[2] = c2_3 - rho * (ct - sqrt3 * st);
  if (r[0] >= r[1]) std::swap(r[0], r[1]);
  if (r[1] >= r[2]) { std::swap(r[1], r[2]); if (r[0] >= r[1]) std::swap(r[0], r[1]); }
  if (r[0] <= 0.f) roots2_f32(c2, c1, r);
}
inline float max_abs33(const float m[3][3]) { float s = 0.f; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s = std::max(s, std::fabs(m[i][j])); return s; }
void largest_eigenvector_f32(const float cov[3][3], float* vec) {
  float scale = max_abs33(cov);
  if (scale <= std::numeric_limits<float>::min()) scale = 1.f;
  float sm[3][3], ev[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sm[i][j] = cov[i][j] / scale;
  roots3_f32(sm, ev);
  const float lambda = ev[2] * scale;                                   // eigen33(mat, evals): evals *= scale
  // computeCorrespondingEigenVector(mat, lambda): its own scaling, (mat / scale - lambda / scale * I), largest of the three row cross products
  for (int i = 0; i < 3; ++i) sm[i][i] -= lambda / scale;
  float c[3][3];
  const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  float len[3];
  for (int k = 0; k < 3; ++k) {
    const float* a = sm[pr[k][0]]; const float* b = sm[pr[k][1]];
    c[k][0] = a[1] * b[2] - a[2] * b[1]; c[k][1] = a[2] * b[0] - a[0] * b[2]; c[k][2] = a[0] * b[1] - a[1] * b[0];
    len[k] = c[k][0] * c[k][0] + c[k][1] * c[k][1] + c[k][2] * c[k][2];
  }
  const int best = (len[0] >= len[1] && len[0] >= len[2]) ? 0 : ((len[1] >= len[0] && len[1] >= len[2]) ? 1 : 2);
  const float n = std::sqrt(len[best]);
  for (int j = 0; j < 3; ++j) vec[j] = c[best][j] / n;
}
}  // namespace

int pvb_pixel_fit_line(const float* xyz, int n, int stride, double dist_threshold, int max_iterations, double probability, float* coeff6, int cap, int* inliers,
                       double* start3, double* end3) {
  if (n < 0 || stride < 3 || (n > 0 && !xyz) || !coeff6 || (cap > 0 && !inliers) || !start3 || !end3 || max_iterations < 0 || !(probability > 0.0 && probability < 1.0))
    return PVB_ERR_ARG;
  auto P = [&](int i) { return xyz + (size_t)i * stride; };
  if (n < 2) return 0;                                                   // getSamples: fewer points than the sample size -> no model
  // --- RandomSampleConsensus::computeModel over SampleConsensusModelLine
  Mt19937 gen(12345u);
  std::vector<int> shuffled(n);
  for (int i = 0; i < n; ++i) shuffled[i] = i;
  const double sqr_thr = dist_threshold * dist_threshold, log_p = std::log(1.0 - probability), one_over_n = 1.0 / (double)n;
  const double eps = std::numeric_limits<double>::epsilon();
  int iterations = 0, best = -std::numeric_limits<int>::max();
  unsigned skipped = 0; const unsigned max_skip = (unsigned)max_iterations * 10u;
  double k = std::numeric_limits<double>::max();
  float best_model[6]; bool have = false;
  while ((double)iterations < k && skipped < max_skip) {
    int s0 = -1, s1 = -1;
    for (int check = 0; check < 1000; ++check) {                         // getSamples: drawIndexSample until isSampleGood (max_sample_checks_)
      for (int i = 0; i < 2; ++i) std::swap(shuffled[i], shuffled[i + (int)((gen.next() >> 1) % (uint32_t)(n - i))]);
      const float* a = P(shuffled[0]); const float* b = P(shuffled[1]);
      if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) { s0 = shuffled[0]; s1 = shuffled[1]; break; }
    }
    if (s0 < 0) break;                                                   // "No samples could be selected"
    const float* a = P(s0); const float* b = P(s1);
    const float fe = std::numeric_limits<float>::epsilon();
    if (std::fabs(a[0] - b[0]) <= fe && std::fabs(a[1] - b[1]) <= fe && std::fabs(a[2] - b[2]) <= fe) { ++skipped; continue; }
    float model[6] = {a[0], a[1], a[2], b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    normalize3_f32(model + 3);
    float dir[3] = {model[3], model[4], model[5]};
    normalize3_f32(dir);                                                 // countWithinDistance normalises the direction once more
    int count = 0;
    for (int i = 0; i < n; ++i) if ((double)line_sqdist_f32(model, dir, P(i)) < sqr_thr) ++count;
    if (count > best) {
      best = count; have = true;
      std::memcpy(best_model, model, sizeof model);
      const double w = (double)best * one_over_n;
      double p_no_outliers = 1.0 - w * w;                                // pow(w, sample size = 2)
      p_no_outliers = std::max(eps, p_no_outliers);
      p_no_outliers = std::min(1.0 - eps, p_no_outliers);
      k = log_p / std::log(p_no_outliers);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if (!have) return 0;
  std::vector<int> in;
  {
    float dir[3] = {best_model[3], best_model[4], best_model[5]};
    normalize3_f32(dir);
    for (int i = 0; i < n; ++i) if ((double)line_sqdist_f32(best_model, dir, P(i)) < sqr_thr) in.push_back(i);
  }
  const int m = (int)in.size();
  if (m < 3) return m;                                                   // FitLineRANSAC :729 -> false
  if (m > cap) return PVB_ERR_NOMEM;
  // --- FitLineRANSAC :731-749: centroid + (unnormalised) covariance of the inliers in float32, direction = eigenvector of the largest eigenvalue
  float cen[3] = {0.f, 0.f, 0.f};
  for (int i : in) { cen[0] += P(i)[0]; cen[1] += P(i)[1]; cen[2] += P(i)[2]; }
  for (int j = 0; j < 3; ++j) cen[j] /= (float)m;
  float cov[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int i : in) {
    float px = P(i)[0] - cen[0]; float py = P(i)[1] - cen[1]; float pz = P(i)[2] - cen[2];
    cov[1][1] += py * py; cov[1][2] += py * pz; cov[2][2] += pz * pz;
    py *= px; pz *= px; px *= px;
    cov[0][0] += px; cov[0][1] += py; cov[0][2] += pz;
  }
  cov[1][0] = cov[0][1]; cov[2][0] = cov[0][2]; cov[2][1] = cov[1][2];
  for (int j = 0; j < 3; ++j) coeff6[j] = cen[j];
  largest_eigenvector_f32(cov, coeff6 + 3);
  for (int i = 0; i < m; ++i) inliers[i] = in[i];
  // --- Associate :117-137: the two inliers farthest apart (float32 squared distance, first maximum) ... whose POSITIONS in the inlier list are then used as
  // indices into the candidate cloud (`line_points[start]`, not `line_points[inliers[start]]`) - reproduced as written
  int start = 0, end = 0; float max_d = -1.f;
  for (int i = 0; i < m; ++i)
    for (int j = i + 1; j < m; ++j) {
      const float* a = P(in[i]); const float* b = P(in[j]);
      const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
      const float d = dx * dx + dy * dy + dz * dz;
      if (d > max_d) { max_d = d; start = i; end = j; }
    }
  const double x0 = coeff6[0], y0 = coeff6[1], z0 = coeff6[2], nx = coeff6[3], ny = coeff6[4], nz = coeff6[5];     // ProjectPoint2Line3D<double>, base/Geometry.hpp:151-161
  const int ends[2] = {start, end}; double* out[2] = {start3, end3};
  for (int e = 0; e < 2; ++e) {
    const float* p = P(ends[e]);
    const double kk = (nx * ((double)p[0] - x0) + ny * ((double)p[1] - y0) + nz * ((double)p[2] - z0)) / (nx * nx + ny * ny + nz * nz);
    out[e][0] = kk * nx + x0; out[e][1] = kk * ny + y0; out[e][2] = kk * nz + z0;
  }
  return m;
}

// All candidate lists of one image (the loop of Associate :88-146): list l = points cloud_cam[lidar_idx[line_off[l] .. line_off[l+1])] (pvb_pixel_line_candidates);
// every reference call builds a fresh SACSegmentation (seed 12345), so the fits are independent and run on all host threads.  n_inliers[l] < 3: no line.
int pvb_pixel_fit_lines(const float* cloud_cam, int n_points, int n_lines, const int* line_off, const int* lidar_idx, double dist_threshold, int max_iterations,
                        double probability, int* n_inliers, float* coeff6, double* start3, double* end3) {
  if (n_lines < 0 || n_points < 0 || !line_off || (n_lines > 0 && (!n_inliers || !coeff6 || !start3 || !end3)) || (n_points > 0 && !cloud_cam)) return PVB_ERR_ARG;
  if (n_lines > 0 && line_off[n_lines] > 0 && !lidar_idx) return PVB_ERR_ARG;
  for (int l = 0; l < n_lines; ++l) if (line_off[l + 1] < line_off[l]) return PVB_ERR_ARG;
  for (int i = 0; i < (n_lines ? line_off[n_lines] : 0); ++i) if (lidar_idx[i] < 0 || lidar_idx[i] >= n_points) return PVB_ERR_ARG;
  int bad = 0;
#pragma omp parallel for schedule(dynamic)
  for (int l = 0; l < n_lines; ++l) {
    const int m = line_off[l + 1] - line_off[l];
    n_inliers[l] = 0;
    if (m == 0) continue;
    std::vector<float> pts((size_t)m * 3);
    for (int i = 0; i < m; ++i) std::memcpy(&pts[(size_t)i * 3], cloud_cam + (size_t)lidar_idx[line_off[l] + i] * 4, 12);
    std::vector<int> inl(m);
    const int rc = pvb_pixel_fit_line(pts.data(), m, 3, dist_threshold, max_iterations, probability, coeff6 + (size_t)l * 6, m, inl.data(), start3 + (size_t)l * 3, end3 + (size_t)l * 3);
    if (rc < 0) {
#pragma omp atomic write
      bad = rc;
    } else n_inliers[l] = rc;
  }
  return bad < 0 ? bad : PVB_OK;
}

// ---- pose interpolation around the sweep undistortion (base/Geometry.hpp:572-583, lidar_mapping/LidarOdometry.cpp:203-243) -------------
namespace {
// 4x4 inverse by Gauss-Jordan elimination with partial pivoting (the reference calls Eigen's general Matrix4d::inverse())
bool inverse4(const double* M, double* out) {
  double a[4][8];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = M[r * 4 + c]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (a[piv][c] == 0.0) return false;
    if (piv != c) for (int k = 0; k < 8; ++k) std::swap(a[piv][k], a[c][k]);
    const double inv = 1.0 / a[c][c];
    for (int k = 0; k < 8; ++k) a[c][k] *= inv;
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      const double f = a[r][c];
      if (f != 0.0) for (int k = 0; k < 8; ++k) a[r][k] -= f * a[c][k];
    }
  }
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[r * 4 + c] = a[r][4 + c];
  return true;
}
void mul4(const double* A, const double* B, double* C) {
  double T[16];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[r * 4 + c] = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c] + A[r * 4 + 3] * B[12 + c];
  memcpy(C, T, sizeof T);
}
bool slerp_pose(const double* pose_w1, const double* pose_w2, double ratio, double* out) {
  double inv2[16], T21[16];
  if (!inverse4(pose_w2, inv2)) return false;
  mul4(inv2, pose_w1, T21);
  double R21[9], q21[4], t21[3];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R21[r * 3 + c] = T21[r * 4 + c]; t21[r] = T21[r * 4 + 3]; }
  pvb::quat_from_matrix_eigen(R21, q21);
  pvb::UndistortPrep u;
  pvb::undistort_prepare(q21, t21, u);
  double w_id, w_q;
  pvb::slerp_weights(u, ratio, w_id, w_q);
  const double qs[4] = {w_q * q21[0], w_q * q21[1], w_q * q21[2], w_id + w_q * q21[3]};
  double Rs[9], Ts[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, invs[16];
  pvb::quat_to_matrix_eigen(qs, Rs);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Ts[r * 4 + c] = Rs[r * 3 + c]; Ts[r * 4 + 3] = t21[r] * ratio; }
  if (!inverse4(Ts, invs)) return false;
  mul4(pose_w1, invs, out);
  return true;
}
}  // namespace

int pvb_slerp_pose(const double* pose_w1_16, const double* pose_w2_16, double ratio, double* out16) {
  if (!pose_w1_16 || !pose_w2_16 || !out16) return PVB_ERR_ARG;
  return slerp_pose(pose_w1_16, pose_w2_16, ratio, out16) ? PVB_OK : PVB_ERR_ARG;
}

int pvb_undistort_end_poses(int n, const double* poses16, const unsigned char* pose_valid, const unsigned char* frame_valid, float gap_time, double* out_pose16,
                            unsigned char* has_end) {
  if (n < 0 || (n > 0 && (!poses16 || !pose_valid || !frame_valid || !out_pose16 || !has_end))) return PVB_ERR_ARG;
  const double sweep = 0.1;                                            // lidar_duration (LidarOdometry.cpp:203)
  const double period = sweep + gap_time;
  for (int i = 0; i < n; ++i) {
    double* dst = out_pose16 + (size_t)i * 16;
    std::fill(dst, dst + 16, 0.0);
    has_end[i] = 0;
    if (!pose_valid[i] || !frame_valid[i]) continue;
    const double* cur = poses16 + (size_t)i * 16;
    if (i + 1 < n) {
      // the next frame that is not (pose-less AND invalid): the reference's loop condition joins the two tests with && (:220)
      int nxt = i + 1;
      while (nxt < n && !pose_valid[nxt] && !frame_valid[nxt]) ++nxt;
      if (nxt == n) continue;
      // a neighbour that is valid but has no pose (R = 0, t = inf) is NOT stepped over by the reference; its SlerpPose then runs on a singular matrix and the
      // sweep comes out as NaN (pinned against the reference's own UndistortLidars, tests/test_reference_pinning.py) - reproduced, not repaired
      if (!slerp_pose(cur, poses16 + (size_t)nxt * 16, sweep / ((nxt - i) * period), dst)) std::fill(dst, dst + 16, std::numeric_limits<double>::quiet_NaN());
    } else {
      // last frame: extrapolate from an earlier frame (:229-240); the loop tests frame i's own `valid` flag and index 0 is rejected (:232)
      int prv = i - 1;
      while (prv >= 0 && !pose_valid[prv] && !frame_valid[i]) --prv;
      if (prv <= 0) continue;
      double mid[16], inv_cur[16], rel[16];
      if (!slerp_pose(poses16 + (size_t)prv * 16, cur, 1.0 - sweep / ((prv - i) * period), mid) || !inverse4(cur, inv_cur)) std::fill(dst, dst + 16, std::numeric_limits<double>::quiet_NaN());
      else { mul4(inv_cur, mid, rel); mul4(cur, rel, dst); }
    }
    has_end[i] = 1;
  }
  return PVB_OK;
}

// Calibration mode: the blocks of CameraLidarOptimizer::Optimize(line_pairs, T_cl) (joint_optimization/CameraLidarOptimizer.cpp:32-64) on ONE
// relative pose block.  Float path of the reference kept: ImageToCam(cv::Point2f, 5.f) and a float cross product for the image plane (:50-56),
// float half angle and middle of the image line inside PlaneRelativeIOUResidual's constructor (base/CostFunction.h:524-529).
int pvb_build_calibration_blocks(int rows, int cols, int n_pairs, const float* line4, const double* start3, const double* end3, int pose_block, long at, long cap,
                                 int* type, int* ref, int* nei, int* normalize, double* huber, double* consts) {
  if (rows <= 0 || cols <= 0 || n_pairs < 0 || (n_pairs > 0 && (!line4 || !start3 || !end3)) || !type || !ref || !nei || !normalize || !huber || !consts) return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref, nei, normalize, huber, consts};
  for (int i = 0; i < n_pairs; ++i) {
    float u[3], v[3];
    image_to_cam_f32(line4[4 * i], line4[4 * i + 1], rows, cols, 5.0f, u);
    image_to_cam_f32(line4[4 * i + 2], line4[4 * i + 3], rows, cols, 5.0f, v);
    // (v - u) x (0 - u) in float, promoted afterwards
    const float w0 = 0.f - u[0], w1 = 0.f - u[1], w2 = 0.f - u[2];
    const double nx = (v[1] - u[1]) * w2 - (v[2] - u[2]) * w1;
    const double ny = (v[2] - u[2]) * w0 - (v[0] - u[0]) * w2;
    const double nz = (v[0] - u[0]) * w1 - (v[1] - u[1]) * w0;
    const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
    const double* s = start3 + 3 * (size_t)i; const double* e = end3 + 3 * (size_t)i;
    const double c1[12] = {nx / len, ny / len, nz / len, e[0], e[1], e[2], s[0], s[1], s[2], 1.0, 0, 0};       // Plane2Plane_Relative(plane, end, start), HuberLoss(2 deg)
    if (!push_block(o, PVB_PLANE2PLANE_RELATIVE, pose_block, pose_block, 1, 2.0 * M_PI / 180.0, c1)) return PVB_ERR_NOMEM;
    float cs = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
    cs /= (std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) * std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    const float arc = cs >= 1.0f ? 0.0f : (cs <= -1.0f ? (float)M_PI : std::acos(cs));
    const double c2[12] = {nx / len, ny / len, nz / len, 0.0 / len, (s[0] + e[0]) / 2.0, (s[1] + e[1]) / 2.0, (s[2] + e[2]) / 2.0,
                           (double)((u[0] + v[0]) / 2.f), (double)((u[1] + v[1]) / 2.f), (double)((u[2] + v[2]) / 2.f), (double)(arc / 2.f), 2.0};
    if (!push_block(o, PVB_PLANE_RELATIVE_IOU, pose_block, pose_block, 1, 0.0, c2)) return PVB_ERR_NOMEM;      // loss == nullptr, weight 2 (:61-63)
  }
  return (int)o.at;
}

// ---- pose text files (util/FileIO.cpp:11-73 ReadPoseT, :168-191 ExportPoseT): the wire format either side of the path -----------------
// One line per frame: [name ]r00 r01 r02 tx r10 r11 r12 ty r20 r21 r22 tz, written with the default ostream precision (6 significant digits,
// "%g" - the reference's own precision loss, kept so that files stay interchangeable); a line holding "inf" / "nan" marks a frame without pose.
int pvb_write_poses_text(const char* path, int n, const double* R9, const double* t3, const char* const* names) {
  if (!path || n < 0 || (n > 0 && (!R9 || !t3))) return PVB_ERR_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return PVB_ERR_STATE;
  for (int i = 0; i < n; ++i) {
    if (names && names[i]) std::fprintf(f, "%s ", names[i]);
    const double* R = R9 + 9 * (size_t)i; const double* t = t3 + 3 * (size_t)i;
    std::fprintf(f, "%g %g %g %g %g %g %g %g %g %g %g %g\n", R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2]);
  }
  std::fclose(f);
  return PVB_OK;
}

// Returns the number of poses read (<= cap) or < 0.  valid[i] = 0: the line held inf / nan (R = 0, t = +inf as ReadPoseT leaves them); such
// lines are skipped unless with_invalid.  names (may be NULL): cap x name_len characters, empty string when a line has no name.
int pvb_read_poses_text(const char* path, int with_invalid, int cap, double* R9, double* t3, unsigned char* valid, char* names, int name_len) {
  if (!path || cap < 0 || (cap > 0 && (!R9 || !t3))) return PVB_ERR_ARG;
  FILE* f = std::fopen(path, "r");
  if (!f) return PVB_ERR_STATE;
  int n = 0;
  char line[4096];
  while (std::fgets(line, sizeof line, f)) {
    std::vector<std::string> tok;
    for (char* p = std::strtok(line, " \t\r\n"); p; p = std::strtok(nullptr, " \t\r\n")) tok.push_back(p);
    std::string name;
    if (tok.size() == 13) { name = tok[0]; tok.erase(tok.begin()); }                 // :31-35
    if (tok.size() != 12) { if (tok.empty()) continue; }
    bool ok = tok.size() == 12;
    double v[12] = {0};
    if (ok) for (const std::string& s : tok) if (s.find("inf") != std::string::npos || s.find("nan") != std::string::npos) { ok = false; break; }   // :41-48
    if (ok) for (int k = 0; k < 12; ++k) v[k] = std::strtod(tok[k].c_str(), nullptr);
    if (!ok && !with_invalid) continue;                                               // :67
    if (n >= cap) { std::fclose(f); return PVB_ERR_NOMEM; }
    double* R = R9 + 9 * (size_t)n; double* t = t3 + 3 * (size_t)n;
    if (ok) { R[0] = v[0]; R[1] = v[1]; R[2] = v[2]; t[0] = v[3]; R[3] = v[4]; R[4] = v[5]; R[5] = v[6]; t[1] = v[7]; R[6] = v[8]; R[7] = v[9]; R[8] = v[10]; t[2] = v[11]; }
    else { for (int k = 0; k < 9; ++k) R[k] = 0.0; t[0] = t[1] = t[2] = INFINITY; }   // :22-23
    if (valid) valid[n] = ok ? 1 : 0;
    if (names && name_len > 0) { std::snprintf(names + (size_t)n * name_len, name_len, "%s", name.c_str()); }
    ++n;
  }
  std::fclose(f);
  return n;
}

int pvb_build_point2plane_blocks(long n, const double* point3, const double* plane4, int ref_block, int nei_block, int angle_residual, int normalize_distance,
                                 double weight, long at, long cap, int* type, int* ref, int* nei, int* normalize, double* huber, double* consts) {
  if (n < 0 || (n > 0 && (!point3 || !plane4)) || !type || !ref || !nei || !normalize || !huber || !consts) return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref, nei, normalize, huber, consts};
  const double hub = angle_residual ? 2 * M_PI / 180.0 : 0.2;                  // Optimization.cpp:513-517
  for (long i = 0; i < n; ++i) {
    double c[12] = {point3[3 * i], point3[3 * i + 1], point3[3 * i + 2], plane4[4 * i], plane4[4 * i + 1], plane4[4 * i + 2], plane4[4 * i + 3], weight, 0, 0, 0, 0};
    if (!push_block(o, angle_residual ? PVB_P2PLANE_ANGLE : PVB_P2PLANE_METER, ref_block, nei_block, normalize_distance, hub, c)) return PVB_ERR_ARG;
  }
  return (int)o.at;
}

// the same for the correspondences of MANY edges at once (pvb_frames_get_point2plane returns them edge-major with their edge index): one call per
// outer iteration instead of one per pose-graph edge
int pvb_build_point2plane_blocks_edges(long n, const int* edge, const double* point3, const double* plane4, int n_edges, const int* edge_ref_block,
                                       const int* edge_nei_block, int angle_residual, int normalize_distance, double weight, long at, long cap, int* type, int* ref,
                                       int* nei, int* normalize, double* huber, double* consts) {
  if (n < 0 || n_edges < 0 || (n > 0 && (!edge || !point3 || !plane4 || !edge_ref_block || !edge_nei_block)) || !type || !ref || !nei || !normalize || !huber || !consts)
    return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref, nei, normalize, huber, consts};
  const double hub = angle_residual ? 2 * M_PI / 180.0 : 0.2;
  for (long i = 0; i < n; ++i) {
    const int e = edge[i];
    if (e < 0 || e >= n_edges) return PVB_ERR_ARG;
    double c[12] = {point3[3 * i], point3[3 * i + 1], point3[3 * i + 2], plane4[4 * i], plane4[4 * i + 1], plane4[4 * i + 2], plane4[4 * i + 3], weight, 0, 0, 0, 0};
    if (!push_block(o, angle_residual ? PVB_P2PLANE_ANGLE : PVB_P2PLANE_METER, edge_ref_block[e], edge_nei_block[e], normalize_distance, hub, c)) return PVB_ERR_ARG;
  }
  return (int)o.at;
}

int pvb_build_point2line_blocks(long n, const double* point3, const double* a3, const double* b3, int ref_block, int nei_block, int angle_residual, int normalize_distance,
                                double weight, long at, long cap, int* type, int* ref, int* nei, int* normalize, double* huber, double* consts) {
  if (n < 0 || (n > 0 && (!point3 || !a3 || !b3)) || !type || !ref || !nei || !normalize || !huber || !consts) return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref, nei, normalize, huber, consts};
  const double hub = angle_residual ? 2 * M_PI / 180.0 : 0.2;                  // Optimization.cpp:449-453 (a loss in both modes)
  for (long i = 0; i < n; ++i) {
    const double* a = a3 + 3 * i; const double* b = b3 + 3 * i;
    double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double c[12] = {point3[3 * i], point3[3 * i + 1], point3[3 * i + 2], a[0], a[1], a[2], d[0] / nn, d[1] / nn, d[2] / nn, weight, 0, 0};
    if (!push_block(o, angle_residual ? PVB_P2LINE_ANGLE : PVB_P2LINE_METER, ref_block, nei_block, normalize_distance, hub, c)) return PVB_ERR_ARG;
  }
  return (int)o.at;
}

int pvb_build_line2line_blocks(const pvb_line_frame* nf, const float* world, int nei_line, const double* a, const double* b, int ref_block, int nei_block,
                               int angle_residual, int normalize_distance, double weight, long at, long cap, int* type, int* ref, int* nei, int* normalize,
                               double* huber, double* consts) {
  if (!nf || !world || !a || !b || !type) return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref, nei, normalize, huber, consts};
  double d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};                       // Point2Line_*: line_direction = (a - b).normalized()
  const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  for (int k = 0; k < 3; ++k) d[k] /= nn;
  for (int i = 0; i < nf->n_corner; ++i) {
    bool member = false;
    for (int e = nf->p2s_off[i]; e < nf->p2s_off[i + 1]; ++e) member = member || nf->p2s_ids[e] == nei_line;
    if (!member) continue;
    const double pw[3] = {world[4 * i], world[4 * i + 1], world[4 * i + 2]};
    double pl[3];
    world2local(nf->R_wl, nf->t_wl, pw, pl);                                   // Optimization.cpp:407 / :422
    double c[12] = {pl[0], pl[1], pl[2], a[0], a[1], a[2], d[0], d[1], d[2], weight, 0, 0};
    // angle residuals are added with loss == nullptr (:417), metre residuals with HuberLoss(0.2) (:430)
    if (!push_block(o, angle_residual ? PVB_P2LINE_ANGLE : PVB_P2LINE_METER, ref_block, nei_block, normalize_distance, angle_residual ? 0.0 : 0.2, c)) return PVB_ERR_ARG;
  }
  return (int)o.at;
}

int pvb_build_camera_lidar_blocks(int rows, int cols, int n_pairs, const float* line4, const double* start3, const double* end3, const float* pair_weight,
                                  int cam_block, int lidar_block, double weight, long at, long cap, int* type, int* ref, int* nei, int* normalize, double* huber,
                                  double* consts) {
  if (n_pairs < 0 || (n_pairs > 0 && (!line4 || !start3 || !end3)) || !type) return PVB_ERR_ARG;
  BlockOut o{at, cap, type, ref, nei, normalize, huber, consts};
  const double hub = 3.0 * M_PI / 180.0;
  for (int i = 0; i < n_pairs; ++i) {
    double p1[3], p2[3], plane[4];
    image_to_cam_f64(line4[4 * i], line4[4 * i + 1], rows, cols, p1);           // Optimization.cpp:587-589
    image_to_cam_f64(line4[4 * i + 2], line4[4 * i + 3], rows, cols, p2);
    plane_through_origin(p1, p2, plane);
    const double nn = std::sqrt(sq(plane[0]) + sq(plane[1]) + sq(plane[2]));
    const double w = (pair_weight ? (double)pair_weight[i] : 1.0) * weight;
    const double* s = start3 + 3 * i; const double* e = end3 + 3 * i;
    // Plane2Plane_Global::Create(plane.head(3), lidar_line_end, lidar_line_start, w): ctor normalises the normal (CostFunction.h:362)
    double c1[12] = {plane[0] / nn, plane[1] / nn, plane[2] / nn, e[0], e[1], e[2], s[0], s[1], s[2], w, 0, 0};
    if (!push_block(o, PVB_PLANE2PLANE_GLOBAL, cam_block, lidar_block, 1, hub, c1)) return PVB_ERR_ARG;
    // PlaneIOUResidual::Create(plane, (end + start)/2, (p1 + p2)/2, VectorAngle3D(p1, p2, true), 2 w): plane / |n| (CostFunction.h:457)
    const double ang = vector_angle(p1, p2, true);
    double c2[12] = {plane[0] / nn, plane[1] / nn, plane[2] / nn, plane[3] / nn, (e[0] + s[0]) / 2.0, (e[1] + s[1]) / 2.0, (e[2] + s[2]) / 2.0,
                     (p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0, (p1[2] + p2[2]) / 2.0, ang, 2.0 * weight};
    if (!push_block(o, PVB_PLANE_IOU, cam_block, lidar_block, 1, hub, c2)) return PVB_ERR_ARG;
  }
  return (int)o.at;
}

}  // extern "C"
