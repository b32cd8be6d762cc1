// ORACLE — TEST INFRASTRUCTURE ONLY (see pvo_math.hpp header).  Parity status: unpinned by the
// reference's own tests; pinned against scipy cKDTree / numpy in tests/test_oracle_assoc.py.
//
// CPU restatement of the association code on the hot path:
//   sensors/Velodyne.cpp:1773-1859        Transform2LidarWorld / World2Local
//   lidar_mapping/LidarFeatureAssociate.cpp:120-197,219-236,442-476,550-630
//   joint_optimization/CameraLidarLineAssociate.cpp:340-475,628-715
//   sensors/Equirectangular.{h,cpp}, util/Visualization.h:408-441
// PCL/FLANN behaviour restated from their documentation: KdTreeFLANN = exact k-NN on
// float32 squared L2 (accumulated x,y,z in float), ascending; ties here are broken by lower index
// (FLANN's tie order is implementation-defined).
#pragma once
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include "pvo_math.hpp"

namespace pvo {

// ---- pcl::transformPointCloud(cloud, out, Matrix4d): double arithmetic, float32 store -----------
// (PCL common/impl/transforms.hpp Transformer<double>::se3; called from Velodyne.cpp:1790-1806)
inline void TransformCloud(const double R[9] /*row-major*/, const double t[3], const float* in, int n, float* out, int stride = 4) {
  for (int i = 0; i < n; ++i) {
    const double p[3] = {in[i * stride], in[i * stride + 1], in[i * stride + 2]};
    for (int r = 0; r < 3; ++r)
      out[i * stride + r] = static_cast<float>(R[r * 3 + 0] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2] + t[r]);
    for (int c = 3; c < stride; ++c) out[i * stride + c] = in[i * stride + c];
  }
}

// Velodyne::World2Local (Velodyne.cpp:1850-1853): R_wl^T * p - R_wl^T * t_wl
inline void World2Local(const double R[9], const double t[3], const double pw[3], double out[3]) {
  for (int r = 0; r < 3; ++r) {
    const double a = R[0 * 3 + r] * pw[0] + R[1 * 3 + r] * pw[1] + R[2 * 3 + r] * pw[2];
    const double b = R[0 * 3 + r] * t[0] + R[1 * 3 + r] * t[1] + R[2 * 3 + r] * t[2];
    out[r] = a - b;
  }
}

// ---- exact k-NN on float32 squared distances ------------------------------------------------
inline float SqDistF32(const float* a, const float* b) {  // flann::L2_Simple<float>
  float r = 0.f;
  const float dx = a[0] - b[0]; r += dx * dx;
  const float dy = a[1] - b[1]; r += dy * dy;
  const float dz = a[2] - b[2]; r += dz * dz;
  return r;
}

struct KnnResult { std::vector<int> idx; std::vector<float> d2; };

struct KnnHeap {  // keeps the k smallest (d2, idx) lexicographically; worst at front of a sorted array
  int k; int n = 0; std::vector<float> d; std::vector<int> id;
  explicit KnnHeap(int k_) : k(k_), d(k_), id(k_) {}
  inline bool full() const { return n == k; }
  inline float worst() const { return d[n - 1]; }
  inline bool accepts(float dd, int ii) const { return n < k || dd < d[n - 1] || (dd == d[n - 1] && ii < id[n - 1]); }
  inline void push(float dd, int ii) {
    if (!accepts(dd, ii)) return;
    int pos = (n < k) ? n : k - 1;
    while (pos > 0 && (d[pos - 1] > dd || (d[pos - 1] == dd && id[pos - 1] > ii))) { d[pos] = d[pos - 1]; id[pos] = id[pos - 1]; --pos; }
    d[pos] = dd; id[pos] = ii;
    if (n < k) ++n;
  }
};

inline void KnnBrute(const float* pts, int n, int stride, const float* q, int k, KnnResult& out) {
  KnnHeap h(k);
  for (int i = 0; i < n; ++i) h.push(SqDistF32(q, pts + (size_t)i * stride), i);
  out.idx.assign(h.id.begin(), h.id.begin() + h.n);
  out.d2.assign(h.d.begin(), h.d.begin() + h.n);
}

// Exact kd-tree (leaf size 15 like pcl::KdTreeFLANN's KDTreeSingleIndexParams(15)); returns the same
// set/order as KnnBrute.  Used for the timed CPU baseline so that it is a fair "reference-like" cost.
struct KdTree {
  struct Node { int lo, hi, left, right, dim; float split_lo, split_hi; };
  const float* pts = nullptr; int stride = 4;
  std::vector<int> order; std::vector<Node> nodes;
  std::vector<float> leaf_xyz;  // points copied in leaf order for locality
  void Build(const float* p, int n, int stride_) {
    pts = p; stride = stride_; order.resize(n); std::iota(order.begin(), order.end(), 0);
    nodes.clear(); nodes.reserve(2 * (n / 8 + 1));
    if (n > 0) BuildRec(0, n);
    leaf_xyz.resize((size_t)n * 3);
    for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) leaf_xyz[(size_t)i * 3 + c] = pts[(size_t)order[i] * stride + c];
  }
  int BuildRec(int lo, int hi) {
    const int id = (int)nodes.size(); nodes.push_back(Node{lo, hi, -1, -1, -1, 0.f, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; ++i) for (int c = 0; c < 3; ++c) { const float v = pts[(size_t)order[i] * stride + c]; mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v); }
    int dim = 0; for (int c = 1; c < 3; ++c) if (mx[c] - mn[c] > mx[dim] - mn[dim]) dim = c;
    if (mx[dim] == mn[dim]) return id;  // all identical: keep as a (large) leaf
    const int mid = (lo + hi) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                     [&](int a, int b) { return pts[(size_t)a * stride + dim] < pts[(size_t)b * stride + dim]; });
    float left_max = -FLT_MAX; for (int i = lo; i < mid; ++i) left_max = std::max(left_max, pts[(size_t)order[i] * stride + dim]);
    const float right_min = pts[(size_t)order[mid] * stride + dim];
    nodes[id].dim = dim; nodes[id].split_lo = left_max; nodes[id].split_hi = right_min;
    const int l = BuildRec(lo, mid); const int r = BuildRec(mid, hi);
    nodes[id].left = l; nodes[id].right = r;
    return id;
  }
  void SearchRec(int nid, const float* q, KnnHeap& h, double mind2) const {
    const Node& nd = nodes[nid];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) h.push(SqDistF32(q, &leaf_xyz[(size_t)i * 3]), order[i]);
      return;
    }
    const double v = q[nd.dim];
    int first = nd.left, second = nd.right; double cut;
    if (v <= 0.5 * ((double)nd.split_lo + nd.split_hi)) { cut = std::max(0.0, (double)nd.split_hi - v); }
    else { first = nd.right; second = nd.left; cut = std::max(0.0, v - (double)nd.split_lo); }
    SearchRec(first, q, h, mind2);
    const double bound = cut * cut * (1.0 - 1e-6);  // conservative vs float32 rounding of the distances
    if (!h.full() || bound <= (double)h.worst()) SearchRec(second, q, h, bound);
  }
  void Knn(const float* q, int k, KnnResult& out) const {
    KnnHeap h(k);
    if (!nodes.empty()) SearchRec(0, q, h, 0.0);
    out.idx.assign(h.id.begin(), h.id.begin() + h.n);
    out.d2.assign(h.d.begin(), h.d.begin() + h.n);
  }
};

// ---- AssociatePoint2Plane (LidarFeatureAssociate.cpp:550-630) ---------------------------------
struct P2PlaneAssoc { int query_idx; double point[3]; double plane[4]; };

// ref_world / nei_world: n x 4 float32 (x,y,z,intensity) already in world coordinates (float32, as
// produced by TransformCloud).  R/t are the frames' R_wl, t_wl.  k is 10 in the reference (:574).
// Quirk C.6: the reference indexes [k-1] even when PCL returns < k results; here such queries are
// rejected (documented difference).
inline void AssociatePoint2Plane(const float* ref_world, int n_ref, const double R_ref[9], const double t_ref[3],
                                 const float* nei_world, int n_nei, const double R_nei[9], const double t_nei[3],
                                 double plane_tolerance, float dist_threshold, int k, bool use_kdtree,
                                 std::vector<P2PlaneAssoc>& out, const KdTree* prebuilt = nullptr) {
  const float sq_thr = dist_threshold * dist_threshold;
  KdTree local_tree;
  const KdTree* tree = prebuilt;
  if (use_kdtree && !tree) { local_tree.Build(ref_world, n_ref, 4); tree = &local_tree; }  // rebuilt per call (:566-567)
  KnnResult res;
  std::vector<double> pl((size_t)k * 3);
  for (int idx = 0; idx < n_nei; ++idx) {
    const float* q = nei_world + (size_t)idx * 4;
    if (use_kdtree) tree->Knn(q, k, res); else KnnBrute(ref_world, n_ref, 4, q, k, res);
    if ((int)res.idx.size() < k) continue;
    if (res.d2[k - 1] > sq_thr) continue;
    int same = 0;
    for (int j = 0; j < k; ++j) {
      const float* pt = ref_world + (size_t)res.idx[j] * 4;
      same += (pt[3] == q[3]);
      const double pw[3] = {pt[0], pt[1], pt[2]};
      World2Local(R_ref, t_ref, pw, &pl[(size_t)j * 3]);
    }
    if (same < k) continue;
    double plane[4], line[6];
    FormPlaneLSQ(k, pl.data(), plane_tolerance, plane);
    const bool is_line = FormLinePCA(k, pl.data(), 3.0, 0.0, line);
    if ((plane[0] == 0 && plane[1] == 0 && plane[2] == 0 && plane[3] == 0) || is_line) continue;
    P2PlaneAssoc a; a.query_idx = idx;
    const double qw[3] = {q[0], q[1], q[2]};
    World2Local(R_nei, t_nei, qw, a.point);
    for (int c = 0; c < 4; ++c) a.plane[c] = plane[c];
    out.push_back(a);
  }
}

// ---- AssociatePoint2Line (LidarFeatureAssociate.cpp:478-548): 5-NN on the corner clouds, PCA line test in WORLD frame
struct P2LineAssoc { int query_idx; double point[3]; double a[3], b[3]; };

inline void AssociatePoint2Line(const float* ref_world, int n_ref, const double R_ref[9], const double t_ref[3],
                                const float* nei_world, int n_nei, const double R_nei[9], const double t_nei[3],
                                float dist_threshold, bool use_kdtree, std::vector<P2LineAssoc>& out) {
  const int k = 5;
  const float sq_thr = dist_threshold * dist_threshold;
  KdTree tree;
  if (use_kdtree) tree.Build(ref_world, n_ref, 4);
  KnnResult res;
  for (int idx = 0; idx < n_nei; ++idx) {
    const float* q = nei_world + (size_t)idx * 4;
    if (use_kdtree) tree.Knn(q, k, res); else KnnBrute(ref_world, n_ref, 4, q, k, res);
    if ((int)res.idx.size() < k) continue;                     // quirk C.6 guard
    if (res.d2[k - 1] > sq_thr) continue;                      // :497
    double pts[15];
    for (int j = 0; j < k; ++j) for (int c = 0; c < 3; ++c) pts[j * 3 + c] = ref_world[(size_t)res.idx[j] * 4 + c];   // :503 world coordinates
    double line[6];
    if (!FormLinePCA(k, pts, 10.0, 0.05, line)) continue;      // :506-509
    double aw[3], bw[3];
    for (int c = 0; c < 3; ++c) { aw[c] = 0.1 * line[3 + c] + line[c]; bw[c] = -0.1 * line[3 + c] + line[c]; }   // :514-515
    P2LineAssoc a; a.query_idx = idx;
    const double qw[3] = {q[0], q[1], q[2]};
    World2Local(R_nei, t_nei, qw, a.point);                    // :517
    World2Local(R_ref, t_ref, aw, a.a);                        // :518-519
    World2Local(R_ref, t_ref, bw, a.b);
    out.push_back(a);
  }
}

// ---- line-to-line (LidarFeatureAssociate.cpp:219-236, 442-476, 120-197) --------------------------
inline void TransformLine(const double R[9], const double t[3], const double in[6], double out[6]) {  // :219-236
  for (int r = 0; r < 3; ++r) {
    out[r] = R[r * 3] * in[0] + R[r * 3 + 1] * in[1] + R[r * 3 + 2] * in[2] + t[r];
    out[3 + r] = R[r * 3] * in[3] + R[r * 3 + 1] * in[4] + R[r * 3 + 2] * in[5];
  }
}

struct L2LAssoc { int nei_line, ref_line; double a[3], b[3]; };

// Vote matrix (row = nei segment, col = ref segment), AssociateLine2Line :459-473.
inline void Line2LineVotes(const double* ref_lines_world, int S_ref, const float* nei_corner_world, int n_pts,
                           const int* p2s_off, const int* p2s_ids, int S_nei, double dist_threshold, std::vector<int>& M) {
  M.assign((size_t)S_nei * S_ref, 0);
  for (int i = 0; i < n_pts; ++i) {
    const double p[3] = {nei_corner_world[i * 4], nei_corner_world[i * 4 + 1], nei_corner_world[i * 4 + 2]};
    for (int s = 0; s < S_ref; ++s) {
      const double d = PointToLineDistance3D(p, ref_lines_world + s * 6);
      if (d > dist_threshold) continue;
      for (int e = p2s_off[i]; e < p2s_off[i + 1]; ++e) M[(size_t)p2s_ids[e] * S_ref + s] += 1;
    }
  }
}

// FindAssociations (:120-197).  seg_sizes_nei[s] = nei.edge_segmented[s].size().
inline void FindAssociations(const double* ref_coeffs_local, const double* ref_lines_world, int S_ref,
                             const double* nei_lines_world, int S_nei, const int* seg_sizes_nei,
                             const std::vector<int>& M, std::vector<L2LAssoc>& out) {
  std::map<int, L2LAssoc> m;
  for (int s = 0; s < S_nei; ++s) {
    if (S_ref == 0) break;
    int max_col = 0, max_count = M[(size_t)s * S_ref];
    for (int c = 1; c < S_ref; ++c) if (M[(size_t)s * S_ref + c] > max_count) { max_count = M[(size_t)s * S_ref + c]; max_col = c; }
    if ((size_t)max_count < (size_t)seg_sizes_nei[s] / 2) continue;
    const double* dr = ref_lines_world + max_col * 6 + 3;
    const double* dn = nei_lines_world + s * 6 + 3;
    if (PlaneAngle(dr, dn) * 180.0 / M_PI > 7) continue;
    const double* cl = ref_coeffs_local + max_col * 6;
    L2LAssoc a; a.nei_line = s; a.ref_line = max_col;
    for (int c = 0; c < 3; ++c) { a.a[c] = 0.1 * cl[3 + c] + cl[c]; a.b[c] = -0.1 * cl[3 + c] + cl[c]; }
    auto it = m.find(max_col);
    if (it == m.end()) m.insert({max_col, a});
    else {
      const double d1 = PointToLineDistance3D(nei_lines_world + it->second.nei_line * 6, ref_lines_world + max_col * 6);
      const double d2 = PointToLineDistance3D(nei_lines_world + s * 6, ref_lines_world + max_col * 6);
      if (d2 < d1) it->second = a;
    }
  }
  for (auto& kv : m) out.push_back(kv.second);
}

// ---- segment-based variants (LidarFeatureAssociate.cpp:238-440; off in the shipped configs) ---------------------
struct P2SegAssoc { int query_idx, ref_line; double point[3], a[3], b[3]; };

// per query: number of its 5 nearest reference corner points that belong to each segment (std::map<size_t,size_t> seg_count, :264-270)
inline bool SegmentCounts(const KdTree* tree, const float* ref_world, int n_ref, const int* ref_p2s_off, const int* ref_p2s_ids,
                          const float* q, float sq_thr, KnnResult& res, std::map<int, int>& seg_count) {
  const int k = 5;
  if (tree) tree->Knn(q, k, res); else KnnBrute(ref_world, n_ref, 4, q, k, res);
  if ((int)res.idx.size() < k) return false;                   // quirk C.6 guard
  if (res.d2[k - 1] > sq_thr) return false;                    // :261, :416
  seg_count.clear();
  for (int j = 0; j < k; ++j) for (int e = ref_p2s_off[res.idx[j]]; e < ref_p2s_off[res.idx[j] + 1]; ++e) seg_count[ref_p2s_ids[e]]++;
  return true;
}

// AssociatePoint2LineSegmentKNN (:238-317): all 5 neighbours on one segment => point-to-line with that segment's coefficients
inline void AssociatePoint2LineSegmentKNN(const float* ref_world, int n_ref, const int* ref_p2s_off, const int* ref_p2s_ids, const double* ref_coeffs_local,
                                          const float* nei_world, int n_nei, const double R_nei[9], const double t_nei[3], float dist_threshold,
                                          bool use_kdtree, std::vector<P2SegAssoc>& out) {
  const float sq_thr = dist_threshold * dist_threshold;
  KdTree tree;
  if (use_kdtree) tree.Build(ref_world, n_ref, 4);
  KnnResult res; std::map<int, int> cnt;
  for (int idx = 0; idx < n_nei; ++idx) {
    const float* q = nei_world + (size_t)idx * 4;
    if (!SegmentCounts(use_kdtree ? &tree : nullptr, ref_world, n_ref, ref_p2s_off, ref_p2s_ids, q, sq_thr, res, cnt)) continue;
    for (auto& kv : cnt) {
      if (kv.second < 5 - 0) continue;                         // :276
      const double* cl = ref_coeffs_local + 6 * kv.first;
      P2SegAssoc a; a.query_idx = idx; a.ref_line = kv.first;
      for (int c = 0; c < 3; ++c) { a.a[c] = 0.1 * cl[3 + c] + cl[c]; a.b[c] = -0.1 * cl[3 + c] + cl[c]; }   // :283-284
      const double qw[3] = {q[0], q[1], q[2]};
      World2Local(R_nei, t_nei, qw, a.point);                  // :286
      out.push_back(a);
    }
  }
}

// AssociatePoint2LineSegment (:319-383): nearest infinite reference line (world), accepted when within dist_threshold
inline void AssociatePoint2LineSegment(const double* ref_lines_world, const double* ref_coeffs_local, int S_ref, const float* nei_world, int n_nei,
                                       const double R_nei[9], const double t_nei[3], float dist_threshold, std::vector<P2SegAssoc>& out) {
  for (int idx = 0; idx < n_nei; ++idx) {
    const double p[3] = {nei_world[(size_t)idx * 4], nei_world[(size_t)idx * 4 + 1], nei_world[(size_t)idx * 4 + 2]};
    double min_distance = std::numeric_limits<double>::max();
    int valid = -1;
    for (int s = 0; s < S_ref; ++s) {
      const double d = PointToLineDistance3D(p, ref_lines_world + 6 * s);
      if (d < min_distance) { min_distance = d; valid = s; }
    }
    if (!(min_distance <= dist_threshold)) continue;           // :343 (float threshold promoted)
    const double* cl = ref_coeffs_local + 6 * valid;
    P2SegAssoc a; a.query_idx = idx; a.ref_line = valid;
    for (int c = 0; c < 3; ++c) { a.a[c] = 0.1 * cl[3 + c] + cl[c]; a.b[c] = -0.1 * cl[3 + c] + cl[c]; }
    World2Local(R_nei, t_nei, p, a.point);
    out.push_back(a);
  }
}

// AssociateLine2LineKNN vote matrix (:401-436): >= 3 of the 5 neighbours on a segment => one vote per segment of the query point
inline void Line2LineKnnVotes(const float* ref_world, int n_ref, const int* ref_p2s_off, const int* ref_p2s_ids, int S_ref,
                              const float* nei_world, int n_nei, const int* nei_p2s_off, const int* nei_p2s_ids, int S_nei, float dist_threshold,
                              bool use_kdtree, std::vector<int>& M) {
  M.assign((size_t)S_nei * S_ref, 0);
  const float sq_thr = dist_threshold * dist_threshold;
  KdTree tree;
  if (use_kdtree) tree.Build(ref_world, n_ref, 4);
  KnnResult res; std::map<int, int> cnt;
  for (int idx = 0; idx < n_nei; ++idx) {
    if (!SegmentCounts(use_kdtree ? &tree : nullptr, ref_world, n_ref, ref_p2s_off, ref_p2s_ids, nei_world + (size_t)idx * 4, sq_thr, res, cnt)) continue;
    for (auto& kv : cnt) {
      if (kv.second < 5 - 2) continue;                         // :428
      for (int e = nei_p2s_off[idx]; e < nei_p2s_off[idx + 1]; ++e) M[(size_t)nei_p2s_ids[e] * S_ref + kv.first] += 1;   // :432-433
    }
  }
}

// ---- Equirectangular (sensors/Equirectangular.h) ------------------------------------------------
struct Equirect {
  int rows, cols;
  template <typename T> void CamToSphere(const T p[3], T s[2]) const {  // :41-72 (USE_FAST_ATAN2 is defined, :6)
    s[0] = FastAtan2(p[0], p[2]);
    s[1] = -FastAtan2(p[1], (T)std::sqrt(Square(p[0]) + Square(p[2])));
  }
  template <typename T> void SphereToImage(const T s[2], T px[2]) const {  // :80-96
    px[0] = cols * (0.5 + s[0] / (2.0 * M_PI));
    px[1] = rows * (0.5 - s[1] / M_PI);
  }
  template <typename T> void ImageToSphere(const T px[2], T s[2]) const {  // :98-114
    s[0] = (2 * px[0] / cols - 1) * M_PI;
    s[1] = (0.5 - px[1] / rows) * M_PI;
  }
  template <typename T> void SphereToCam(const T s[2], T r, T cam[3]) const {  // :116-146
    T cy = std::cos(s[1]);
    cam[0] = r * cy * std::sin(s[0]);
    cam[1] = -r * std::sin(s[1]);
    cam[2] = r * cy * std::cos(s[0]);
  }
  template <typename T> void ImageToCam(const T px[2], T r, T cam[3]) const { T s[2]; ImageToSphere(px, s); SphereToCam(s, r, cam); }  // :148-170
  template <typename T> void CamToImage(const T cam[3], T px[2]) const { T s[2]; CamToSphere(cam, s); SphereToImage(s, px); }      // :172-182
  bool IsInsideI(int x, int y) const { return x >= 0 && y >= 0 && x + 1 <= cols && y + 1 <= rows; }                               // :184-187

  // BreakToSegments (Equirectangular.cpp:20-58)
  std::vector<std::pair<float, float>> BreakToSegments(const float start[2], const float end[2], float seg_length) const {
    float p1[3], p2[3];
    ImageToCam(start, 5.0f, p1); ImageToCam(end, 5.0f, p2);
    const float sl[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    const float length = std::sqrt((start[0] - end[0]) * (start[0] - end[0]) + (start[1] - end[1]) * (start[1] - end[1]));
    const int count = length / seg_length + 1;
    std::vector<std::pair<float, float>> seg = {{start[0], start[1]}};
    for (int i = 1; i < count; ++i) {
      const float f = i * 1.f / count;
      float p[3] = {p1[0] + f * sl[0], p1[1] + f * sl[1], p1[2] + f * sl[2]};
      float px[2]; CamToImage(p, px);
      if (std::abs(px[0] - seg.back().first) > 0.8 * cols) {
        const float g = p1[0] / (p1[0] - p2[0]);
        float pb[3] = {p1[0] + g * sl[0], p1[1] + g * sl[1], p1[2] + g * sl[2]};
        float left[2]; CamToImage(pb, left); left[0] = 0;
        const std::pair<float, float> L{left[0], left[1]}, Rr{float(cols - 1), left[1]};
        if (px[0] > seg.back().first) { seg.push_back(L); seg.push_back(Rr); }
        else { seg.push_back(Rr); seg.push_back(L); }
      }
      seg.push_back({px[0], px[1]});
    }
    seg.push_back({end[0], end[1]});
    return seg;
  }
};

// ---- CameraLidarLineAssociate::Filter (CameraLidarLineAssociate.cpp:628-715), both branches, on pairs given in the CAMERA frame -------------
// keep[i] = the pair survives; angle[i] receives the plane angle in DEGREES when filter_by_angle is set (:652), else it is left untouched.
inline void FilterLinePairs(const Equirect& eq, int n, const float* image_line4, const double* start3, const double* end3, bool filter_by_angle, bool filter_by_length,
                            unsigned char* keep, float* angle) {
  const double zero[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) {
    keep[i] = 0;
    const double* ls = start3 + 3 * i; const double* le = end3 + 3 * i;
    if (filter_by_angle) {
      double plane_lidar[4];
      FormPlane3(ls, le, zero, plane_lidar);                                                         // :641
      { const double nn = std::sqrt(Square(plane_lidar[0]) + Square(plane_lidar[1]) + Square(plane_lidar[2]) + Square(plane_lidar[3])); for (double& v : plane_lidar) v /= nn; }
      const double px1[2] = {image_line4[4 * i], image_line4[4 * i + 1]}, px2[2] = {image_line4[4 * i + 2], image_line4[4 * i + 3]};
      double p1[3], p2[3];
      eq.ImageToCam(px1, 1.0, p1); eq.ImageToCam(px2, 1.0, p2);                                      // :645-646
      double plane_img[4];
      FormPlane3(p1, p2, zero, plane_img);
      { const double nn = std::sqrt(Square(plane_img[0]) + Square(plane_img[1]) + Square(plane_img[2]) + Square(plane_img[3])); for (double& v : plane_img) v /= nn; }
      const double plane_angle = PlaneAngle(plane_lidar, plane_img, true) * 180.0 / M_PI;           // :650
      if (plane_angle > 5) continue;
      angle[i] = (float)plane_angle;                                                                 // :652
      const double image_line_angle = VectorAngle3D(p1, p2) / 2.0;
      double sp[3], ep[3];
      ProjectPointToPlane(ls, plane_img, sp, true);
      ProjectPointToPlane(le, plane_img, ep, true);
      const double mid[3] = {(p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0, (p1[2] + p2[2]) / 2.0};
      if (VectorAngle3D(sp, mid) > image_line_angle) continue;                                       // :660
      if (VectorAngle3D(ep, mid) > image_line_angle) continue;
      const double ns = std::sqrt(Square(ls[0]) + Square(ls[1]) + Square(ls[2])), ne = std::sqrt(Square(le[0]) + Square(le[1]) + Square(le[2]));
      const double a5[3] = {ls[0] / ns * 5, ls[1] / ns * 5, ls[2] / ns * 5}, b5[3] = {le[0] / ne * 5, le[1] / ne * 5, le[2] / ne * 5};   // :665-666
      const float distance = (float)std::min(PointToPlaneDistance(plane_img, a5, true), PointToPlaneDistance(plane_img, b5, true));
      if (distance > 0.4) continue;                                                                  // :669
    }
    if (filter_by_length) {                                                                          // :673-692
      const float a[3] = {(float)ls[0], (float)ls[1], (float)ls[2]}, b[3] = {(float)le[0], (float)le[1], (float)le[2]};
      float pa[2], pb[2]; eq.CamToImage(a, pa); eq.CamToImage(b, pb);
      auto seg = eq.BreakToSegments(pa, pb, 100);
      float len = 0;
      for (size_t k = 0; k + 1 < seg.size(); ++k) {
        if (std::abs(seg[k].first - seg[k + 1].first) > 0.8 * eq.cols) continue;
        const float dx = seg[k].first - seg[k + 1].first, dy = seg[k].second - seg[k + 1].second;
        len += std::sqrt(dx * dx + dy * dy);
      }
      if (len < 100.f || len > 2000.f) continue;
    }
    keep[i] = 1;
  }
}

// ---- calibration mode: the residual blocks of CameraLidarOptimizer::Optimize(line_pairs, T_cl) (CameraLidarOptimizer.cpp:32-64) -------------
// Per line pair two blocks on the ONE relative pose (aa_cl, t_cl): Plane2Plane_Relative (HuberLoss(2 deg)) and PlaneRelativeIOUResidual (no loss,
// weight 2).  The image line's end points go through ImageToCam(cv::Point2f, 5.f) in float and the plane normal is a float cross product (:50-56).
// consts layout = the oracle's block types 6 / 7 (pvo_solver.hpp); out arrays hold 2 * n rows.
inline void BuildCalibrationBlocks(const Equirect& eq, int n, const float* image_line4, const double* start3, const double* end3, int* type, double* huber, double* consts) {
  for (int i = 0; i < n; ++i) {
    const float px1[2] = {image_line4[4 * i], image_line4[4 * i + 1]}, px2[2] = {image_line4[4 * i + 2], image_line4[4 * i + 3]};
    float p1[3], p2[3];
    eq.ImageToCam(px1, 5.0f, p1); eq.ImageToCam(px2, 5.0f, p2);                                 // :50-51
    const float p3[3] = {0, 0, 0};
    const double a = ((p2[1] - p1[1]) * (p3[2] - p1[2]) - (p2[2] - p1[2]) * (p3[1] - p1[1]));      // :54-56, float arithmetic
    const double b = ((p2[2] - p1[2]) * (p3[0] - p1[0]) - (p2[0] - p1[0]) * (p3[2] - p1[2]));
    const double c = ((p2[0] - p1[0]) * (p3[1] - p1[1]) - (p2[1] - p1[1]) * (p3[0] - p1[0]));
    const double nn = std::sqrt(a * a + b * b + c * c);
    const double* ls = start3 + 3 * i; const double* le = end3 + 3 * i;
    double* c1 = consts + 24 * i; double* c2 = c1 + 12;
    for (int k = 0; k < 24; ++k) c1[k] = 0.0;
    // Plane2Plane_Relative::Create(Vector3d(a,b,c), lidar_line_end, lidar_line_start): ctor normalises the plane (CostFunction.h:305), weight 1
    c1[0] = a / nn; c1[1] = b / nn; c1[2] = c / nn;
    for (int k = 0; k < 3; ++k) { c1[3 + k] = le[k]; c1[6 + k] = ls[k]; }
    c1[9] = 1.0;
    type[2 * i] = 6; huber[2 * i] = 2.0 * M_PI / 180.0;                                           // :36, :59
    // PlaneRelativeIOUResidual::Create(Vector4d(a,b,c,0), (start + end)/2, p1, p2, 2): plane / |n|, angle and middle in float (CostFunction.h:524-529)
    c2[0] = a / nn; c2[1] = b / nn; c2[2] = c / nn; c2[3] = 0.0 / nn;
    for (int k = 0; k < 3; ++k) c2[4 + k] = (ls[k] + le[k]) / 2.0;
    float dot = p1[0] * p2[0] + p1[1] * p2[1] + p1[2] * p2[2];
    const float n1 = std::sqrt(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]), n2 = std::sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
    dot /= (n1 * n2);
    const float ang = dot >= 1.0f ? 0.0f : (dot <= -1.0f ? (float)M_PI : std::acos(dot));
    c2[10] = (double)(ang / 2.f);
    for (int k = 0; k < 3; ++k) c2[7 + k] = (double)((p1[k] + p2[k]) / 2.f);
    c2[11] = 2.0;
    type[2 * i + 1] = 7; huber[2 * i + 1] = 0.0;                                                   // :63 loss == nullptr
  }
}

// ---- pixel-space Associate, first stage (CameraLidarLineAssociate.cpp:22-91): the fallback for frames without LiDAR segments ----------
// image lines -> sub-line mid points (BreakToSegments(line, 70), seam pieces skipped, :38-54); every LiDAR point -> camera frame
// (pcl::transformPointCloud, float32) -> pixel (CamToImage, float + FastAtan2, :75-76) -> its 3 nearest mid points (cv::flann exact search,
// float squared L2, :78), kept within 60 px (:81).  Output per point: image line of the k-th nearest mid point or -1, and the line -> LiDAR
// point lists (`line_lidar`, :83) with lines holding fewer than `min_points` (6, :92) emptied.  The RANSAC line fit that follows (:105-110,
// PCL's randomised SACSegmentation) has no deterministic answer and is not restated.
inline void PixelSubLines(const Equirect& eq, const float* lines, int L, std::vector<float>& mid, std::vector<int>& sub_to_line) {
  mid.clear(); sub_to_line.clear();
  for (int l = 0; l < L; ++l) {
    const float a[2] = {lines[l * 4], lines[l * 4 + 1]}, b[2] = {lines[l * 4 + 2], lines[l * 4 + 3]};
    const auto seg = eq.BreakToSegments(a, b, 70);
    for (size_t i = 0; i + 1 < seg.size(); ++i) {
      if (std::abs(seg[i].first - seg[i + 1].first) > 0.8 * eq.cols) continue;                        // :44
      mid.push_back((float)((seg[i + 1].first + seg[i].first) / 2.0));                                // :47-48
      mid.push_back((float)((seg[i + 1].second + seg[i].second) / 2.0));
      sub_to_line.push_back(l);
    }
  }
}

inline void PixelLineNeighbors(const Equirect& eq, const float* lines, int L, const float* cloud_local, int P, const double T_cl[16], int* line3, float* d2_3,
                               float* pixel2) {
  std::vector<float> mid; std::vector<int> s2l;
  PixelSubLines(eq, lines, L, mid, s2l);
  const int M = (int)s2l.size();
  std::vector<float> cam((size_t)std::max(P, 1) * 4);
  double R[9], t[3];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = T_cl[r * 4 + c]; t[r] = T_cl[r * 4 + 3]; }
  TransformCloud(R, t, cloud_local, P, cam.data(), 4);
  for (int i = 0; i < P; ++i) {
    const float pc[3] = {cam[(size_t)i * 4], cam[(size_t)i * 4 + 1], cam[(size_t)i * 4 + 2]};
    float px[2]; eq.CamToImage(pc, px);
    if (pixel2) { pixel2[2 * i] = px[0]; pixel2[2 * i + 1] = px[1]; }
    float bd[3] = {INFINITY, INFINITY, INFINITY}; int bi[3] = {-1, -1, -1};
    for (int m = 0; m < M; ++m) {
      const float dx = px[0] - mid[2 * m], dy = px[1] - mid[2 * m + 1];
      const float d = dx * dx + dy * dy;                                                                // FLANN L2<float>
      if (d < bd[2]) {
        int k = 2;
        while (k > 0 && d < bd[k - 1]) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; --k; }
        bd[k] = d; bi[k] = m;
      }
    }
    for (int k = 0; k < 3; ++k) {
      const bool keep = bi[k] >= 0 && !(bd[k] > 60 * 60);                                               // :81
      line3[3 * i + k] = keep ? s2l[bi[k]] : -1;
      if (d2_3) d2_3[3 * i + k] = bd[k];
    }
  }
}

// ---- AssociateByAngle (CameraLidarLineAssociate.cpp:340-475) + Filter(false,true) (:628-715) ----------
struct CamLidarPair { int image_line, lidar_line; double start[3], end[3]; float angle; };

inline void Transform4(const double T[16], const double p[3], double out[3]) {  // (T * p.homogeneous()).hnormalized(), affine
  for (int r = 0; r < 3; ++r) out[r] = T[r * 4] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3];
}

// Per (image line, LiDAR segment) vote counts of AssociateByAngle's inner loop (:389-414); exposed so
// the device counters can be compared directly.  counts is L x S.
inline void AngleVotes(const Equirect& eq, const float* lines /*L x 4*/, int L, const float* cloud_local /*P x 4*/, int P,
                       const int* p2s_off, const int* p2s_ids, int S, const double T_cl[16], std::vector<int>& counts) {
  counts.assign((size_t)L * S, 0);
  std::vector<float> range(P), cam((size_t)P * 4);
  for (int i = 0; i < P; ++i) { const float* p = cloud_local + i * 4; range[i] = p[0] * p[0] + p[1] * p[1] + p[2] * p[2]; }
  double R[9], t[3];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = T_cl[r * 4 + c]; t[r] = T_cl[r * 4 + 3]; }
  TransformCloud(R, t, cloud_local, P, cam.data(), 4);
  const double thr = 3.0 / 180.0 * M_PI;
  for (int l = 0; l < L; ++l) {
    const double px1[2] = {lines[l * 4], lines[l * 4 + 1]}, px2[2] = {lines[l * 4 + 2], lines[l * 4 + 3]};
    double p1[3], p2[3]; eq.ImageToCam(px1, 1.0, p1); eq.ImageToCam(px2, 1.0, p2);
    const double zero[3] = {0, 0, 0}; double plane[4];
    FormPlane3(p1, p2, zero, plane);
    const double nn = std::sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2] + plane[3] * plane[3]);
    for (int c = 0; c < 4; ++c) plane[c] /= nn;
    const double p4[3] = {(p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0, (p1[2] + p2[2]) / 2.0};
    const double scope = VectorAngle3D(p1, p4);
    for (int i = 0; i < P; ++i) {
      if (range[i] > 15 * 15) continue;
      const double p[3] = {cam[i * 4], cam[i * 4 + 1], cam[i * 4 + 2]};
      double pp[3]; ProjectPointToPlane(p, plane, pp, true);
      if (VectorAngle3D(p, pp) >= thr) continue;
      if (VectorAngle3D(p4, pp) >= scope + thr) continue;
      for (int e = p2s_off[i]; e < p2s_off[i + 1]; ++e) counts[(size_t)l * S + p2s_ids[e]]++;
    }
  }
}

// UniqueLinePair (CameraLidarLineAssociate.cpp:754-876): one-to-one pairs, the smaller score (angle error) wins; processed in
// input order, output ascending by image line.  The reference evaluates its cases 1-3 as consecutive `if`s on iterators that case 1
// has just erased; with the strict inequalities those later conditions can never hold after case 1, so they are `else if` here.
inline void UniqueLinePair(std::vector<CamLidarPair>& pairs) {
  struct PS { int idx; float score; };
  std::map<int, PS> i2l, l2i;
  for (const CamLidarPair& pr : pairs) {
    const int il = pr.image_line, ll = pr.lidar_line; const float sc = pr.angle;
    auto a = i2l.find(il); auto b = l2i.find(ll);
    const bool ha = a != i2l.end(), hb = b != l2i.end();
    if (!ha && !hb) { i2l.insert({il, PS{ll, sc}}); l2i.insert({ll, PS{il, sc}}); }
    else if (ha && !hb) {
      if (sc < a->second.score) { l2i.erase(l2i.find(a->second.idx)); a->second = PS{ll, sc}; l2i.insert({ll, PS{il, sc}}); }
    } else if (!ha && hb) {
      if (sc < b->second.score) { i2l.erase(i2l.find(b->second.idx)); b->second = PS{il, sc}; i2l.insert({il, PS{ll, sc}}); }
    } else {
      if (sc < std::min(a->second.score, b->second.score)) {                      // case 1
        i2l.erase(b->second.idx); l2i.erase(a->second.idx); i2l.erase(a); l2i.erase(b);
        i2l.insert({il, PS{ll, sc}}); l2i.insert({ll, PS{il, sc}});
      } else if (sc > a->second.score && sc < b->second.score) {                  // case 2
        i2l.erase(i2l.find(b->second.idx)); l2i.erase(b);
      } else if (sc < a->second.score && sc > b->second.score) {                  // case 3
        l2i.erase(l2i.find(a->second.idx)); i2l.erase(a);
      }
    }
  }
  std::vector<CamLidarPair> out;
  for (auto& kv : i2l) { CamLidarPair p{}; p.image_line = kv.first; p.lidar_line = kv.second.idx; p.angle = kv.second.score; out.push_back(p); }
  pairs.swap(out);
}

inline void AssociateByAngle(const Equirect& eq, const float* lines, int L, const float* cloud_local, int P,
                             const int* p2s_off, const int* p2s_ids, int S, const int* seg_sizes,
                             const double* end_points /*S x 2 x 3, lidar frame*/, const double T_cl[16],
                             bool filter_by_length, std::vector<CamLidarPair>& out, bool multiple_association = true,
                             const unsigned char* image_mask = nullptr, const unsigned char* lidar_mask = nullptr) {
  std::vector<int> counts;
  AngleVotes(eq, lines, L, cloud_local, P, p2s_off, p2s_ids, S, T_cl, counts);
  const double thr = 3.0 / 180.0 * M_PI;
  std::vector<double> ep((size_t)S * 6), lplane((size_t)S * 4);
  const double zero[3] = {0, 0, 0};
  for (int s = 0; s < S; ++s) {
    Transform4(T_cl, end_points + s * 6, &ep[s * 6]);
    Transform4(T_cl, end_points + s * 6 + 3, &ep[s * 6 + 3]);
    double pl[4]; FormPlane3(&ep[s * 6], &ep[s * 6 + 3], zero, pl);
    const double nn = std::sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2] + pl[3] * pl[3]);
    for (int c = 0; c < 4; ++c) lplane[s * 4 + c] = pl[c] / nn;
  }
  std::vector<CamLidarPair> pairs;
  for (int l = 0; l < L; ++l) {
    if (image_mask && !image_mask[l]) continue;               // :396
    const double px1[2] = {lines[l * 4], lines[l * 4 + 1]}, px2[2] = {lines[l * 4 + 2], lines[l * 4 + 3]};
    double p1[3], p2[3]; eq.ImageToCam(px1, 1.0, p1); eq.ImageToCam(px2, 1.0, p2);
    double plane[4]; FormPlane3(p1, p2, zero, plane);
    const double nn = std::sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2] + plane[3] * plane[3]);
    for (int c = 0; c < 4; ++c) plane[c] /= nn;
    const double p4[3] = {(p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0, (p1[2] + p2[2]) / 2.0};
    const double scope = VectorAngle3D(p1, p4);
    for (int s = 0; s < S; ++s) {  // std::map iteration = ascending segment id, only segments with votes
      const int cnt = counts[(size_t)l * S + s];
      if (cnt == 0) continue;
      if ((size_t)cnt < (size_t)seg_sizes[s] / 2) continue;
      if (lidar_mask && !lidar_mask[s]) continue;             // :420
      const double angle = PlaneAngle(plane, &lplane[s * 4], true);
      if (angle > thr) continue;
      double mid[3], midp[3];
      for (int c = 0; c < 3; ++c) mid[c] = (ep[s * 6 + c] + ep[s * 6 + 3 + c]) / 2.f;
      ProjectPointToPlane(mid, plane, midp, true);
      if (VectorAngle3D(midp, p4) > scope) continue;
      const float angle2 = VectorAngle3D(mid, midp);
      if (angle2 > thr / 2.0) continue;
      CamLidarPair pr; pr.image_line = l; pr.lidar_line = s; pr.angle = angle + angle2;
      for (int c = 0; c < 3; ++c) { pr.start[c] = ep[s * 6 + c]; pr.end[c] = ep[s * 6 + 3 + c]; }
      pairs.push_back(pr);
    }
  }
  // Filter(false, true): projected LiDAR line length in [100, 2000] px (:676-692)
  for (const CamLidarPair& p : pairs) {
    if (filter_by_length) {
      const float a[3] = {(float)p.start[0], (float)p.start[1], (float)p.start[2]};
      const float b[3] = {(float)p.end[0], (float)p.end[1], (float)p.end[2]};
      float pa[2], pb[2]; eq.CamToImage(a, pa); eq.CamToImage(b, pb);
      auto seg = eq.BreakToSegments(pa, pb, 100);
      float len = 0;
      for (size_t i = 0; i + 1 < seg.size(); ++i) {
        if (std::abs(seg[i].first - seg[i + 1].first) > 0.8 * eq.cols) continue;
        const float dx = seg[i].first - seg[i + 1].first, dy = seg[i].second - seg[i + 1].second;
        len += std::sqrt(dx * dx + dy * dy);
      }
      if (len < 100.f || len > 2000.f) continue;
    }
    out.push_back(p);
  }
  if (!multiple_association) {                                // :465-466; pairs are rebuilt from the unfiltered camera-frame end points (:866-875)
    UniqueLinePair(out);
    for (CamLidarPair& p : out) for (int c = 0; c < 3; ++c) { p.start[c] = ep[p.lidar_line * 6 + c]; p.end[c] = ep[p.lidar_line * 6 + 3 + c]; }
  }
  // back to the LiDAR frame (:469-474): T_lc = T_cl^-1 (rigid)
  double Tlc[16] = {0};
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Tlc[r * 4 + c] = T_cl[c * 4 + r];
  for (int r = 0; r < 3; ++r) Tlc[r * 4 + 3] = -(Tlc[r * 4] * T_cl[3] + Tlc[r * 4 + 1] * T_cl[7] + Tlc[r * 4 + 2] * T_cl[11]);
  Tlc[15] = 1;
  for (CamLidarPair& p : out) { double s[3], e[3]; Transform4(Tlc, p.start, s); Transform4(Tlc, p.end, e); for (int c = 0; c < 3; ++c) { p.start[c] = s[c]; p.end[c] = e[c]; } }
}

// ---- ProjectLidar2PanoramaDepth (util/Visualization.h:408-441) ----------------------------------
// uvd (optional, n x 3): per-point pixel.x, pixel.y (float, FastAtan2 path) and depth (float).
inline void ProjectLidar2PanoramaDepth(const float* cloud, int n, int stride, int rows, int cols, const double T_cl[16],
                                       int size, uint16_t* img /*rows*cols, zeroed by caller*/, float* uvd) {
  Equirect eq{rows, cols};
  double R[9], t[3];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r * 3 + c] = T_cl[r * 4 + c]; t[r] = T_cl[r * 4 + 3]; }
  for (int i = 0; i < n; ++i) {
    float p[4] = {0, 0, 0, 0}, in4[4] = {cloud[(size_t)i * stride], cloud[(size_t)i * stride + 1], cloud[(size_t)i * stride + 2], 0};
    TransformCloud(R, t, in4, 1, p, 4);
    float s[2], px[2];
    eq.CamToSphere(p, s); eq.SphereToImage(s, px);
    const float depth = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if (uvd) { uvd[(size_t)i * 3] = px[0]; uvd[(size_t)i * 3 + 1] = px[1]; uvd[(size_t)i * 3 + 2] = depth; }
    if (!img) continue;
    const int rbx = (int)std::ceil(px[0]) + size / 2, rby = (int)std::ceil(px[1]) + size / 2;
    const int ltx = (int)std::floor(px[0]) - size / 2, lty = (int)std::floor(px[1]) - size / 2;
    if (!eq.IsInsideI(rbx, rby) || !eq.IsInsideI(ltx, lty)) continue;
    const uint16_t rel = static_cast<uint16_t>(depth * 256.0);
    for (int u = lty; u <= rby; ++u) for (int v = ltx; v <= rbx; ++v) img[(size_t)u * cols + v] = rel;
  }
}

}  // namespace pvo
