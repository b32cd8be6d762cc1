// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  Stand-in for the PCL 1.10 types that the reference's sensors/Velodyne.h (class declaration only) and
// lidar_mapping/LidarFeatureAssociate.{h,cpp} use: point structs, pcl::PointCloud, and pcl::KdTreeFLANN as an EXACT search - what FLANN's
// KDTreeSingleIndex returns with epsilon 0: the k nearest points by squared L2 distance accumulated in float32 over (x, y, z), ascending, ties by index.
// The kd-tree itself is not reproduced (an exact search has one answer up to ties).  Everything else named by those headers (VoxelGrid, IO, ICP,
// RANSAC, normals, region growing) is declared only as far as the compiler needs to see a type; none of it is called on the association path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; PointXYZI() {} explicit PointXYZI(float i) : intensity(i) {} };
struct PointXYZRGB { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0; };
struct PointXYZRGBL { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0; unsigned int label = 0; };
struct Normal { float normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct ModelCoefficients { std::vector<float> values; };
struct PointIndices { typedef std::shared_ptr<PointIndices> Ptr; std::vector<int> indices; };

template <typename P> struct PointCloud {
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
  typedef typename std::vector<P>::iterator iterator;
  typedef typename std::vector<P>::const_iterator const_iterator;
  std::vector<P> points; unsigned width = 0, height = 1; bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; }
  void resize(size_t n) { points.resize(n); width = (unsigned)n; }
  void reserve(size_t n) { points.reserve(n); }
  void push_back(const P& p) { points.push_back(p); width = (unsigned)points.size(); }
  template <typename... A> void emplace_back(A&&... a) { points.emplace_back(std::forward<A>(a)...); width = (unsigned)points.size(); }
  void swap(PointCloud& o) { points.swap(o.points); std::swap(width, o.width); std::swap(height, o.height); std::swap(is_dense, o.is_dense); }
  P& operator[](size_t i) { return points[i]; }
  const P& operator[](size_t i) const { return points[i]; }
  P& at(size_t i) { return points.at(i); }
  const P& at(size_t i) const { return points.at(i); }
  iterator begin() { return points.begin(); } iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); } const_iterator end() const { return points.end(); }
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (unsigned)points.size(); return *this; }
  Ptr makeShared() const { return Ptr(new PointCloud<P>(*this)); }
};

template <typename P> class KdTreeFLANN {
  typename PointCloud<P>::ConstPtr cloud_;
 public:
  typedef std::shared_ptr<KdTreeFLANN<P>> Ptr;
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { cloud_ = c; }
  static float dist2(const P& a, const P& b) { float r = 0; float d = a.x - b.x; r += d * d; d = a.y - b.y; r += d * d; d = a.z - b.z; r += d * d; return r; }
  int nearestKSearch(const P& q, int k, std::vector<int>& idx, std::vector<float>& sqd) const {
    const int n = (int)cloud_->points.size();
    std::vector<std::pair<float, int>> all(n);
    for (int i = 0; i < n; ++i) all[i] = std::make_pair(dist2(cloud_->points[i], q), i);
    const int kk = std::min(k, n);
    std::partial_sort(all.begin(), all.begin() + kk, all.end());
    idx.resize(kk); sqd.resize(kk);
    for (int i = 0; i < kk; ++i) { idx[i] = all[i].second; sqd[i] = all[i].first; }
    return kk;
  }
  int radiusSearch(const P& q, double radius, std::vector<int>& idx, std::vector<float>& sqd, unsigned max_nn = 0) const {
    const int n = (int)cloud_->points.size();
    const float r2 = (float)(radius * radius);
    std::vector<std::pair<float, int>> in;
    for (int i = 0; i < n; ++i) { const float d = dist2(cloud_->points[i], q); if (d <= r2) in.push_back(std::make_pair(d, i)); }
    std::sort(in.begin(), in.end());
    if (max_nn && in.size() > max_nn) in.resize(max_nn);
    idx.resize(in.size()); sqd.resize(in.size());
    for (size_t i = 0; i < in.size(); ++i) { idx[i] = in[i].second; sqd[i] = in[i].first; }
    return (int)in.size();
  }
};
// pcl::VoxelGrid / RANSAC plane model / NaN removal: named by the feature-extraction code of sensors/Velodyne.cpp, which is NOT on the compiled path and is not
// reproduced: filter() copies its input, the RANSAC finds nothing.
template <typename P> class VoxelGrid {
  typename PointCloud<P>::ConstPtr in_;
 public:
  void setLeafSize(float, float, float) {}
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) { in_ = c; }
  void filter(PointCloud<P>& out) { if (in_) out = *in_; }
};
template <typename P> class SampleConsensusModelPlane {
 public:
  typedef std::shared_ptr<SampleConsensusModelPlane<P>> Ptr;
  explicit SampleConsensusModelPlane(const typename PointCloud<P>::ConstPtr&) {}
};
template <typename P> class RandomSampleConsensus {
 public:
  explicit RandomSampleConsensus(const typename SampleConsensusModelPlane<P>::Ptr&) {}
  void setDistanceThreshold(double) {}
  bool computeModel() { return false; }
  void getInliers(std::vector<int>& v) { v.clear(); }
  template <typename V> void getModelCoefficients(V&) {}
};
template <typename P> inline void removeNaNFromPointCloud(const PointCloud<P>& in, PointCloud<P>& out, std::vector<int>& index) {
  PointCloud<P> tmp; index.clear();
  for (size_t i = 0; i < in.points.size(); ++i) { const P& p = in.points[i]; if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) { tmp.push_back(p); index.push_back((int)i); } }
  out = tmp;
}
// pcl::transformPointCloud(in, out, Matrix4): PCL 1.10 computes x' = T(0,0) x + T(0,1) y + T(0,2) z + T(0,3) in the MATRIX's scalar type and stores float32;
// the other fields are copied.  M is any 4x4 type with operator()(i, j).
template <typename P, typename M> inline void transformPointCloud(const PointCloud<P>& in, PointCloud<P>& out, const M& T) {
  PointCloud<P> tmp = in;
  for (size_t i = 0; i < in.points.size(); ++i) {
    const P& p = in.points[i];
    tmp.points[i].x = static_cast<float>(T(0, 0) * p.x + T(0, 1) * p.y + T(0, 2) * p.z + T(0, 3));
    tmp.points[i].y = static_cast<float>(T(1, 0) * p.x + T(1, 1) * p.y + T(1, 2) * p.z + T(1, 3));
    tmp.points[i].z = static_cast<float>(T(2, 0) * p.x + T(2, 1) * p.y + T(2, 2) * p.z + T(2, 3));
  }
  out = tmp;
}
// pcl::SACSegmentation (RANSAC line fit behind CameraLidarLineAssociate::FitLineRANSAC, the fallback for frames without LiDAR segments): NOT reproduced -
// its sample sequence depends on PCL's internal random generator.  segment() reports no inliers (FitLineRANSAC returns false) unless a harness scripts them.
enum { SACMODEL_LINE = 1, SAC_RANSAC = 0 };
// A test harness may ask the stand-in to RECORD the clouds it is handed (the candidate points of every image line = the result of the first, deterministic
// stage of CameraLidarLineAssociate::Associate): sac_recorder() points at a vector of (x, y, z) lists, or is null.
inline std::vector<std::vector<float>>*& sac_recorder() { static std::vector<std::vector<float>>* r = nullptr; return r; }
// sac_script(): (inlier lists in call order, next call) or null - see SACSegmentation::segment
inline std::pair<std::vector<std::vector<int>>, size_t>*& sac_script() { static std::pair<std::vector<std::vector<int>>, size_t>* r = nullptr; return r; }
template <typename P> class SACSegmentation {
 public:
  void setOptimizeCoefficients(bool) {} void setModelType(int) {} void setMethodType(int) {} void setDistanceThreshold(double) {}
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) {
    if (!sac_recorder()) return;
    std::vector<float> xyz; for (const P& p : c->points) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
    sac_recorder()->push_back(xyz);
  }
  // Scripted mode (sac_script() set by a harness): the k-th call reports the k-th scripted inlier list and a 6-value coefficient vector (the caller overwrites
  // it), so that the REFERENCE'S OWN code after the RANSAC runs (FitLineRANSAC :729-751, Associate :117-187).  Otherwise: no inliers.
  void segment(PointIndices& inliers, ModelCoefficients& coeff) {
    inliers.indices.clear();
    if (!sac_script()) return;
    size_t& k = sac_script()->second;
    if (k < sac_script()->first.size()) inliers.indices = sac_script()->first[k];
    ++k;
    coeff.values.assign(6, 0.f);
  }
};
// stand-ins for pcl/common/centroid.hpp and pcl/common/eigen.hpp (PCL 1.10's published float32 algorithms restated; only operator[] / operator()(i, j) of the
// vector / matrix types are used).  They serve the scripted mode above; PCL itself is not available here.
template <typename P, typename V> inline unsigned compute3DCentroid(const PointCloud<P>& cloud, const PointIndices& ind, V& c) {
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i : ind.indices) { sx += cloud.points[i].x; sy += cloud.points[i].y; sz += cloud.points[i].z; }
  const float n = static_cast<float>(ind.indices.size());
  c[0] = sx / n; c[1] = sy / n; c[2] = sz / n; c[3] = 1.f;
  return static_cast<unsigned>(ind.indices.size());
}
template <typename P, typename V, typename M> inline unsigned computeCovarianceMatrix(const PointCloud<P>& cloud, const PointIndices& ind, const V& c, M& cov) {
  float m00 = 0.f, m01 = 0.f, m02 = 0.f, m11 = 0.f, m12 = 0.f, m22 = 0.f;
  for (int i : ind.indices) {
    float x = cloud.points[i].x - c[0], y = cloud.points[i].y - c[1], z = cloud.points[i].z - c[2];
    m11 += y * y; m12 += y * z; m22 += z * z;
    y *= x; z *= x; x *= x;
    m00 += x; m01 += y; m02 += z;
  }
  cov(0, 0) = m00; cov(0, 1) = m01; cov(0, 2) = m02; cov(1, 0) = m01; cov(1, 1) = m11; cov(1, 2) = m12; cov(2, 0) = m02; cov(2, 1) = m12; cov(2, 2) = m22;
  return static_cast<unsigned>(ind.indices.size());
}
namespace shim_detail {
inline void roots2(float b, float c, float* r) { r[0] = 0.f; float d = b * b - 4.0f * c; if (d < 0.f) d = 0.f; const float sd = std::sqrt(d); r[2] = 0.5f * (b + sd); r[1] = 0.5f * (b - sd); }
inline void roots3(const float m[3][3], float* r) {
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) { roots2(c2, c1, r); return; }
  const float inv3 = 1.0f / 3.0f, sqrt3 = std::sqrt(3.0f), c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3; if (a_3 > 0.f) a_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_3 * (2.f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3; if (q > 0.f) q = 0.f;
  const float rho = std::sqrt(-a_3), theta = std::atan2(std::sqrt(-q), half_b) * inv3, ct = std::cos(theta), st = std::sin(theta);
  r[0] = c2_3 + 2.f * rho * ct; r[1] = c2_3 - rho * (ct + sqrt3 * st); r[2] = c2_3 - rho * (ct - sqrt3 * st);
  if (r[0] >= r[1]) std::swap(r[0], r[1]);
  if (r[1] >= r[2]) { std::swap(r[1], r[2]); if (r[0] >= r[1]) std::swap(r[0], r[1]); }
  if (r[0] <= 0.f) roots2(c2, c1, r);
}
template <typename M> inline float scaled(const M& mat, float sm[3][3]) {
  float scale = 0.f;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(static_cast<float>(mat(i, j))));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.f;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sm[i][j] = static_cast<float>(mat(i, j)) / scale;
  return scale;
}
}  // namespace shim_detail
template <typename M, typename V> inline void eigen33(const M& mat, V& evals) {
  float sm[3][3], r[3];
  const float scale = shim_detail::scaled(mat, sm);
  shim_detail::roots3(sm, r);
  for (int k = 0; k < 3; ++k) evals[k] = r[k] * scale;
}
template <typename M, typename S, typename V> inline void computeCorrespondingEigenVector(const M& mat, const S& eigenvalue, V& vec) {
  float sm[3][3];
  const float scale = shim_detail::scaled(mat, sm);
  for (int i = 0; i < 3; ++i) sm[i][i] -= static_cast<float>(eigenvalue) / scale;
  const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  float c[3][3], len[3];
  for (int k = 0; k < 3; ++k) {
    const float* a = sm[pr[k][0]]; const float* b = sm[pr[k][1]];
    c[k][0] = a[1] * b[2] - a[2] * b[1]; c[k][1] = a[2] * b[0] - a[0] * b[2]; c[k][2] = a[0] * b[1] - a[1] * b[0];
    len[k] = c[k][0] * c[k][0] + c[k][1] * c[k][1] + c[k][2] * c[k][2];
  }
  const int best = (len[0] >= len[1] && len[0] >= len[2]) ? 0 : ((len[1] >= len[0] && len[1] >= len[2]) ? 1 : 2);
  const float n = std::sqrt(len[best]);
  for (int j = 0; j < 3; ++j) vec[j] = c[best][j] / n;
}
namespace io {
template <typename C> inline int savePCDFileASCII(const std::string&, const C&) { return 0; }   // debug dumps (visualization = true only)
template <typename C> inline int savePCDFileBinary(const std::string&, const C&) { return 0; }
template <typename C> inline int savePCDFile(const std::string&, const C&) { return 0; }
template <typename C> inline int loadPCDFile(const std::string&, C&) { return -1; }   // no file IO in the stand-in
template <typename C> inline int loadPLYFile(const std::string&, C&) { return -1; }
}  // namespace io
}  // namespace pcl
