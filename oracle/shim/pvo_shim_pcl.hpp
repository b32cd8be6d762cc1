// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  Stand-in for the PCL 1.10 types that the reference's sensors/Velodyne.h (class declaration only) and
// lidar_mapping/LidarFeatureAssociate.{h,cpp} use: point structs, pcl::PointCloud, and pcl::KdTreeFLANN as an EXACT search - what FLANN's
// KDTreeSingleIndex returns with epsilon 0: the k nearest points by squared L2 distance accumulated in float32 over (x, y, z), ascending, ties by index.
// The kd-tree itself is not reproduced (an exact search has one answer up to ties).  Everything else named by those headers (VoxelGrid, IO, ICP,
// RANSAC, normals, region growing) is declared only as far as the compiler needs to see a type; none of it is called on the association path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
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
// its sample sequence depends on PCL's internal random generator.  segment() reports no inliers, so FitLineRANSAC returns false.
enum { SACMODEL_LINE = 1, SAC_RANSAC = 0 };
// A test harness may ask the stand-in to RECORD the clouds it is handed (the candidate points of every image line = the result of the first, deterministic
// stage of CameraLidarLineAssociate::Associate): sac_recorder() points at a vector of (x, y, z) lists, or is null.
inline std::vector<std::vector<float>>*& sac_recorder() { static std::vector<std::vector<float>>* r = nullptr; return r; }
template <typename P> class SACSegmentation {
 public:
  void setOptimizeCoefficients(bool) {} void setModelType(int) {} void setMethodType(int) {} void setDistanceThreshold(double) {}
  void setInputCloud(const typename PointCloud<P>::ConstPtr& c) {
    if (!sac_recorder()) return;
    std::vector<float> xyz; for (const P& p : c->points) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
    sac_recorder()->push_back(xyz);
  }
  void segment(PointIndices& inliers, ModelCoefficients&) { inliers.indices.clear(); }
};
template <typename P, typename V> inline unsigned compute3DCentroid(const PointCloud<P>&, const PointIndices&, V&) { return 0; }
template <typename P, typename V, typename M> inline unsigned computeCovarianceMatrix(const PointCloud<P>&, const PointIndices&, const V&, M&) { return 0; }
template <typename M, typename V> inline void eigen33(const M&, V&) {}
template <typename M, typename S, typename V> inline void computeCorrespondingEigenVector(const M&, const S&, V&) {}
namespace io {
template <typename C> inline int savePCDFileASCII(const std::string&, const C&) { return 0; }   // debug dumps (visualization = true only)
template <typename C> inline int savePCDFileBinary(const std::string&, const C&) { return 0; }
template <typename C> inline int savePCDFile(const std::string&, const C&) { return 0; }
template <typename C> inline int loadPCDFile(const std::string&, C&) { return -1; }   // no file IO in the stand-in
template <typename C> inline int loadPLYFile(const std::string&, C&) { return -1; }
}  // namespace io
}  // namespace pcl
