// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  Stand-in for the OpenCV value types the reference's geometry / projection headers use (cv::Point_,
// cv::Point3_, cv::Vec, a cv::Mat that can only be empty or a float3 table), plus glog's LOG() (the PCL stand-ins live in pvo_shim_pcl.hpp).
// See pvo_shim_eigen.hpp for why this exists.
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#include <cassert>
typedef unsigned char uchar; typedef unsigned short ushort;   // OpenCV declares these at global scope
#define CV_GRAY2BGR 8
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_LOAD_IMAGE_COLOR 1
#define CV_LOAD_IMAGE_UNCHANGED -1
#define CV_BGR2GRAY 6
#define CV_32FC3 21
#define CV_16U 2
#define CV_16UC1 2
namespace cv {
// cv::saturate_cast: floating point -> integer rounds to nearest (cvRound = lrint, ties to even); everything else is a plain conversion
template <typename To, typename From> inline typename std::enable_if<std::is_integral<To>::value && std::is_floating_point<From>::value, To>::type saturate_cast(From v) { return (To)std::lrint(v); }
template <typename To, typename From> inline typename std::enable_if<!(std::is_integral<To>::value && std::is_floating_point<From>::value), To>::type saturate_cast(From v) { return (To)v; }
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <typename U> Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}  // NOLINT: OpenCV's implicit Point_ conversion
  T dot(const Point_& o) const { return x * o.x + y * o.y; }
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, float k) { return Point_<T>(T(a.x * k), T(a.y * k)); }
template <typename T> inline Point_<T> operator*(float k, const Point_<T>& a) { return Point_<T>(T(a.x * k), T(a.y * k)); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, double k) { return Point_<T>(T(a.x * k), T(a.y * k)); }
template <typename T> inline Point_<T> operator*(double k, const Point_<T>& a) { return Point_<T>(T(a.x * k), T(a.y * k)); }
template <typename T> inline Point_<T> operator/(const Point_<T>& a, double k) { return Point_<T>(T(a.x / k), T(a.y / k)); }
template <typename T, int n> struct Vec {
  T val[n];
  Vec() { for (int i = 0; i < n; ++i) val[i] = T(0); }
  Vec(T a, T b) { static_assert(n == 2, "n"); val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) { static_assert(n == 3, "n"); val[0] = a; val[1] = b; val[2] = c; }
  Vec(T a, T b, T c, T d) { static_assert(n == 4, "n"); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  T& operator()(int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  T dot(const Vec& o) const { T s = 0; for (int i = 0; i < n; ++i) s += val[i] * o.val[i]; return s; }
};
template <typename T, int n> inline double norm(const Vec<T, n>& v) { double s = 0; for (int i = 0; i < n; ++i) s += (double)v.val[i] * v.val[i]; return std::sqrt(s); }
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
  template <typename U> Point3_(const Point3_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)), z(saturate_cast<T>(o.z)) {}  // NOLINT
  Point3_(const Vec<T, 3>& v) : x(v[0]), y(v[1]), z(v[2]) {}  // NOLINT
  operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
  T dot(const Point3_& o) const { return x * o.x + y * o.y + z * o.z; }
  Point3_ cross(const Point3_& o) const { return Point3_(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
};
// OpenCV: a*b with a float/double/int scalar is computed in the scalar's type and saturate_cast back to T
template <typename T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a) { return Point3_<T>(-a.x, -a.y, -a.z); }
template <typename T> inline Point3_<T> operator*(const Point3_<T>& a, float k) { return Point3_<T>(T(a.x * k), T(a.y * k), T(a.z * k)); }
template <typename T> inline Point3_<T> operator*(float k, const Point3_<T>& a) { return Point3_<T>(T(a.x * k), T(a.y * k), T(a.z * k)); }
template <typename T> inline Point3_<T> operator*(const Point3_<T>& a, double k) { return Point3_<T>(T(a.x * k), T(a.y * k), T(a.z * k)); }
template <typename T> inline Point3_<T> operator*(double k, const Point3_<T>& a) { return Point3_<T>(T(a.x * k), T(a.y * k), T(a.z * k)); }
template <typename T> inline Point3_<T> operator*(const Point3_<T>& a, int k) { return Point3_<T>(T(a.x * k), T(a.y * k), T(a.z * k)); }
template <typename T> inline Point3_<T> operator*(int k, const Point3_<T>& a) { return Point3_<T>(T(a.x * k), T(a.y * k), T(a.z * k)); }
template <typename T> inline Point3_<T> operator/(const Point3_<T>& a, double k) { return Point3_<T>(T(a.x / k), T(a.y / k), T(a.z / k)); }
template <typename T> inline Point3_<T> operator/(const Point3_<T>& a, float k) { return Point3_<T>(T(a.x / k), T(a.y / k), T(a.z / k)); }
template <typename T> inline double norm(const Point3_<T>& v) { return std::sqrt((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z); }
typedef Point_<int> Point2i; typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point2i Point;
typedef Point3_<int> Point3i; typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;
typedef Vec<float, 2> Vec2f; typedef Vec<float, 3> Vec3f; typedef Vec<float, 4> Vec4f; typedef Vec<float, 6> Vec6f;
typedef Vec<double, 2> Vec2d; typedef Vec<double, 3> Vec3d; typedef Vec<double, 4> Vec4d; typedef Vec<double, 6> Vec6d;
typedef Vec<int, 2> Vec2i; typedef Vec<int, 3> Vec3i; typedef Vec<int, 4> Vec4i;
typedef Vec<uchar, 3> Vec3b;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } double& operator[](int i) { return val[i]; } const double& operator[](int i) const { return val[i]; } };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };
namespace line_descriptor { struct KeyLine { float angle = 0; int class_id = 0, octave = 0; Point2f pt; float response = 0, size = 0, startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0,
  sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0, lineLength = 0; int numOfPixels = 0; }; }
// a dense 2-D array of `type` elements (CV_8U, CV_16U, CV_32F, CV_8UC3, CV_32FC3 ...): enough for zeros / at<T> / clone / empty / convertTo-free code
#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
struct Mat {
  int rows = 0, cols = 0, type_ = 0; std::vector<unsigned char> d;
  static int elem(int type) { const int depth = type & 7, ch = (type >> 3) + 1; const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2}; return sz[depth] * ch; }
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), d((size_t)r * c * elem(type), 0) {}
  Mat(int r, int c, int type, const Scalar&) : rows(r), cols(c), type_(type), d((size_t)r * c * elem(type), 0) {}
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  bool empty() const { return d.empty(); }
  bool isContinuous() const { return true; }
  size_t elemSize() const { return (size_t)elem(type_); }
  void create(int r, int c, int type) { *this = Mat(r, c, type); }
  unsigned char* ptr(int i = 0) { return d.data() + (size_t)i * cols * elem(type_); }
  const unsigned char* ptr(int i = 0) const { return d.data() + (size_t)i * cols * elem(type_); }
  Mat clone() const { return *this; }
  void release() { d.clear(); rows = cols = 0; }
  void convertTo(Mat& out, int, double = 1, double = 0) const { out = *this; }     // image arithmetic is only named by image IO helpers: not reproduced
  Mat& operator*=(double) { return *this; } Mat& operator/=(double) { return *this; } Mat& operator-=(double) { return *this; } Mat& operator+=(double) { return *this; }
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  Size size() const { return Size(cols, rows); }
  template <typename V> V& at(int i, int j) { return *reinterpret_cast<V*>(d.data() + ((size_t)i * cols + j) * sizeof(V)); }
  template <typename V> const V& at(int i, int j) const { return *reinterpret_cast<const V*>(d.data() + ((size_t)i * cols + j) * sizeof(V)); }
  template <typename V> V& at(const Point2i& p) { return at<V>(p.y, p.x); }
  template <typename V> const V& at(const Point2i& p) const { return at<V>(p.y, p.x); }
  template <typename V> V* ptr(int i) { return reinterpret_cast<V*>(d.data() + (size_t)i * cols * elem(type_)); }
};
// drawing / IO named by the reference's debug branches (visualization flags off in every test): no-ops
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline Mat imread(const std::string&, int = 1) { return Mat(); }          // no file IO in the stand-in
inline Mat operator+(const Mat& a, double) { return a; }
inline Mat operator*(const Mat& a, double) { return a; }
inline void cvtColor(const Mat& a, Mat& b, int) { b = a; }
enum { COLOR_GRAY2BGR = 8, COLOR_BGR2GRAY = 6, COLOR_GRAY2RGB = 8 };
inline void circle(Mat&, Point2i, int, const Scalar&, int = 1) {}
inline void pyrUp(const Mat& a, Mat& b) { b = a; }      // image pyramids: named by Frame's image loaders only, not reproduced
inline void pyrDown(const Mat& a, Mat& b) { b = a; }
inline void line(Mat&, Point2i, Point2i, const Scalar&, int = 1) {}
namespace flann {
struct KDTreeIndexParams { explicit KDTreeIndexParams(int = 4) {} };
struct SearchParams { explicit SearchParams(int = 32, float = 0, bool = true) {} };
// cv::flann::Index over the rows of a CV_32F matrix; SearchParams(-1) = unlimited checks = exact search: float32 squared L2, ascending
struct Index {
  Mat feat;
  Index() {}
  Index(const Mat& features, const KDTreeIndexParams&) : feat(features) {}
  void knnSearch(const std::vector<float>& q, std::vector<int>& idx, std::vector<float>& dist, int knn, const SearchParams&) const {
    const int n = feat.rows, dim = feat.cols;
    std::vector<std::pair<float, int>> all(n);
    for (int i = 0; i < n; ++i) { float s = 0; for (int k = 0; k < dim; ++k) { const float df = feat.at<float>(i, k) - q[k]; s += df * df; } all[i] = std::make_pair(s, i); }
    const int kk = std::min(knn, n);
    std::partial_sort(all.begin(), all.begin() + kk, all.end());
    idx.assign(knn, -1); dist.assign(knn, 0.f);
    for (int i = 0; i < kk; ++i) { idx[i] = all[i].second; dist[i] = all[i].first; }
  }
};
}  // namespace flann
}  // namespace cv

#include "pvo_shim_pcl.hpp"

#ifndef LOG
namespace pvo_shim { struct NullLog { template <typename T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } }; }
#define LOG(severity) ::pvo_shim::NullLog()
#endif
