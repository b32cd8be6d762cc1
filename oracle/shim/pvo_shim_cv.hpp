// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  Stand-in for the OpenCV value types the reference's geometry / projection headers use (cv::Point_,
// cv::Point3_, cv::Vec, a cv::Mat that can only be empty or a float3 table), plus glog's LOG() (the PCL stand-ins live in pvo_shim_pcl.hpp).
// See pvo_shim_eigen.hpp for why this exists.
#pragma once
#include <cmath>
#include <iostream>
#include <vector>

#define CV_32FC3 21
#define CV_16U 2
namespace cv {
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <typename U> Point_(const Point_<U>& o) : x(T(o.x)), y(T(o.y)) {}  // NOLINT
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
  template <typename U> Point3_(const Point3_<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}  // NOLINT
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
// only what Equirectangular's optional pixel -> bearing table needs
struct Mat {
  int rows = 0, cols = 0; std::vector<Vec3f> d;
  static Mat zeros(int r, int c, int) { Mat m; m.rows = r; m.cols = c; m.d.assign((size_t)r * c, Vec3f()); return m; }
  bool empty() const { return d.empty(); }
  template <typename V> V& at(int i, int j) { return d[(size_t)i * cols + j]; }
  template <typename V> const V& at(int i, int j) const { return d[(size_t)i * cols + j]; }
  template <typename V> const V& at(const Point2i& p) const { return d[(size_t)p.y * cols + p.x]; }
};
}  // namespace cv

#include "pvo_shim_pcl.hpp"

#ifndef LOG
namespace pvo_shim { struct NullLog { template <typename T> NullLog& operator<<(const T&) { return *this; } NullLog& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } }; }
#define LOG(severity) ::pvo_shim::NullLog()
#endif
