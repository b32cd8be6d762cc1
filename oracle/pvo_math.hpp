// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the math on PanoVLM's correspondence-and-residual hot path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything
// under oracle/.  Every function cites the reference file:line (relative to /root/reference) it
// follows.  Third-party arithmetic the reference calls but does not vendor (Ceres 2.0 rotation.h /
// jet.h / loss_function.h, Eigen 3.4 QR + eigen solver, PCL 1.10 + FLANN) is restated from the
// libraries' documented behaviour.
//
// PARITY STATUS: *unpinned by the reference's own tests* - the reference ships no unit tests, golden vectors or KATs for this path and cannot be
// built as a whole here (Eigen / Ceres / PCL / OpenCV / Boost absent).  What DOES run here is the reference's own SOURCE (oracle/_ref, `make -C oracle ref`): base/Math.h
// as it is, and 14 translation units (CostFunction.h / Geometry.hpp / Equirectangular, Velodyne.cpp, LidarFeatureAssociate.cpp, LidarLineMatch.cpp, Tracks.cpp,
// Optimization.cpp, LidarOdometry.cpp, CameraLidarLineAssociate.cpp, CameraLidarOptimizer.cpp, FileIO.cpp, Frame.cpp ...) compiled where they lie against the stand-in
// container / solver types of oracle/shim.  This restatement is bit-identical to it for FastAtan2, all functors (residuals and Jacobians), the projection, BreakToSegments,
// Transform2LidarWorld and UndistortCloud, and returns the same correspondences, tracks, pairs and residual-block lists in the same order
// (tests/test_reference_pinning.py, tests/golden/ref_*.npz).  Not pinned anywhere: the arithmetic INSIDE Eigen, Ceres and FLANN and the Ceres solver,
// which are restated from documentation; those parts are checked against independent implementations (scipy Rotation, torch float64 autograd, numpy lstsq / eigh, scipy
// cKDTree, central finite differences) in tests/test_oracle_*.py and the other fixtures under tests/golden/.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace pvo {

// ----------------------------------------------------------------------------------------------
// Forward-mode dual number, the work ceres::Jet<double,N> does inside AutoDiffCostFunction
// (reference: base/CostFunction.h:615-617 instantiates AutoDiffCostFunction<F,1,3,3,3,3> => N=12).
// Branch semantics follow SURVEY.md App. A.4: comparisons use the scalar part; abs' = sign with +1
// at 0 (ceres/jet.h: abs(f) = f < 0 ? -f : f); sqrt/acos derivatives unguarded.
// ----------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT (implicit like ceres::Jet)
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; const double inv = 1.0 / y.a; r.a = x.a * inv; const double q = r.a;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - q * y.v[i]) * inv; return r;
}
#define PVO_MIXED(op) \
  template <int N> inline Jet<N> operator op(const Jet<N>& x, double s) { return x op Jet<N>(s); } \
  template <int N> inline Jet<N> operator op(double s, const Jet<N>& y) { return Jet<N>(s) op y; }
PVO_MIXED(+) PVO_MIXED(-) PVO_MIXED(*) PVO_MIXED(/)
#undef PVO_MIXED
template <int N> inline Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
template <int N> inline Jet<N>& operator*=(Jet<N>& x, const Jet<N>& y) { x = x * y; return x; }
#define PVO_CMP(op) \
  template <int N> inline bool operator op(const Jet<N>& x, const Jet<N>& y) { return x.a op y.a; } \
  template <int N> inline bool operator op(const Jet<N>& x, double s) { return x.a op s; } \
  template <int N> inline bool operator op(double s, const Jet<N>& y) { return s op y.a; }
PVO_CMP(<) PVO_CMP(>) PVO_CMP(<=) PVO_CMP(>=) PVO_CMP(==)
#undef PVO_CMP
template <int N> inline Jet<N> sqrt(const Jet<N>& x) { Jet<N> r; r.a = std::sqrt(x.a); const double d = 1.0 / (2.0 * r.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> inline Jet<N> sin(const Jet<N>& x) { Jet<N> r; r.a = std::sin(x.a); const double d = std::cos(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> inline Jet<N> cos(const Jet<N>& x) { Jet<N> r; r.a = std::cos(x.a); const double d = -std::sin(x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> inline Jet<N> acos(const Jet<N>& x) { Jet<N> r; r.a = std::acos(x.a); const double d = -1.0 / std::sqrt(1.0 - x.a * x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> inline Jet<N> atan2(const Jet<N>& y, const Jet<N>& x) {
  Jet<N> r; r.a = std::atan2(y.a, x.a); const double d = 1.0 / (x.a * x.a + y.a * y.a);
  for (int i = 0; i < N; ++i) r.v[i] = (x.a * y.v[i] - y.a * x.v[i]) * d; return r;
}
template <int N> inline Jet<N> abs(const Jet<N>& x) { return x.a < 0.0 ? -x : x; }
inline double abs(double x) { return std::fabs(x); }
using std::sqrt; using std::sin; using std::cos; using std::acos; using std::atan2;

template <typename T> inline T Square(const T& a) { return a * a; }  // base/Math.h:31-35

// ----------------------------------------------------------------------------------------------
// Ceres 2.0 rotation.h, restated.  Matrices are column-major 3x3 like the
// Eigen::Matrix<T,3,3>::data() buffers the functors pass (base/CostFunction.h:595-598).
// ----------------------------------------------------------------------------------------------
template <typename T>
inline void AngleAxisRotatePoint(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > DBL_EPSILON) {
    const T theta = sqrt(theta2);
    const T c = cos(theta), s = sin(theta), inv = T(1.0) / theta;
    const T w[3] = {aa[0] * inv, aa[1] * inv, aa[2] * inv};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - c);
    T r[3];
    for (int i = 0; i < 3; ++i) r[i] = pt[i] * c + wxp[i] * s + w[i] * tmp;
    for (int i = 0; i < 3; ++i) out[i] = r[i];
  } else {
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    T r[3];
    for (int i = 0; i < 3; ++i) r[i] = pt[i] + wxp[i];
    for (int i = 0; i < 3; ++i) out[i] = r[i];
  }
}

template <typename T>
inline void AngleAxisToRotationMatrix(const T aa[3], T R[9]) {  // R column-major: R[c*3+r]
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > DBL_EPSILON) {
    const T theta = sqrt(theta2);
    const T wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const T c = cos(theta), s = sin(theta), k = T(1.0) - c;
    R[0] = c + wx * wx * k;      R[1] = wz * s + wx * wy * k;  R[2] = -wy * s + wx * wz * k;
    R[3] = wx * wy * k - wz * s; R[4] = c + wy * wy * k;       R[5] = wx * s + wy * wz * k;
    R[6] = wy * s + wx * wz * k; R[7] = -wx * s + wy * wz * k; R[8] = c + wz * wz * k;
  } else {
    R[0] = T(1.0); R[1] = aa[2];  R[2] = -aa[1];
    R[3] = -aa[2]; R[4] = T(1.0); R[5] = aa[0];
    R[6] = aa[1];  R[7] = -aa[0]; R[8] = T(1.0);
  }
}

template <typename T>
inline void RotationMatrixToQuaternion(const T R[9], T q[4]) {  // column-major: R(r,c) = R[c*3+r]
  auto M = [&](int r, int c) -> const T& { return R[c * 3 + r]; };
  const T trace = M(0, 0) + M(1, 1) + M(2, 2);
  if (trace >= 0.0) {
    T t = sqrt(trace + T(1.0));
    q[0] = T(0.5) * t;
    t = T(0.5) / t;
    q[1] = (M(2, 1) - M(1, 2)) * t;
    q[2] = (M(0, 2) - M(2, 0)) * t;
    q[3] = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    T t = sqrt(M(i, i) - M(j, j) - M(k, k) + T(1.0));
    q[i + 1] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (M(k, j) - M(j, k)) * t;
    q[j + 1] = (M(j, i) + M(i, j)) * t;
    q[k + 1] = (M(k, i) + M(i, k)) * t;
  }
}

template <typename T>
inline void QuaternionToAngleAxis(const T q[4], T aa[3]) {
  const T s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const T s = sqrt(s2);
    const T two_theta = T(2.0) * ((q[0] < 0.0) ? atan2(-s, -q[0]) : atan2(s, q[0]));
    const T k = two_theta / s;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
  } else {
    aa[0] = q[1] * T(2.0); aa[1] = q[2] * T(2.0); aa[2] = q[3] * T(2.0);
  }
}

template <typename T>
inline void RotationMatrixToAngleAxis(const T R[9], T aa[3]) {
  T q[4];
  RotationMatrixToQuaternion(R, q);
  QuaternionToAngleAxis(q, aa);
}

// ----------------------------------------------------------------------------------------------
// base/Geometry.hpp helpers used by the functors and the association code.
// ----------------------------------------------------------------------------------------------
template <typename T>
inline T PointToPlaneDistance(const T* plane, const T* point, bool normalized = false) {  // Geometry.hpp:275-283
  if (!normalized)
    return abs(plane[0] * point[0] + plane[1] * point[1] + plane[2] * point[2] + plane[3]) /
           sqrt(Square(plane[0]) + Square(plane[1]) + Square(plane[2]));
  return abs(plane[0] * point[0] + plane[1] * point[1] + plane[2] * point[2] + plane[3]);
}

template <typename T>
inline void ProjectPointToPlane(const T* point, const T* plane, T* out, bool normalized = false) {  // Geometry.hpp:301-316
  T dis = PointToPlaneDistance(plane, point, normalized);
  T t = normalized ? dis : dis / sqrt(Square(plane[0]) + Square(plane[1]) + Square(plane[2]));
  out[0] = point[0] - t * plane[0];
  out[1] = point[1] - t * plane[1];
  out[2] = point[2] - t * plane[2];
  if (abs(plane[0] * out[0] + plane[1] * out[1] + plane[2] * out[2] + plane[3]) > 1e-4) {
    out[0] = point[0] + t * plane[0];
    out[1] = point[1] + t * plane[1];
    out[2] = point[2] + t * plane[2];
  }
}

template <typename T>
inline T PointToLineDistance3D(const T* point, const T* line) {  // Geometry.hpp:198-211
  T x0 = line[0], y0 = line[1], z0 = line[2], nx = line[3], ny = line[4], nz = line[5];
  T k = (nx * (point[0] - x0) + ny * (point[1] - y0) + nz * (point[2] - z0)) / (Square(nx) + Square(ny) + Square(nz));
  T pp[3] = {k * nx + x0, k * ny + y0, k * nz + z0};
  return sqrt(Square(pp[0] - point[0]) + Square(pp[1] - point[1]) + Square(pp[2] - point[2]));
}

template <typename T>
inline T VectorAngle3D(const T* v1, const T* v2, bool normalized = false) {  // Geometry.hpp:450-466
  T cos_angle = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
  if (!normalized) {
    T n1 = sqrt(Square(v1[0]) + Square(v1[1]) + Square(v1[2]));
    T n2 = sqrt(Square(v2[0]) + Square(v2[1]) + Square(v2[2]));
    cos_angle = cos_angle / (n1 * n2);
  }
  if (cos_angle >= T(1.0)) return T(0);
  if (cos_angle <= T(-1.0)) return T(M_PI);
  return acos(cos_angle);
}

template <typename T>
inline T PlaneAngle(const T* p1, const T* p2, bool normalized = false) {  // Geometry.hpp:471-485
  T cos_angle = abs(p1[0] * p2[0] + p1[1] * p2[1] + p1[2] * p2[2]);
  if (!normalized) {
    T n1 = sqrt(Square(p1[0]) + Square(p1[1]) + Square(p1[2]));
    T n2 = sqrt(Square(p2[0]) + Square(p2[1]) + Square(p2[2]));
    cos_angle = cos_angle / (n1 * n2);
  }
  if (cos_angle >= T(1.0)) return T(0);
  return acos(cos_angle);
}

// FormPlane(p1,p2,p3): Geometry.hpp:328-336 (un-normalised cross product, d = -n.p1)
inline void FormPlane3(const double p1[3], const double p2[3], const double p3[3], double out[4]) {
  out[0] = (p2[1] - p1[1]) * (p3[2] - p1[2]) - (p2[2] - p1[2]) * (p3[1] - p1[1]);
  out[1] = (p2[2] - p1[2]) * (p3[0] - p1[0]) - (p2[0] - p1[0]) * (p3[2] - p1[2]);
  out[2] = (p2[0] - p1[0]) * (p3[1] - p1[1]) - (p2[1] - p1[1]) * (p3[0] - p1[0]);
  out[3] = -(out[0] * p1[0] + out[1] * p1[1] + out[2] * p1[2]);
}

// Least-squares solve of an m x 3 system by Householder QR with column pivoting — restates what
// Eigen::ColPivHouseholderQR::solve does for Geometry.hpp:361 (largest remaining column norm first,
// rank decided with Eigen's default threshold eps*min(m,n) relative to the largest pivot).
inline void LstsqColPivQR3(int m, const double* A_rowmajor, const double* b, double x[3]) {
  std::vector<double> A(A_rowmajor, A_rowmajor + 3 * m), rhs(b, b + m);
  int perm[3] = {0, 1, 2};
  double diag[3] = {0, 0, 0};
  int rank = 0;
  double maxpivot = 0.0;
  const int n = 3;
  for (int k = 0; k < n && k < m; ++k) {
    int best = k; double bestn = -1.0;
    for (int c = k; c < n; ++c) {
      double s = 0; for (int r = k; r < m; ++r) s += A[r * 3 + c] * A[r * 3 + c];
      if (s > bestn) { bestn = s; best = c; }
    }
    if (best != k) { for (int r = 0; r < m; ++r) std::swap(A[r * 3 + k], A[r * 3 + best]); std::swap(perm[k], perm[best]); }
    double norm = std::sqrt(bestn);
    if (k == 0) maxpivot = norm;
    if (norm <= maxpivot * DBL_EPSILON * std::min(m, n)) break;
    ++rank;
    const double alpha = (A[k * 3 + k] > 0) ? -norm : norm;
    std::vector<double> vv(m, 0.0);
    vv[k] = A[k * 3 + k] - alpha;
    for (int r = k + 1; r < m; ++r) vv[r] = A[r * 3 + k];
    double vtv = 0; for (int r = k; r < m; ++r) vtv += vv[r] * vv[r];
    if (vtv > 0) {
      for (int c = k; c < n; ++c) {
        double dot = 0; for (int r = k; r < m; ++r) dot += vv[r] * A[r * 3 + c];
        const double f = 2.0 * dot / vtv;
        for (int r = k; r < m; ++r) A[r * 3 + c] -= f * vv[r];
      }
      double dot = 0; for (int r = k; r < m; ++r) dot += vv[r] * rhs[r];
      const double f = 2.0 * dot / vtv;
      for (int r = k; r < m; ++r) rhs[r] -= f * vv[r];
    }
    diag[k] = A[k * 3 + k];
  }
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) {
    double s = rhs[k];
    for (int c = k + 1; c < rank; ++c) s -= A[k * 3 + c] * y[c];
    y[k] = s / diag[k];
  }
  x[0] = x[1] = x[2] = 0.0;
  for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

// FormPlane(points, tolerance): Geometry.hpp:345-373.  A x = -1, d = 1/|x|, n = x/|x|; zero vector if
// any |n.p + d| > tolerance (tolerance > 0).
inline void FormPlaneLSQ(int m, const double* pts, double tolerance, double out[4]) {
  std::vector<double> b(m, -1.0);
  double x[3];
  LstsqColPivQR3(m, pts, b.data(), x);
  const double nrm = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const double d = 1.0 / nrm;
  const double n[3] = {x[0] / nrm, x[1] / nrm, x[2] / nrm};
  if (tolerance > 0) {
    for (int i = 0; i < m; ++i) {
      if (std::fabs(n[0] * pts[i * 3] + n[1] * pts[i * 3 + 1] + n[2] * pts[i * 3 + 2] + d) > tolerance) {
        out[0] = out[1] = out[2] = out[3] = 0.0;
        return;
      }
    }
  }
  out[0] = n[0]; out[1] = n[1]; out[2] = n[2]; out[3] = d;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi sweeps (ascending eigenvalues, like
// Eigen::SelfAdjointEigenSolver used at Geometry.hpp:237).  evec is column k = vec[k][*].
inline void SymEig3(const double Ain[9], double eval[3], double evec[3][3]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = Ain[i * 3 + j];
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * dg || off == 0.0) break;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      if (A[p][q] == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
      for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
      for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
    }
  }
  int idx[3] = {0, 1, 2};
  std::sort(idx, idx + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
  for (int k = 0; k < 3; ++k) { eval[k] = A[idx[k]][idx[k]]; for (int r = 0; r < 3; ++r) evec[k][r] = V[r][idx[k]]; }
}

// FormLine(points, tolerance, dis_threshold): Geometry.hpp:220-260.  Returns true and fills line6
// (centroid, unit direction) when the points form a line, else zero vector + false.
inline bool FormLinePCA(int m, const double* pts, double tolerance, double dis_threshold, double line6[6]) {
  double c[3] = {0, 0, 0};
  for (int i = 0; i < m; ++i) for (int k = 0; k < 3; ++k) c[k] = c[k] + pts[i * 3 + k];
  for (int k = 0; k < 3; ++k) c[k] = c[k] / double(m);
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < m; ++i) {
    const double z[3] = {pts[i * 3] - c[0], pts[i * 3 + 1] - c[1], pts[i * 3 + 2] - c[2]};
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cov[r * 3 + k] += z[r] * z[k];
  }
  double ev[3], evec[3][3];
  SymEig3(cov, ev, evec);
  for (int k = 0; k < 6; ++k) line6[k] = 0.0;
  if (ev[2] > tolerance * ev[1]) {
    double d[3] = {evec[2][0], evec[2][1], evec[2][2]};
    const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double cand[6] = {c[0], c[1], c[2], d[0] / n, d[1] / n, d[2] / n};
    if (dis_threshold > 0.0) {
      for (int i = 0; i < m; ++i)
        if (PointToLineDistance3D(pts + i * 3, cand) > dis_threshold) return false;
    }
    for (int k = 0; k < 6; ++k) line6[k] = cand[k];
    return true;
  }
  return false;
}

// base/Math.h:15-29.  T=float: the double literals promote intermediates to double exactly as the
// C++ expression does (float*double -> double), the result is rounded to T on return / assignment.
template <typename T>
inline T FastAtan2(const T& y, const T& x) {
  T ax = std::abs(x), ay = std::abs(y);
  T a = std::min(ax, ay) / (std::max(ax, ay) + (T)DBL_EPSILON);
  T s = a * a;
  T r = ((-0.04432655554792128 * s + 0.1555786518463281) * s - 0.3258083974640975) * s * a + 0.9997878412794807 * a;
  if (ay > ax) r = M_PI_2 - r;
  if (x < 0) r = M_PI - r;
  if (y < 0) r = -r;
  return r;
}

// ----------------------------------------------------------------------------------------------
// Cost functors (base/CostFunction.h).  Each is a template on T exactly like the reference so the
// same body is evaluated with T=double and T=Jet<12>.
// ----------------------------------------------------------------------------------------------
// Common 4-block transform P_r = R_rw R_wn P_n - R_rw R_wn t_nw + t_rw via the angle-axis detour
// (CostFunction.h:584-604, 648-668, 791-811, 858-878).
template <typename T>
inline void TransformNeiToRef(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, const double p[3], T out[3]) {
  T point[3] = {T(p[0]), T(p[1]), T(p[2])};
  T aa_wn[3] = {T(-1.0) * aa_nw[0], T(-1.0) * aa_nw[1], T(-1.0) * aa_nw[2]};
  T R_rw[9], R_wn[9], R_rn[9];
  AngleAxisToRotationMatrix(aa_rw, R_rw);
  AngleAxisToRotationMatrix(aa_wn, R_wn);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r)
      R_rn[c * 3 + r] = R_rw[0 * 3 + r] * R_wn[c * 3 + 0] + R_rw[1 * 3 + r] * R_wn[c * 3 + 1] + R_rw[2 * 3 + r] * R_wn[c * 3 + 2];
  T aa_rn[3];
  RotationMatrixToAngleAxis(R_rn, aa_rn);
  T vec_tmp[3];
  AngleAxisRotatePoint(aa_rn, point, out);
  AngleAxisRotatePoint(aa_rn, t_nw, vec_tmp);
  out[0] = out[0] - vec_tmp[0] + t_rw[0];
  out[1] = out[1] - vec_tmp[1] + t_rw[1];
  out[2] = out[2] - vec_tmp[2] + t_rw[2];
}

// Two-step transform of Plane2Plane_Global / PlaneIOUResidual (CostFunction.h:369-400, 466-483).
template <typename T>
inline void TransformTwoStep(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, const double p[3], T out[3]) {
  T aa_wn[3] = {T(-1.0) * aa_nw[0], T(-1.0) * aa_nw[1], T(-1.0) * aa_nw[2]};
  T t_wn[3];
  AngleAxisRotatePoint(aa_wn, t_nw, t_wn);
  t_wn[0] = t_wn[0] * T(-1.0); t_wn[1] = t_wn[1] * T(-1.0); t_wn[2] = t_wn[2] * T(-1.0);
  T pt[3] = {T(p[0]), T(p[1]), T(p[2])};
  T pw[3];
  AngleAxisRotatePoint(aa_wn, pt, pw);
  pw[0] = pw[0] + t_wn[0]; pw[1] = pw[1] + t_wn[1]; pw[2] = pw[2] + t_wn[2];
  AngleAxisRotatePoint(aa_rw, pw, out);
  out[0] = out[0] + t_rw[0]; out[1] = out[1] + t_rw[1]; out[2] = out[2] + t_rw[2];
}

// The `normalize_distance` tail shared by Point2Plane_Angle (699-715) and Point2Line_Angle (901-917).
template <typename T>
inline T AngleTail(const T* point_ref, const T* point_projected, bool normalize_distance) {
  if (normalize_distance) {
    T norm = sqrt(point_projected[0] * point_projected[0] + point_projected[1] * point_projected[1] + point_projected[2] * point_projected[2]);
    T ratio = (norm - T(1.0)) / norm;
    T c[3] = {ratio * point_projected[0], ratio * point_projected[1], ratio * point_projected[2]};
    T vec1[3] = {point_projected[0] - c[0], point_projected[1] - c[1], point_projected[2] - c[2]};
    T vec2[3] = {point_ref[0] - c[0], point_ref[1] - c[1], point_ref[2] - c[2]};
    return VectorAngle3D(vec1, vec2);
  }
  return VectorAngle3D(point_ref, point_projected);
}

struct Point2Plane_Meter {  // CostFunction.h:567-619
  double plane[4], p[3], weight;
  template <typename T> bool operator()(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, T* residual) const {
    T point_ref[3];
    TransformNeiToRef(aa_rw, t_rw, aa_nw, t_nw, p, point_ref);
    T pc[4] = {T(plane[0]), T(plane[1]), T(plane[2]), T(plane[3])};
    residual[0] = T(weight) * PointToPlaneDistance(pc, point_ref, true);
    return true;
  }
};

struct Point2Plane_Angle {  // CostFunction.h:630-729 (weight is stored but never used, 714-717)
  double plane[4], p[3], weight; bool normalize_distance;
  template <typename T> bool operator()(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, T* residual) const {
    T point_ref[3];
    TransformNeiToRef(aa_rw, t_rw, aa_nw, t_nw, p, point_ref);
    T pc[4] = {T(plane[0]), T(plane[1]), T(plane[2]), T(plane[3])};
    T pp[3];
    T dis = PointToPlaneDistance(pc, point_ref, true);
    if (dis < T(1e-3)) { residual[0] = T(0.0); return true; }
    pp[0] = point_ref[0] - dis * pc[0]; pp[1] = point_ref[1] - dis * pc[1]; pp[2] = point_ref[2] - dis * pc[2];
    if (abs(pc[0] * pp[0] + pc[1] * pp[1] + pc[2] * pp[2] + pc[3]) > 1e-4) {
      pp[0] = point_ref[0] + dis * pc[0]; pp[1] = point_ref[1] + dis * pc[1]; pp[2] = point_ref[2] + dis * pc[2];
    }
    residual[0] = AngleTail(point_ref, pp, normalize_distance);
    return true;
  }
};

struct Point2Line_Meter {  // CostFunction.h:769-829; ctor normalises (a-b) (781-784)
  double line_point[3], line_dir[3], p[3], weight;
  template <typename T> bool operator()(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, T* residual) const {
    T point_ref[3];
    TransformNeiToRef(aa_rw, t_rw, aa_nw, t_nw, p, point_ref);
    T line[6] = {T(line_point[0]), T(line_point[1]), T(line_point[2]), T(line_dir[0]), T(line_dir[1]), T(line_dir[2])};
    residual[0] = T(weight) * PointToLineDistance3D(point_ref, line);
    return true;
  }
};

struct Point2Line_Angle {  // CostFunction.h:836-934 (weight unused, 916-920)
  double line_point[3], line_dir[3], p[3], weight; bool normalize_distance;
  template <typename T> bool operator()(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, T* residual) const {
    T point_ref[3];
    TransformNeiToRef(aa_rw, t_rw, aa_nw, t_nw, p, point_ref);
    T x0 = T(line_point[0]), y0 = T(line_point[1]), z0 = T(line_point[2]);
    T nx = T(line_dir[0]), ny = T(line_dir[1]), nz = T(line_dir[2]);
    T k = nx * (point_ref[0] - x0) + ny * (point_ref[1] - y0) + nz * (point_ref[2] - z0);
    T pp[3] = {k * nx + x0, k * ny + y0, k * nz + z0};
    T dis = sqrt((point_ref[0] - pp[0]) * (point_ref[0] - pp[0]) + (point_ref[1] - pp[1]) * (point_ref[1] - pp[1]) +
                 (point_ref[2] - pp[2]) * (point_ref[2] - pp[2]));
    if (dis < T(1e-3)) { residual[0] = T(0.0); return true; }
    residual[0] = AngleTail(point_ref, pp, normalize_distance);
    return true;
  }
};

struct Plane2Plane_Global {  // CostFunction.h:350-425; ctor normalises plane_ref (362)
  double plane_ref[3], point_a[3], point_b[3], weight;
  template <typename T> bool operator()(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, T* residual) const {
    T A[3], B[3];
    TransformTwoStep(aa_rw, t_rw, aa_nw, t_nw, point_a, A);
    TransformTwoStep(aa_rw, t_rw, aa_nw, t_nw, point_b, B);
    T plane1[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
    T plane2[3] = {T(plane_ref[0]), T(plane_ref[1]), T(plane_ref[2])};
    residual[0] = T(weight) * PlaneAngle<T>(plane2, plane1);
    return true;
  }
};

struct PlaneIOUResidual {  // CostFunction.h:433-507 (LiDAR-LiDAR ctor 452-459 normalises the plane by |n|)
  double ref_plane[4], middle_neighbor[3], middle_ref[3], angle, weight;
  template <typename T> bool operator()(const T* aa_rw, const T* t_rw, const T* aa_nw, const T* t_nw, T* residual) const {
    T mid[3];
    TransformTwoStep(aa_rw, t_rw, aa_nw, t_nw, middle_neighbor, mid);
    T pc[4] = {T(ref_plane[0]), T(ref_plane[1]), T(ref_plane[2]), T(ref_plane[3])};
    T ip[3] = {T(middle_ref[0]), T(middle_ref[1]), T(middle_ref[2])};
    T proj[3];
    ProjectPointToPlane(mid, pc, proj, true);
    T curr = VectorAngle3D(proj, ip);
    if (curr < T(angle)) residual[0] = T(0.0);
    else residual[0] = T(weight) * (curr - T(angle));
    return true;
  }
};

struct PairWisePoint2Plane_Meter {  // CostFunction.h:732-766 (2 blocks)
  double plane[4], p[3], weight;
  template <typename T> bool operator()(const T* aa_21, const T* t_21, T* residual) const {
    T point[3] = {T(p[0]), T(p[1]), T(p[2])}, q[3];
    AngleAxisRotatePoint(aa_21, point, q);
    q[0] = q[0] + t_21[0]; q[1] = q[1] + t_21[1]; q[2] = q[2] + t_21[2];
    T pc[4] = {T(plane[0]), T(plane[1]), T(plane[2]), T(plane[3])};
    residual[0] = T(weight) * PointToPlaneDistance(pc, q, true);
    return true;
  }
};

struct PairWisePoint2Line_Meter {  // CostFunction.h:939-980 (2 blocks)
  double line_point[3], line_dir[3], p[3], weight;
  template <typename T> bool operator()(const T* aa_21, const T* t_21, T* residual) const {
    T point[3] = {T(p[0]), T(p[1]), T(p[2])}, q[3];
    AngleAxisRotatePoint(aa_21, point, q);
    q[0] = q[0] + t_21[0]; q[1] = q[1] + t_21[1]; q[2] = q[2] + t_21[2];
    T line[6] = {T(line_point[0]), T(line_point[1]), T(line_point[2]), T(line_dir[0]), T(line_dir[1]), T(line_dir[2])};
    residual[0] = T(weight) * PointToLineDistance3D(q, line);
    return true;
  }
};

struct Plane2Plane_Relative {  // CostFunction.h:294-346 (2 blocks aa_cl, t_cl; residual in DEGREES, :334)
  double plane_ref[3], point_a[3], point_b[3], weight;
  template <typename T> bool operator()(const T* aa_cl, const T* t_cl, T* residual) const {
    T pa[3] = {T(point_a[0]), T(point_a[1]), T(point_a[2])}, pb[3] = {T(point_b[0]), T(point_b[1]), T(point_b[2])}, A[3], B[3];
    AngleAxisRotatePoint(aa_cl, pa, A);
    A[0] = A[0] + t_cl[0]; A[1] = A[1] + t_cl[1]; A[2] = A[2] + t_cl[2];
    AngleAxisRotatePoint(aa_cl, pb, B);
    B[0] = B[0] + t_cl[0]; B[1] = B[1] + t_cl[1]; B[2] = B[2] + t_cl[2];
    T lidar_plane[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
    T img_plane[3] = {T(plane_ref[0]), T(plane_ref[1]), T(plane_ref[2])};
    residual[0] = T(weight) * PlaneAngle<T>(img_plane, lidar_plane) * T(180.0) / T(M_PI);
    return true;
  }
};

struct PlaneRelativeIOUResidual {  // CostFunction.h:509-563 (2 blocks aa_cl, t_cl)
  double ref_plane[4], middle_neighbor[3], middle_ref[3], angle, weight;
  template <typename T> bool operator()(const T* aa_cl, const T* t_cl, T* residual) const {
    T pn[3] = {T(middle_neighbor[0]), T(middle_neighbor[1]), T(middle_neighbor[2])}, mid[3];
    AngleAxisRotatePoint(aa_cl, pn, mid);
    mid[0] = mid[0] + t_cl[0]; mid[1] = mid[1] + t_cl[1]; mid[2] = mid[2] + t_cl[2];
    T pc[4] = {T(ref_plane[0]), T(ref_plane[1]), T(ref_plane[2]), T(ref_plane[3])};
    T ip[3] = {T(middle_ref[0]), T(middle_ref[1]), T(middle_ref[2])};
    T proj[3];
    ProjectPointToPlane(mid, pc, proj, true);
    T curr = VectorAngle3D(proj, ip);
    if (curr < T(angle)) residual[0] = T(0.0);
    else residual[0] = T(weight) * (curr - T(angle));
    return true;
  }
};

struct Line2Line_Angle {  // CostFunction.h:984-1022 (2 rotation blocks; < 1e-3 => 0)
  double dir_ref[3], dir_nei[3];
  template <typename T> bool operator()(const T* aa_rw, const T* aa_nw, T* residual) const {
    T aa_wn[3] = {-aa_nw[0], -aa_nw[1], -aa_nw[2]};
    T dn[3] = {T(dir_nei[0]), T(dir_nei[1]), T(dir_nei[2])}, dw[3], dr[3];
    AngleAxisRotatePoint(aa_wn, dn, dw);
    AngleAxisRotatePoint(aa_rw, dw, dr);
    T ref[3] = {T(dir_ref[0]), T(dir_ref[1]), T(dir_ref[2])};
    residual[0] = PlaneAngle<T>(dr, ref, true);
    if (residual[0] < T(1e-3)) residual[0] = T(0.0);
    return true;
  }
};

// ceres::AutoDiffCostFunction<F,1,3,3,3,3>::Evaluate: residual (+ 1x12 Jacobian, row-major per block
// = [d/daa_rw | d/dt_rw | d/daa_nw | d/dt_nw]) from one Jet<12> pass.
template <typename F>
inline void EvaluateAutoDiff4(const F& f, const double* aa_r, const double* t_r, const double* aa_n, const double* t_n,
                              double* residual, double* jac12) {
  if (!jac12) { f(aa_r, t_r, aa_n, t_n, residual); return; }
  using J = Jet<12>;
  J a[3], b[3], c[3], d[3], r;
  for (int i = 0; i < 3; ++i) { a[i] = J(aa_r[i], i); b[i] = J(t_r[i], 3 + i); c[i] = J(aa_n[i], 6 + i); d[i] = J(t_n[i], 9 + i); }
  f(a, b, c, d, &r);
  *residual = r.a;
  for (int i = 0; i < 12; ++i) jac12[i] = r.v[i];
}
template <typename F>
inline void EvaluateAutoDiff2(const F& f, const double* p0, const double* p1, double* residual, double* jac6) {
  if (!jac6) { f(p0, p1, residual); return; }
  using J = Jet<6>;
  J a[3], b[3], r;
  for (int i = 0; i < 3; ++i) { a[i] = J(p0[i], i); b[i] = J(p1[i], 3 + i); }
  f(a, b, &r);
  *residual = r.a;
  for (int i = 0; i < 6; ++i) jac6[i] = r.v[i];
}

// ceres::HuberLoss + Corrector.  rho'' <= 0 => residual and Jacobian row are
// scaled by sqrt(rho'); cost contribution is 0.5*rho(s).  a <= 0 means "loss == nullptr".
inline void HuberCorrect(double a, double* r, double* jac, int njac, double* cost) {
  const double s = (*r) * (*r);
  if (a <= 0.0 || s <= a * a) { *cost = 0.5 * s; return; }
  const double sq = std::sqrt(s);
  const double rho = 2.0 * a * sq - a * a, rho1 = a / sq;
  const double k = std::sqrt(rho1);
  *r *= k;
  if (jac) for (int i = 0; i < njac; ++i) jac[i] *= k;
  *cost = 0.5 * rho;
}

}  // namespace pvo
