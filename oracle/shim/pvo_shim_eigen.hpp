// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  A stand-in for the handful of Eigen 3.4 types and members that the reference's base/Geometry.hpp,
// base/CostFunction.h and sensors/Equirectangular.h touch, so that those REFERENCE files compile in this container (no Eigen installed, no network)
// exactly where they lie under /root/reference (oracle/Makefile: `ref`).  What this pins: the reference's own formulas, branch thresholds, argument
// order and constructor normalisations.  What it does NOT pin: Eigen's arithmetic (column-pivoting Householder QR, the 3x3 symmetric eigen solver,
// Quaternion(Matrix3) / slerp) - those are written here from Eigen's documented algorithms, eagerly evaluated, column-major like Eigen's default.
// Not a general library: only what the compiled reference code uses.
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <limits>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16
namespace Eigen {
const int Dynamic = -1;
typedef std::ptrdiff_t Index;
template <class T> struct aligned_allocator : std::allocator<T> {
  aligned_allocator() = default;
  template <class U> aligned_allocator(const aligned_allocator<U>&) {}
  template <class U> struct rebind { typedef aligned_allocator<U> other; };
};
template <typename T> class Quaternion;
template <typename T> class AngleAxis;
template <typename T, int R, int C> class Matrix;

namespace shim {
template <typename T, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)> struct Storage {
  T d[R * C];
  Storage() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }  // Eigen leaves it uninitialised; zero is harmless
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int, int) {}
  T* data() { return d; }
  const T* data() const { return d; }
};
template <typename T, int R, int C> struct Storage<T, R, C, true> {
  std::vector<T> d; int r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
  int rows() const { return r; }
  int cols() const { return c; }
  void resize(int rr, int cc) { r = rr; c = cc; d.assign((size_t)rr * cc, T(0)); }
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
};
template <typename T, int C> struct ColPivQR;
template <typename T, int R, int C> struct CommaInit {
  Matrix<T, R, C>& m; int k;
  CommaInit& operator,(const T& v) { m.coeffLinearRowMajor(k++) = v; return *this; }
};
template <typename T, int BR, int BC, int R, int C> struct BlockRef;
}  // namespace shim

template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  typedef Eigen::Index Index;
  shim::Storage<T, R, C> s;
  Matrix() {}
  // element-list constructors (fixed-size vectors)
  Matrix(const T& a, const T& b) { init2(a, b, std::integral_constant<bool, (R == Dynamic || C == Dynamic)>()); }
  Matrix(const T& a, const T& b, const T& c) { static_assert(R * C == 3, "size"); s.d[0] = a; s.d[1] = b; s.d[2] = c; }
  Matrix(const T& a, const T& b, const T& c, const T& d) { static_assert(R * C == 4, "size"); s.d[0] = a; s.d[1] = b; s.d[2] = c; s.d[3] = d; }
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
  Matrix(U a, U b, U c) { static_assert(R * C == 3, "size"); s.d[0] = T(a); s.d[1] = T(b); s.d[2] = T(c); }
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
  Matrix(U a, U b, U c, U d) { static_assert(R * C == 4, "size"); s.d[0] = T(a); s.d[1] = T(b); s.d[2] = T(c); s.d[3] = T(d); }
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
  Matrix(U a, U b) { init2(a, b, std::integral_constant<bool, (R == Dynamic || C == Dynamic)>()); }   // (x, y) of a 2-vector, or (rows, cols) of a dynamic matrix
 private:
  template <typename U> void init2(U a, U b, std::false_type) { static_assert(R * C == 2, "size"); s.d[0] = T(a); s.d[1] = T(b); }
  template <typename U> void init2(U rows, U cols, std::true_type) { s.resize((int)rows, (int)cols); }
 public:
  template <int R2, int C2, typename = typename std::enable_if<(R2 != R || C2 != C) && (R2 == Dynamic || C2 == Dynamic || R == Dynamic || C == Dynamic)>::type>
  Matrix(const Matrix<T, R2, C2>& o) { s.resize(o.rows(), o.cols()); for (int i = 0; i < o.size() && i < size(); ++i) s.data()[i] = o.data()[i]; }   // dynamic <-> fixed of the same shape
  explicit Matrix(const Quaternion<T>& q) { static_assert(R == 3 && C == 3, "3x3"); *this = q.toRotationMatrix(); }
  explicit Matrix(const AngleAxis<T>& a) { static_assert(R == 3 && C == 3, "3x3"); *this = a.toRotationMatrix(); }

  int rows() const { return s.rows(); }
  int cols() const { return s.cols(); }
  int size() const { return rows() * cols(); }
  void resize(int r, int c) { s.resize(r, c); }
  T* data() { return s.data(); }
  const T* data() const { return s.data(); }
  T& operator()(int i, int j) { return s.data()[(size_t)j * rows() + i]; }
  const T& operator()(int i, int j) const { return s.data()[(size_t)j * rows() + i]; }
  T& operator()(int i) { return s.data()[i]; }
  const T& operator()(int i) const { return s.data()[i]; }
  T& operator[](int i) { return s.data()[i]; }
  const T& operator[](int i) const { return s.data()[i]; }
  T& coeffLinearRowMajor(int k) { return (*this)(k / cols(), k % cols()); }
  T& x() { return s.data()[0]; } const T& x() const { return s.data()[0]; }
  T& y() { return s.data()[1]; } const T& y() const { return s.data()[1]; }
  T& z() { return s.data()[2]; } const T& z() const { return s.data()[2]; }
  T& w() { return s.data()[3]; } const T& w() const { return s.data()[3]; }
  void fill(const T& v) { for (int i = 0; i < size(); ++i) s.data()[i] = v; }
  static Matrix Zero() { Matrix m; m.fill(T(0)); return m; }
  static Matrix Ones() { Matrix m; m.fill(T(1)); return m; }
  T trace() const { T t = (*this)(0, 0); for (int i = 1; i < std::min(rows(), cols()); ++i) t = t + (*this)(i, i); return t; }
  bool isZero(double prec = 1e-12) const { using std::abs; for (int i = 0; i < size(); ++i) if (!(abs(s.data()[i]) <= prec)) return false; return true; }
  Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> v; v.resize(1, cols()); for (int j = 0; j < cols(); ++j) v(0, j) = (*this)(i, j); return v; }
  template <typename I> T maxCoeff(I* where) const { int b = 0; for (int k = 1; k < size(); ++k) if (s.data()[k] > s.data()[b]) b = k; *where = (I)b; return s.data()[b]; }
  T maxCoeff() const { int b; return maxCoeff(&b); }
  static Matrix Identity() { Matrix m; m.fill(T(0)); for (int i = 0; i < std::min(m.rows(), m.cols()); ++i) m(i, i) = T(1); return m; }
  shim::CommaInit<T, R, C> operator<<(const T& v) { coeffLinearRowMajor(0) = v; return shim::CommaInit<T, R, C>{*this, 1}; }

  Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; t.resize(cols(), rows()); for (int i = 0; i < rows(); ++i) for (int j = 0; j < cols(); ++j) t(j, i) = (*this)(i, j); return t; }
  T dot(const Matrix& o) const { T acc = s.data()[0] * o.s.data()[0]; for (int i = 1; i < size(); ++i) acc = acc + s.data()[i] * o.s.data()[i]; return acc; }
  T squaredNorm() const { return dot(*this); }
  T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  Matrix normalized() const { const T n = norm(); Matrix m = *this; if (n > T(0)) for (int i = 0; i < size(); ++i) m.s.data()[i] = m.s.data()[i] / n; return m; }
  void normalize() { *this = normalized(); }
  Matrix cross(const Matrix& o) const {
    static_assert(R * C == 3, "cross");
    const T *a = data(), *b = o.data();
    return Matrix(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
  }
  Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> v; v.resize(rows(), 1); for (int i = 0; i < rows(); ++i) v(i) = (*this)(i, j); return v; }
  template <int BR, int BC> Matrix<T, BR, BC> block(int i0, int j0) const {
    Matrix<T, BR, BC> b; for (int i = 0; i < BR; ++i) for (int j = 0; j < BC; ++j) b(i, j) = (*this)(i0 + i, j0 + j); return b;
  }
  template <int BR, int BC> shim::BlockRef<T, BR, BC, R, C> block(int i0, int j0) { return shim::BlockRef<T, BR, BC, R, C>(*this, i0, j0); }
  Matrix<T, (R == Dynamic ? Dynamic : R + 1), 1> homogeneous() const { static_assert(C == 1, "vector"); Matrix<T, (R == Dynamic ? Dynamic : R + 1), 1> h; h.resize(rows() + 1, 1); for (int i = 0; i < rows(); ++i) h(i) = (*this)(i); h(rows()) = T(1); return h; }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m; m.resize(rows(), cols()); for (int i = 0; i < size(); ++i) m.data()[i] = U(s.data()[i]); return m; }
  template <int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> h; for (int i = 0; i < N; ++i) h(i) = (*this)(i); return h; }
  Matrix<T, Dynamic, 1> head(int n) const { Matrix<T, Dynamic, 1> h; h.resize(n, 1); for (int i = 0; i < n; ++i) h(i) = (*this)(i); return h; }
  Matrix<T, (R == Dynamic ? Dynamic : R - 1), 1> hnormalized() const { static_assert(C == 1, "vector"); Matrix<T, (R == Dynamic ? Dynamic : R - 1), 1> h; h.resize(rows() - 1, 1); for (int i = 0; i + 1 < rows(); ++i) h(i) = (*this)(i) / (*this)(rows() - 1); return h; }
  Matrix inverse() const;  // square, fixed size: Gauss-Jordan with partial pivoting (Eigen uses cofactors for <= 4x4; equal up to rounding)
  shim::ColPivQR<T, C> colPivHouseholderQr() const { return shim::ColPivQR<T, C>(*this); }
};

namespace shim {
// a writable view of a block that also IS its current value (so it takes part in template argument deduction as a Matrix)
template <typename T, int BR, int BC, int R, int C>
struct BlockRef : Matrix<T, BR, BC> {
  Matrix<T, R, C>& parent; int i0, j0;
  BlockRef(Matrix<T, R, C>& p, int i, int j) : parent(p), i0(i), j0(j) {
    for (int a = 0; a < BR; ++a) for (int b = 0; b < BC; ++b) (*static_cast<Matrix<T, BR, BC>*>(this))(a, b) = p(i + a, j + b);
  }
  BlockRef& operator=(const Matrix<T, BR, BC>& m) {
    for (int a = 0; a < BR; ++a) for (int b = 0; b < BC; ++b) { parent(i0 + a, j0 + b) = m(a, b); (*static_cast<Matrix<T, BR, BC>*>(this))(a, b) = m(a, b); }
    return *this;
  }
  BlockRef& operator=(const BlockRef& m) { return *this = static_cast<const Matrix<T, BR, BC>&>(m); }
};
}  // namespace shim

template <typename T, int R, int C> inline Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m = a; for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] + b.data()[i]; return m; }
template <typename T, int R, int C> inline Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m = a; for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] - b.data()[i]; return m; }
template <typename T, int R, int C> inline Matrix<T, R, C> operator-(const Matrix<T, R, C>& a) { Matrix<T, R, C> m = a; for (int i = 0; i < a.size(); ++i) m.data()[i] = -a.data()[i]; return m; }
template <typename T, int R, int C> inline Matrix<T, R, C>& operator+=(Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { a = a + b; return a; }
template <typename T, int R, int C> inline Matrix<T, R, C>& operator-=(Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { a = a - b; return a; }
// scalar products / quotients; the scalar may be any arithmetic type or T itself (Eigen promotes int / float literals the same way)
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_convertible<S, T>::value && !std::is_class<typename std::remove_reference<S>::type>::value || std::is_same<S, T>::value>::type>
inline Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, const S& k) { Matrix<T, R, C> m = a; for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] * T(k); return m; }
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_convertible<S, T>::value && !std::is_class<typename std::remove_reference<S>::type>::value || std::is_same<S, T>::value>::type>
inline Matrix<T, R, C> operator*(const S& k, const Matrix<T, R, C>& a) { Matrix<T, R, C> m = a; for (int i = 0; i < a.size(); ++i) m.data()[i] = T(k) * a.data()[i]; return m; }
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_convertible<S, T>::value && !std::is_class<typename std::remove_reference<S>::type>::value || std::is_same<S, T>::value>::type>
inline Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, const S& k) { Matrix<T, R, C> m = a; for (int i = 0; i < a.size(); ++i) m.data()[i] = a.data()[i] / T(k); return m; }
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_convertible<S, T>::value && !std::is_class<typename std::remove_reference<S>::type>::value || std::is_same<S, T>::value>::type>
inline Matrix<T, R, C>& operator/=(Matrix<T, R, C>& a, const S& k) { a = a / k; return a; }
template <typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_convertible<S, T>::value && !std::is_class<typename std::remove_reference<S>::type>::value || std::is_same<S, T>::value>::type>
inline Matrix<T, R, C>& operator*=(Matrix<T, R, C>& a, const S& k) { a = a * k; return a; }
template <typename T, int R, int K, int C>
inline Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
  Matrix<T, R, C> m; m.resize(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < b.cols(); ++j) { T acc = a(i, 0) * b(0, j); for (int k = 1; k < a.cols(); ++k) acc = acc + a(i, k) * b(k, j); m(i, j) = acc; }
  return m;
}

template <typename T, int R, int C>
inline Matrix<T, R, C> Matrix<T, R, C>::inverse() const {
  static_assert(R == C && R > 0, "square fixed size");
  T a[R][2 * R];
  for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) { a[i][j] = (*this)(i, j); a[i][R + j] = (i == j) ? T(1) : T(0); }
  for (int c = 0; c < R; ++c) {
    int p = c; using std::abs;
    for (int r = c + 1; r < R; ++r) if (abs(a[r][c]) > abs(a[p][c])) p = r;
    if (p != c) for (int j = 0; j < 2 * R; ++j) std::swap(a[p][j], a[c][j]);
    const T inv = T(1) / a[c][c];
    for (int j = 0; j < 2 * R; ++j) a[c][j] = a[c][j] * inv;
    for (int r = 0; r < R; ++r) if (r != c) { const T f = a[r][c]; if (f != T(0)) for (int j = 0; j < 2 * R; ++j) a[r][j] = a[r][j] - f * a[c][j]; }
  }
  Matrix m; for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) m(i, j) = a[i][R + j];
  return m;
}

namespace shim {
// x = argmin |A x - b| by Householder QR with column pivoting (Eigen::ColPivHouseholderQR: pivot = column of largest remaining norm, rank decided by
// |R_kk| <= eps * size * |R_00|, basic solution with zeros for the dropped columns).  A is m x C.
template <typename T, int C>
struct ColPivQR {
  int m; std::vector<T> A;  // column-major m x C
  template <int R> explicit ColPivQR(const Matrix<T, R, C>& M) : m(M.rows()), A(M.data(), M.data() + (size_t)M.rows() * C) {}
  template <int RB> Matrix<T, C, 1> solve(const Matrix<T, RB, 1>& B) const {
    using std::sqrt; using std::abs;
    std::vector<T> a = A, b(B.data(), B.data() + m);
    int perm[C]; for (int j = 0; j < C; ++j) perm[j] = j;
    T rdiag[C]; int rank = 0; T maxpivot = T(0);
    const int steps = std::min(m, C);
    for (int k = 0; k < steps; ++k) {
      int p = k; T best = T(-1);
      for (int j = k; j < C; ++j) { T n2 = T(0); for (int i = k; i < m; ++i) n2 += a[(size_t)j * m + i] * a[(size_t)j * m + i]; if (n2 > best) { best = n2; p = j; } }
      if (p != k) { for (int i = 0; i < m; ++i) std::swap(a[(size_t)p * m + i], a[(size_t)k * m + i]); std::swap(perm[p], perm[k]); }
      // Householder vector for column k, rows k..m-1
      T tail2 = T(0); for (int i = k + 1; i < m; ++i) tail2 += a[(size_t)k * m + i] * a[(size_t)k * m + i];
      const T c0 = a[(size_t)k * m + k];
      T beta, tau;
      if (tail2 <= std::numeric_limits<T>::min()) { tau = T(0); beta = c0; }
      else {
        beta = sqrt(c0 * c0 + tail2); if (c0 >= T(0)) beta = -beta;
        for (int i = k + 1; i < m; ++i) a[(size_t)k * m + i] /= (c0 - beta);
        tau = (beta - c0) / beta;
      }
      // apply H = I - tau v v^T (v = [1, essential]) to the remaining columns and to b
      if (tau != T(0)) {
        for (int j = k + 1; j < C; ++j) {
          T w = a[(size_t)j * m + k]; for (int i = k + 1; i < m; ++i) w += a[(size_t)k * m + i] * a[(size_t)j * m + i];
          w *= tau; a[(size_t)j * m + k] -= w; for (int i = k + 1; i < m; ++i) a[(size_t)j * m + i] -= w * a[(size_t)k * m + i];
        }
        T w = b[k]; for (int i = k + 1; i < m; ++i) w += a[(size_t)k * m + i] * b[i];
        w *= tau; b[k] -= w; for (int i = k + 1; i < m; ++i) b[i] -= w * a[(size_t)k * m + i];
      }
      a[(size_t)k * m + k] = beta; rdiag[k] = beta;
      if (abs(beta) > maxpivot) maxpivot = abs(beta);
    }
    const T thr = maxpivot * std::numeric_limits<T>::epsilon() * T(std::min(m, C));
    for (int k = 0; k < steps; ++k) if (abs(rdiag[k]) > thr) ++rank;
    T y[C]; for (int j = 0; j < C; ++j) y[j] = T(0);
    for (int k = rank - 1; k >= 0; --k) { T acc = b[k]; for (int j = k + 1; j < rank; ++j) acc -= a[(size_t)j * m + k] * y[j]; y[k] = acc / a[(size_t)k * m + k]; }
    Matrix<T, C, 1> x; for (int j = 0; j < C; ++j) x(perm[j]) = y[j];
    return x;
  }
};
}  // namespace shim

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi), eigenvalues ascending like Eigen::SelfAdjointEigenSolver, unit eigenvectors in the columns.
template <typename M> class SelfAdjointEigenSolver;
template <typename T>
class SelfAdjointEigenSolver<Matrix<T, 3, 3>> {
  Matrix<T, 3, 1> val; Matrix<T, 3, 3> vec;
 public:
  explicit SelfAdjointEigenSolver(const Matrix<T, 3, 3>& A0) {
    using std::sqrt; using std::abs;
    T a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a[i][j] = A0(i, j); v[i][j] = (i == j) ? T(1) : T(0); }
    for (int sweep = 0; sweep < 64; ++sweep) {
      const T off = abs(a[0][1]) + abs(a[0][2]) + abs(a[1][2]);
      if (off == T(0)) break;
      for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == T(0)) continue;
        const T theta = (a[q][q] - a[p][p]) / (T(2) * a[p][q]);
        const T t = (theta >= T(0) ? T(1) : T(-1)) / (abs(theta) + sqrt(theta * theta + T(1)));
        const T c = T(1) / sqrt(t * t + T(1)), s = t * c;
        for (int k = 0; k < 3; ++k) { const T akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 3; ++k) { const T apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 3; ++k) { const T vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
      }
    }
    int idx[3] = {0, 1, 2};
    std::sort(idx, idx + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
    for (int j = 0; j < 3; ++j) { val(j) = a[idx[j]][idx[j]]; for (int i = 0; i < 3; ++i) vec(i, j) = v[i][idx[j]]; }
  }
  const Matrix<T, 3, 1>& eigenvalues() const { return val; }
  const Matrix<T, 3, 3>& eigenvectors() const { return vec; }
};

// Eigen::Quaternion: (w, x, y, z); Quaternion(Matrix3) is Eigen's trace / largest-diagonal branch; slerp as Eigen 3.4 (threshold 1 - eps, shortest arc).
template <typename T>
class Quaternion {
 public:
  T w_, x_, y_, z_;
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(const T& w, const T& x, const T& y, const T& z) : w_(w), x_(x), y_(y), z_(z) {}
  explicit Quaternion(const Matrix<T, 3, 3>& m) {
    using std::sqrt;
    T t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > T(0)) {
      t = sqrt(t + T(1.0)); w_ = T(0.5) * t; t = T(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t; y_ = (m(0, 2) - m(2, 0)) * t; z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
      T q[3]; q[i] = T(0.5) * t; t = T(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t; q[j] = (m(j, i) + m(i, j)) * t; q[k] = (m(k, i) + m(i, k)) * t;
      x_ = q[0]; y_ = q[1]; z_ = q[2];
    }
  }
  static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
  T w() const { return w_; } T x() const { return x_; } T y() const { return y_; } T z() const { return z_; }
  Quaternion slerp(const T& t, const Quaternion& o) const {
    using std::abs; using std::acos; using std::sin;
    const T one = T(1) - std::numeric_limits<T>::epsilon();
    const T d = w_ * o.w_ + x_ * o.x_ + y_ * o.y_ + z_ * o.z_;
    const T ad = abs(d);
    T s0, s1;
    if (ad >= one) { s0 = T(1) - t; s1 = t; }
    else { const T theta = acos(ad), st = sin(theta); s0 = sin((T(1) - t) * theta) / st; s1 = sin(t * theta) / st; }
    if (d < T(0)) s1 = -s1;
    return Quaternion(s0 * w_ + s1 * o.w_, s0 * x_ + s1 * o.x_, s0 * y_ + s1 * o.y_, s0 * z_ + s1 * o.z_);
  }
  // q * v: Eigen's _transformVector: uv = 2 * (vec x v); v + w * uv + vec x uv
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const {
    const Matrix<T, 3, 1> u(x_, y_, z_);
    Matrix<T, 3, 1> uv = u.cross(v);
    uv = uv + uv;
    return v + uv * w_ + u.cross(uv);
  }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> r;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    r(0, 0) = T(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = T(1) - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = T(1) - (txx + tyy);
    return r;
  }
};

// Eigen::AngleAxis -> rotation matrix (Rodrigues, Eigen's toRotationMatrix formula); axis expected normalised
template <typename T>
class AngleAxis {
  T angle_; Matrix<T, 3, 1> axis_;
 public:
  AngleAxis(const T& angle, const Matrix<T, 3, 1>& axis) : angle_(angle), axis_(axis) {}
  Matrix<T, 3, 3> toRotationMatrix() const {
    using std::sin; using std::cos;
    Matrix<T, 3, 3> res; const T s = sin(angle_), c = cos(angle_);
    const Matrix<T, 3, 1> sin_axis = axis_ * s, cos1_axis = axis_ * (T(1) - c);
    T tmp;
    tmp = cos1_axis.x() * axis_.y(); res(0, 1) = tmp - sin_axis.z(); res(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * axis_.z(); res(0, 2) = tmp + sin_axis.y(); res(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * axis_.z(); res(1, 2) = tmp - sin_axis.x(); res(2, 1) = tmp + sin_axis.x();
    res(0, 0) = cos1_axis.x() * axis_.x() + c; res(1, 1) = cos1_axis.y() * axis_.y() + c; res(2, 2) = cos1_axis.z() * axis_.z() + c;
    return res;
  }
};
typedef AngleAxis<double> AngleAxisd; typedef AngleAxis<float> AngleAxisf;
typedef Quaternion<double> Quaterniond; typedef Quaternion<float> Quaternionf;
typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 2, 1> Vector2f;  typedef Matrix<float, 3, 1> Vector3f;  typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 4, 4> Matrix4d; typedef Matrix<float, 3, 3> Matrix3f; typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd; typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<int, 3, 1> Vector3i; typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf; typedef Matrix<int, Dynamic, Dynamic> MatrixXi; typedef Matrix<float, Dynamic, 1> VectorXf;
}  // namespace Eigen
