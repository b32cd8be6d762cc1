// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  Stand-in for the part of Ceres Solver 2.0 that base/CostFunction.h uses: ceres::Jet<T,N> with the
// derivative rules of ceres/jet.h, the angle-axis helpers of ceres/rotation.h (column-major 3x3 through raw pointers), ceres::CostFunction and an
// AutoDiffCostFunction whose Evaluate() runs the functor once on doubles (residuals only) or once on Jets seeded with the identity (residuals +
// row-major num_residuals x block_size Jacobians, null entries skipped) - the interface of ceres::CostFunction::Evaluate.  Written from Ceres'
// published behaviour (BSD licence, documented formulas); the solver, loss functions and Problem are not part of it.
#pragma once
#include <cfloat>
#include <cmath>
#include <limits>
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

namespace ceres {
template <typename T, int N>
struct Jet {
  T a; T v[N];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); }  // NOLINT: implicit like ceres::Jet
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>::type>
  Jet(U s) : a(T(s)) { for (int i = 0; i < N; ++i) v[i] = T(); }  // NOLINT
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1); }
};
#define PVS_J template <typename T, int N> inline Jet<T, N>
PVS_J operator+(const Jet<T, N>& f) { return f; }
PVS_J operator-(const Jet<T, N>& f) { Jet<T, N> r; r.a = -f.a; for (int i = 0; i < N; ++i) r.v[i] = -f.v[i]; return r; }
PVS_J operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> r; r.a = f.a + g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] + g.v[i]; return r; }
PVS_J operator+(const Jet<T, N>& f, T s) { Jet<T, N> r = f; r.a = f.a + s; return r; }
PVS_J operator+(T s, const Jet<T, N>& f) { Jet<T, N> r = f; r.a = s + f.a; return r; }
PVS_J operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> r; r.a = f.a - g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] - g.v[i]; return r; }
PVS_J operator-(const Jet<T, N>& f, T s) { Jet<T, N> r = f; r.a = f.a - s; return r; }
PVS_J operator-(T s, const Jet<T, N>& f) { Jet<T, N> r; r.a = s - f.a; for (int i = 0; i < N; ++i) r.v[i] = -f.v[i]; return r; }
PVS_J operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> r; r.a = f.a * g.a; for (int i = 0; i < N; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
PVS_J operator*(const Jet<T, N>& f, T s) { Jet<T, N> r; r.a = f.a * s; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * s; return r; }
PVS_J operator*(T s, const Jet<T, N>& f) { Jet<T, N> r; r.a = f.a * s; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * s; return r; }
PVS_J operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  const T g_a_inverse = T(1.0) / g.a; const T f_a_by_g_a = f.a * g_a_inverse;
  Jet<T, N> r; r.a = f_a_by_g_a; for (int i = 0; i < N; ++i) r.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse; return r;
}
PVS_J operator/(T s, const Jet<T, N>& g) { const T k = -s / (g.a * g.a); Jet<T, N> r; r.a = s / g.a; for (int i = 0; i < N; ++i) r.v[i] = g.v[i] * k; return r; }
PVS_J operator/(const Jet<T, N>& f, T s) { const T inv = T(1.0) / s; Jet<T, N> r; r.a = f.a * inv; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * inv; return r; }
template <typename T, int N> inline Jet<T, N>& operator+=(Jet<T, N>& f, const Jet<T, N>& g) { f = f + g; return f; }
template <typename T, int N> inline Jet<T, N>& operator-=(Jet<T, N>& f, const Jet<T, N>& g) { f = f - g; return f; }
template <typename T, int N> inline Jet<T, N>& operator*=(Jet<T, N>& f, const Jet<T, N>& g) { f = f * g; return f; }
template <typename T, int N> inline Jet<T, N>& operator/=(Jet<T, N>& f, const Jet<T, N>& g) { f = f / g; return f; }
template <typename T, int N> inline Jet<T, N>& operator+=(Jet<T, N>& f, T s) { f = f + s; return f; }
template <typename T, int N> inline Jet<T, N>& operator-=(Jet<T, N>& f, T s) { f = f - s; return f; }
template <typename T, int N> inline Jet<T, N>& operator*=(Jet<T, N>& f, T s) { f = f * s; return f; }
template <typename T, int N> inline Jet<T, N>& operator/=(Jet<T, N>& f, T s) { f = f / s; return f; }
// comparisons look at the scalar part only
#define PVS_CMP(op) \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, T s) { return f.a op s; } \
  template <typename T, int N> inline bool operator op(T s, const Jet<T, N>& g) { return s op g.a; }
PVS_CMP(<) PVS_CMP(<=) PVS_CMP(>) PVS_CMP(>=) PVS_CMP(==) PVS_CMP(!=)
#undef PVS_CMP
PVS_J abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
PVS_J sqrt(const Jet<T, N>& f) { const T tmp = std::sqrt(f.a); const T k = T(1.0) / (T(2.0) * tmp); Jet<T, N> r; r.a = tmp; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J sin(const Jet<T, N>& f) { const T k = std::cos(f.a); Jet<T, N> r; r.a = std::sin(f.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J cos(const Jet<T, N>& f) { const T k = -std::sin(f.a); Jet<T, N> r; r.a = std::cos(f.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J acos(const Jet<T, N>& f) { const T k = -T(1.0) / std::sqrt(T(1.0) - f.a * f.a); Jet<T, N> r; r.a = std::acos(f.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J asin(const Jet<T, N>& f) { const T k = T(1.0) / std::sqrt(T(1.0) - f.a * f.a); Jet<T, N> r; r.a = std::asin(f.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J atan(const Jet<T, N>& f) { const T k = T(1.0) / (T(1.0) + f.a * f.a); Jet<T, N> r; r.a = std::atan(f.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  const T k = T(1.0) / (f.a * f.a + g.a * g.a); Jet<T, N> r; r.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; ++i) r.v[i] = k * (-g.a * f.v[i] + f.a * g.v[i]); return r;
}
PVS_J exp(const Jet<T, N>& f) { const T k = std::exp(f.a); Jet<T, N> r; r.a = k; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
PVS_J log(const Jet<T, N>& f) { const T k = T(1.0) / f.a; Jet<T, N> r; r.a = std::log(f.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * k; return r; }
#undef PVS_J
using std::abs; using std::sqrt; using std::sin; using std::cos; using std::acos; using std::asin; using std::atan; using std::atan2; using std::exp; using std::log;

// ---- ceres/rotation.h (angle-axis <-> rotation matrix through quaternions; matrices column-major) ----
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  const T theta2 = angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta), sintheta = sin(theta), theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    const T r0 = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    const T r1 = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    const T r2 = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
    result[0] = r0; result[1] = r1; result[2] = r2;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2], angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    const T r0 = pt[0] + w_cross_pt[0], r1 = pt[1] + w_cross_pt[1], r2 = pt[2] + w_cross_pt[2];
    result[0] = r0; result[1] = r1; result[2] = r2;
  }
}
template <typename T>
inline void AngleAxisToRotationMatrix(const T* angle_axis, T* R) {  // R(r,c) = R[c * 3 + r]
  const T kOne = T(1.0);
  const T theta2 = angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T wx = angle_axis[0] / theta, wy = angle_axis[1] / theta, wz = angle_axis[2] / theta;
    const T costheta = cos(theta), sintheta = sin(theta);
    R[0] = costheta + wx * wx * (kOne - costheta);
    R[1] = wz * sintheta + wx * wy * (kOne - costheta);
    R[2] = -wy * sintheta + wx * wz * (kOne - costheta);
    R[3] = wx * wy * (kOne - costheta) - wz * sintheta;
    R[4] = costheta + wy * wy * (kOne - costheta);
    R[5] = wx * sintheta + wy * wz * (kOne - costheta);
    R[6] = wy * sintheta + wx * wz * (kOne - costheta);
    R[7] = -wx * sintheta + wy * wz * (kOne - costheta);
    R[8] = costheta + wz * wz * (kOne - costheta);
  } else {
    R[0] = kOne; R[1] = angle_axis[2]; R[2] = -angle_axis[1];
    R[3] = -angle_axis[2]; R[4] = kOne; R[5] = angle_axis[0];
    R[6] = angle_axis[1]; R[7] = -angle_axis[0]; R[8] = kOne;
  }
}
template <typename T>
inline void RotationMatrixToQuaternion(const T* R, T* q) {
#define PVS_R(r, c) R[(c) * 3 + (r)]
  const T trace = PVS_R(0, 0) + PVS_R(1, 1) + PVS_R(2, 2);
  if (trace >= 0.0) {
    T t = sqrt(trace + T(1.0));
    q[0] = T(0.5) * t; t = T(0.5) / t;
    q[1] = (PVS_R(2, 1) - PVS_R(1, 2)) * t; q[2] = (PVS_R(0, 2) - PVS_R(2, 0)) * t; q[3] = (PVS_R(1, 0) - PVS_R(0, 1)) * t;
  } else {
    int i = 0;
    if (PVS_R(1, 1) > PVS_R(0, 0)) i = 1;
    if (PVS_R(2, 2) > PVS_R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    T t = sqrt(PVS_R(i, i) - PVS_R(j, j) - PVS_R(k, k) + T(1.0));
    q[i + 1] = T(0.5) * t; t = T(0.5) / t;
    q[0] = (PVS_R(k, j) - PVS_R(j, k)) * t; q[j + 1] = (PVS_R(j, i) + PVS_R(i, j)) * t; q[k + 1] = (PVS_R(k, i) + PVS_R(i, k)) * t;
  }
#undef PVS_R
}
template <typename T>
inline void QuaternionToAngleAxis(const T* q, T* angle_axis) {
  const T &q1 = q[1], &q2 = q[2], &q3 = q[3];
  const T sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3;
  if (sin_squared_theta > T(0.0)) {
    const T sin_theta = sqrt(sin_squared_theta);
    const T& cos_theta = q[0];
    const T two_theta = T(2.0) * ((cos_theta < T(0.0)) ? atan2(-sin_theta, -cos_theta) : atan2(sin_theta, cos_theta));
    const T k = two_theta / sin_theta;
    angle_axis[0] = q1 * k; angle_axis[1] = q2 * k; angle_axis[2] = q3 * k;
  } else {
    const T k(2.0);
    angle_axis[0] = q1 * k; angle_axis[1] = q2 * k; angle_axis[2] = q3 * k;
  }
}
template <typename T>
inline void RotationMatrixToAngleAxis(const T* R, T* angle_axis) { T q[4]; RotationMatrixToQuaternion(R, q); QuaternionToAngleAxis(q, angle_axis); }

// ---- cost function interface ----
class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int> sizes_; int num_residuals_ = 0;
};
// rho(s), rho'(s), rho''(s) of the squared residual norm s (ceres/loss_function.h)
// ceres::SizedCostFunction / ceres::EvaluationCallback: the two classes a user-side bridge derives from (include/panovlm_b200_ceres_adapter.hpp)
template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() { const int n[] = {Ns...}; sizes_.assign(n, n + sizeof...(Ns)); num_residuals_ = kNumResiduals; }
};
class EvaluationCallback {
 public:
  virtual ~EvaluationCallback() {}
  virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class HuberLoss : public LossFunction {
  const double a_, b_;
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  double a() const { return a_; }
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r); rho[2] = -rho[1] / (2.0 * s); }
    else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
};

template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  std::unique_ptr<Functor> functor_;
  static constexpr int kBlocks = sizeof...(Ns);
  static constexpr int Total() { int s = 0; const int n[] = {Ns...}; for (int i = 0; i < kBlocks; ++i) s += n[i]; return s; }
  template <typename T, size_t... I> bool Call(const T* const* p, T* r, std::index_sequence<I...>) const { return (*functor_)(p[I]..., r); }
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) { const int n[] = {Ns...}; sizes_.assign(n, n + kBlocks); num_residuals_ = kNumResiduals; }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return Call<double>(parameters, residuals, std::make_index_sequence<kBlocks>());
    constexpr int kTotal = Total();
    typedef Jet<double, kTotal> J;
    const int n[] = {Ns...};
    J x[kTotal]; const J* px[kBlocks]; J out[kNumResiduals];
    int o = 0;
    for (int b = 0; b < kBlocks; ++b) { px[b] = x + o; for (int i = 0; i < n[b]; ++i, ++o) x[o] = J(parameters[b][i], o); }
    if (!Call<J>(px, out, std::make_index_sequence<kBlocks>())) return false;
    for (int r = 0; r < kNumResiduals; ++r) residuals[r] = out[r].a;
    o = 0;
    for (int b = 0; b < kBlocks; ++b) {
      if (jacobians[b]) for (int r = 0; r < kNumResiduals; ++r) for (int i = 0; i < n[b]; ++i) jacobians[b][r * n[b] + i] = out[r].v[o + i];
      o += n[b];
    }
    return true;
  }
};

// ---- ceres::Problem as a RECORDER: it keeps the residual blocks exactly as the caller registered them (cost function, loss function or null, parameter block
// pointers in call order) and the constant-block marks, so a test can replay them.  ceres::Solve is NOT reproduced: it returns without touching the parameters
// and reports an unusable solution, which is what the stand-in deserves.
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum PreconditionerType { IDENTITY, JACOBI, SCHUR_JACOBI, CLUSTER_JACOBI, CLUSTER_TRIDIAGONAL };
enum SparseLinearAlgebraLibraryType { SUITE_SPARSE, CX_SPARSE, EIGEN_SPARSE, ACCELERATE_SPARSE, NO_SPARSE };
inline bool IsSparseLinearAlgebraLibraryTypeAvailable(SparseLinearAlgebraLibraryType) { return false; }
class Problem {
 public:
  struct Options { EvaluationCallback* evaluation_callback = nullptr; };
  Options options;
  Problem() {}
  explicit Problem(const Options& o) : options(o) {}
  struct Block { CostFunction* cost; LossFunction* loss; std::vector<double*> params; };
  std::vector<Block> blocks;
  std::vector<const double*> constant_blocks;
  template <typename... P> void* AddResidualBlock(CostFunction* cost, LossFunction* loss, P*... p) { blocks.push_back(Block{cost, loss, std::vector<double*>{p...}}); return nullptr; }
  void AddParameterBlock(double*, int) {}
  void SetParameterBlockConstant(const double* p) { constant_blocks.push_back(p); }
  void SetParameterBlockVariable(double*) {}
  int NumResidualBlocks() const { return (int)blocks.size(); }
  int NumResiduals() const { int n = 0; for (const Block& b : blocks) n += b.cost->num_residuals(); return n; }
  ~Problem() { for (Block& b : blocks) delete b.cost; }   // Ceres owns the cost functions (loss functions are shared between blocks: left to leak here)
};
class Solver {
 public:
  struct Options {
    bool minimizer_progress_to_stdout = false, update_state_every_iteration = false;
    int num_threads = 1, max_num_iterations = 50, max_linear_solver_iterations = 500;
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY; PreconditionerType preconditioner_type = JACOBI;
    SparseLinearAlgebraLibraryType sparse_linear_algebra_library_type = NO_SPARSE;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  };
  struct Summary {
    double initial_cost = -1, final_cost = -1; int num_successful_steps = 0, num_unsuccessful_steps = 0;
    bool usable = false;
    bool IsSolutionUsable() const { return usable; }
    std::string BriefReport() const { return "oracle/shim: ceres::Solve is not reproduced"; }
    std::string FullReport() const { return BriefReport(); }
  };
};
// ceres::Solve: the solver is NOT reproduced.  A test harness may install a hook that LOOKS at the problem the caller assembled (and fills the summary);
// without a hook the call returns an unusable summary and leaves the parameters alone.
typedef void (*SolveHook)(const Solver::Options&, Problem*, Solver::Summary*);
inline SolveHook& solve_hook() { static SolveHook h = nullptr; return h; }
inline void Solve(const Solver::Options& o, Problem* p, Solver::Summary* s) { if (solve_hook()) solve_hook()(o, p, s); }
}  // namespace ceres
