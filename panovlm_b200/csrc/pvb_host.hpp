// panovlm_b200 — host-side (C++) helpers of the hot path: pose preparation and the trust-region loop that
// consumes the device-reduced normal equations.  Mirrors the roles of lidar_mapping/LidarOdometry.cpp:15-114
// (RefinePose: pose -> angle-axis blocks, first valid frame constant, ceres::Solve, write back) and
// util/Optimization.cpp:638-666 (SetOptionsLidar: 20 iterations, dense/sparse Schur by frame count).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>
#include "pvb_math.cuh"

namespace pvb {

// Rodrigues, same closed form Ceres documents for AngleAxisToRotationMatrix; output row-major R(aa).
inline void aa_to_R(const double aa[3], double R[9]) {
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (th2 > DBL_EPSILON) {
    const double th = std::sqrt(th2);
    const double wx = aa[0] / th, wy = aa[1] / th, wz = aa[2] / th;
    const double c = std::cos(th), s = std::sin(th), k = 1.0 - c;
    R[0] = c + wx * wx * k;      R[1] = wx * wy * k - wz * s; R[2] = wy * s + wx * wz * k;
    R[3] = wz * s + wx * wy * k; R[4] = c + wy * wy * k;      R[5] = -wx * s + wy * wz * k;
    R[6] = -wy * s + wx * wz * k; R[7] = wx * s + wy * wz * k; R[8] = c + wz * wz * k;
  } else {
    R[0] = 1.0;    R[1] = -aa[2]; R[2] = aa[1];
    R[3] = aa[2];  R[4] = 1.0;    R[5] = -aa[0];
    R[6] = -aa[1]; R[7] = aa[0];  R[8] = 1.0;
  }
}

// SO(3) left Jacobian J_l(a) = I + (1-cos t)/t^2 [a]x + (t - sin t)/t^3 [a]x^2 with a series below 0.1 rad.
inline void left_jacobian(const double aa[3], double Jl[9]) {
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  double A, B;
  if (th2 < 1e-2) {
    A = 0.5 - th2 / 24.0 + th2 * th2 / 720.0 - th2 * th2 * th2 / 40320.0;
    B = 1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0 - th2 * th2 * th2 / 362880.0;
  } else {
    const double th = std::sqrt(th2);
    A = (1.0 - std::cos(th)) / th2;
    B = (th - std::sin(th)) / (th2 * th);
  }
  const double x = aa[0], y = aa[1], z = aa[2];
  // K = [a]x ; K^2 = a a^T - |a|^2 I
  Jl[0] = 1.0 + B * (x * x - th2); Jl[1] = -A * z + B * x * y;      Jl[2] = A * y + B * x * z;
  Jl[3] = A * z + B * x * y;       Jl[4] = 1.0 + B * (y * y - th2); Jl[5] = -A * x + B * y * z;
  Jl[6] = -A * y + B * x * z;      Jl[7] = A * x + B * y * z;       Jl[8] = 1.0 + B * (z * z - th2);
}

inline void prepare_pose(const double* pose6, PosePrep& p) {
  aa_to_R(pose6, p.R);
  left_jacobian(pose6, p.Jl);
  p.t[0] = pose6[3]; p.t[1] = pose6[4]; p.t[2] = pose6[5];
}

// T_wl from the (aa_lw, t_lw) block: R_wl = R_lw^T, t_wl = -R_wl t_lw (LidarOdometry.cpp:104-108)
inline void world_pose(const PosePrep& p, double R_wl[9], double t_wl[3]) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R_wl[r * 3 + c] = p.R[c * 3 + r];
  for (int r = 0; r < 3; ++r) t_wl[r] = -(R_wl[r * 3] * p.t[0] + R_wl[r * 3 + 1] * p.t[1] + R_wl[r * 3 + 2] * p.t[2]);
}

// ---- reduced system handed back by the device: per pose-graph edge 12x12 upper + gradient -------------------
// edge_sys layout (92 doubles): H upper row-major (78) | g (12) | cost | n_residuals
constexpr int kEdgeSys = 92;
constexpr int kFrameSys = 29;   // single-pose (ref constant) form: H upper 6x6 (21) | g (6) | cost | n_residuals

struct LMOptions { int max_iterations = 20; double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8; };
struct LMSummary { double initial_cost = 0, final_cost = 0; int iterations = 0, successful = 0, unsuccessful = 0, termination = 0; };

// In-place lower Cholesky + solve.  Right-looking blocked factorisation (64-wide panels, OpenMP over the rows of the panel
// solve and of the trailing update) so that pose graphs of a few hundred frames (6 N unknowns) factor in well under a second.
inline bool cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  const int NB = 64;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int k1 = std::min(n, k0 + NB);
    for (int j = k0; j < k1; ++j) {                      // factor the diagonal block (unblocked, rows k0..k1)
      double d = A[(size_t)j * n + j];
      for (int k = k0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
      if (!(d > 0.0)) return false;
      d = std::sqrt(d); A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k1; ++i) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / d;
      }
    }
#pragma omp parallel for schedule(static) if (n > 512)
    for (int i = k1; i < n; ++i) {                       // panel: L[i, k0:k1] = A[i, k0:k1] * L[k0:k1, k0:k1]^-T
      for (int j = k0; j < k1; ++j) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / A[(size_t)j * n + j];
      }
    }
#pragma omp parallel for schedule(dynamic, 8) if (n > 512)
    for (int i = k1; i < n; ++i) {                       // trailing update (lower part): A[i, j] -= L[i, k0:k1] . L[j, k0:k1]
      const double* li = &A[(size_t)i * n + k0];
      for (int j = k1; j <= i; ++j) {
        const double* lj = &A[(size_t)j * n + k0];
        double s = 0.0;
        for (int k = 0; k < k1 - k0; ++k) s += li[k] * lj[k];
        A[(size_t)i * n + j] -= s;
      }
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return true;
}

// Dense trust-region Levenberg-Marquardt with Ceres 2.0's documented defaults (radius 1e4, Jacobi scaling fixed
// from the first Jacobian, diagonal clamp [1e-6, 1e32], radius update 1/max(1/3, 1-(2rho-1)^3), halving on
// rejection with doubling decrease factor, min_relative_decrease 1e-3).
// eval(poses, H, g) returns the cost and, when H/g are non-null, fills the dense (6nb x 6nb) JtJ and Jt r.
typedef std::function<double(const double* poses, double* H, double* g)> EvalFn;

inline LMSummary solve_lm(const EvalFn& eval, double* poses, int nb, const unsigned char* is_const, const LMOptions& opt) {
  const int D = 6 * nb;
  std::vector<int> fi;
  for (int b = 0; b < nb; ++b) if (!is_const || !is_const[b]) for (int k = 0; k < 6; ++k) fi.push_back(6 * b + k);
  const int n = (int)fi.size();
  std::vector<double> H((size_t)D * D), g(D), Hs((size_t)n * n), gs(n), sc(n), A((size_t)n * n), Hsc((size_t)n * n), y(n), cand(D);
  LMSummary S;
  double cost = eval(poses, H.data(), g.data());
  S.initial_cost = cost;
  auto gather = [&]() {
#pragma omp parallel for schedule(static) if (n > 512)
    for (int i = 0; i < n; ++i) { gs[i] = g[fi[i]]; for (int j = 0; j < n; ++j) Hs[(size_t)i * n + j] = H[(size_t)fi[i] * D + fi[j]]; }
  };
  auto gmax = [&]() { double m = 0; for (int i = 0; i < n; ++i) m = std::max(m, std::fabs(gs[i])); return m; };
  gather();
  for (int i = 0; i < n; ++i) sc[i] = 1.0 / (1.0 + std::sqrt(Hs[(size_t)i * n + i]));
  double radius = 1e4, decrease = 2.0;
  int invalid = 0;
  if (n == 0 || gmax() <= opt.gradient_tolerance) { S.final_cost = cost; S.termination = 2; return S; }
  for (int it = 1; it <= opt.max_iterations; ++it) {
    S.iterations = it;
#pragma omp parallel for schedule(static) if (n > 512)
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { const double v = Hs[(size_t)i * n + j] * sc[i] * sc[j]; Hsc[(size_t)i * n + j] = v; A[(size_t)i * n + j] = v; }
    for (int i = 0; i < n; ++i) { y[i] = -gs[i] * sc[i]; A[(size_t)i * n + i] += std::min(std::max(A[(size_t)i * n + i], 1e-6), 1e32) / radius; }
    bool ok = cholesky_solve(A, n, y);
    double model = 0;
    if (ok) {
      std::vector<double> term(n);
#pragma omp parallel for schedule(static) if (n > 512)
      for (int i = 0; i < n; ++i) { double hy = 0; for (int j = 0; j < n; ++j) hy += Hsc[(size_t)i * n + j] * y[j]; term[i] = y[i] * (gs[i] * sc[i] + 0.5 * hy); }
      for (int i = 0; i < n; ++i) model -= term[i];        // fixed summation order
      ok = model > 0.0;
    }
    if (!ok) { radius *= 0.5; S.unsuccessful++; if (++invalid >= 5 || radius < 1e-32) { S.termination = 4; break; } continue; }
    invalid = 0;
    double sn = 0, xn = 0;
    std::copy(poses, poses + D, cand.begin());
    for (int i = 0; i < n; ++i) { const double d = y[i] * sc[i]; cand[fi[i]] += d; sn += d * d; xn += poses[fi[i]] * poses[fi[i]]; }
    sn = std::sqrt(sn); xn = std::sqrt(xn);
    const double new_cost = eval(cand.data(), nullptr, nullptr);
    if (sn <= opt.parameter_tolerance * (xn + opt.parameter_tolerance)) { S.termination = 3; break; }
    const double change = cost - new_cost;
    if (std::fabs(change) <= opt.function_tolerance * cost) { S.termination = 1; break; }
    const double rho = change / model;
    if (rho > 1e-3) {
      std::copy(cand.begin(), cand.end(), poses);
      cost = eval(poses, H.data(), g.data());
      gather();
      S.successful++;
      radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease = 2.0;
      if (gmax() <= opt.gradient_tolerance) { S.termination = 2; break; }
    } else {
      radius /= decrease; decrease *= 2.0; S.unsuccessful++;
      if (radius < 1e-32) { S.termination = 4; break; }
    }
  }
  S.final_cost = cost;
  return S;
}

}  // namespace pvb
