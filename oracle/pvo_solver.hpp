// ORACLE — TEST INFRASTRUCTURE ONLY (see pvo_math.hpp header).  Parity status: unpinned (Ceres is not
// available here); restated from Ceres 2.0's documented trust-region Levenberg-Marquardt behaviour.
//
// Residual-block list + dense Levenberg-Marquardt, restating what the reference hands to Ceres:
//   util/Optimization.cpp:506-562 (blocks), :638-666 (SetOptionsLidar: max_num_iterations = 20),
//   lidar_mapping/LidarOdometry.cpp:59-66 (first valid frame constant), :78-80 (ceres::Solve).
// Every residual is evaluated ONE AutoDiffCostFunction AT A TIME with a 12-wide Jet, as Ceres does.
#pragma once
#include "pvo_math.hpp"

namespace pvo {

enum BlockType { P2PLANE_METER = 0, P2PLANE_ANGLE = 1, P2LINE_METER = 2, P2LINE_ANGLE = 3, PLANE2PLANE_GLOBAL = 4, PLANE_IOU = 5,
                 PLANE2PLANE_RELATIVE = 6, PLANE_RELATIVE_IOU = 7, LINE2LINE_ANGLE = 8 };

// consts layout (12 doubles):
//  P2PLANE_*          : p[0..2] plane[3..6] weight[7]
//  P2LINE_*           : p[0..2] a[3..5] dir[6..8] weight[9]        (dir already normalised (a-b)/|a-b|)
//  PLANE2PLANE_GLOBAL : n[0..2] (normalised) a[3..5] b[6..8] weight[9]
//  PLANE_IOU          : plane[0..3] (normalised) mid_nei[4..6] mid_ref[7..9] angle[10] weight[11]
//  PLANE2PLANE_RELATIVE / PLANE_RELATIVE_IOU : consts as PLANE2PLANE_GLOBAL / PLANE_IOU; parameters = the `ref` block only
//                       (aa_cl, t_cl); the `nei` block is ignored and its Jacobian columns are zero
//  LINE2LINE_ANGLE    : dir_ref[0..2] dir_nei[3..5] (unit); parameters = aa of the ref block and aa of the nei block
struct Block { int type, ref, nei, normalize; double huber; double c[12]; };

inline void EvalBlockRaw(const Block& b, const double* pr, const double* pn, double* r, double* J) {
  const double *aa_r = pr, *t_r = pr + 3, *aa_n = pn, *t_n = pn + 3;
  switch (b.type) {
    case P2PLANE_METER: { Point2Plane_Meter f; std::memcpy(f.p, b.c, 24); std::memcpy(f.plane, b.c + 3, 32); f.weight = b.c[7]; EvaluateAutoDiff4(f, aa_r, t_r, aa_n, t_n, r, J); break; }
    case P2PLANE_ANGLE: { Point2Plane_Angle f; std::memcpy(f.p, b.c, 24); std::memcpy(f.plane, b.c + 3, 32); f.weight = b.c[7]; f.normalize_distance = b.normalize != 0; EvaluateAutoDiff4(f, aa_r, t_r, aa_n, t_n, r, J); break; }
    case P2LINE_METER: { Point2Line_Meter f; std::memcpy(f.p, b.c, 24); std::memcpy(f.line_point, b.c + 3, 24); std::memcpy(f.line_dir, b.c + 6, 24); f.weight = b.c[9]; EvaluateAutoDiff4(f, aa_r, t_r, aa_n, t_n, r, J); break; }
    case P2LINE_ANGLE: { Point2Line_Angle f; std::memcpy(f.p, b.c, 24); std::memcpy(f.line_point, b.c + 3, 24); std::memcpy(f.line_dir, b.c + 6, 24); f.weight = b.c[9]; f.normalize_distance = b.normalize != 0; EvaluateAutoDiff4(f, aa_r, t_r, aa_n, t_n, r, J); break; }
    case PLANE2PLANE_GLOBAL: { Plane2Plane_Global f; std::memcpy(f.plane_ref, b.c, 24); std::memcpy(f.point_a, b.c + 3, 24); std::memcpy(f.point_b, b.c + 6, 24); f.weight = b.c[9]; EvaluateAutoDiff4(f, aa_r, t_r, aa_n, t_n, r, J); break; }
    case PLANE_IOU: { PlaneIOUResidual f; std::memcpy(f.ref_plane, b.c, 32); std::memcpy(f.middle_neighbor, b.c + 4, 24); std::memcpy(f.middle_ref, b.c + 7, 24); f.angle = b.c[10]; f.weight = b.c[11]; EvaluateAutoDiff4(f, aa_r, t_r, aa_n, t_n, r, J); break; }
    case PLANE2PLANE_RELATIVE: { Plane2Plane_Relative f; std::memcpy(f.plane_ref, b.c, 24); std::memcpy(f.point_a, b.c + 3, 24); std::memcpy(f.point_b, b.c + 6, 24); f.weight = b.c[9];
      if (J) for (int i = 6; i < 12; ++i) J[i] = 0; EvaluateAutoDiff2(f, aa_r, t_r, r, J); break; }
    case PLANE_RELATIVE_IOU: { PlaneRelativeIOUResidual f; std::memcpy(f.ref_plane, b.c, 32); std::memcpy(f.middle_neighbor, b.c + 4, 24); std::memcpy(f.middle_ref, b.c + 7, 24); f.angle = b.c[10]; f.weight = b.c[11];
      if (J) for (int i = 6; i < 12; ++i) J[i] = 0; EvaluateAutoDiff2(f, aa_r, t_r, r, J); break; }
    case LINE2LINE_ANGLE: { Line2Line_Angle f; std::memcpy(f.dir_ref, b.c, 24); std::memcpy(f.dir_nei, b.c + 3, 24);
      double j6[6]; EvaluateAutoDiff2(f, aa_r, aa_n, r, J ? j6 : nullptr);
      if (J) { for (int i = 0; i < 12; ++i) J[i] = 0; for (int i = 0; i < 3; ++i) { J[i] = j6[i]; J[6 + i] = j6[3 + i]; } } break; }
    default: *r = 0; if (J) for (int i = 0; i < 12; ++i) J[i] = 0;
  }
}

// residual + Jacobian after the robust-loss corrector; returns the block's cost 0.5*rho(r^2)
inline double EvalBlock(const Block& b, const double* poses, double* r, double* J, bool apply_loss) {
  EvalBlockRaw(b, poses + 6 * b.ref, poses + 6 * b.nei, r, J);
  double cost;
  HuberCorrect(apply_loss ? b.huber : 0.0, r, J, 12, &cost);
  return cost;
}

// Dense normal equations over all blocks: H (6nb x 6nb, row-major, full symmetric), g = J^T r, cost.
inline double NormalEquations(const Block* blocks, long n, const double* poses, int nb, double* H, double* g) {
  const int D = 6 * nb;
  if (H) std::fill(H, H + (size_t)D * D, 0.0);
  if (g) std::fill(g, g + D, 0.0);
  double cost = 0;
  for (long i = 0; i < n; ++i) {
    double r, J[12];
    cost += EvalBlock(blocks[i], poses, &r, (H || g) ? J : nullptr, true);
    if (!(H || g)) continue;
    const int o[2] = {6 * blocks[i].ref, 6 * blocks[i].nei};
    for (int a = 0; a < 12; ++a) {
      const int ia = o[a / 6] + a % 6;
      if (g) g[ia] += J[a] * r;
      if (H) for (int c = 0; c < 12; ++c) H[(size_t)ia * D + o[c / 6] + c % 6] += J[a] * J[c];
    }
  }
  return cost;
}

inline bool CholeskySolve(std::vector<double>& A, int n, std::vector<double>& b) {  // in place, lower
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d); A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return true;
}

struct LMSummary { double initial_cost, final_cost; int iterations, successful, unsuccessful, termination; };
// termination: 0 = max iterations, 1 = function tolerance, 2 = gradient tolerance, 3 = parameter tolerance, 4 = failure

// Ceres 2.0 TrustRegionMinimizer + LevenbergMarquardtStrategy defaults: initial radius 1e4, max 1e16,
// min 1e-32, min_relative_decrease 1e-3, min/max_lm_diagonal 1e-6/1e32, jacobi_scaling (fixed from the
// first Jacobian: 1/(1+sqrt(colnorm^2))), function/gradient/parameter tolerance 1e-6/1e-10/1e-8.
// Generic form: D parameters, param_const[i] != 0 pins parameter i (Problem::SetParameterBlockConstant on the block holding it).
template <typename EvalFn>  // double eval(const double* x, double* H, double* g)  (H,g may be null; H is D x D, g is D)
inline LMSummary SolveLMParams(EvalFn eval, double* poses, int D, const unsigned char* param_const, int max_iter) {
  std::vector<int> freeidx;
  for (int i = 0; i < D; ++i) if (!param_const || !param_const[i]) freeidx.push_back(i);
  const int n = (int)freeidx.size();
  std::vector<double> H((size_t)D * D), g(D), Hs((size_t)n * n), gs(n), scale(n), A((size_t)n * n), step(n), cand(poses, poses + D);
  LMSummary S{}; S.termination = 0;
  double cost = eval(poses, H.data(), g.data());
  S.initial_cost = cost;
  auto gather = [&]() {
    for (int i = 0; i < n; ++i) { gs[i] = g[freeidx[i]]; for (int j = 0; j < n; ++j) Hs[(size_t)i * n + j] = H[(size_t)freeidx[i] * D + freeidx[j]]; }
  };
  gather();
  for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(Hs[(size_t)i * n + i]));
  auto gmax = [&]() { double m = 0; for (int i = 0; i < n; ++i) m = std::max(m, std::fabs(gs[i])); return m; };
  double radius = 1e4, decrease_factor = 2.0;
  int invalid = 0;
  if (n == 0 || gmax() <= 1e-10) { S.final_cost = cost; S.termination = 2; return S; }
  // Loop order follows Ceres 2.0 TrustRegionMinimizer::Minimize: step -> (invalid? shrink) -> candidate
  // cost -> parameter tolerance -> function tolerance -> accept/reject -> gradient tolerance.
  for (int iter = 1; iter <= max_iter; ++iter) {
    S.iterations = iter;
    // scaled system  (S H S + diag(clamp(diag(S H S))/radius)) y = -S g
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = Hs[(size_t)i * n + j] * scale[i] * scale[j];
    std::vector<double> Hscaled(A);
    for (int i = 0; i < n; ++i) { step[i] = -gs[i] * scale[i]; A[(size_t)i * n + i] += std::min(std::max(A[(size_t)i * n + i], 1e-6), 1e32) / radius; }
    bool ok = CholeskySolve(A, n, step);
    double model_change = 0;
    if (ok) {
      // model_cost_change = -y^T (S g + 0.5 * S H S y)
      for (int i = 0; i < n; ++i) { double hy = 0; for (int j = 0; j < n; ++j) hy += Hscaled[(size_t)i * n + j] * step[j]; model_change -= step[i] * (gs[i] * scale[i] + 0.5 * hy); }
      ok = model_change > 0.0;
    }
    if (!ok) {  // LevenbergMarquardtStrategy::StepIsInvalid: radius *= 0.5; 5 consecutive => failure
      radius *= 0.5; S.unsuccessful++;
      if (++invalid >= 5 || radius < 1e-32) { S.termination = 4; break; }
      continue;
    }
    invalid = 0;
    double step_norm = 0, x_norm = 0;
    cand.assign(poses, poses + D);
    for (int i = 0; i < n; ++i) { const double d = step[i] * scale[i]; cand[freeidx[i]] += d; step_norm += d * d; x_norm += poses[freeidx[i]] * poses[freeidx[i]]; }
    step_norm = std::sqrt(step_norm); x_norm = std::sqrt(x_norm);
    const double new_cost = eval(cand.data(), nullptr, nullptr);
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { S.termination = 3; break; }
    const double cost_change = cost - new_cost;
    if (std::fabs(cost_change) <= 1e-6 * cost) { S.termination = 1; break; }
    const double rho = cost_change / model_change;
    if (rho > 1e-3) {
      std::copy(cand.begin(), cand.end(), poses);
      cost = eval(poses, H.data(), g.data());
      gather();
      S.successful++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(1e16, radius); decrease_factor = 2.0;
      if (gmax() <= 1e-10) { S.termination = 2; break; }
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0; S.unsuccessful++;
      if (radius < 1e-32) { S.termination = 4; break; }
    }
  }
  S.final_cost = cost;
  return S;
}

// pose-block form used by the LiDAR problems: nb blocks of 6, is_const per block
template <typename EvalFn>
inline LMSummary SolveLM(EvalFn eval, double* poses, int nb, const unsigned char* is_const, int max_iter) {
  std::vector<unsigned char> mask((size_t)6 * nb, 0);
  if (is_const) for (int b = 0; b < nb; ++b) for (int k = 0; k < 6; ++k) mask[6 * b + k] = is_const[b];
  return SolveLMParams(eval, poses, 6 * nb, mask.data(), max_iter);
}

}  // namespace pvo
