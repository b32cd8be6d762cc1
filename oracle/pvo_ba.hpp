// ORACLE — TEST INFRASTRUCTURE ONLY (see pvo_math.hpp header).  Parity status: unpinned by the reference (no tests exist); the functor is
// pinned against a torch float64 autograd twin and finite differences in tests/test_reproj.py.
//
// Camera-camera reprojection residuals of the joint problem:
//   PanoramaReprojResidual_1Angle   base/CostFunction.h:218-247  (AutoDiffCostFunction<F,1,3,3,3>: aa_cw, t_cw, point_3d)
//   AddCameraResidual               util/Optimization.cpp:172-222 (ANGLE_RESIDUAL_1 branch: one block per (track, observation), HuberLoss(4 deg))
//   SfMGlobalBA                     util/Optimization.cpp:10-82   (first valid camera constant; rotations / translations / structure optional)
// evaluated one autodiff functor at a time with a 9-wide Jet, assembled into dense normal equations over [cameras (6 each) | points (3 each)].
#pragma once
#include "pvo_solver.hpp"

namespace pvo {

struct PanoramaReprojResidual_1Angle {  // CostFunction.h:218-247; the constructor normalises the bearing (:225)
  double point_sphere[3];
  double weight;
  template <typename T>
  bool operator()(const T* const angleAxis_cw, const T* const t_cw, const T* const point_3d, T* residuals) const {
    T point_c[3];
    AngleAxisRotatePoint(angleAxis_cw, point_3d, point_c);
    point_c[0] += t_cw[0];
    point_c[1] += t_cw[1];
    point_c[2] += t_cw[2];
    T norm = sqrt(point_c[0] * point_c[0] + point_c[1] * point_c[1] + point_c[2] * point_c[2]);
    T dot_product = point_c[0] * T(point_sphere[0]) + point_c[1] * T(point_sphere[1]) + point_c[2] * T(point_sphere[2]);
    residuals[0] = T(weight) * acos(dot_product / norm);
    return true;
  }
};

struct ReprojObs { int cam, point; double bearing[3]; };   // bearing = eq.ImageToCam(keypoint) (Optimization.cpp:205), normalised by the functor's constructor

inline void MakeReprojFunctor(const ReprojObs& o, double weight, PanoramaReprojResidual_1Angle& f) {
  const double n = std::sqrt(o.bearing[0] * o.bearing[0] + o.bearing[1] * o.bearing[1] + o.bearing[2] * o.bearing[2]);
  for (int k = 0; k < 3; ++k) f.point_sphere[k] = o.bearing[k] / n;       // Vector3d::normalize()
  f.weight = weight;
}

// residual + 1x9 Jacobian [d aa_cw | d t_cw | d point] of one observation (before the loss)
inline void EvalReprojRaw(const ReprojObs& o, double weight, const double* cams, const double* points, double* r, double* J9) {
  PanoramaReprojResidual_1Angle f; MakeReprojFunctor(o, weight, f);
  const double* aa = cams + 6 * o.cam; const double* t = aa + 3; const double* X = points + 3 * o.point;
  if (!J9) { double v; f(aa, t, X, &v); *r = v; return; }
  Jet<9> ja[3], jt[3], jx[3], res;
  for (int k = 0; k < 3; ++k) { ja[k] = Jet<9>(aa[k], k); jt[k] = Jet<9>(t[k], 3 + k); jx[k] = Jet<9>(X[k], 6 + k); }
  f(ja, jt, jx, &res);
  *r = res.a;
  for (int k = 0; k < 9; ++k) J9[k] = res.v[k];
}

inline double EvalReproj(const ReprojObs& o, double weight, double huber, const double* cams, const double* points, double* r, double* J9) {
  EvalReprojRaw(o, weight, cams, points, r, J9);
  double cost;
  HuberCorrect(huber, r, J9, 9, &cost);
  return cost;
}

// dense normal equations over x = [cams (6 nc) | points (3 np)]
inline double ReprojNormalEquations(const ReprojObs* obs, long n, double weight, double huber, const double* x, int nc, long np, double* H, double* g) {
  const long D = 6L * nc + 3L * np;
  if (H) std::fill(H, H + (size_t)D * D, 0.0);
  if (g) std::fill(g, g + D, 0.0);
  double cost = 0;
  for (long i = 0; i < n; ++i) {
    double r, J[9];
    cost += EvalReproj(obs[i], weight, huber, x, x + 6L * nc, &r, (H || g) ? J : nullptr);
    if (!(H || g)) continue;
    long idx[9];
    for (int k = 0; k < 6; ++k) idx[k] = 6L * obs[i].cam + k;
    for (int k = 0; k < 3; ++k) idx[6 + k] = 6L * nc + 3L * obs[i].point + k;
    for (int a = 0; a < 9; ++a) {
      if (g) g[idx[a]] += J[a] * r;
      if (H) for (int c = 0; c < 9; ++c) H[(size_t)idx[a] * D + idx[c]] += J[a] * J[c];
    }
  }
  return cost;
}

}  // namespace pvo
