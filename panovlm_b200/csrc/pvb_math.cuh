// panovlm_b200 — device math for the correspondence-and-residual hot path (sm_100a).
//
// Everything here is `__host__ __device__` so that tests/host_harness.cpp can exercise the exact same
// arithmetic on the CPU (test infrastructure only; the shipped library has no CPU path).
//
// What the reference does (base/CostFunction.h:567-1022, one ceres::AutoDiffCostFunction<F,1,3,3,3,3> at a
// time with Jet<double,12>) is re-derived here analytically:
//   P = R_r R_n^T (p - t_n) + t_r                                  (CostFunction.h:584-604 via angle-axis detour)
//   dP/da_r = -[R_r u]x J_l(a_r)   dP/dt_r = I   dP/da_n = R_r R_n^T [q]x J_l(a_n)   dP/dt_n = -R_r R_n^T
// with q = p - t_n, u = R_n^T q and J_l the SO(3) left Jacobian (Ceres differentiates w.r.t. the raw
// angle-axis vector, no manifold).  The scalar tail r(P) is evaluated with a 3-wide (6-wide for the two
// point Plane2Plane_Global) forward dual so every autodiff branch (zero rows below 1e-3, abs' = +-1,
// clamped acos) is reproduced exactly.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PVB_HD __host__ __device__ __forceinline__
#else
#define PVB_HD inline
#endif

namespace pvb {

enum BlockType : int { P2PLANE_METER = 0, P2PLANE_ANGLE = 1, P2LINE_METER = 2, P2LINE_ANGLE = 3, PLANE2PLANE_GLOBAL = 4, PLANE_IOU = 5,
                       PLANE2PLANE_RELATIVE = 6, PLANE_RELATIVE_IOU = 7, LINE2LINE_ANGLE = 8 };

// ---- rounding-exact float/double helpers (no FMA contraction: the reference's PCL/FLANN/OpenCV float paths
//      are plain mul+add; the host build uses -ffp-contract=off) ------------------------------------------
PVB_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
PVB_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
PVB_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
PVB_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
PVB_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
PVB_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}

// ---- per pose-block quantities prepared on the host once per evaluation -----------------------------------
struct PosePrep {
  double R[9];    // R_lw (world -> frame), row-major
  double Jl[9];   // SO(3) left Jacobian of aa_lw, row-major
  double t[3];    // t_lw
};

// pcl::transformPointCloud(cloud, out, Matrix4d) (sensors/Velodyne.cpp:1790-1806): double math, float32 store.
PVB_HD void transform_point_f32(const double* R_rowmajor, const double* t, float x, float y, float z, float& ox, float& oy, float& oz) {
  const double px = x, py = y, pz = z;
  ox = (float)dadd(dadd(dadd(dmul(R_rowmajor[0], px), dmul(R_rowmajor[1], py)), dmul(R_rowmajor[2], pz)), t[0]);
  oy = (float)dadd(dadd(dadd(dmul(R_rowmajor[3], px), dmul(R_rowmajor[4], py)), dmul(R_rowmajor[5], pz)), t[1]);
  oz = (float)dadd(dadd(dadd(dmul(R_rowmajor[6], px), dmul(R_rowmajor[7], py)), dmul(R_rowmajor[8], pz)), t[2]);
}

// Velodyne::World2Local (sensors/Velodyne.cpp:1850-1853): R_wl^T p - R_wl^T t_wl, each product summed x,y,z.
PVB_HD void world2local(const double* R_wl, const double* t_wl, const double pw[3], double out[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double a = dadd(dadd(dmul(R_wl[0 * 3 + r], pw[0]), dmul(R_wl[1 * 3 + r], pw[1])), dmul(R_wl[2 * 3 + r], pw[2]));
    const double b = dadd(dadd(dmul(R_wl[0 * 3 + r], t_wl[0]), dmul(R_wl[1 * 3 + r], t_wl[1])), dmul(R_wl[2 * 3 + r], t_wl[2]));
    out[r] = dsub(a, b);
  }
}

// flann::L2_Simple<float> squared distance (the metric behind pcl::KdTreeFLANN::nearestKSearch)
PVB_HD float sqdist_f32(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = fsub(ax, bx), dy = fsub(ay, by), dz = fsub(az, bz);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// ---- forward dual number (N = 3: d/dP, N = 6: d/d(A,B)) ---------------------------------------------------
template <int N>
struct Dual {
  double v;
  double d[N];
};
template <int N> PVB_HD Dual<N> mk(double v) { Dual<N> r; r.v = v; for (int i = 0; i < N; ++i) r.d[i] = 0.0; return r; }
template <int N> PVB_HD Dual<N> seed(double v, int k) { Dual<N> r = mk<N>(v); r.d[k] = 1.0; return r; }
template <int N> PVB_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> PVB_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> PVB_HD Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> PVB_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v; return r; }
template <int N> PVB_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; const double inv = 1.0 / b.v; r.v = a.v * inv;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int N> PVB_HD Dual<N> operator+(const Dual<N>& a, double s) { Dual<N> r = a; r.v += s; return r; }
template <int N> PVB_HD Dual<N> operator-(const Dual<N>& a, double s) { Dual<N> r = a; r.v -= s; return r; }
template <int N> PVB_HD Dual<N> operator*(const Dual<N>& a, double s) { Dual<N> r; r.v = a.v * s; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> PVB_HD Dual<N> operator*(double s, const Dual<N>& a) { return a * s; }
template <int N> PVB_HD Dual<N> dsqrt(const Dual<N>& a) { Dual<N> r; r.v = sqrt(a.v); const double k = 1.0 / (2.0 * r.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
template <int N> PVB_HD Dual<N> dacos(const Dual<N>& a) { Dual<N> r; r.v = acos(a.v); const double k = -1.0 / sqrt(1.0 - a.v * a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k; return r; }
template <int N> PVB_HD Dual<N> dabs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }   // ceres::abs(Jet): +1 slope at 0

// base/Geometry.hpp:450-466
template <int N> PVB_HD Dual<N> vector_angle(const Dual<N> a[3], const Dual<N> b[3]) {
  Dual<N> c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  const Dual<N> n1 = dsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const Dual<N> n2 = dsqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  c = c / (n1 * n2);
  if (c.v >= 1.0) return mk<N>(0.0);
  if (c.v <= -1.0) return mk<N>(M_PI);
  return dacos(c);
}

// normalize_distance tail of Point2Plane_Angle / Point2Line_Angle (CostFunction.h:699-715, 901-917)
template <int N> PVB_HD Dual<N> angle_tail(const Dual<N> P[3], const Dual<N> Pp[3], bool normalize) {
  if (normalize) {
    const Dual<N> nrm = dsqrt(Pp[0] * Pp[0] + Pp[1] * Pp[1] + Pp[2] * Pp[2]);
    const Dual<N> ratio = (nrm - 1.0) / nrm;
    Dual<N> c[3], v1[3], v2[3];
    for (int k = 0; k < 3; ++k) { c[k] = ratio * Pp[k]; v1[k] = Pp[k] - c[k]; v2[k] = P[k] - c[k]; }
    return vector_angle(v1, v2);
  }
  return vector_angle(P, Pp);
}

// ---- residual tails: r and g = dr/dP (before the robust loss) ---------------------------------------------
// consts layout per type: see include/panovlm_b200.h (identical to the reference functors' members).
PVB_HD double tail_point_plane(int type, bool normalize, const double* c, const double P[3], double g[3]) {
  typedef Dual<3> D;
  D p[3] = {seed<3>(P[0], 0), seed<3>(P[1], 1), seed<3>(P[2], 2)};
  const double n0 = c[3], n1 = c[4], n2 = c[5], d = c[6];
  D s = p[0] * n0 + p[1] * n1 + p[2] * n2 + d;
  D dis = dabs(s);                                              // PointToPlaneDistance(..., normalized=true)
  D r;
  if (type == P2PLANE_METER) {
    r = dis * c[7];                                             // CostFunction.h:607
  } else {
    if (dis.v < 1e-3) { g[0] = g[1] = g[2] = 0.0; return 0.0; }  // :680-684 zero residual AND zero row
    D pp[3] = {p[0] - dis * n0, p[1] - dis * n1, p[2] - dis * n2};
    const double chk = n0 * pp[0].v + n1 * pp[1].v + n2 * pp[2].v + d;
    if (fabs(chk) > 1e-4) { pp[0] = p[0] + dis * n0; pp[1] = p[1] + dis * n1; pp[2] = p[2] + dis * n2; }   // :688-693
    r = angle_tail(p, pp, normalize);
  }
  g[0] = r.d[0]; g[1] = r.d[1]; g[2] = r.d[2];
  return r.v;
}

PVB_HD double tail_point_line(int type, bool normalize, const double* c, const double P[3], double g[3]) {
  typedef Dual<3> D;
  D p[3] = {seed<3>(P[0], 0), seed<3>(P[1], 1), seed<3>(P[2], 2)};
  const double x0 = c[3], y0 = c[4], z0 = c[5], nx = c[6], ny = c[7], nz = c[8];
  D k = (p[0] - x0) * nx + (p[1] - y0) * ny + (p[2] - z0) * nz;
  if (type == P2LINE_METER) k = k / mk<3>(nx * nx + ny * ny + nz * nz);   // PointToLineDistance3D divides by |n|^2 (Geometry.hpp:207)
  D pp[3] = {k * nx + x0, k * ny + y0, k * nz + z0};
  D dx = p[0] - pp[0], dy = p[1] - pp[1], dz = p[2] - pp[2];
  D r;
  if (type == P2LINE_METER) {
    D ex = pp[0] - p[0], ey = pp[1] - p[1], ez = pp[2] - p[2];
    r = dsqrt(ex * ex + ey * ey + ez * ez) * c[9];              // CostFunction.h:815
  } else {
    D dis = dsqrt(dx * dx + dy * dy + dz * dz);                 // :889-891
    if (dis.v < 1e-3) { g[0] = g[1] = g[2] = 0.0; return 0.0; }
    r = angle_tail(p, pp, normalize);
  }
  g[0] = r.d[0]; g[1] = r.d[1]; g[2] = r.d[2];
  return r.v;
}

// Plane2Plane_Global (CostFunction.h:350-425): r = w * PlaneAngle(n_ref, A x B); gA/gB = dr/dA, dr/dB
PVB_HD double tail_plane2plane(const double* c, const double A[3], const double B[3], double gA[3], double gB[3]) {
  typedef Dual<6> D;
  D a[3] = {seed<6>(A[0], 0), seed<6>(A[1], 1), seed<6>(A[2], 2)};
  D b[3] = {seed<6>(B[0], 3), seed<6>(B[1], 4), seed<6>(B[2], 5)};
  D n1[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  D cs = dabs(n1[0] * c[0] + n1[1] * c[1] + n1[2] * c[2]);      // PlaneAngle(plane2, plane1): |dot|
  const double nr = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  D nn = dsqrt(n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2]);
  cs = cs / (nn * nr);                                          // Geometry.hpp:477-480 (norm1 * norm2, plane2 first)
  D r;
  if (cs.v >= 1.0) r = mk<6>(0.0); else r = dacos(cs);
  r = r * c[9];
  for (int k = 0; k < 3; ++k) { gA[k] = r.d[k]; gB[k] = r.d[3 + k]; }
  return r.v;
}

// PlaneIOUResidual (CostFunction.h:433-507)
PVB_HD double tail_plane_iou(const double* c, const double M[3], double g[3]) {
  typedef Dual<3> D;
  D m[3] = {seed<3>(M[0], 0), seed<3>(M[1], 1), seed<3>(M[2], 2)};
  const double n0 = c[0], n1 = c[1], n2 = c[2], d = c[3];
  D dis = dabs(m[0] * n0 + m[1] * n1 + m[2] * n2 + d);          // ProjectPointToPlane(normalized=true)
  D pp[3] = {m[0] - dis * n0, m[1] - dis * n1, m[2] - dis * n2};
  if (fabs(n0 * pp[0].v + n1 * pp[1].v + n2 * pp[2].v + d) > 1e-4) { pp[0] = m[0] + dis * n0; pp[1] = m[1] + dis * n1; pp[2] = m[2] + dis * n2; }
  D ip[3] = {mk<3>(c[7]), mk<3>(c[8]), mk<3>(c[9])};
  D ang = vector_angle(pp, ip);
  if (ang.v < c[10]) { g[0] = g[1] = g[2] = 0.0; return 0.0; }
  D r = (ang - c[10]) * c[11];
  g[0] = r.d[0]; g[1] = r.d[1]; g[2] = r.d[2];
  return r.v;
}

// ---- the common transform + analytic Jacobian row --------------------------------------------------------
PVB_HD void transform_nei_to_ref(const PosePrep& pr, const PosePrep& pn, const double p[3], double q[3], double P[3]) {
  q[0] = p[0] - pn.t[0]; q[1] = p[1] - pn.t[1]; q[2] = p[2] - pn.t[2];
  double u[3];
  for (int k = 0; k < 3; ++k) u[k] = pn.R[0 * 3 + k] * q[0] + pn.R[1 * 3 + k] * q[1] + pn.R[2 * 3 + k] * q[2];   // R_n^T q
  for (int k = 0; k < 3; ++k) P[k] = pr.R[k * 3 + 0] * u[0] + pr.R[k * 3 + 1] * u[1] + pr.R[k * 3 + 2] * u[2] + pr.t[k];
}

// accumulate J += g^T dP/d(params) for one transformed point (q = p - t_n, P = transformed)
PVB_HD void accumulate_row(const PosePrep& pr, const PosePrep& pn, const double q[3], const double P[3], const double g[3], double J[12]) {
  const double v[3] = {P[0] - pr.t[0], P[1] - pr.t[1], P[2] - pr.t[2]};
  const double a[3] = {v[1] * g[2] - v[2] * g[1], v[2] * g[0] - v[0] * g[2], v[0] * g[1] - v[1] * g[0]};          // v x g
  double w[3], h[3];
  for (int k = 0; k < 3; ++k) w[k] = pr.R[0 * 3 + k] * g[0] + pr.R[1 * 3 + k] * g[1] + pr.R[2 * 3 + k] * g[2];    // R_r^T g
  for (int k = 0; k < 3; ++k) h[k] = pn.R[k * 3 + 0] * w[0] + pn.R[k * 3 + 1] * w[1] + pn.R[k * 3 + 2] * w[2];    // R_n R_r^T g
  const double b[3] = {h[1] * q[2] - h[2] * q[1], h[2] * q[0] - h[0] * q[2], h[0] * q[1] - h[1] * q[0]};          // h x q
  for (int j = 0; j < 3; ++j) {
    J[j] += a[0] * pr.Jl[0 * 3 + j] + a[1] * pr.Jl[1 * 3 + j] + a[2] * pr.Jl[2 * 3 + j];
    J[3 + j] += g[j];
    J[6 + j] += b[0] * pn.Jl[0 * 3 + j] + b[1] * pn.Jl[1 * 3 + j] + b[2] * pn.Jl[2 * 3 + j];
    J[9 + j] -= h[j];
  }
}

PVB_HD PosePrep identity_prep() {
  PosePrep p;
  for (int k = 0; k < 9; ++k) { p.R[k] = (k % 4 == 0) ? 1.0 : 0.0; p.Jl[k] = p.R[k]; }
  p.t[0] = p.t[1] = p.t[2] = 0.0;
  return p;
}

// One residual block: raw residual + 1x12 Jacobian row [d/daa_r | d/dt_r | d/daa_n | d/dt_n].
PVB_HD double eval_block(int type, bool normalize, const double* c, const PosePrep& pr, const PosePrep& pn, double J[12]) {
  for (int k = 0; k < 12; ++k) J[k] = 0.0;
  double q[3], P[3], g[3], r;
  switch (type) {
    case P2PLANE_METER:
    case P2PLANE_ANGLE:
      transform_nei_to_ref(pr, pn, c, q, P);
      r = tail_point_plane(type, normalize, c, P, g);
      accumulate_row(pr, pn, q, P, g, J);
      return r;
    case P2LINE_METER:
    case P2LINE_ANGLE:
      transform_nei_to_ref(pr, pn, c, q, P);
      r = tail_point_line(type, normalize, c, P, g);
      accumulate_row(pr, pn, q, P, g, J);
      return r;
    case PLANE2PLANE_GLOBAL: {
      double qb[3], A[3], B[3], gA[3], gB[3];
      transform_nei_to_ref(pr, pn, c + 3, q, A);
      transform_nei_to_ref(pr, pn, c + 6, qb, B);
      r = tail_plane2plane(c, A, B, gA, gB);
      accumulate_row(pr, pn, q, A, gA, J);
      accumulate_row(pr, pn, qb, B, gB, J);
      return r;
    }
    case PLANE_IOU:
      transform_nei_to_ref(pr, pn, c + 4, q, P);
      r = tail_plane_iou(c, P, g);
      accumulate_row(pr, pn, q, P, g, J);
      return r;
    // Calibration-mode functors (CostFunction.h:294-346, 509-563): one relative pose (aa_cl, t_cl) = the `ref` block,
    // P = R p + t, i.e. the common transform with the neighbour block pinned at identity; its columns stay zero.
    case PLANE2PLANE_RELATIVE: {
      const PosePrep id = identity_prep();
      double qb[3], A[3], B[3], gA[3], gB[3];
      transform_nei_to_ref(pr, id, c + 3, q, A);
      transform_nei_to_ref(pr, id, c + 6, qb, B);
      r = tail_plane2plane(c, A, B, gA, gB);
      accumulate_row(pr, id, q, A, gA, J);
      accumulate_row(pr, id, qb, B, gB, J);
      const double deg = 180.0 / M_PI;                          // :334 residual in degrees ((w * angle) * 180) / pi
      r = r * 180.0 / M_PI;
      for (int k = 0; k < 6; ++k) { J[k] *= deg; J[6 + k] = 0.0; }
      return r;
    }
    case PLANE_RELATIVE_IOU: {
      const PosePrep id = identity_prep();
      transform_nei_to_ref(pr, id, c + 4, q, P);
      r = tail_plane_iou(c, P, g);
      accumulate_row(pr, id, q, P, g, J);
      for (int k = 6; k < 12; ++k) J[k] = 0.0;
      return r;
    }
    // Line2Line_Angle (CostFunction.h:984-1022): rotations only, r = PlaneAngle(R_r R_n^T d_nei, d_ref, normalized), < 1e-3 => 0
    case LINE2LINE_ANGLE: {
      PosePrep r0 = pr, n0 = pn;
      r0.t[0] = r0.t[1] = r0.t[2] = 0.0; n0.t[0] = n0.t[1] = n0.t[2] = 0.0;
      transform_nei_to_ref(r0, n0, c + 3, q, P);
      const double dot = P[0] * c[0] + P[1] * c[1] + P[2] * c[2];
      const double cs = fabs(dot);
      if (cs >= 1.0) return 0.0;
      r = acos(cs);
      if (r < 1e-3) return 0.0;
      const double k = (dot < 0.0 ? 1.0 : -1.0) / sqrt(1.0 - cs * cs);
      g[0] = k * c[0]; g[1] = k * c[1]; g[2] = k * c[2];
      accumulate_row(r0, n0, q, P, g, J);
      for (int j = 0; j < 3; ++j) J[3 + j] = J[9 + j] = 0.0;
      return r;
    }
    default:
      return 0.0;
  }
}

// ceres::HuberLoss + Corrector (rho'' <= 0 branch): scales r and the row by sqrt(rho'); returns 0.5*rho(r^2).
PVB_HD double huber_correct(double a, double& r, double* J, int n) {
  const double s = r * r;
  if (a <= 0.0 || s <= a * a) return 0.5 * s;
  const double sq = sqrt(s);
  const double k = sqrt(a / sq);
  r *= k;
  for (int i = 0; i < n; ++i) J[i] *= k;
  return 0.5 * (2.0 * a * sq - a * a);
}

// ---- streaming plane fit (registers: 9 doubles instead of a K x 3 matrix) ------------------------------------------
// Same least-squares problem as Geometry.hpp:345-373 (A x = -1): Gram matrix G = A^T A and h = A^T 1 are accumulated
// point by point, G x = -h is solved by a 3x3 Cholesky and corrected by one step of iterative refinement with the
// residual evaluated from the points (corrected semi-normal equations), which restores QR-level accuracy for the
// conditioning met here (cond(A) ~ range / neighbourhood size ~ 1e3..1e4).  The scatter matrix of the collinearity
// test (Geometry.hpp:229-235) is G - h h^T / n.
struct PlaneAcc { double g00, g01, g02, g11, g12, g22, h0, h1, h2; };
PVB_HD void plane_acc_clear(PlaneAcc& a) { a.g00 = a.g01 = a.g02 = a.g11 = a.g12 = a.g22 = a.h0 = a.h1 = a.h2 = 0.0; }
PVB_HD void plane_acc_add(PlaneAcc& a, const double p[3]) {
  a.g00 += p[0] * p[0]; a.g01 += p[0] * p[1]; a.g02 += p[0] * p[2]; a.g11 += p[1] * p[1]; a.g12 += p[1] * p[2]; a.g22 += p[2] * p[2];
  a.h0 += p[0]; a.h1 += p[1]; a.h2 += p[2];
}
struct Chol3 { double l00, l10, l11, l20, l21, l22, i00, i11, i22; };      // i.. = reciprocals of the diagonal (one division each; the solves only multiply)
PVB_HD bool chol3_factor(const PlaneAcc& a, Chol3& L) {
  if (!(a.g00 > 0.0)) return false;
  L.l00 = sqrt(a.g00); L.i00 = 1.0 / L.l00; L.l10 = a.g01 * L.i00; L.l20 = a.g02 * L.i00;
  const double d1 = a.g11 - L.l10 * L.l10;
  if (!(d1 > 0.0)) return false;
  L.l11 = sqrt(d1); L.i11 = 1.0 / L.l11; L.l21 = (a.g12 - L.l20 * L.l10) * L.i11;
  const double d2 = a.g22 - L.l20 * L.l20 - L.l21 * L.l21;
  if (!(d2 > 0.0)) return false;
  L.l22 = sqrt(d2); L.i22 = 1.0 / L.l22;
  return true;
}
PVB_HD void chol3_solve(const Chol3& L, const double b[3], double x[3]) {
  const double y0 = b[0] * L.i00;
  const double y1 = (b[1] - L.l10 * y0) * L.i11;
  const double y2 = (b[2] - L.l20 * y0 - L.l21 * y1) * L.i22;
  x[2] = y2 * L.i22;
  x[1] = (y1 - L.l21 * x[2]) * L.i11;
  x[0] = (y0 - L.l10 * x[1] - L.l20 * x[2]) * L.i00;
}
// lambda_max > tol * lambda_mid of the scatter matrix (FormLine's "is a line" test)
PVB_HD bool collinear_from_gram(const PlaneAcc& a, int n, double tol) {
  const double inv = 1.0 / (double)n;
  const double a00 = a.g00 - a.h0 * a.h0 * inv, a01 = a.g01 - a.h0 * a.h1 * inv, a02 = a.g02 - a.h0 * a.h2 * inv;
  const double a11 = a.g11 - a.h1 * a.h1 * inv, a12 = a.g12 - a.h1 * a.h2 * inv, a22 = a.g22 - a.h2 * a.h2 * inv;
  const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
  const double qm = (a00 + a11 + a22) / 3.0;
  const double p2 = (a00 - qm) * (a00 - qm) + (a11 - qm) * (a11 - qm) + (a22 - qm) * (a22 - qm) + 2.0 * p1;
  if (p2 <= 0.0) return 1.0 > tol;      // all eigenvalues equal
  const double p = sqrt(p2 / 6.0), ip = 1.0 / p;
  const double b00 = (a00 - qm) * ip, b11 = (a11 - qm) * ip, b22 = (a22 - qm) * ip, b01 = a01 * ip, b02 = a02 * ip, b12 = a12 * ip;
  double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
  r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  const double phi = acos(r) / 3.0;
  const double l2 = qm + 2.0 * p * cos(phi);
  const double l0 = qm + 2.0 * p * cos(phi + 2.0943951023931953);
  const double l1 = 3.0 * qm - l0 - l2;
  return l2 > tol * l1;
}

// FormLine(points, tolerance, dis_threshold) (Geometry.hpp:220-260) for the k = 5 neighbours of AssociatePoint2Line:
// centroid + unit direction (eigenvector of the largest eigenvalue of the scatter matrix) when lambda_max > tolerance *
// lambda_mid and every point is within dis_threshold of the line.  Eigenvalues by the trigonometric closed form, the
// eigenvector as the largest cross product of two rows of (S - lambda_max I) (well conditioned: the test guarantees
// a spectral gap of 10x).  The sign of the direction is arbitrary (the residuals only use the line).
template <int K>
PVB_HD bool form_line_pca(const double (*pts)[3], double tolerance, double dis_threshold, double line6[6]) {
  double c[3] = {0, 0, 0};
  for (int i = 0; i < K; ++i) { c[0] = c[0] + pts[i][0]; c[1] = c[1] + pts[i][1]; c[2] = c[2] + pts[i][2]; }
  c[0] = c[0] / double(K); c[1] = c[1] / double(K); c[2] = c[2] / double(K);
  double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
  for (int i = 0; i < K; ++i) {
    const double x = pts[i][0] - c[0], y = pts[i][1] - c[1], z = pts[i][2] - c[2];
    a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z;
  }
  const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
  const double qm = (a00 + a11 + a22) / 3.0;
  const double p2 = (a00 - qm) * (a00 - qm) + (a11 - qm) * (a11 - qm) + (a22 - qm) * (a22 - qm) + 2.0 * p1;
  if (p2 <= 0.0) return false;                         // isotropic scatter: lambda_max == lambda_mid
  const double p = sqrt(p2 / 6.0), ip = 1.0 / p;
  const double b00 = (a00 - qm) * ip, b11 = (a11 - qm) * ip, b22 = (a22 - qm) * ip, b01 = a01 * ip, b02 = a02 * ip, b12 = a12 * ip;
  double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
  r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  const double phi = acos(r) / 3.0;
  const double l2 = qm + 2.0 * p * cos(phi);
  const double l0 = qm + 2.0 * p * cos(phi + 2.0943951023931953);
  const double l1 = 3.0 * qm - l0 - l2;
  if (!(l2 > tolerance * l1)) return false;
  const double r0[3] = {a00 - l2, a01, a02}, r1[3] = {a01, a11 - l2, a12}, r2[3] = {a02, a12, a22 - l2};
  double v[3][3] = {{r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]},
                    {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]},
                    {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]}};
  double n0 = v[0][0] * v[0][0] + v[0][1] * v[0][1] + v[0][2] * v[0][2];
  const double n1 = v[1][0] * v[1][0] + v[1][1] * v[1][1] + v[1][2] * v[1][2];
  const double n2 = v[2][0] * v[2][0] + v[2][1] * v[2][1] + v[2][2] * v[2][2];
  double d[3] = {v[0][0], v[0][1], v[0][2]};
  if (n1 > n0) { n0 = n1; d[0] = v[1][0]; d[1] = v[1][1]; d[2] = v[1][2]; }
  if (n2 > n0) { n0 = n2; d[0] = v[2][0]; d[1] = v[2][1]; d[2] = v[2][2]; }
  if (!(n0 > 0.0)) return false;
  const double inv = 1.0 / sqrt(n0);
  d[0] *= inv; d[1] *= inv; d[2] *= inv;
  if (dis_threshold > 0.0) {
    for (int i = 0; i < K; ++i) {                       // PointToLineDistance3D (Geometry.hpp:198-211)
      const double k = d[0] * (pts[i][0] - c[0]) + d[1] * (pts[i][1] - c[1]) + d[2] * (pts[i][2] - c[2]);
      const double ex = k * d[0] + c[0] - pts[i][0], ey = k * d[1] + c[1] - pts[i][1], ez = k * d[2] + c[2] - pts[i][2];
      if (sqrt(ex * ex + ey * ey + ez * ez) > dis_threshold) return false;
    }
  }
  line6[0] = c[0]; line6[1] = c[1]; line6[2] = c[2]; line6[3] = d[0]; line6[4] = d[1]; line6[5] = d[2];
  return true;
}

// Rank-deficient fallback of the plane fit (Cholesky of the Gram matrix failed: e.g. every neighbour has one
// coordinate exactly 0).  Eigen's colPivHouseholderQr().solve (Geometry.hpp:361) then returns the basic solution of
// the leading rank x rank block with the remaining components 0; this rolled-loop version (local memory, few
// registers, rare path) reproduces it.  A: K x 3 row-major, destroyed.
PVB_HD void lstsq_minus_one_rolled(int K, double* A, double x[3]) {
  double b[16];
  for (int i = 0; i < K; ++i) b[i] = -1.0;
  int perm[3] = {0, 1, 2};
  double diag[3] = {0, 0, 0};
  int rank = 0;
  double maxpivot = 0.0;
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    int best = k; double bn = -1.0;
#pragma unroll 1
    for (int c = k; c < 3; ++c) {
      double s = 0;
      for (int r = k; r < K; ++r) s += A[r * 3 + c] * A[r * 3 + c];
      if (s > bn) { bn = s; best = c; }
    }
    if (best != k) {
      for (int r = 0; r < K; ++r) { const double tmp = A[r * 3 + k]; A[r * 3 + k] = A[r * 3 + best]; A[r * 3 + best] = tmp; }
      const int tp = perm[k]; perm[k] = perm[best]; perm[best] = tp;
    }
    const double norm = sqrt(bn);
    if (k == 0) maxpivot = norm;
    if (norm <= maxpivot * DBL_EPSILON * 3.0) break;
    ++rank;
    const double akk = A[k * 3 + k];
    const double alpha = (akk > 0) ? -norm : norm;
    const double vk = akk - alpha;
    double vtv = vk * vk;
    for (int r = k + 1; r < K; ++r) vtv += A[r * 3 + k] * A[r * 3 + k];
    if (vtv > 0) {
      const double inv = 2.0 / vtv;
#pragma unroll 1
      for (int c = k + 1; c < 3; ++c) {
        double dot = vk * A[k * 3 + c];
        for (int r = k + 1; r < K; ++r) dot += A[r * 3 + k] * A[r * 3 + c];
        const double f = dot * inv;
        A[k * 3 + c] -= f * vk;
        for (int r = k + 1; r < K; ++r) A[r * 3 + c] -= f * A[r * 3 + k];
      }
      double dot = vk * b[k];
      for (int r = k + 1; r < K; ++r) dot += A[r * 3 + k] * b[r];
      const double f = dot * inv;
      b[k] -= f * vk;
      for (int r = k + 1; r < K; ++r) b[r] -= f * A[r * 3 + k];
    }
    diag[k] = alpha;
  }
  double y[3] = {0, 0, 0};
#pragma unroll 1
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int c = k + 1; c < rank; ++c) s -= A[k * 3 + c] * y[c];
    y[k] = s / diag[k];
  }
  x[0] = x[1] = x[2] = 0.0;
  for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

// ---- Equirectangular float path (sensors/Equirectangular.h:41-96 with USE_FAST_ATAN2, base/Math.h:15-29) ----
PVB_HD float fast_atan2_f32(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = mn / fadd(mx, (float)DBL_EPSILON);
  const float s = fmul(a, a);
  const double sd = s, ad = a;
  // ((c3*s + c2)*s - c1)*s*a + c0*a evaluated in double (the literals are double), rounded once to float
  double r = dmul(dmul(dsub(dmul(dadd(dmul(-0.04432655554792128, sd), 0.1555786518463281), sd), 0.3258083974640975), sd), ad);
  r = dadd(r, dmul(0.9997878412794807, ad));
  float rf = (float)r;
  if (ay > ax) rf = (float)dsub(M_PI_2, (double)rf);
  if (x < 0) rf = (float)dsub(M_PI, (double)rf);
  if (y < 0) rf = -rf;
  return rf;
}

PVB_HD double fast_atan2_f64(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double a = fmin(ax, ay) / dadd(fmax(ax, ay), DBL_EPSILON);
  const double s = dmul(a, a);
  double r = dadd(dmul(dmul(dsub(dmul(dadd(dmul(-0.04432655554792128, s), 0.1555786518463281), s), 0.3258083974640975), s), a), dmul(0.9997878412794807, a));
  if (ay > ax) r = dsub(M_PI_2, r);
  if (x < 0) r = dsub(M_PI, r);
  if (y < 0) r = -r;
  return r;
}

// CamToImage in float (CamToSphere + SphereToImage): pixel of a camera-frame point
PVB_HD void cam_to_image_f32(float x, float y, float z, int rows, int cols, float& u, float& v) {
  const float lon = fast_atan2_f32(x, z);
  const float lat = -fast_atan2_f32(y, sqrtf(fadd(fmul(x, x), fmul(z, z))));
  u = (float)dmul((double)cols, dadd(0.5, (double)lon / (2.0 * M_PI)));
  v = (float)dmul((double)rows, dsub(0.5, (double)lat / M_PI));
}

// ImageToCam in double, exact sin/cos (Equirectangular.h:98-170)
PVB_HD void image_to_cam_f64(double px, double py, int rows, int cols, double cam[3]) {
  const double lon = (2 * px / cols - 1) * M_PI;
  const double lat = (0.5 - py / rows) * M_PI;
  const double cy = cos(lat);
  cam[0] = cy * sin(lon);
  cam[1] = -sin(lat);
  cam[2] = cy * cos(lon);
}

// ---- sweep undistortion (sensors/Velodyne.cpp:1642-1674) ------------------------------------------------------------------------
// One sweep = one rigid motion (q_se, t_se) from the pose at the start to the pose at the end; point i of n is moved by the fraction
// ratio = float(i) / float(n) of it: rotation = slerp(identity, q_se, ratio) (Eigen's formula, not renormalised), translation = ratio * t_se.
// Everything that does not depend on the point is prepared once per frame on the host (UndistortPrep).
struct UndistortPrep {
  double q[4];          // q_se as x, y, z, w
  double theta, sin_theta;
  double t_se[3];
  int linear;           // |w| >= 1 - eps: Eigen's slerp degenerates to linear weights
  int enabled;          // 0: the frame has no usable end pose, points pass through unchanged (LidarOdometry.cpp:212-225 `goto save_undistort`)
};

PVB_HD void quat_from_matrix_eigen(const double R[9], double q[4]) {          // Eigen::Quaterniond(Matrix3d), R row-major
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0), k = 0.5 / s;
    q[3] = 0.5 * s; q[0] = (R[7] - R[5]) * k; q[1] = (R[2] - R[6]) * k; q[2] = (R[3] - R[1]) * k;
    return;
  }
  int i = R[4] > R[0] ? 1 : 0;
  if (R[8] > R[i * 4]) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  const double s = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0), h = 0.5 / s;
  q[i] = 0.5 * s;
  q[3] = (R[k * 3 + j] - R[j * 3 + k]) * h;
  q[j] = (R[j * 3 + i] + R[i * 3 + j]) * h;
  q[k] = (R[k * 3 + i] + R[i * 3 + k]) * h;
}

PVB_HD void quat_to_matrix_eigen(const double q[4], double R[9]) {           // QuaternionBase::toRotationMatrix
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

PVB_HD void undistort_prepare(const double q_se[4], const double t_se[3], UndistortPrep& u) {
  for (int a = 0; a < 4; ++a) u.q[a] = q_se[a];
  for (int a = 0; a < 3; ++a) u.t_se[a] = t_se[a];
  const double ad = fabs(q_se[3]);
  u.linear = ad >= 1.0 - 2.220446049250313e-16 ? 1 : 0;
  u.theta = u.linear ? 0.0 : acos(ad);
  u.sin_theta = u.linear ? 1.0 : sin(u.theta);
  u.enabled = 1;
}

// weights of identity and q_se in the interpolated quaternion
PVB_HD void slerp_weights(const UndistortPrep& u, double ratio, double& w_id, double& w_q) {
  if (u.linear) { w_id = 1.0 - ratio; w_q = ratio; }
  else { w_id = sin((1.0 - ratio) * u.theta) / u.sin_theta; w_q = sin(ratio * u.theta) / u.sin_theta; }
  if (u.q[3] < 0.0) w_q = -w_q;
}

PVB_HD void undistort_point_f32(const UndistortPrep& u, long long i, long long n, float x, float y, float z, float& ox, float& oy, float& oz) {
  const float ratio_f = (float)i / (float)n;                                 // `1.f * i / cloud.points.size()` (:1656) is float arithmetic
  const double ratio = (double)ratio_f;
  double w_id, w_q;
  slerp_weights(u, ratio, w_id, w_q);
  const double qx = dmul(w_q, u.q[0]), qy = dmul(w_q, u.q[1]), qz = dmul(w_q, u.q[2]), qw = dadd(w_id, dmul(w_q, u.q[3]));
  const double vx = (double)x, vy = (double)y, vz = (double)z;
  // Eigen's q * v: uv = 2 (q.vec x v); v + w uv + q.vec x uv  (products and sums kept un-fused like the host build of the reference)
  double ux = dsub(dmul(qy, vz), dmul(qz, vy)), uy = dsub(dmul(qz, vx), dmul(qx, vz)), uz = dsub(dmul(qx, vy), dmul(qy, vx));
  ux = dadd(ux, ux); uy = dadd(uy, uy); uz = dadd(uz, uz);
  const double cx = dsub(dmul(qy, uz), dmul(qz, uy)), cy = dsub(dmul(qz, ux), dmul(qx, uz)), cz = dsub(dmul(qx, uy), dmul(qy, ux));
  ox = (float)dadd(dadd(dadd(vx, dmul(qw, ux)), cx), dmul(ratio, u.t_se[0]));
  oy = (float)dadd(dadd(dadd(vy, dmul(qw, uy)), cy), dmul(ratio, u.t_se[1]));
  oz = (float)dadd(dadd(dadd(vz, dmul(qw, uz)), cz), dmul(ratio, u.t_se[2]));
}

}  // namespace pvb
