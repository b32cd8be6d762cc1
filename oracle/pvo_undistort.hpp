// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the per-frame preprocessing that feeds the hot path: pose interpolation and the per-point
// motion undistortion of a LiDAR sweep.
//   SlerpPose            base/Geometry.hpp:572-583
//   UndistortCloud       sensors/Velodyne.cpp:1642-1674
//   UndistortLidars      lidar_mapping/LidarOdometry.cpp:189-243 (choice of the sweep-end pose per frame)
// Third-party arithmetic restated from Eigen 3.4's documented source behaviour (Geometry/Quaternion.h): Quaternion(Matrix3),
// QuaternionBase::slerp, QuaternionBase::_transformVector, toRotationMatrix; Matrix4d::inverse() as adjugate / determinant
// (Eigen's 4x4 kernel groups the same cofactors differently; the results agree to rounding, ~1e-16 relative).
// PARITY STATUS: unpinned by the reference's own tests (none exist); pinned against scipy.spatial.transform (Slerp, Rotation) in
// tests/test_undistort.py.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>

namespace pvo {

// Eigen::Quaterniond(Matrix3d) — coefficients x, y, z, w.  R row-major.
inline void EigenQuatFromMatrix(const double R[9], double q[4]) {
  auto M = [&](int r, int c) { return R[r * 3 + c]; };
  double t = M(0, 0) + M(1, 1) + M(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M(2, 1) - M(1, 2)) * t;
    q[1] = (M(0, 2) - M(2, 0)) * t;
    q[2] = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (M(k, j) - M(j, k)) * t;
    q[j] = (M(j, i) + M(i, j)) * t;
    q[k] = (M(k, i) + M(i, k)) * t;
  }
}

// Quaterniond::Identity().slerp(t, q): scale0 * identity + scale1 * q
inline void EigenSlerpFromIdentity(double t, const double q[4], double out[4]) {
  const double one = 1.0 - std::numeric_limits<double>::epsilon();
  const double d = q[3];                       // identity . q = w
  const double absD = std::fabs(d);
  double scale0, scale1;
  if (absD >= one) { scale0 = 1.0 - t; scale1 = t; }
  else {
    const double theta = std::acos(absD);
    const double sinTheta = std::sin(theta);
    scale0 = std::sin((1.0 - t) * theta) / sinTheta;
    scale1 = std::sin(t * theta) / sinTheta;
  }
  if (d < 0.0) scale1 = -scale1;
  out[0] = scale0 * 0.0 + scale1 * q[0];
  out[1] = scale0 * 0.0 + scale1 * q[1];
  out[2] = scale0 * 0.0 + scale1 * q[2];
  out[3] = scale0 * 1.0 + scale1 * q[3];
}

// q * v (QuaternionBase::_transformVector): v + w * uv + vec x uv, uv = 2 (vec x v)
inline void EigenQuatRotate(const double q[4], const double v[3], double out[3]) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int a = 0; a < 3; ++a) out[a] = v[a] + q[3] * uv[a] + c[a];
}

inline void EigenQuatToMatrix(const double q[4], double R[9]) {
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

inline void Mat4Mul(const double A[16], const double B[16], double C[16]) {
  double T[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { double s = 0.0; for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * B[k * 4 + c]; T[r * 4 + c] = s; }
  std::memcpy(C, T, sizeof(T));
}

// general 4x4 inverse (adjugate / determinant), row-major
inline void Mat4Inverse(const double m[16], double inv[16]) {
  double a[16];
  a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const double det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
  const double id = 1.0 / det;
  for (int i = 0; i < 16; ++i) inv[i] = a[i] * id;
}

// base/Geometry.hpp:572-583.  Poses are 4x4 row-major T_w<-local.
inline void SlerpPose(const double pose_w1[16], const double pose_w2[16], double ratio, double out[16]) {
  double inv2[16], T_21[16];
  Mat4Inverse(pose_w2, inv2);
  Mat4Mul(inv2, pose_w1, T_21);                                       // :575
  double R21[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R21[r * 3 + c] = T_21[r * 4 + c];
  double q_21[4], q_s1[4];
  EigenQuatFromMatrix(R21, q_21);                                     // :576
  EigenSlerpFromIdentity(ratio, q_21, q_s1);                          // :577
  double T_s1[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double Rs[9];
  EigenQuatToMatrix(q_s1, Rs);                                        // :580
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T_s1[r * 4 + c] = Rs[r * 3 + c]; T_s1[r * 4 + 3] = T_21[r * 4 + 3] * ratio; }   // :578, :581
  double inv_s1[16];
  Mat4Inverse(T_s1, inv_s1);
  Mat4Mul(pose_w1, inv_s1, out);                                      // :582
}

// sensors/Velodyne.cpp:1642-1674: points n x 4 float32 (x, y, z, intensity), in place semantics restated as in -> out.
inline void UndistortCloud(const double R_wl[9], const double t_wl[3], const double R_we[9], const double t_we[3], const float* in, size_t n, float* out) {
  double R_se[9], t_se[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { double s = 0.0; for (int k = 0; k < 3; ++k) s += R_wl[k * 3 + r] * R_we[k * 3 + c]; R_se[r * 3 + c] = s; }   // :1647
    double s = 0.0; for (int k = 0; k < 3; ++k) s += R_wl[k * 3 + r] * (t_we[k] - t_wl[k]); t_se[r] = s;                                     // :1648
  }
  double q_se[4];
  EigenQuatFromMatrix(R_se, q_se);                                    // :1649
  for (size_t i = 0; i < n; ++i) {
    const float ratio_f = 1.f * (float)i / (float)n;                  // :1656 `1.f * i / cloud.points.size()` is float arithmetic
    const double ratio = (double)ratio_f;
    double q_sc[4];
    EigenSlerpFromIdentity(ratio, q_se, q_sc);                        // :1657
    const double t_sc[3] = {ratio * t_se[0], ratio * t_se[1], ratio * t_se[2]};
    const double p[3] = {(double)in[i * 4], (double)in[i * 4 + 1], (double)in[i * 4 + 2]};
    double r[3];
    EigenQuatRotate(q_sc, p, r);                                      // :1660
    out[i * 4] = (float)(r[0] + t_sc[0]); out[i * 4 + 1] = (float)(r[1] + t_sc[1]); out[i * 4 + 2] = (float)(r[2] + t_sc[2]);
    out[i * 4 + 3] = in[i * 4 + 3];
  }
}

// lidar_mapping/LidarOdometry.cpp:203-243: the pose of the END of sweep i, or has[i] = 0 when the frame is saved undistorted.
// poses: n x 16 (T_wl row-major); pose_valid / frame_valid as Velodyne::IsPoseValid() / Velodyne::valid.  The loop conditions are
// the reference's, including `!IsPoseValid() && !valid` (both must fail for a candidate to be skipped) and `idx <= 0` for the last frame.
inline void UndistortEndPoses(int n, const double* poses, const unsigned char* pose_valid, const unsigned char* frame_valid, float gap_time, double* out_pose,
                              unsigned char* has) {
  const double lidar_duration = 0.1;                                  // :203
  for (int i = 0; i < n; ++i) {
    has[i] = 0;
    double* pose = out_pose + (size_t)i * 16;
    std::memset(pose, 0, 16 * sizeof(double));
    if (!pose_valid[i] || !frame_valid[i]) continue;                  // :212
    if (i < n - 1) {                                                  // :217
      int idx = i + 1;
      while (idx < n && !pose_valid[idx] && !frame_valid[idx]) idx++;
      if (idx >= n) continue;
      SlerpPose(poses + (size_t)i * 16, poses + (size_t)idx * 16, lidar_duration / ((idx - i) * (lidar_duration + gap_time)), pose);   // :224
    } else {                                                          // :227
      int idx = i - 1;
      while (idx >= 0 && !pose_valid[idx] && !frame_valid[i]) idx--;
      if (idx <= 0) continue;
      double tmp[16], inv_i[16], T_cs[16];
      SlerpPose(poses + (size_t)idx * 16, poses + (size_t)i * 16, 1.0 - lidar_duration / ((idx - i) * (lidar_duration + gap_time)), tmp);   // :236
      Mat4Inverse(poses + (size_t)i * 16, inv_i);
      Mat4Mul(inv_i, tmp, T_cs);                                      // :238
      Mat4Mul(poses + (size_t)i * 16, T_cs, pose);                    // :240
    }
    has[i] = 1;
  }
}

}  // namespace pvo
