// ORACLE/_ref — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE'S OWN base/CostFunction.h, base/Geometry.hpp, base/Math.h,
// sensors/Equirectangular.h and sensors/Equirectangular.cpp, compiled from the files where they lie under /root/reference (never copied into this
// repository).  Those files include Eigen / Ceres / OpenCV / PCL / glog / Boost headers, none of which exist in this container; oracle/shim/ provides
// stand-ins for the few value types and functions they touch (see shim/pvo_shim_eigen.hpp for what that does and does not pin).  The cost functors
// are evaluated through the reference's own `Functor::Create(...)` -> ceres::CostFunction::Evaluate(parameters, residuals, jacobians) surface.
// Built by `make -C oracle ref` into oracle/_ref/libpvo_ref_path.so; used by tests/test_oracle_pinning.py and tests/make_golden.py only.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include REF_COSTFUNCTION_H
#include REF_EQUIRECT_H
#include REF_EQUIRECT_CPP

namespace {
typedef Eigen::Vector3d V3;
typedef Eigen::Vector4d V4;
inline V3 v3(const double* p) { return V3(p[0], p[1], p[2]); }
inline V4 v4(const double* p) { return V4(p[0], p[1], p[2], p[3]); }
inline cv::Point3f p3f(const double* p) { return cv::Point3f((float)p[0], (float)p[1], (float)p[2]); }
template <typename V> inline void put(double* o, const V& v, int n) { for (int i = 0; i < n; ++i) o[i] = v[i]; }
}  // namespace

extern "C" {
// One functor per row.  type: 0..8 as oracle/pvo.py (P2PLANE_METER .. LINE2LINE_ANGLE); 9 / 10 = PairWisePoint2Plane_Meter / PairWisePoint2Line_Meter;
// 11 = PanoramaReprojResidual_1Angle; 12 = PlaneIOUResidual through its camera-LiDAR constructor.  raw: 16 doubles per row = the CONSTRUCTOR arguments
// in declaration order (vectors flattened, then the weight); params: 12 doubles per row = the parameter blocks in call order, 3 doubles each.
// Outputs: r[n]; J[n x 12] (row-major 1x3 per block, blocks in call order; may be null); consts[n x 12] = the functor's members AFTER its constructor
// ran, in the oracle's block-constant layout (may be null).  Returns 0, or -1 - row on an Evaluate() failure / unknown type.
long ref_eval_functors(long n, const int* type, const int* normalize, const double* raw, const double* params, double* r, double* J, double* consts) {
  for (long i = 0; i < n; ++i) {
    const double* a = raw + 16 * i;
    const double* p = params + 12 * i;
    double* c = consts ? consts + 12 * i : nullptr;
    if (c) std::fill(c, c + 12, 0.0);
    std::unique_ptr<ceres::CostFunction> f;
    switch (type[i]) {
      case 0: { f.reset(Point2Plane_Meter::Create(v3(a), v4(a + 3), a[7]));
        if (c) { Point2Plane_Meter m(v3(a), v4(a + 3), a[7]); put(c, m.curr_point, 3); put(c + 3, m.plane, 4); c[7] = m.weight; } break; }
      case 1: { f.reset(Point2Plane_Angle::Create(v3(a), v4(a + 3), normalize[i] != 0, a[7]));
        if (c) { Point2Plane_Angle m(v3(a), v4(a + 3), normalize[i] != 0, a[7]); put(c, m.curr_point, 3); put(c + 3, m.plane, 4); c[7] = m.weight; } break; }
      case 2: { f.reset(Point2Line_Meter::Create(v3(a), v3(a + 3), v3(a + 6), a[9]));
        if (c) { Point2Line_Meter m(v3(a), v3(a + 3), v3(a + 6), a[9]); put(c, m.curr_point, 3); put(c + 3, m.line_point, 3); put(c + 6, m.line_direction, 3); c[9] = m.weight; } break; }
      case 3: { f.reset(Point2Line_Angle::Create(v3(a), v3(a + 3), v3(a + 6), normalize[i] != 0, a[9]));
        if (c) { Point2Line_Angle m(v3(a), v3(a + 3), v3(a + 6), normalize[i] != 0, a[9]); put(c, m.curr_point, 3); put(c + 3, m.line_point, 3); put(c + 6, m.line_direction, 3); c[9] = m.weight; } break; }
      case 4: { f.reset(Plane2Plane_Global::Create(v3(a), v3(a + 3), v3(a + 6), a[9]));
        if (c) { Plane2Plane_Global m(v3(a), v3(a + 3), v3(a + 6), a[9]); put(c, m.plane_ref, 3); put(c + 3, m.point_a, 3); put(c + 6, m.point_b, 3); c[9] = m.weight; } break; }
      case 5: { f.reset(PlaneIOUResidual::Create(v4(a), v3(a + 4), v3(a + 7), a[10], a[11]));
        if (c) { PlaneIOUResidual m(v4(a), v3(a + 4), v3(a + 7), a[10], a[11]); put(c, m.ref_plane, 4); put(c + 4, m.middle_neighbor, 3); put(c + 7, m.middle_ref, 3); c[10] = m.angle; c[11] = m.weight; } break; }
      case 12: { PlaneIOUResidual* m = new PlaneIOUResidual(v4(a), v3(a + 4), v3(a + 7), v3(a + 10), a[13]);
        if (c) { put(c, m->ref_plane, 4); put(c + 4, m->middle_neighbor, 3); put(c + 7, m->middle_ref, 3); c[10] = m->angle; c[11] = m->weight; }
        f.reset(new ceres::AutoDiffCostFunction<PlaneIOUResidual, 1, 3, 3, 3, 3>(m)); break; }
      case 6: { f.reset(Plane2Plane_Relative::Create(v3(a), v3(a + 3), v3(a + 6), a[9]));
        if (c) { Plane2Plane_Relative m(v3(a), v3(a + 3), v3(a + 6), a[9]); put(c, m.plane_ref, 3); put(c + 3, m.point_a, 3); put(c + 6, m.point_b, 3); c[9] = m.weight; } break; }
      case 7: { f.reset(PlaneRelativeIOUResidual::Create(v4(a), v3(a + 4), p3f(a + 7), p3f(a + 10), a[13]));
        if (c) { PlaneRelativeIOUResidual m(v4(a), v3(a + 4), p3f(a + 7), p3f(a + 10), a[13]); put(c, m.ref_plane, 4); put(c + 4, m.middle_neighbor, 3); put(c + 7, m.middle_ref, 3); c[10] = m.angle; c[11] = m.weight; } break; }
      case 8: { f.reset(Line2Line_Angle::Create(v3(a), v3(a + 3), a[6]));
        if (c) { Line2Line_Angle m(v3(a), v3(a + 3), a[6]); put(c, m.line_direction_ref, 3); put(c + 3, m.line_direction_nei, 3); } break; }
      case 9: { f.reset(PairWisePoint2Plane_Meter::Create(v3(a), v4(a + 3), a[7]));
        if (c) { put(c, a, 8); } break; }
      case 10: { f.reset(PairWisePoint2Line_Meter::Create(v3(a), v3(a + 3), v3(a + 6), a[9]));
        if (c) { PairWisePoint2Line_Meter m(v3(a), v3(a + 3), v3(a + 6), a[9]); put(c, m.curr_point, 3); put(c + 3, m.line_point, 3); put(c + 6, m.line_direction, 3); c[9] = m.weight; } break; }
      case 11: { f.reset(PanoramaReprojResidual_1Angle::Create(v3(a), a[3]));
        if (c) { PanoramaReprojResidual_1Angle m(v3(a), a[3]); put(c, m.point_sphere, 3); c[3] = m.weight; } break; }
      default: return -1 - i;
    }
    const int nb = (int)f->parameter_block_sizes().size();
    const double* blocks[4] = {p, p + 3, p + 6, p + 9};
    double jb[4][3]; double* jp[4] = {jb[0], jb[1], jb[2], jb[3]};
    double res = 0;
    if (!f->Evaluate(blocks, &res, J ? jp : nullptr)) return -1 - i;
    r[i] = res;
    if (J) { std::fill(J + 12 * i, J + 12 * i + 12, 0.0); for (int b = 0; b < nb; ++b) std::memcpy(J + 12 * i + 3 * b, jb[b], 24); }
  }
  return 0;
}

// ---- base/Geometry.hpp ----
void ref_form_plane(int n, const double* pts, double tol, double* out4) {
  eigen_vector<V3> v; for (int i = 0; i < n; ++i) v.push_back(v3(pts + 3 * i));
  put(out4, FormPlane<double>(v, tol), 4);
}
void ref_form_plane3(const double* p1, const double* p2, const double* p3, double* out4) { put(out4, FormPlane<double>(v3(p1), v3(p2), v3(p3)), 4); }
void ref_form_line(int n, const double* pts, double tol, double dis_thr, double* out6) {
  eigen_vector<V3> v; for (int i = 0; i < n; ++i) v.push_back(v3(pts + 3 * i));
  put(out6, FormLine<double>(v, tol, dis_thr), 6);
}
double ref_point_to_line_distance3d(const double* point, const double* line6) { return PointToLineDistance3D<double>(point, line6); }
double ref_point_to_plane_distance(const double* plane4, const double* point, int normalized) { return PointToPlaneDistance<double>(plane4, point, normalized != 0); }
void ref_project_point_to_plane(const double* point, const double* plane4, double* out3, int normalized) { ProjectPointToPlane<double>(point, plane4, out3, normalized != 0); }
double ref_vector_angle3d(const double* a, const double* b, int normalized) { return VectorAngle3D<double>(a, b, normalized != 0); }
double ref_plane_angle(const double* a, const double* b, int normalized) { return PlaneAngle<double>(a, b, normalized != 0); }
void ref_slerp_pose(const double* w1, const double* w2, double ratio, double* out) {  // 4x4 row-major
  Eigen::Matrix4d a, b; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a(i, j) = w1[4 * i + j]; b(i, j) = w2[4 * i + j]; }
  const Eigen::Matrix4d o = SlerpPose<double>(a, b, ratio);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[4 * i + j] = o(i, j);
}

// ---- sensors/Equirectangular ----
void ref_cam_to_image_f(int rows, int cols, long n, const float* cam, float* px) {
  Equirectangular eq(rows, cols);
  for (long i = 0; i < n; ++i) { const cv::Point2f q = eq.CamToImage(cv::Point3f(cam[3 * i], cam[3 * i + 1], cam[3 * i + 2])); px[2 * i] = q.x; px[2 * i + 1] = q.y; }
}
void ref_cam_to_image_d(int rows, int cols, long n, const double* cam, double* px) {
  Equirectangular eq(rows, cols);
  for (long i = 0; i < n; ++i) { const cv::Point2d q = eq.CamToImage(cv::Point3d(cam[3 * i], cam[3 * i + 1], cam[3 * i + 2])); px[2 * i] = q.x; px[2 * i + 1] = q.y; }
}
void ref_cam_to_image_eigen_d(int rows, int cols, long n, const double* cam, double* px) {
  Equirectangular eq(rows, cols);
  for (long i = 0; i < n; ++i) { const Eigen::Vector2d q = eq.CamToImage(Eigen::Vector3d(cam[3 * i], cam[3 * i + 1], cam[3 * i + 2])); px[2 * i] = q.x(); px[2 * i + 1] = q.y(); }
}
void ref_image_to_cam_f(int rows, int cols, long n, const float* px, float r, float* cam) {
  Equirectangular eq(rows, cols);
  for (long i = 0; i < n; ++i) { const cv::Point3f q = eq.ImageToCam(cv::Point2f(px[2 * i], px[2 * i + 1]), r); cam[3 * i] = q.x; cam[3 * i + 1] = q.y; cam[3 * i + 2] = q.z; }
}
void ref_image_to_cam_d(int rows, int cols, long n, const double* px, double r, double* cam) {
  Equirectangular eq(rows, cols);
  for (long i = 0; i < n; ++i) { const cv::Point3d q = eq.ImageToCam(cv::Point2d(px[2 * i], px[2 * i + 1]), r); cam[3 * i] = q.x; cam[3 * i + 1] = q.y; cam[3 * i + 2] = q.z; }
}
void ref_image_to_cam_eigen_d(int rows, int cols, long n, const double* px, double r, double* cam) {
  Equirectangular eq(rows, cols);
  for (long i = 0; i < n; ++i) { const Eigen::Vector3d q = eq.ImageToCam(Eigen::Vector2d(px[2 * i], px[2 * i + 1]), r); put(cam + 3 * i, q, 3); }
}
int ref_break_to_segments(int rows, int cols, const float* line4, float seg_length, int cap, float* out2) {
  Equirectangular eq(rows, cols);
  const std::vector<cv::Point2f> s = eq.BreakToSegments(cv::Vec4f(line4[0], line4[1], line4[2], line4[3]), seg_length);
  if ((int)s.size() > cap) return -1;
  for (size_t i = 0; i < s.size(); ++i) { out2[2 * i] = s[i].x; out2[2 * i + 1] = s[i].y; }
  return (int)s.size();
}
int ref_is_inside_i(int rows, int cols, int x, int y) { Equirectangular eq(rows, cols); return eq.IsInside(cv::Point2i(x, y)) ? 1 : 0; }
int ref_is_inside_f(int rows, int cols, float x, float y) { Equirectangular eq(rows, cols); return eq.IsInside(cv::Point2f(x, y)) ? 1 : 0; }
}
