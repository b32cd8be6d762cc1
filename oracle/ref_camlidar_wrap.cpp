// ORACLE/_ref — TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE'S OWN joint_optimization/CameraLidarLineAssociate.cpp
// (AssociateByAngle with its Filter(false, true) and UniqueLinePair tails) and the template ProjectLidar2PanoramaDepth of util/Visualization.h, compiled from
// the files where they lie under /root/reference (never copied) together with every header they include (Visualization.h, Frame.h, Velodyne.h,
// Equirectangular.h, Geometry.hpp, Serialization.h, DepthCompletion.h, SIFT.h ...).  None of OpenCV / PCL / Eigen / Boost / CGAL / glog exists in this
// image; oracle/shim/ provides stand-ins (value types, an exact float32 k-NN behind cv::flann::Index, pcl::transformPointCloud as double math + float
// store, no-op drawing / IO).  pcl::SACSegmentation is NOT reproduced (shim: no inliers), so the pixel-space Associate() overloads, whose result
// depends on PCL's RANSAC, are compiled but not exported.  Functions the reference defines in other translation units and only calls from debug
// branches (drawing, Frame accessors) are given empty bodies below so that the library loads.
// Built by `make -C oracle ref` into oracle/_ref/libpvo_ref_camlidar.so; used by tests/test_reference_pinning.py and tests/make_golden.py only.
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include REF_CAMERA_LIDAR_LINE_ASSOCIATE_CPP
#include REF_EQUIRECT_CPP            // sensors/Equirectangular.cpp: BreakToSegments, used by Filter's length branch

// defined by the reference in util/Visualization.cpp / sensors/Frame.cpp and reached only from debug / visualisation branches: empty bodies so the library loads
cv::Vec3b Gray2Color(uchar) { return cv::Vec3b(); }
cv::Mat DrawLinePairsOnImage(const cv::Mat& img_gray, const vector<CameraLidarLinePair>&, const Eigen::Matrix4d&, const int, const bool) { return img_gray; }
const cv::Mat Frame::GetImageGray() const { return cv::Mat(); }

extern "C" {
// AssociateByAngle.  lines: L x 4 float (image line end points, pixels); cloud_local: n x 4 float = cornerLessSharp in the LiDAR frame; point -> segment
// sets (CSR); S segments with coefficients (S x 6) and projected end points (2 S x 3, LiDAR frame); T_cl row-major 4x4; masks may be null.
// Outputs per pair: image_line_id, lidar_line_id, start / end (LiDAR frame, as the function leaves them), score (`angle` member).  Returns the count or -1.
int ref_associate_by_angle(int rows, int cols, const float* lines, int L, const float* cloud_local, int n, const int* p2s_off, const int* p2s_ids, int S,
                           const double* coeffs, const double* end_points, const double* T_cl_rowmajor, int multiple_association, const unsigned char* image_mask,
                           const unsigned char* lidar_mask, int cap, int* image_line_id, int* lidar_line_id, double* start3, double* end3, float* score) {
  std::vector<cv::Vec4f> ln;
  for (int i = 0; i < L; ++i) ln.push_back(cv::Vec4f(lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]));
  pcl::PointCloud<pcl::PointXYZI> cloud;
  for (int i = 0; i < n; ++i) { pcl::PointXYZI p; p.x = cloud_local[4 * i]; p.y = cloud_local[4 * i + 1]; p.z = cloud_local[4 * i + 2]; p.intensity = cloud_local[4 * i + 3]; cloud.push_back(p); }
  std::vector<std::set<int>> p2s(n);
  for (int i = 0; i < n; ++i) for (int k = p2s_off[i]; k < p2s_off[i + 1]; ++k) p2s[i].insert(p2s_ids[k]);
  std::vector<pcl::PointCloud<pcl::PointXYZI>> seg(S);
  for (int i = 0; i < n; ++i) for (int s : p2s[i]) seg[s].push_back(cloud.points[i]);
  eigen_vector<Vector6d> co;
  for (int s = 0; s < S; ++s) { Vector6d c; for (int k = 0; k < 6; ++k) c[k] = coeffs[6 * s + k]; co.push_back(c); }
  eigen_vector<Eigen::Vector3d> ends;
  for (int i = 0; i < 2 * S; ++i) ends.push_back(Eigen::Vector3d(end_points[3 * i], end_points[3 * i + 1], end_points[3 * i + 2]));
  Eigen::Matrix4d T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T_cl_rowmajor[4 * i + j];
  std::vector<bool> im, lm;
  if (image_mask) im.assign(image_mask, image_mask + L);
  if (lidar_mask) lm.assign(lidar_mask, lidar_mask + S);
  CameraLidarLineAssociate a(rows, cols);
  a.AssociateByAngle(ln, seg, co, cloud, p2s, ends, T, multiple_association != 0, im, lm);
  const std::vector<CameraLidarLinePair> pairs = a.GetAssociatedPairs();
  if ((int)pairs.size() > cap) return -1;
  for (size_t i = 0; i < pairs.size(); ++i) {
    image_line_id[i] = pairs[i].image_line_id; lidar_line_id[i] = pairs[i].lidar_line_id; score[i] = pairs[i].angle;
    for (int k = 0; k < 3; ++k) { start3[3 * i + k] = pairs[i].lidar_line_start[k]; end3[3 * i + k] = pairs[i].lidar_line_end[k]; }
  }
  return (int)pairs.size();
}

// The deterministic FIRST stage of the pixel-space Associate(lines, point_cloud, T_cl) (:22-102): cloud -> camera frame -> CamToImage pixel -> 3 nearest sub-line mid points
// (BreakToSegments(line, 70), 60 px gate) -> per image line the list of LiDAR points, lists shorter than 6 dropped.  The reference then hands every list to
// FitLineRANSAC -> pcl::SACSegmentation; the stand-in records what it receives (camera-frame x, y, z, in the order of the line -> points map) and reports no inliers,
// so nothing after the RANSAC runs.  Output: CSR of the recorded lists.  Returns the number of lists or -1.
int ref_pixel_associate_candidates(int rows, int cols, const float* lines, int L, const float* cloud_local, int n, const double* T_cl_rowmajor, int cap_lists, long cap_points,
                                   int* off, float* xyz) {
  std::vector<cv::Vec4f> ln;
  for (int i = 0; i < L; ++i) ln.push_back(cv::Vec4f(lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]));
  pcl::PointCloud<pcl::PointXYZI> cloud;
  for (int i = 0; i < n; ++i) { pcl::PointXYZI p; p.x = cloud_local[4 * i]; p.y = cloud_local[4 * i + 1]; p.z = cloud_local[4 * i + 2]; p.intensity = cloud_local[4 * i + 3]; cloud.push_back(p); }
  Eigen::Matrix4d T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T_cl_rowmajor[4 * i + j];
  std::vector<std::vector<float>> rec;
  pcl::sac_recorder() = &rec;
  CameraLidarLineAssociate a(rows, cols);
  a.Associate(ln, cloud, T);
  pcl::sac_recorder() = nullptr;
  if ((int)rec.size() > cap_lists) return -1;
  long total = 0; off[0] = 0;
  for (size_t k = 0; k < rec.size(); ++k) {
    if (total + (long)rec[k].size() / 3 > cap_points) return -1;
    std::memcpy(xyz + 3 * total, rec[k].data(), rec[k].size() * sizeof(float));
    total += (long)rec[k].size() / 3; off[k + 1] = (int)total;
  }
  return (int)rec.size();
}

// The WHOLE pixel-space Associate(lines, point_cloud, T_cl) (:22-188) with the RANSAC's answer scripted: the k-th FitLineRANSAC call (= the k-th recorded candidate
// list of the function above) is told inlier list k (indices into that candidate list).  Everything else is the reference's own code: the `< 3 inliers` test, the
// least-squares refit (:731-749, through the stand-ins for pcl's centroid / covariance / eigen33), the farthest-pair search and its use of inlier POSITIONS as
// indices (:117-137), ProjectPoint2Line3D, Filter(true, true), the transform back to the LiDAR frame.  Output: the surviving pairs.  Returns their number or -1.
int ref_pixel_associate_scripted(int rows, int cols, const float* lines, int L, const float* cloud_local, int n, const double* T_cl_rowmajor, int n_lists, const int* inl_off,
                                 const int* inl_idx, int cap, float* image_line4, double* start3, double* end3, float* angle) {
  std::vector<cv::Vec4f> ln;
  for (int i = 0; i < L; ++i) ln.push_back(cv::Vec4f(lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]));
  pcl::PointCloud<pcl::PointXYZI> cloud;
  for (int i = 0; i < n; ++i) { pcl::PointXYZI p; p.x = cloud_local[4 * i]; p.y = cloud_local[4 * i + 1]; p.z = cloud_local[4 * i + 2]; p.intensity = cloud_local[4 * i + 3]; cloud.push_back(p); }
  Eigen::Matrix4d T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T_cl_rowmajor[4 * i + j];
  std::pair<std::vector<std::vector<int>>, size_t> script;
  for (int k = 0; k < n_lists; ++k) script.first.push_back(std::vector<int>(inl_idx + inl_off[k], inl_idx + inl_off[k + 1]));
  script.second = 0;
  pcl::sac_script() = &script;
  CameraLidarLineAssociate a(rows, cols);
  a.Associate(ln, cloud, T);
  pcl::sac_script() = nullptr;
  if ((int)script.second != n_lists) return -2;                                    // the script must match the number of FitLineRANSAC calls
  const std::vector<CameraLidarLinePair> pairs = a.GetAssociatedPairs();
  if ((int)pairs.size() > cap) return -1;
  for (size_t i = 0; i < pairs.size(); ++i) {
    for (int k = 0; k < 4; ++k) image_line4[4 * i + k] = pairs[i].image_line[k];
    for (int k = 0; k < 3; ++k) { start3[3 * i + k] = pairs[i].lidar_line_start[k]; end3[3 * i + k] = pairs[i].lidar_line_end[k]; }
    angle[i] = pairs[i].angle;
  }
  return (int)pairs.size();
}

// The SEGMENTED overload, Associate(lines, segmented_cloud, T_cl) (CameraLidarLineAssociate.cpp:191-338; no live caller in the reference, the call sites are
// commented out), run the same way: segment k = points [seg_off[k], seg_off[k+1]) of cloud_local (LiDAR frame), the k-th FitLineRANSAC call (one per image line whose
// candidates pass the 6-point and the 70 % single-segment tests; it fits the WHOLE majority segment, in the LiDAR frame) answers with inlier list k.
int ref_pixel_associate_segmented_scripted(int rows, int cols, const float* lines, int L, const float* cloud_local, int n_seg, const int* seg_off, const double* T_cl_rowmajor,
                                           int n_lists, const int* inl_off, const int* inl_idx, int cap, float* image_line4, double* start3, double* end3, float* angle) {
  std::vector<cv::Vec4f> ln;
  for (int i = 0; i < L; ++i) ln.push_back(cv::Vec4f(lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3]));
  std::vector<pcl::PointCloud<pcl::PointXYZI>> segs(n_seg);
  for (int k = 0; k < n_seg; ++k)
    for (int i = seg_off[k]; i < seg_off[k + 1]; ++i) { pcl::PointXYZI p; p.x = cloud_local[4 * i]; p.y = cloud_local[4 * i + 1]; p.z = cloud_local[4 * i + 2]; p.intensity = cloud_local[4 * i + 3]; segs[k].push_back(p); }
  Eigen::Matrix4d T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T_cl_rowmajor[4 * i + j];
  std::pair<std::vector<std::vector<int>>, size_t> script;
  for (int k = 0; k < n_lists; ++k) script.first.push_back(std::vector<int>(inl_idx + inl_off[k], inl_idx + inl_off[k + 1]));
  script.second = 0;
  pcl::sac_script() = &script;
  CameraLidarLineAssociate a(rows, cols);
  a.Associate(ln, segs, T);
  pcl::sac_script() = nullptr;
  if ((int)script.second != n_lists) return -2 - (int)script.second * 0;            // the script must match the number of FitLineRANSAC calls
  const std::vector<CameraLidarLinePair> pairs = a.GetAssociatedPairs();
  if ((int)pairs.size() > cap) return -1;
  for (size_t i = 0; i < pairs.size(); ++i) {
    for (int k = 0; k < 4; ++k) image_line4[4 * i + k] = pairs[i].image_line[k];
    for (int k = 0; k < 3; ++k) { start3[3 * i + k] = pairs[i].lidar_line_start[k]; end3[3 * i + k] = pairs[i].lidar_line_end[k]; }
    angle[i] = pairs[i].angle;
  }
  return (int)pairs.size();
}

// ProjectLidar2PanoramaDepth<pcl::PointXYZI> (util/Visualization.h:407-441): rows x cols uint16 image
void ref_project_lidar2panorama_depth(const float* cloud_xyzi, long n, int rows, int cols, const double* T_cl_rowmajor, int size, unsigned short* image) {
  pcl::PointCloud<pcl::PointXYZI> cloud;
  for (long i = 0; i < n; ++i) { pcl::PointXYZI p; p.x = cloud_xyzi[4 * i]; p.y = cloud_xyzi[4 * i + 1]; p.z = cloud_xyzi[4 * i + 2]; p.intensity = cloud_xyzi[4 * i + 3]; cloud.push_back(p); }
  Eigen::Matrix4d T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T(i, j) = T_cl_rowmajor[4 * i + j];
  const cv::Mat img = ProjectLidar2PanoramaDepth(cloud, rows, cols, T, (size_t)size);
  for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) image[(size_t)i * cols + j] = img.at<uint16_t>(i, j);
}
}
