// panovlm_b200 — header-only adapter that plugs the C ABI (panovlm_b200.h) into an existing ceres::Problem.
//
// NOT compiled in this repository (Ceres is not installed in the build image); it is the reference-side binding a
// PanoVLM maintainer adds.  It keeps the surface the reference uses today:
//   * util/Optimization.cpp:549,555,417,430,592,602  problem.AddResidualBlock(cost, loss, aa_rw, t_rw, aa_nw, t_nw)
//   * lidar_mapping/LidarOdometry.cpp:78-80           ceres::Solve(options, &problem, &summary)
// and replaces the per-correspondence AutoDiffCostFunction<F,1,3,3,3,3> evaluation (base/CostFunction.h:567-1022) by
// one device evaluation per ceres evaluation point, driven by a ceres::EvaluationCallback (Ceres >= 2.0).
//
// Usage (inside LidarOdometry::RefinePose, replacing the Add*Residual calls):
//
//   pvb::CeresBridge bridge(ctx, angleAxis_lw_list, t_lw_list);        // pose arrays stay where they are
//   ceres::Problem::Options popt; popt.evaluation_callback = &bridge;
//   ceres::Problem problem(popt);
//   ...                                                                  // associations via pvb_frames_* / pvb_line2line_*
//   bridge.AddBlocks(n, type, ref, nei, normalize, huber, consts, &problem);   // one thin CostFunction per correspondence
//   problem.SetParameterBlockConstant(...first valid frame...);         // unchanged (LidarOdometry.cpp:59-66)
//   ceres::Solve(SetOptionsLidar(threads, n), &problem, &summary);     // unchanged
//
// Robust loss.  CeresBridge (one Row per correspondence) serves RAW residuals / Jacobians and registers the reference's own
// ceres::HuberLoss (one shared object per distinct width, as util/Optimization.cpp:513-517 does), so Ceres' Corrector, cost, step-quality
// ratio and final_cost are exactly the reference's.  (Serving device-corrected rows with a null loss would give the right r and J but
// the cost 0.5 * rho' * r^2 instead of 0.5 * rho(r^2) for outliers.)
//
// ReducedBridge is the north-star form: Ceres sees ONE 13-residual block per pose-graph edge built from the per-edge normal equations the
// device reduces (include/panovlm_b200_reduced.hpp): the same J^T J, J^T r and robust cost as the rows give, n_edges blocks instead of n_rows.
#pragma once
#include <ceres/ceres.h>
#include <Eigen/Core>
#include <map>
#include <memory>
#include <vector>
#include "panovlm_b200.h"
#include "panovlm_b200_reduced.hpp"

namespace pvb {

class CeresBridge : public ceres::EvaluationCallback {
 public:
  // pose arrays: the eigen_vector<Eigen::Vector3d> lists of lidar_mapping/LidarOdometry.cpp:23-24 (stable addresses)
  CeresBridge(pvb_ctx* ctx, std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>>& aa,
              std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>>& t)
      : ctx_(ctx), aa_(aa), t_(t), poses_(6 * aa.size()) {}

  // Called once, single-threaded, before Ceres evaluates all residual blocks at a point.
  void PrepareForEvaluation(bool evaluate_jacobians, bool /*new_evaluation_point*/) override {
    for (size_t i = 0; i < aa_.size(); ++i) {
      for (int k = 0; k < 3; ++k) { poses_[6 * i + k] = aa_[i][k]; poses_[6 * i + 3 + k] = t_[i][k]; }
    }
    ok_ = pvb_blocks_evaluate(ctx_, poses_.data(), /*want_rows=*/2 /* raw: the loss is Ceres' */, /*want_system=*/0) == PVB_OK;
    r_ = pvb_blocks_residuals(ctx_);
    J_ = evaluate_jacobians ? pvb_blocks_jacobians(ctx_) : nullptr;
  }

  // Thin per-correspondence cost function: copies one device-computed row.  Thread-safe (read-only).
  class Row : public ceres::SizedCostFunction<1, 3, 3, 3, 3> {
   public:
    Row(const CeresBridge* b, long i) : b_(b), i_(i) {}
    bool Evaluate(double const* const*, double* residuals, double** jacobians) const override {
      if (!b_->ok_) return false;
      residuals[0] = b_->r_[i_];
      if (jacobians) {
        const double* row = b_->Jrow(i_);
        for (int blk = 0; blk < 4; ++blk)
          if (jacobians[blk]) for (int k = 0; k < 3; ++k) jacobians[blk][k] = row ? row[3 * blk + k] : 0.0;
      }
      return true;
    }
   private:
    const CeresBridge* b_; long i_;
  };

  // Registers the blocks with the library and adds one Row per block to the problem, with ceres::HuberLoss(huber[i]) (null for huber[i] <= 0,
  // like AddLidarLineToLineResidual2's angle residual, util/Optimization.cpp:417); one loss object per distinct width, shared, owned by the problem.
  bool AddBlocks(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts,
                 ceres::Problem* problem) {
    if (pvb_blocks_set(ctx_, n, type, ref, nei, normalize, huber, consts, (int)aa_.size()) != PVB_OK) return false;
    std::map<double, ceres::LossFunction*> losses;
    for (long i = 0; i < n; ++i) {
      ceres::LossFunction* loss = nullptr;
      if (huber[i] > 0.0) {
        auto it = losses.find(huber[i]);
        if (it == losses.end()) it = losses.emplace(huber[i], new ceres::HuberLoss(huber[i])).first;
        loss = it->second;
      }
      problem->AddResidualBlock(new Row(this, i), loss, aa_[ref[i]].data(), t_[ref[i]].data(), aa_[nei[i]].data(), t_[nei[i]].data());
    }
    return true;
  }

 private:
  friend class Row;
  const double* Jrow(long i) const { return J_ ? J_ + 12 * i : nullptr; }
  pvb_ctx* ctx_;
  std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>>& aa_;
  std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>>& t_;
  std::vector<double> poses_;
  const double* r_ = nullptr;
  const double* J_ = nullptr;
  bool ok_ = false;
};

// ---- the reduced form: one residual block per pose-graph edge ---------------------------------------------------------------------------
// Usage is CeresBridge's; the problem then holds pvb_blocks_num_edges() blocks of 13 residuals on (aa_ref, t_ref, aa_nei, t_nei) instead of one block
// per correspondence.  Per evaluation point: one device evaluation with the per-edge reduction (k_eval_blocks + k_sum_partials), n_edges x 92 doubles
// back to the host, one 12 x 12 pivoted Cholesky per edge (panovlm_b200_reduced.hpp).  Edges with ref == nei cannot be expressed as a Ceres block
// with four distinct parameter blocks: AddBlocks refuses them (use CeresBridge for such problems).
class ReducedBridge : public ceres::EvaluationCallback {
 public:
  typedef std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> Vec3List;
  ReducedBridge(pvb_ctx* ctx, Vec3List& aa, Vec3List& t) : ctx_(ctx), aa_(aa), t_(t), poses_(6 * aa.size()) {}

  void PrepareForEvaluation(bool /*evaluate_jacobians*/, bool /*new_evaluation_point*/) override {
    for (size_t i = 0; i < aa_.size(); ++i)
      for (int k = 0; k < 3; ++k) { poses_[6 * i + k] = aa_[i][k]; poses_[6 * i + 3 + k] = t_[i][k]; }
    ok_ = pvb_blocks_evaluate(ctx_, poses_.data(), /*want_rows=*/0, /*want_system=*/1) == PVB_OK;
    const double* S = ok_ ? pvb_blocks_edge_systems_ptr(ctx_) : nullptr;
    ok_ = ok_ && S != nullptr;
    if (!ok_) return;
    const long ne = (long)blocks_.size();
#if defined(_OPENMP)
#pragma omp parallel for schedule(static)
#endif
    for (long e = 0; e < ne; ++e) reduce_edge_system(S + 92 * e, &blocks_[e]);
  }

  class EdgeCost : public ceres::SizedCostFunction<13, 3, 3, 3, 3> {
   public:
    EdgeCost(const ReducedBridge* b, long e) : b_(b), e_(e) {}
    bool Evaluate(double const* const*, double* residuals, double** jacobians) const override {
      if (!b_->ok_) return false;
      const ReducedEdgeBlock& B = b_->blocks_[e_];
      for (int k = 0; k < 13; ++k) residuals[k] = B.r[k];
      if (jacobians) {
        for (int blk = 0; blk < 4; ++blk) {
          if (!jacobians[blk]) continue;                                   // constant parameter block
          for (int k = 0; k < 12; ++k) for (int c = 0; c < 3; ++c) jacobians[blk][k * 3 + c] = B.Jt[k * 12 + 3 * blk + c];
          for (int c = 0; c < 3; ++c) jacobians[blk][12 * 3 + c] = 0.0;    // the cost-completing residual has no Jacobian
        }
      }
      return true;
    }
   private:
    const ReducedBridge* b_; long e_;
  };

  bool AddBlocks(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts,
                 ceres::Problem* problem) {
    if (pvb_blocks_set(ctx_, n, type, ref, nei, normalize, huber, consts, (int)aa_.size()) != PVB_OK) return false;
    const int ne = pvb_blocks_num_edges(ctx_);
    if (ne < 0) return false;
    std::vector<int> er(ne), en(ne);
    if (ne > 0 && pvb_blocks_edges(ctx_, er.data(), en.data()) != PVB_OK) return false;
    for (int e = 0; e < ne; ++e) if (er[e] == en[e]) return false;
    blocks_.assign(ne, ReducedEdgeBlock());
    for (int e = 0; e < ne; ++e)
      problem->AddResidualBlock(new EdgeCost(this, e), nullptr, aa_[er[e]].data(), t_[er[e]].data(), aa_[en[e]].data(), t_[en[e]].data());
    return true;
  }
  long num_edges() const { return (long)blocks_.size(); }

 private:
  friend class EdgeCost;
  pvb_ctx* ctx_;
  Vec3List& aa_;
  Vec3List& t_;
  std::vector<double> poses_;
  std::vector<ReducedEdgeBlock> blocks_;
  bool ok_ = false;
};

// The camera-camera term: AddCameraResidual (util/Optimization.cpp:172-222, ANGLE_RESIDUAL_1) hands Ceres one
// AutoDiffCostFunction<PanoramaReprojResidual_1Angle,1,3,3,3> per (track, observation) on the blocks (aa_cw, t_cw, point_3d).  This bridge evaluates
// all of them in one launch per evaluation point (pvb_reproj_evaluate) and serves the raw 1x9 rows; the reference's HuberLoss(4 deg) is registered with Ceres.
// The structure points are read where they live (PointTrack::point_3d): `points` holds their addresses in track order.
// A ceres::Problem takes ONE evaluation callback: when both bridges are used (CameraLidarOptimizer::Optimize) register a small callback that forwards
// PrepareForEvaluation to the two of them.
class ReprojBridge : public ceres::EvaluationCallback {
 public:
  typedef std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> Vec3List;
  ReprojBridge(pvb_ctx* ctx, Vec3List& aa_cw, Vec3List& t_cw, std::vector<double*> points)
      : ctx_(ctx), aa_(aa_cw), t_(t_cw), points_(std::move(points)), cams_(6 * aa_cw.size()), xyz_(3 * points_.size()) {}

  void PrepareForEvaluation(bool evaluate_jacobians, bool /*new_evaluation_point*/) override {
    for (size_t i = 0; i < aa_.size(); ++i)
      for (int k = 0; k < 3; ++k) { cams_[6 * i + k] = aa_[i][k]; cams_[6 * i + 3 + k] = t_[i][k]; }
    for (size_t p = 0; p < points_.size(); ++p)
      for (int k = 0; k < 3; ++k) xyz_[3 * p + k] = points_[p][k];
    ok_ = pvb_reproj_evaluate(ctx_, cams_.data(), xyz_.data(), /*want_rows=*/2 /* raw: the loss is Ceres' */, /*want_system=*/0) == PVB_OK;
    r_ = pvb_reproj_residuals(ctx_);
    J_ = evaluate_jacobians ? pvb_reproj_jacobians(ctx_) : nullptr;
  }

  class Row : public ceres::SizedCostFunction<1, 3, 3, 3> {
   public:
    Row(const ReprojBridge* b, long i) : b_(b), i_(i) {}
    bool Evaluate(double const* const*, double* residuals, double** jacobians) const override {
      if (!b_->ok_) return false;
      residuals[0] = b_->r_[i_];
      if (jacobians) {
        const double* row = b_->J_ ? b_->J_ + 9 * i_ : nullptr;       // [d aa_cw | d t_cw | d point_3d]
        for (int blk = 0; blk < 3; ++blk)
          if (jacobians[blk]) for (int k = 0; k < 3; ++k) jacobians[blk][k] = row ? row[3 * blk + k] : 0.0;
      }
      return true;
    }
   private:
    const ReprojBridge* b_; long i_;
  };

  // observation i: camera cam[i] sees point[i] along bearing[i] (unit sphere, eq.ImageToCam of the key point; see pvb_build_reproj_observations)
  bool AddObservations(long n, const int* cam, const int* point, const double* bearing3, double weight, double huber, ceres::Problem* problem) {
    if (pvb_reproj_set(ctx_, n, cam, point, bearing3, weight, huber, (int)aa_.size(), (long)points_.size()) != PVB_OK) return false;
    ceres::LossFunction* loss = huber > 0.0 ? new ceres::HuberLoss(huber) : nullptr;        // HuberLoss(4 deg), util/Optimization.cpp:185; shared, owned by the problem
    for (long i = 0; i < n; ++i) problem->AddResidualBlock(new Row(this, i), loss, aa_[cam[i]].data(), t_[cam[i]].data(), points_[point[i]]);
    return true;
  }

 private:
  friend class Row;
  pvb_ctx* ctx_;
  Vec3List& aa_;
  Vec3List& t_;
  std::vector<double*> points_;
  std::vector<double> cams_, xyz_;
  const double* r_ = nullptr;
  const double* J_ = nullptr;
  bool ok_ = false;
};

}  // namespace pvb
