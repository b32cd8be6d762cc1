// TEST INFRASTRUCTURE.  The reference-side Ceres bridge (include/panovlm_b200_ceres_adapter.hpp) exercised the way Ceres drives it, with the ceres stand-in of
// oracle/shim (a ceres::Problem that records AddResidualBlock calls and carries Problem::Options::evaluation_callback): blocks are registered through
// CeresBridge::AddBlocks on pose lists laid out like lidar_mapping/LidarOdometry.cpp:23-24, then - as the solver would at every evaluation point -
// EvaluationCallback::PrepareForEvaluation runs once and every registered ceres::CostFunction::Evaluate is called with the blocks' own parameter pointers.
// Built by tests/test_zz_gpu_reference_fixtures.py with g++ against libpanovlm_b200.so; needs a CUDA device to run.
#include <cstring>
#include "panovlm_b200_ceres_adapter.hpp"

extern "C" int adapter_run(int device, long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts, int nb,
                           const double* poses6, int with_null_jacobian_block, double* r_out, double* J_out) {
  pvb_ctx* ctx = nullptr;
  if (pvb_create(device, &ctx) != PVB_OK) return -1;
  std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> aa(nb), t(nb);
  for (int i = 0; i < nb; ++i) for (int k = 0; k < 3; ++k) { aa[i][k] = poses6[6 * i + k]; t[i][k] = poses6[6 * i + 3 + k]; }
  int rc = 0;
  {
    pvb::CeresBridge bridge(ctx, aa, t);
    ceres::Problem::Options popt; popt.evaluation_callback = &bridge;
    ceres::Problem problem(popt);
    if (!bridge.AddBlocks(n, type, ref, nei, normalize, huber, consts, &problem)) rc = -2;
    if (rc == 0 && (long)problem.blocks.size() != n) rc = -3;
    if (rc == 0) {
      problem.options.evaluation_callback->PrepareForEvaluation(/*evaluate_jacobians=*/true, /*new_evaluation_point=*/true);
      for (long i = 0; i < n && rc == 0; ++i) {
        const ceres::Problem::Block& b = problem.blocks[i];
        if (b.loss != nullptr || b.params.size() != 4 || b.params[0] != aa[ref[i]].data() || b.params[3] != t[nei[i]].data()) { rc = -4; break; }
        double jb[4][3]; double* jp[4] = {jb[0], jb[1], jb[2], jb[3]};
        std::memset(jb, 0, sizeof(jb));
        if (with_null_jacobian_block) jp[i % 4] = nullptr;                    // Ceres passes null for constant parameter blocks
        if (!b.cost->Evaluate(b.params.data(), r_out + i, jp)) { rc = -5; break; }
        std::memcpy(J_out + 12 * i, jb, sizeof(jb));
      }
    }
  }
  pvb_destroy(ctx);
  return rc;
}
