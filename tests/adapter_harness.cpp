// TEST INFRASTRUCTURE.  The reference-side Ceres bridges (include/panovlm_b200_ceres_adapter.hpp) exercised the way Ceres drives them, with the ceres stand-in of
// oracle/shim (a ceres::Problem that records AddResidualBlock calls and carries Problem::Options::evaluation_callback): blocks are registered through
// CeresBridge::AddBlocks / ReducedBridge::AddBlocks on pose lists laid out like lidar_mapping/LidarOdometry.cpp:23-24, then - as the solver would at every evaluation
// point - EvaluationCallback::PrepareForEvaluation runs once and every registered ceres::CostFunction::Evaluate is called with the blocks' own parameter pointers; the
// registered ceres::LossFunction is applied the way Ceres' ResidualBlock::Evaluate + Corrector do (rho'' <= 0: residual and Jacobian scaled by sqrt(rho'), cost rho / 2).
// Built by the tests with g++ against libpanovlm_b200.so; needs a CUDA device to run.
#include <cmath>
#include <cstring>
#include "panovlm_b200_ceres_adapter.hpp"

typedef std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>> Vec3List;

extern "C" int adapter_run(int device, long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts, int nb,
                           const double* poses6, int with_null_jacobian_block, double* r_out, double* J_out) {
  pvb_ctx* ctx = nullptr;
  if (pvb_create(device, &ctx) != PVB_OK) return -1;
  Vec3List aa(nb), t(nb);
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
        // the reference's loss objects: HuberLoss(huber[i]) shared between blocks of the same width, null where the reference passes nullptr
        const ceres::HuberLoss* hl = dynamic_cast<const ceres::HuberLoss*>(b.loss);
        if ((huber[i] > 0.0) != (b.loss != nullptr) || (hl && hl->a() != huber[i])) { rc = -4; break; }
        if (b.params.size() != 4 || b.params[0] != aa[ref[i]].data() || b.params[3] != t[nei[i]].data()) { rc = -4; break; }
        double jb[4][3]; double* jp[4] = {jb[0], jb[1], jb[2], jb[3]};
        std::memset(jb, 0, sizeof(jb));
        if (with_null_jacobian_block) jp[i % 4] = nullptr;                    // Ceres passes null for constant parameter blocks
        if (!b.cost->Evaluate(b.params.data(), r_out + i, jp)) { rc = -5; break; }
        if (b.loss) {                                                         // ceres::internal::Corrector, rho'' <= 0 branch
          double rho[3]; b.loss->Evaluate(r_out[i] * r_out[i], rho);
          const double s = std::sqrt(rho[1]);
          r_out[i] *= s;
          for (int q = 0; q < 12; ++q) jb[q / 3][q % 3] *= s;
        }
        std::memcpy(J_out + 12 * i, jb, sizeof(jb));
      }
    }
  }
  pvb_destroy(ctx);
  return rc;
}

// The same blocks through both bridges: the normal equations (dense 6 nb x 6 nb J^T J, J^T r) and the cost Ceres assembles from (a) the rows with their loss functions
// and (b) the reduced 13-residual blocks, one per pose-graph edge.  const_block (or -1): a parameter block held constant (null Jacobian pointers, zero rows / columns).
// out: H_rows, g_rows (D x D, D), H_red, g_red, costs[2], counts[2] = residual blocks seen by Ceres in (a) and (b).
extern "C" int adapter_reduced_run(int device, long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts, int nb,
                                   const double* poses6, int const_block, double* H_rows, double* g_rows, double* H_red, double* g_red, double* costs, long* counts) {
  pvb_ctx* ctx = nullptr;
  if (pvb_create(device, &ctx) != PVB_OK) return -1;
  Vec3List aa(nb), t(nb);
  for (int i = 0; i < nb; ++i) for (int k = 0; k < 3; ++k) { aa[i][k] = poses6[6 * i + k]; t[i][k] = poses6[6 * i + 3 + k]; }
  const int D = 6 * nb;
  int rc = 0;
  auto index_of = [&](const double* p) {                                      // parameter pointer -> first column of its 3-block
    for (int i = 0; i < nb; ++i) { if (p == aa[i].data()) return 6 * i; if (p == t[i].data()) return 6 * i + 3; }
    return -1;
  };
  auto accumulate = [&](ceres::Problem& problem, double* H, double* g, double* cost) {
    std::memset(H, 0, sizeof(double) * D * D); std::memset(g, 0, sizeof(double) * D); *cost = 0;
    problem.options.evaluation_callback->PrepareForEvaluation(true, true);
    for (const ceres::Problem::Block& b : problem.blocks) {
      const int nr = b.cost->num_residuals();
      std::vector<double> r(nr), J(4 * nr * 3, 0.0);
      double* jp[4]; int col[4];
      for (int k = 0; k < 4; ++k) {
        col[k] = index_of(b.params[k]);
        if (col[k] < 0) return -6;
        jp[k] = (const_block >= 0 && col[k] / 6 == const_block) ? nullptr : &J[(size_t)k * nr * 3];
      }
      if (!b.cost->Evaluate(b.params.data(), r.data(), jp)) return -5;
      double sq = 0; for (int q = 0; q < nr; ++q) sq += r[q] * r[q];
      double scale = 1.0, rho0 = sq;
      if (b.loss) { double rho[3]; b.loss->Evaluate(sq, rho); scale = std::sqrt(rho[1]); rho0 = rho[0]; }
      *cost += 0.5 * rho0;
      for (int q = 0; q < nr; ++q) {
        const double rq = r[q] * scale;
        for (int ka = 0; ka < 4; ++ka) for (int ca = 0; ca < 3; ++ca) {
          const double ja = J[((size_t)ka * nr + q) * 3 + ca] * scale;
          if (ja == 0.0) continue;
          g[col[ka] + ca] += ja * rq;
          for (int kb = 0; kb < 4; ++kb) for (int cb = 0; cb < 3; ++cb) H[(size_t)(col[ka] + ca) * D + col[kb] + cb] += ja * J[((size_t)kb * nr + q) * 3 + cb] * scale;
        }
      }
    }
    return 0;
  };
  {
    pvb::CeresBridge rows(ctx, aa, t);
    ceres::Problem::Options popt; popt.evaluation_callback = &rows;
    ceres::Problem problem(popt);
    if (!rows.AddBlocks(n, type, ref, nei, normalize, huber, consts, &problem)) rc = -2;
    if (rc == 0) rc = accumulate(problem, H_rows, g_rows, &costs[0]);
    counts[0] = (long)problem.blocks.size();
  }
  if (rc == 0) {
    pvb::ReducedBridge red(ctx, aa, t);
    ceres::Problem::Options popt; popt.evaluation_callback = &red;
    ceres::Problem problem(popt);
    if (!red.AddBlocks(n, type, ref, nei, normalize, huber, consts, &problem)) rc = -3;
    if (rc == 0) for (const ceres::Problem::Block& b : problem.blocks) if (b.loss != nullptr || b.cost->num_residuals() != 13) rc = -4;
    if (rc == 0) rc = accumulate(problem, H_red, g_red, &costs[1]);
    counts[1] = (long)problem.blocks.size();
  }
  pvb_destroy(ctx);
  return rc;
}

// panovlm_b200_reduced.hpp alone (no device): S92 -> the 13-residual block
extern "C" void reduced_block(const double* S92, double* Jt144, double* r13, int* rank) {
  pvb::ReducedEdgeBlock b;
  pvb::reduce_edge_system(S92, &b);
  std::memcpy(Jt144, b.Jt, sizeof(b.Jt)); std::memcpy(r13, b.r, sizeof(b.r)); *rank = b.rank;
}
