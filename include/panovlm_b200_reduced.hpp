// panovlm_b200 — the reduced per-edge system as a residual block (header-only, no dependencies).
//
// north_star: "a second warp-reduce kernel that assembles the normal equations so Ceres only sees the reduced system".  The device hands back,
// per pose-graph edge e = (ref, nei), S_e = { H_e = sum_i J_i^T J_i (12 x 12, upper triangle), g_e = sum_i J_i^T r_i (12), c_e = sum_i rho_i / 2, n_e }
// over the loss-corrected rows of the edge (pvb_blocks_evaluate(want_system = 1), pvb_blocks_edge_systems).  A least-squares solver such as Ceres does
// not take normal equations, it takes residual blocks; this header turns S_e into ONE residual block with 13 residuals that reproduces it exactly:
//
//     P H_e P^T = L L^T             (Cholesky with diagonal pivoting; rank-deficient edges stop at their rank, the remaining rows are zero)
//     Jt = L^T P                    12 x 12   =>  Jt^T Jt = H_e
//     rt = L^+ P g_e                12        =>  Jt^T rt = g_e      (g_e lies in the range of H_e)
//     r13 = sqrt(max(0, 2 c_e - |rt|^2))      =>  (|rt|^2 + r13^2) / 2 = c_e,  zero Jacobian row
//
// so the solver's J^T J, J^T r and cost summed over the edges are the ones it would have formed from the n_e rows (util/Optimization.cpp:549-557 adds
// those rows one ceres::AutoDiffCostFunction at a time), while it touches n_edges blocks instead of n_rows (Room: 3.7 k instead of 1.1 M).
// The robust loss is already inside S_e (Ceres' Corrector applied per row before the products, rho'' <= 0 branch), so the block is added with a null loss.
#pragma once
#include <cmath>
#include <cstring>

namespace pvb {

struct ReducedEdgeBlock {
  double Jt[12 * 12];   // row-major: residual k, parameter column c (order aa_ref, t_ref, aa_nei, t_nei)
  double r[13];
  int rank;
};

// S92 = H upper row-major (78) | g (12) | cost | n   (include/panovlm_b200.h: pvb_blocks_edge_systems)
inline void reduce_edge_system(const double* S92, ReducedEdgeBlock* out) {
  double A[12][12];
  int q = 0;
  for (int a = 0; a < 12; ++a)
    for (int b = a; b < 12; ++b, ++q) { A[a][b] = S92[q]; A[b][a] = S92[q]; }
  int perm[12];
  for (int k = 0; k < 12; ++k) perm[k] = k;
  double L[12][12];
  std::memset(L, 0, sizeof(L));
  double dmax = 0.0;
  for (int k = 0; k < 12; ++k) dmax = A[k][k] > dmax ? A[k][k] : dmax;
  const double tol = dmax * 1e-13;
  int rank = 0;
  for (int k = 0; k < 12; ++k) {
    int piv = k;
    for (int j = k + 1; j < 12; ++j) if (A[j][j] > A[piv][piv]) piv = j;
    if (!(A[piv][piv] > tol)) break;
    if (piv != k) {            // symmetric swap of rows / columns k and piv of the trailing matrix, and of the finished part of L
      for (int j = 0; j < 12; ++j) { const double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
      for (int j = 0; j < 12; ++j) { const double t = A[j][k]; A[j][k] = A[j][piv]; A[j][piv] = t; }
      for (int j = 0; j < k; ++j) { const double t = L[k][j]; L[k][j] = L[piv][j]; L[piv][j] = t; }
      const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    const double d = std::sqrt(A[k][k]);
    L[k][k] = d;
    for (int i = k + 1; i < 12; ++i) L[i][k] = A[i][k] / d;
    for (int i = k + 1; i < 12; ++i)
      for (int j = k + 1; j <= i; ++j) { A[i][j] -= L[i][k] * L[j][k]; A[j][i] = A[i][j]; }
    ++rank;
  }
  out->rank = rank;
  // Jt = L^T P : Jt[k][perm[j]] = L[j][k]
  std::memset(out->Jt, 0, sizeof(out->Jt));
  for (int k = 0; k < rank; ++k)
    for (int j = k; j < 12; ++j) out->Jt[k * 12 + perm[j]] = L[j][k];
  // rt: L y = P g (first `rank` equations)
  double nrm = 0.0;
  for (int k = 0; k < 12; ++k) out->r[k] = 0.0;
  for (int k = 0; k < rank; ++k) {
    double s = S92[78 + perm[k]];
    for (int j = 0; j < k; ++j) s -= L[k][j] * out->r[j];
    out->r[k] = s / L[k][k];
    nrm += out->r[k] * out->r[k];
  }
  const double rest = 2.0 * S92[90] - nrm;
  out->r[12] = rest > 0.0 ? std::sqrt(rest) : 0.0;
}

}  // namespace pvb
