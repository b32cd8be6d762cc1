// panovlm_b200 — device side of the trust-region step: the reduced pose-graph normal equations stay in
// HBM, are assembled into a dense symmetric matrix, Jacobi-scaled, damped, factored (blocked right-looking Cholesky, FP64) and
// solved on the GPU; only scalars and 6N-vectors travel to the host loop that mirrors Ceres' Levenberg-Marquardt
// (util/Optimization.cpp:638-666 SetOptionsLidar -> DENSE_SCHUR / SPARSE_SCHUR on the CPU in the reference).
//
// Layout: matrices are row-major N x N with N = n rounded up to kNB (padding rows/columns are identity, so no kernel has an
// edge case); the factor overwrites the lower triangle.  Every sum runs in a fixed order => bit-reproducible results.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pvb {

constexpr int kNB = 64;            // panel width / tile edge of the factorisation

// One contribution of an edge system to a 6x6 destination block of the dense matrix (or to a 6-vector of the gradient).
// kind: bit 1 = row half (0 ref, 1 nei), bit 0 = column half of the edge's 12x12 system.
struct HContrib { int dest_r, dest_c, edge, kind; };

// index of (a, b), a <= b, in the row-major upper triangle of a 12 x 12 matrix
__host__ __device__ inline int upper12(int a, int b) { return a * 12 - a * (a - 1) / 2 + (b - a); }

// grid = number of destination blocks; seg[d]..seg[d+1] = its contributions (ascending edge order); 36 active threads
__global__ void __launch_bounds__(64) k_assemble_H(const HContrib* __restrict__ con, const int* __restrict__ seg, const double* __restrict__ esys, int N,
                                                   double* __restrict__ H) {
  const int t = threadIdx.x;
  if (t >= 36) return;
  const int i = t / 6, j = t % 6;
  const int s0 = seg[blockIdx.x], s1 = seg[blockIdx.x + 1];
  double acc = 0.0;
  int br = 0, bc = 0;
  for (int s = s0; s < s1; ++s) {
    const HContrib c = con[s];
    br = c.dest_r; bc = c.dest_c;
    const int a = ((c.kind >> 1) & 1) * 6 + i, b = (c.kind & 1) * 6 + j;
    acc += esys[(size_t)c.edge * 92 + (a <= b ? upper12(a, b) : upper12(b, a))];
  }
  H[(size_t)(6 * br + i) * N + 6 * bc + j] = acc;
}

// gradient: contributions (dest_r = free block, edge, kind = half); grid = number of free blocks with contributions, 6 active threads
__global__ void __launch_bounds__(32) k_assemble_g(const HContrib* __restrict__ con, const int* __restrict__ seg, const double* __restrict__ esys, double* __restrict__ g) {
  const int t = threadIdx.x;
  if (t >= 6) return;
  const int s0 = seg[blockIdx.x], s1 = seg[blockIdx.x + 1];
  double acc = 0.0;
  int br = 0;
  for (int s = s0; s < s1; ++s) {
    const HContrib c = con[s];
    br = c.dest_r;
    acc += esys[(size_t)c.edge * 92 + 78 + (c.kind & 1) * 6 + t];
  }
  g[6 * br + t] = acc;
}

// Jacobi scaling fixed from the first Jacobian: sc = 1 / (1 + sqrt(H_ii))  (Ceres jacobi_scaling)
__global__ void k_jacobi_scale(const double* __restrict__ H, int n, int N, double* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) sc[i] = i < n ? 1.0 / (1.0 + sqrt(H[(size_t)i * N + i])) : 1.0;
}

// A = diag(sc) H diag(sc) + D / radius on the lower triangle, D = clamp(diag, 1e-6, 1e32); rhs = -g * sc; padding -> identity
__global__ void __launch_bounds__(256) k_build_damped(const double* __restrict__ H, const double* __restrict__ g, const double* __restrict__ sc, int n, int N,
                                                      double radius, double* __restrict__ A, double* __restrict__ rhs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= N || j > i) return;
  double v;
  if (i >= n) v = (i == j) ? 1.0 : 0.0;
  else {
    v = H[(size_t)i * N + j] * sc[i] * sc[j];
    if (i == j) { v += fmin(fmax(v, 1e-6), 1e32) / radius; }
  }
  A[(size_t)i * N + j] = v;
  if (j == 0) rhs[i] = i < n ? -g[i] * sc[i] : 0.0;
}

// ---- blocked right-looking Cholesky ---------------------------------------------------------------------------------------------
// diagonal block: A[k0:k0+NB, k0:k0+NB] -> L11 (in place); *fail set when a pivot is not positive
__global__ void __launch_bounds__(256) k_potrf_diag(double* __restrict__ A, int N, int k0, int* __restrict__ fail) {
  __shared__ double s[kNB][kNB + 1];
  const int t = threadIdx.x;
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; s[i][j] = j <= i ? A[(size_t)(k0 + i) * N + k0 + j] : 0.0; }
  __syncthreads();
  for (int j = 0; j < kNB; ++j) {
    // column j: every row i >= j subtracts its left part in ascending k (the order of the host's left-looking loop)
    if (t < kNB && t >= j) {
      double v = s[t][j];
      for (int k = 0; k < j; ++k) v -= s[t][k] * s[j][k];
      s[t][j] = v;
    }
    __syncthreads();
    const double d = s[j][j];
    if (!(d > 0.0)) { if (t == 0) *fail = 1; return; }
    const double r = sqrt(d);
    __syncthreads();
    if (t < kNB && t > j) s[t][j] /= r;
    if (t == j) s[j][j] = r;
    __syncthreads();
  }
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; if (j <= i) A[(size_t)(k0 + i) * N + k0 + j] = s[i][j]; }
}

// panel: rows below the diagonal block, L21 = A21 L11^-T; one thread per row, 64 rows per block
__global__ void __launch_bounds__(kNB) k_trsm_panel(double* __restrict__ A, int N, int k0) {
  extern __shared__ double trsm_smem[];                         // 2 x 64 x 65 doubles (dynamic: above the 48 KB static limit)
  double (*L)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(trsm_smem);
  double (*R)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(trsm_smem + kNB * (kNB + 1));
  const int t = threadIdx.x;
  const int r0 = k0 + kNB + blockIdx.x * kNB;
  for (int e = t; e < kNB * kNB; e += kNB) {
    const int i = e / kNB, j = e % kNB;
    L[i][j] = A[(size_t)(k0 + i) * N + k0 + j];
    R[i][j] = A[(size_t)(r0 + i) * N + k0 + j];
  }
  __syncthreads();
  for (int j = 0; j < kNB; ++j) {
    double v = R[t][j];
    for (int k = 0; k < j; ++k) v -= R[t][k] * L[j][k];
    R[t][j] = v / L[j][j];
  }
  __syncthreads();
  for (int e = t; e < kNB * kNB; e += kNB) { const int i = e / kNB, j = e % kNB; A[(size_t)(r0 + i) * N + k0 + j] = R[i][j]; }
}

// trailing update: for every lower-triangular tile (ti >= tj) of the rows/columns after the panel,
// C[ti][tj] -= P[ti] P[tj]^T with P = the 64-wide panel; 256 threads, 4 x 4 outputs each, k ascending
__global__ void __launch_bounds__(256) k_syrk_update(double* __restrict__ A, int N, int k0) {
  constexpr int KH = 32;                                        // the 64-wide panel goes through shared memory in two halves
  __shared__ double Pa[kNB][KH + 1];
  __shared__ double Pb[kNB][KH + 1];
  // linear block id -> (ti, tj), tj <= ti
  const int b = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((long long)(ti + 1) * (ti + 2) / 2 <= b) ++ti;
  while ((long long)ti * (ti + 1) / 2 > b) --ti;
  const int tj = b - ti * (ti + 1) / 2;
  const int base = k0 + kNB;
  const int ra = base + ti * kNB, rb = base + tj * kNB;
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
  for (int kh = 0; kh < kNB; kh += KH) {
    __syncthreads();
    for (int e = t; e < kNB * KH; e += 256) {
      const int i = e / KH, j = e % KH;
      Pa[i][j] = A[(size_t)(ra + i) * N + k0 + kh + j];
      Pb[i][j] = A[(size_t)(rb + i) * N + k0 + kh + j];
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KH; ++k) {
      double a[4], c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = Pa[ty + 16 * u][k]; c[u] = Pb[tx + 16 * u][k]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * c[v];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = ra + ty + 16 * u, j = rb + tx + 16 * v;
      if (j <= i) A[(size_t)i * N + j] -= acc[u][v];
    }
}

// forward + backward substitution with the factor, one thread block (1024 threads): x overwrites rhs
__global__ void __launch_bounds__(1024) k_chol_solve(const double* __restrict__ A, int N, double* __restrict__ x) {
  __shared__ double L[kNB][kNB + 1];
  __shared__ double xs[kNB];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nblk = N / kNB;
  // L y = b
  for (int b = 0; b < nblk; ++b) {
    const int k0 = b * kNB;
    for (int e = t; e < kNB * kNB; e += 1024) { const int i = e / kNB, j = e % kNB; L[i][j] = A[(size_t)(k0 + i) * N + k0 + j]; }
    if (t < kNB) xs[t] = x[k0 + t];
    __syncthreads();
    for (int j = 0; j < kNB; ++j) {
      if (t == j) xs[j] = xs[j] / L[j][j];
      __syncthreads();
      if (t < kNB && t > j) xs[t] -= L[t][j] * xs[j];
      __syncthreads();
    }
    if (t < kNB) x[k0 + t] = xs[t];
    // rows below: x[i] -= L[i, k0:k0+NB] . xs  (one warp per row)
    for (int i = k0 + kNB + warp; i < N; i += 32) {
      const double* row = A + (size_t)i * N + k0;
      double s = row[lane] * xs[lane] + row[lane + 32] * xs[lane + 32];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) x[i] -= s;
    }
    __syncthreads();
  }
  // L^T x = y
  for (int b = nblk - 1; b >= 0; --b) {
    const int k0 = b * kNB;
    for (int e = t; e < kNB * kNB; e += 1024) { const int i = e / kNB, j = e % kNB; L[i][j] = A[(size_t)(k0 + i) * N + k0 + j]; }
    if (t < kNB) xs[t] = x[k0 + t];
    __syncthreads();
    for (int j = kNB - 1; j >= 0; --j) {
      if (t == j) xs[j] = xs[j] / L[j][j];
      __syncthreads();
      if (t < j) xs[t] -= L[j][t] * xs[j];
      __syncthreads();
    }
    if (t < kNB) x[k0 + t] = xs[t];
    // columns to the left: x[c] -= sum_i L[k0 + i][c] xs[i]  (one thread per column, rows ascending)
    for (int c = t; c < k0; c += 1024) {
      double s = 0.0;
      for (int i = 0; i < kNB; ++i) s += A[(size_t)(k0 + i) * N + c] * xs[i];
      x[c] -= s;
    }
    __syncthreads();
  }
}

// model-cost terms of the step: term[i] = y_i (g_i sc_i + 0.5 (Hsc y)_i) with Hsc = diag(sc) H diag(sc); one warp per row
__global__ void __launch_bounds__(256) k_model_terms(const double* __restrict__ H, const double* __restrict__ g, const double* __restrict__ sc, const double* __restrict__ y,
                                                     int n, int N, double* __restrict__ term) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* h = H + (size_t)row * N;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += h[j] * sc[row] * sc[j] * y[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) term[row] = y[row] * (g[row] * sc[row] + 0.5 * s);
}

}  // namespace pvb
