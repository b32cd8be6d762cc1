// panovlm_b200 — device side of the trust-region step: the reduced pose-graph normal equations stay in
// HBM, are assembled into a dense symmetric matrix, Jacobi-scaled, damped, factored (blocked right-looking Cholesky, FP64) and
// solved on the GPU; only scalars and 6N-vectors travel to the host loop that mirrors Ceres' Levenberg-Marquardt
// (util/Optimization.cpp:638-666 SetOptionsLidar -> DENSE_SCHUR / SPARSE_SCHUR on the CPU in the reference).
//
// Layout: matrices are row-major N x N with N = n rounded up to kNB (padding rows/columns are identity, so no kernel has an
// edge case); the factor overwrites the lower triangle.  Every sum runs in a fixed order => bit-reproducible results.
// Per 64-wide step: k_potrf_diag (1 block) -> k_trsm_panel (64 rows per block) -> k_fwd_step (forward substitution of this block, fused
// into the factorisation sweep) -> k_syrk_update_mma (FP64 tensor cores); then one k_bwd_step per block from the last to the first.
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
// diagonal block: A[k0:k0+NB, k0:k0+NB] -> L11 (in place); *fail set when a pivot is not positive.  Right-looking inside the block:
// column j is scaled, then the whole trailing triangle takes its rank-1 update in parallel (256 threads); every element still
// receives its subtractions in ascending j, the order of a left-looking loop.
__global__ void __launch_bounds__(256) k_potrf_diag(double* __restrict__ A, int N, int k0, int* __restrict__ fail) {
  __shared__ double s[kNB][kNB + 1];
  const int t = threadIdx.x;
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; s[i][j] = j <= i ? A[(size_t)(k0 + i) * N + k0 + j] : 0.0; }
  __syncthreads();
  const int ui = t >> 4, uk = t & 15;                  // update mapping: rows ui + 16 a, columns uk + 16 b
  for (int j = 0; j < kNB; ++j) {
    const double d = s[j][j];
    if (!(d > 0.0)) { if (t == 0) *fail = 1; return; }
    const double r = sqrt(d), rinv = 1.0 / r;
    __syncthreads();
    if (t < kNB && t > j) s[t][j] *= rinv;
    if (t == j) s[j][j] = r;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ui + 16 * a;
      if (i <= j) continue;
      const double sij = s[i][j];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int k = uk + 16 * b;
        if (k > j && k <= i) s[i][k] -= sij * s[k][j];
      }
    }
    __syncthreads();
  }
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; if (j <= i) A[(size_t)(k0 + i) * N + k0 + j] = s[i][j]; }
}

// panel: rows below the diagonal block, L21 = A21 L11^-T; 64 rows per block, 256 threads.  Column j is divided by L[j][j], then the
// columns to its right take their update in parallel (thread = row x column group); per element the subtractions run in ascending j.
__global__ void __launch_bounds__(256) k_trsm_panel(double* __restrict__ A, int N, int k0) {
  extern __shared__ double trsm_smem[];                         // 2 x 64 x 65 doubles (dynamic: above the 48 KB static limit)
  double (*L)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(trsm_smem);
  double (*R)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(trsm_smem + kNB * (kNB + 1));
  const int t = threadIdx.x;
  const int r0 = k0 + kNB + blockIdx.x * kNB;
  for (int e = t; e < kNB * kNB; e += 256) {
    const int i = e / kNB, j = e % kNB;
    L[i][j] = A[(size_t)(k0 + i) * N + k0 + j];
    R[i][j] = A[(size_t)(r0 + i) * N + k0 + j];
  }
  __syncthreads();
  __shared__ double dinv[kNB];
  if (t < kNB) dinv[t] = 1.0 / L[t][t];
  __syncthreads();
  const int row = t & (kNB - 1), cg = t >> 6;
  for (int j = 0; j < kNB; ++j) {
    if (t < kNB) R[t][j] *= dinv[j];
    __syncthreads();
    const double rj = R[row][j];
    for (int k = j + 1 + cg; k < kNB; k += 4) R[row][k] -= rj * L[k][j];
    __syncthreads();
  }
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; A[(size_t)(r0 + i) * N + k0 + j] = R[i][j]; }
}

// trailing update on the FP64 tensor cores (the one dense contraction of this library: C -= P P^T with the 64-wide panel P).
// mma.sync.m8n8k4.f64: A fragment = P[row0 + (lane >> 2)][k + (lane & 3)], B fragment (col layout) = the same pattern on the column tile,
// C fragment = 2 doubles per lane at [lane >> 2][2 (lane & 3) + {0, 1}].  Block = 128 rows x 64 columns of the lower triangle, 8 warps of
// 32 x 32 (4 x 4 mma tiles); the two panel tiles are staged once in shared memory with a row stride of 68 doubles (conflict-free fragment
// loads).  k runs ascending inside every accumulator => bit-reproducible.
constexpr int kSyrkLd = kNB + 4;
constexpr size_t kSyrkSmem = (size_t)(128 + 64) * kSyrkLd * sizeof(double);
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k_syrk_update_mma(double* __restrict__ A, int N, int k0) {
  extern __shared__ double syrk_smem[];
  double* Pa = syrk_smem;                       // 128 x 68
  double* Pb = syrk_smem + 128 * kSyrkLd;       // 64 x 68
  // linear block id -> (ti, tj): row tile ti of 128 rows owns the column tiles tj = 0 .. 2 ti + 1 of 64 columns
  const int b = blockIdx.x;
  int ti = (int)((sqrt(4.0 * (double)b + 1.0) - 1.0) * 0.5);
  while ((long long)(ti + 1) * (ti + 2) <= b) ++ti;
  while ((long long)ti * (ti + 1) > b) --ti;
  const int tj = b - ti * (ti + 1);
  const int base = k0 + kNB;
  const int ra = base + ti * 128, rb = base + tj * 64;
  if (rb >= N) return;                          // odd number of 64-tiles: the last row tile has one column tile less
  const int t = threadIdx.x;
  for (int e = t; e < 128 * kNB; e += 256) {
    const int i = e / kNB, j = e % kNB;
    Pa[i * kSyrkLd + j] = ra + i < N ? A[(size_t)(ra + i) * N + k0 + j] : 0.0;
  }
  for (int e = t; e < 64 * kNB; e += 256) {
    const int i = e / kNB, j = e % kNB;
    Pb[i * kSyrkLd + j] = A[(size_t)(rb + i) * N + k0 + j];
  }
  __syncthreads();
  const int w = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
  const int wr = (w >> 1) * 32, wc = (w & 1) * 32;
  if (rb + wc > ra + wr + 31) return;           // warp tile entirely above the diagonal
  double c[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) c[mi][ni][0] = c[mi][ni][1] = 0.0;
  const double* pa = Pa + (wr + g) * kSyrkLd + q;
  const double* pb = Pb + (wc + g) * kSyrkLd + q;
#pragma unroll 4
  for (int k = 0; k < kNB; k += 4) {
    double a[4], bb[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = pa[mi * 8 * kSyrkLd + k];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) bb[ni] = pb[ni * 8 * kSyrkLd + k];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma_m8n8k4(c[mi][ni][0], c[mi][ni][1], a[mi], bb[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    const int i = ra + wr + mi * 8 + g;
    if (i >= N) continue;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int j = rb + wc + ni * 8 + 2 * q;
      double* dst = A + (size_t)i * N + j;
      if (j + 1 <= i) { double2 v = *reinterpret_cast<double2*>(dst); v.x -= c[mi][ni][0]; v.y -= c[mi][ni][1]; *reinterpret_cast<double2*>(dst) = v; }
      else if (j <= i) dst[0] -= c[mi][ni][0];
    }
  }
}

// ---- substitution, one launch per 64-block and direction, spread over many thread blocks ----------------------------------------------
// forward (L y = b), launched right after the panel of step k0 is final: every block solves the 64 x 64 diagonal system redundantly in shared
// memory (no grid-wide barrier needed), block 0 stores y[k0 .. k0+63], and the blocks subtract L21 y from their 64 rows of b below
// (one warp per row, coalesced, fixed-order shuffle reduction).  b is updated in place below the block; y is a separate vector.
__global__ void __launch_bounds__(256) k_fwd_step(const double* __restrict__ A, int N, int k0, double* __restrict__ b, double* __restrict__ y) {
  __shared__ double L[kNB][kNB + 1];
  __shared__ double ys[kNB];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; L[i][j] = A[(size_t)(k0 + i) * N + k0 + j]; }
  __syncthreads();
  if (warp == 0) {                                     // 64 unknowns in one warp (two per lane): no block barrier inside the chain
    double v0 = b[k0 + lane], v1 = b[k0 + 32 + lane];
    const double d0 = 1.0 / L[lane][lane], d1 = 1.0 / L[lane + 32][lane + 32];
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const double yj = __shfl_sync(0xffffffffu, v0 * d0, j);
      if (lane == j) v0 = yj;
      if (lane > j) v0 -= L[lane][j] * yj;
      v1 -= L[lane + 32][j] * yj;
    }
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const double yj = __shfl_sync(0xffffffffu, v1 * d1, j);
      if (lane == j) v1 = yj;
      if (lane > j) v1 -= L[lane + 32][j + 32] * yj;
    }
    ys[lane] = v0; ys[lane + 32] = v1;
  }
  __syncthreads();
  if (blockIdx.x == 0 && t < kNB) y[k0 + t] = ys[t];
  const int r0 = k0 + kNB + blockIdx.x * kNB;
  for (int i = r0 + warp; i < r0 + kNB && i < N; i += 8) {
    const double* row = A + (size_t)i * N + k0;
    double s = row[lane] * ys[lane] + row[lane + 32] * ys[lane + 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) b[i] -= s;
  }
}

// backward (L^T x = y), blocks visited from the last to the first: the diagonal system is again solved redundantly, block 0 stores
// x[k0 .. k0+63] (a separate vector), and every block subtracts this block's contribution from its 256 entries of y to the left
// (one thread per column, rows ascending).
__global__ void __launch_bounds__(256) k_bwd_step(const double* __restrict__ A, int N, int k0, double* __restrict__ y, double* __restrict__ x) {
  __shared__ double L[kNB][kNB + 1];
  __shared__ double xs[kNB];
  const int t = threadIdx.x;
  for (int e = t; e < kNB * kNB; e += 256) { const int i = e / kNB, j = e % kNB; L[i][j] = A[(size_t)(k0 + i) * N + k0 + j]; }
  __syncthreads();
  if (t < 32) {                                        // one warp, two unknowns per lane, last to first
    const int lane = t;
    double v0 = y[k0 + lane], v1 = y[k0 + 32 + lane];
    const double d0 = 1.0 / L[lane][lane], d1 = 1.0 / L[lane + 32][lane + 32];
#pragma unroll 8
    for (int j = 31; j >= 0; --j) {
      const double xj = __shfl_sync(0xffffffffu, v1 * d1, j);
      if (lane == j) v1 = xj;
      if (lane < j) v1 -= L[j + 32][lane + 32] * xj;
      v0 -= L[j + 32][lane] * xj;
    }
#pragma unroll 8
    for (int j = 31; j >= 0; --j) {
      const double xj = __shfl_sync(0xffffffffu, v0 * d0, j);
      if (lane == j) v0 = xj;
      if (lane < j) v0 -= L[j][lane] * xj;
    }
    xs[lane] = v0; xs[lane + 32] = v1;
  }
  __syncthreads();
  if (blockIdx.x == 0 && t < kNB) x[k0 + t] = xs[t];
  const int c = blockIdx.x * 256 + t;
  if (c < k0) {
    double s = 0.0;
    for (int i = 0; i < kNB; ++i) s += A[(size_t)(k0 + i) * N + c] * xs[i];
    y[c] -= s;
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

// ---- block-sparse form of the same step: preconditioned conjugate gradients (SetOptionsLidar's policy above 2000 frames is an iterative solver,
//      util/Optimization.cpp:652-658; here also selectable for any size, converged to rounding) ------------------------------------------------------
// The reduced pose-graph matrix is kept as 6x6 blocks (BSR, full symmetric storage, rows and columns = free pose blocks, blocks of a row sorted by
// column): one block per distinct (row block, column block) pair, the sum of its edge contributions in ascending edge order - exactly the blocks
// k_assemble_H scatters into the dense matrix, without the dense matrix (37 MB instead of 730 MB for Floor, and no O(n^3) factorisation).
// System of a trust-region step: (S H S + D) y = -S g with S = diag(sc) (Jacobi scaling), D = clamp(diag(S H S), 1e-6, 1e32) / radius.
// Preconditioner: block Jacobi (the inverse of every damped 6x6 diagonal block).  Every sum runs in a fixed order (per-CTA partials added in CTA
// order by every thread) => bit-reproducible, and identical on every rank of a sharded run.
constexpr int kPcgThreads = 192;            // 32 block rows per CTA: a block row never straddles two CTAs

__global__ void __launch_bounds__(64) k_assemble_bsr(const HContrib* __restrict__ con, const int* __restrict__ seg, const double* __restrict__ esys, double* __restrict__ val) {
  const int t = threadIdx.x;
  if (t >= 36) return;
  const int i = t / 6, j = t % 6;
  const int s0 = seg[blockIdx.x], s1 = seg[blockIdx.x + 1];
  double acc = 0.0;
  for (int s = s0; s < s1; ++s) {
    const HContrib c = con[s];
    const int a = ((c.kind >> 1) & 1) * 6 + i, b = (c.kind & 1) * 6 + j;
    acc += esys[(size_t)c.edge * 92 + (a <= b ? upper12(a, b) : upper12(b, a))];
  }
  val[(size_t)blockIdx.x * 36 + t] = acc;
}

__global__ void k_bsr_jacobi_scale(const double* __restrict__ val, const int* __restrict__ diag, int n, double* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sc[i] = 1.0 / (1.0 + sqrt(val[(size_t)diag[i / 6] * 36 + (i % 6) * 7]));
}

// per LM iteration: rhs = -g sc, damping D, inverse of the damped scaled diagonal blocks (one thread per block row; Gauss-Jordan on the SPD 6x6 block)
__global__ void k_bsr_prepare_step(const double* __restrict__ val, const int* __restrict__ diag, const double* __restrict__ g, const double* __restrict__ sc, int nfb, double radius,
                                   double* __restrict__ rhs, double* __restrict__ dmp, double* __restrict__ minv, int* __restrict__ fail) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nfb) return;
  double M[6][6], I[6][6];
  const double* v = val + (size_t)diag[b] * 36;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { M[i][j] = v[i * 6 + j] * sc[6 * b + i] * sc[6 * b + j]; I[i][j] = i == j ? 1.0 : 0.0; }
  for (int i = 0; i < 6; ++i) {
    const double d = fmin(fmax(M[i][i], 1e-6), 1e32) / radius;
    dmp[6 * b + i] = d; M[i][i] += d;
    rhs[6 * b + i] = -g[6 * b + i] * sc[6 * b + i];
  }
  for (int k = 0; k < 6; ++k) {                                   // SPD: no pivoting needed
    const double piv = M[k][k];
    if (!(piv > 0.0)) { *fail = 1; return; }
    const double ip = 1.0 / piv;
    for (int j = 0; j < 6; ++j) { M[k][j] *= ip; I[k][j] *= ip; }
    for (int i = 0; i < 6; ++i) {
      if (i == k) continue;
      const double f = M[i][k];
      for (int j = 0; j < 6; ++j) { M[i][j] -= f * M[k][j]; I[i][j] -= f * I[k][j]; }
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) minv[(size_t)b * 36 + i * 6 + j] = I[i][j];
}

// Sum over the CTA with a fixed-shape tree (deterministic: the shape depends on blockDim only); result in every thread.  sm: >= 512 doubles.
__device__ __forceinline__ double cta_sum_tree(double v, double* sm) {
  const int t = threadIdx.x, bd = blockDim.x;
  __syncthreads();
  sm[t] = v;
  __syncthreads();
#pragma unroll
  for (int s = 256; s > 0; s >>= 1) {
    if (t < s && t + s < bd) sm[t] += sm[t + s];
    __syncthreads();
  }
  return sm[0];
}
// Grid-wide total of per-CTA partial sums without a second launch and without every thread re-reading all partials: every CTA stores its partial, the CTA
// that arrives LAST (atomic counter) adds all of them with the same fixed-shape tree - the result does not depend on which CTA is last - and publishes it.
// Up to two totals at once.  Returns nothing: consumers are the NEXT kernel on the stream.
__device__ __forceinline__ void grid_total(double p0, double p1, double* part0, double* part1, int n_cta, unsigned* counter, double* out0, double* out1, double* sm) {
  __shared__ int s_last;
  const int t = threadIdx.x, bd = blockDim.x;
  if (t == 0) {
    part0[blockIdx.x] = p0;
    if (part1) part1[blockIdx.x] = p1;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == (unsigned)(n_cta - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double a = 0.0, b = 0.0;
  for (int c = t; c < n_cta; c += bd) { a += reinterpret_cast<volatile double*>(part0)[c]; if (part1) b += reinterpret_cast<volatile double*>(part1)[c]; }
  a = cta_sum_tree(a, sm);
  if (part1) b = cta_sum_tree(b, sm);
  if (t == 0) { *out0 = a; if (part1) *out1 = b; *counter = 0u; }
}

// Device scalars of the CG loop (doubles): [0], [1] = r.z after the latest two updates (update number k writes slot k & 1; the initial one is number 0),
// [2] = r.r after the latest update, [3] = p.q of the latest product.
// One CG iteration = two launches.
// (1) k_pcg_spmv (iteration it = 0, 1, ...): direction + matrix-vector product.  beta = rz[it & 1] / rz[(it + 1) & 1] (it == 0: beta = 0; plain != 0: beta = 0,
//     no scalars touched: q = S H S z, used for the model terms); the new direction p = z + beta p_old is formed on the fly for the gathered entries and stored
//     for the CTA's own rows (p ping-pongs between two buffers, so no launch separates "update p" from "use all of p"); q = (S H S + D) p.
//     One warp per block row (see the kernel).  [3] = p.q.
// (2) k_pcg_update: alpha = rz[it & 1] / [3]; x += alpha p; r -= alpha q; z = Minv r; rz[(it + 1) & 1] = r.z, [2] = r.r.
constexpr int kSpmvRows = 4;                // one warp per block row, 4 rows per CTA
__global__ void __launch_bounds__(kSpmvRows * 32) k_pcg_spmv(int plain, int it, int nfb, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                             const double* __restrict__ val, const double* __restrict__ sc, const double* __restrict__ dmp, double* scal,
                                                             const double* __restrict__ z, const double* __restrict__ p_old, double* __restrict__ p_new, double* __restrict__ q,
                                                             double* part_pq, unsigned* counter) {
  __shared__ double sm[512];
  __shared__ double swarp[kSpmvRows];
  double beta = 0.0;
  if (!plain && it > 0) {
    const double rz = scal[it & 1], rz_old = scal[(it + 1) & 1];
    beta = rz_old > 0.0 ? rz / rz_old : 0.0;
  }
  const bool mix = !plain && it > 0;                             // it == 0: p_old is uninitialised memory (0 * NaN)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kSpmvRows + w;
  // lane l multiplies the blocks d0 + l, d0 + l + 32, ... of its row (a whole 6x6 block per lane: 18 independent 16-byte loads + 6 gathered vector entries, so a
  // row of ~40 blocks is two short dependent steps instead of a chain of 40), then the 6 outputs are summed over the lanes with a fixed-shape butterfly
  double out[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (b < nfb) {
    const int d1 = rowptr[b + 1];
    for (int d = rowptr[b] + lane; d < d1; d += 32) {
      const int c = 6 * col[d];
      double pv[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) pv[j] = sc[c + j] * (mix ? z[c + j] + beta * p_old[c + j] : z[c + j]);
      const double2* v = reinterpret_cast<const double2*>(val + (size_t)d * 36);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double2 v0 = v[3 * i], v1 = v[3 * i + 1], v2 = v[3 * i + 2];
        out[i] += ((((v0.x * pv[0] + v0.y * pv[1]) + v1.x * pv[2]) + v1.y * pv[3]) + v2.x * pv[4]) + v2.y * pv[5];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 6; ++i) out[i] += __shfl_xor_sync(0xffffffffu, out[i], o);
  }
  double pq = 0.0;
  if (b < nfb && lane < 6) {
    const int row = 6 * b + lane;
    double s = out[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) s = lane == i ? out[i] : s;
    s *= sc[row];
    const double pv = mix ? z[row] + beta * p_old[row] : z[row];
    if (!plain) { p_new[row] = pv; s += dmp[row] * pv; }
    q[row] = s;
    pq = pv * s;
  }
  if (plain) return;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) pq += __shfl_xor_sync(0xffffffffu, pq, o);      // lanes 0..5 hold values, 6 and 7 zeros: lane 0 gets the sum of the row
  if (lane == 0) swarp[w] = pq;
  __syncthreads();
  const double cta = ((swarp[0] + swarp[1]) + swarp[2]) + swarp[3];
  grid_total(cta, 0.0, part_pq, nullptr, (int)gridDim.x, counter, scal + 3, nullptr, sm);
}

__global__ void __launch_bounds__(kPcgThreads) k_pcg_update(int first, int it, int n, const double* __restrict__ rhs, const double* __restrict__ minv, double* scal,
                                                            double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, const double* __restrict__ p,
                                                            const double* __restrict__ q, double* part_rz, double* part_rr, unsigned* counter) {
  __shared__ double sm[512];
  __shared__ double sr[kPcgThreads];
  const int i = blockIdx.x * kPcgThreads + threadIdx.x;
  double alpha = 0.0;
  if (!first) {
    const double pq = scal[3], rz = scal[it & 1];
    alpha = (pq > 0.0 && rz > 0.0) ? rz / pq : 0.0;             // breakdown / converged to zero: freeze
  }
  double ri = 0.0;
  if (i < n) {
    if (first) { ri = rhs[i]; x[i] = 0.0; }
    else { ri = r[i] - alpha * q[i]; x[i] += alpha * p[i]; }
    r[i] = ri;
  }
  sr[threadIdx.x] = ri;
  __syncthreads();
  double zi = 0.0;
  if (i < n) {
    const int b = i / 6, k = i - 6 * b, t0 = threadIdx.x - k;
    const double* m = minv + (size_t)b * 36 + k * 6;
#pragma unroll
    for (int j = 0; j < 6; ++j) zi += m[j] * sr[t0 + j];
    z[i] = zi;
  }
  const double s1 = cta_sum_tree(ri * zi, sm);
  const double s2 = cta_sum_tree(ri * ri, sm);
  grid_total(s1, s2, part_rz, part_rr, (int)gridDim.x, counter, scal + (first ? 0 : ((it + 1) & 1)), scal + 2, sm);
}

// model-cost terms from the block-sparse matrix: term[i] = y_i (g_i sc_i + 0.5 (S H S y)_i), hy = S H S y from k_bsr_spmv
__global__ void k_bsr_model_terms(const double* __restrict__ g, const double* __restrict__ sc, const double* __restrict__ y, const double* __restrict__ hy, int n, double* __restrict__ term) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) term[i] = y[i] * (g[i] * sc[i] + 0.5 * hy[i]);
}

}  // namespace pvb
