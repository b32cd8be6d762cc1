// panovlm_b200 — camera-camera reprojection residuals of the joint problem and their bundle adjustment.
//
// Replaces, for the ANGLE_RESIDUAL_1 form the joint optimisation uses (joint_optimization/CameraLidarOptimizer.cpp:431-432, :734):
//   PanoramaReprojResidual_1Angle   base/CostFunction.h:218-247   r = w * acos(s . P / |P|),  P = R(aa_cw) X + t_cw
//   AddCameraResidual               util/Optimization.cpp:172-222 one block per (track, observation), HuberLoss(4 deg), parameter blocks
//                                                                 (aa_cw[3], t_cw[3], point_3d[3])
//   SfMGlobalBA + SetOptionsSfM     util/Optimization.cpp:10-82, 611-636 (DENSE_SCHUR / SPARSE_SCHUR: the points are eliminated first)
//
// Layout in HBM: observations sorted by point (a track's observations are contiguous, as in the reference's loop), 80-byte work rows
// {r, J[9]} per observation; per point the 3x3 block C_p (6) and gradient (3); per observation the 6x3 coupling block E_o; per camera the
// 6x6 block B_c (21) and gradient (6).  Every sum runs in a fixed order (no atomics), so results are bit-reproducible run to run.
// The trust-region step eliminates the points on the device (Schur complement), factors the reduced camera system with the blocked
// FP64 Cholesky of the LiDAR LM loop (pvb_solver.cuh) and back-substitutes the points; only vectors cross PCIe.
#include <algorithm>
#include <cmath>
#include <numeric>
#include "pvb_ctx.hpp"
#include "pvb_host.hpp"

using namespace pvb;

namespace {

constexpr int kPad = 64;        // the Cholesky works on 64-wide panels (kNB in pvb_solver.cuh)
constexpr int kW = 11;          // work row per observation: r, J[9] (d aa | d t | d X), cost
#ifndef PVB_K5_MINB
#define PVB_K5_MINB 7         // resident blocks per SM the reprojection kernel is compiled for
#endif

struct PairEntry { int o1, o2; };             // contribution T_o1 * E'_o2^T to one 6x6 block of the reduced camera system
struct PairDest { int c1, c2, begin, end; };  // block (c1 <= c2) = entries [begin, end)

struct BAState {
  long n_obs = 0, n_pts = 0; int n_cam = 0; double weight = 1.0, huber = 0.0;
  std::vector<int> cam, pt, orig;             // sorted by point
  std::vector<int> pt_off, cam_off, cam_obs;
  int n_dest = 0;
  DevBuf d_cam, d_pt, d_orig, d_bearing, d_pt_off, d_cam_off, d_cam_obs, d_prep, d_X, d_W, d_r, d_J;
  DevBuf d_C, d_gp, d_E, d_B, d_gc, d_ccost;
  DevBuf d_entries, d_dest, d_fidx, d_scc, d_scp, d_ptconst, d_Cinv, d_gsp, d_T, d_Es, d_yp;
  DevBuf d_camblk, d_pinned;                    // joint solve: first unknown of every camera's block, pinned (constant) unknowns of free blocks
  PinBuf h_r, h_J, h_blocks, h_stage;
  bool has_rows = false, has_sys = false, pairs_built = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool ev_valid = false;
  void release() {
    DevBuf* bs[] = {&d_cam, &d_pt, &d_orig, &d_bearing, &d_pt_off, &d_cam_off, &d_cam_obs, &d_prep, &d_X, &d_W, &d_r, &d_J, &d_C, &d_gp, &d_E, &d_B, &d_gc, &d_ccost,
                    &d_entries, &d_dest, &d_fidx, &d_scc, &d_scp, &d_ptconst, &d_Cinv, &d_gsp, &d_T, &d_Es, &d_yp, &d_camblk, &d_pinned};
    for (DevBuf* b : bs) b->release();
    h_r.release(); h_J.release(); h_blocks.release(); h_stage.release();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

void ba_free(void* p) { BAState* s = static_cast<BAState*>(p); s->release(); delete s; }

BAState* ba_get(pvb_ctx* ctx, bool create) {
  if (!ctx->ba_state && create) { ctx->ba_state = new BAState(); ctx->ba_free = ba_free; }
  return static_cast<BAState*>(ctx->ba_state);
}

// ---- K5: residual + 1x9 Jacobian of every observation -------------------------------------------------------------------------------
struct ReprojArgs {
  const int* cam; const int* pt; const int* orig; const double* bearing; const PosePrep* prep; const double* X;
  double weight, huber; long n;
  double* W; double* r_rows; double* J_rows;     // W: sorted order (may be null); rows: caller's order (may be null)
};

__global__ void __launch_bounds__(128, PVB_K5_MINB) k_reproj_rows(ReprojArgs a) {
  __shared__ double stage[128 * kW];               // the tile's work rows {r, J[9], cost}: written out with coalesced stores
  const int tid = threadIdx.x;
  const long base = (long)blockIdx.x * blockDim.x;
  const long o = base + tid;
  const int count = (int)min((long)blockDim.x, a.n - base);
  const bool act = tid < count;
  double r = 0.0, cost = 0.0, J[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) J[j] = 0.0;
  if (act) {
    const PosePrep* pp = a.prep + a.cam[o];
    double R[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = __ldg(pp->R + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = __ldg(pp->t + k);
    const double* Xp = a.X + 3 * (long)a.pt[o];
    const double X[3] = {__ldg(Xp), __ldg(Xp + 1), __ldg(Xp + 2)};
    const double s[3] = {a.bearing[3 * o], a.bearing[3 * o + 1], a.bearing[3 * o + 2]};
    double v[3], P[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k] = R[k * 3] * X[0] + R[k * 3 + 1] * X[1] + R[k * 3 + 2] * X[2]; P[k] = v[k] + t[k]; }
    const double n2 = P[0] * P[0] + P[1] * P[1] + P[2] * P[2];
    const double n = sqrt(n2);
    const double c = (P[0] * s[0] + P[1] * s[1] + P[2] * s[2]) / n;
    r = a.weight * acos(c);
    // d r / d P = -w / sqrt(1 - c^2) * (s - c P / |P|) / |P|   (unguarded at c = 1 like the autodiff functor)
    const double k = -a.weight / (sqrt(1.0 - c * c) * n);
    const double g[3] = {k * (s[0] - c * P[0] / n), k * (s[1] - c * P[1] / n), k * (s[2] - c * P[2] / n)};
    const double w[3] = {v[1] * g[2] - v[2] * g[1], v[2] * g[0] - v[0] * g[2], v[0] * g[1] - v[1] * g[0]};     // (R X) x g
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[j] = w[0] * __ldg(pp->Jl + j) + w[1] * __ldg(pp->Jl + 3 + j) + w[2] * __ldg(pp->Jl + 6 + j);             // d aa: through the left Jacobian of SO(3)
      J[3 + j] = g[j];
      J[6 + j] = g[0] * R[j] + g[1] * R[3 + j] + g[2] * R[6 + j];                                               // R^T g
    }
    cost = huber_correct(a.huber, r, J, 9);
  }
  double* mine = stage + tid * kW;
  mine[0] = r;
#pragma unroll
  for (int j = 0; j < 9; ++j) mine[1 + j] = J[j];
  mine[10] = cost;
  // rows in the caller's order: when the tile maps to consecutive caller rows (point-major input, the reference's loop order) the 72-byte rows
  // leave as one contiguous coalesced block, otherwise each thread scatters its own row
  const long d0 = a.r_rows ? (long)a.orig[base] : 0;
  const int contiguous = __syncthreads_and(!a.r_rows || !act || (long)a.orig[o] == d0 + tid);
  if (a.W) {
    double* dst = a.W + (size_t)base * kW;
    for (int e = tid; e < count * kW; e += 128) dst[e] = stage[e];
  }
  if (a.r_rows) {
    if (contiguous) {
      if (act) a.r_rows[d0 + tid] = r;
      double* dst = a.J_rows + (size_t)d0 * 9;
      for (int e = tid; e < count * 9; e += 128) { const int row = e / 9; dst[e] = stage[row * kW + 1 + (e - row * 9)]; }
    } else if (act) {
      const long d = a.orig[o];
      a.r_rows[d] = r;
#pragma unroll
      for (int j = 0; j < 9; ++j) a.J_rows[d * 9 + j] = J[j];
    }
  }
}

// ---- K6: per point C_p = sum Jp^T Jp, g_p = sum Jp^T r and per observation E_o = Jc^T Jp -----------------------------------------------
__global__ void __launch_bounds__(128) k_reproj_point_blocks(const double* __restrict__ W, const int* __restrict__ pt_off, long n_pts, double* __restrict__ C,
                                                             double* __restrict__ gp, double* __restrict__ E) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  double c[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
  for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) {
    const double* w = W + (size_t)o * kW;
    const double r = w[0], jx = w[7], jy = w[8], jz = w[9];
    c[0] += jx * jx; c[1] += jx * jy; c[2] += jx * jz; c[3] += jy * jy; c[4] += jy * jz; c[5] += jz * jz;
    g[0] += jx * r; g[1] += jy * r; g[2] += jz * r;
    double* e = E + (size_t)o * 18;
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double jc = w[1 + i]; e[i * 3] = jc * jx; e[i * 3 + 1] = jc * jy; e[i * 3 + 2] = jc * jz; }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) C[p * 6 + k] = c[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) gp[p * 3 + k] = g[k];
}

// ---- K7: per camera B_c = sum Jc^T Jc (upper, 21), g_c = sum Jc^T r, cost: strided partial sums, then a fixed-order reduction -------------
__global__ void __launch_bounds__(128) k_reproj_cam_blocks(const double* __restrict__ W, const int* __restrict__ cam_off, const int* __restrict__ cam_obs,
                                                           double* __restrict__ B, double* __restrict__ gc, double* __restrict__ ccost) {
  __shared__ double part[28][129];
  const int c = blockIdx.x, tid = threadIdx.x;
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  for (int i = cam_off[c] + tid; i < cam_off[c + 1]; i += 128) {
    const double* w = W + (size_t)cam_obs[i] * kW;
    double j[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) j[k] = w[1 + k];
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) acc[q++] += j[a] * j[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += j[a] * w[0];
    acc[27] += w[10];
  }
#pragma unroll
  for (int k = 0; k < 28; ++k) part[k][tid] = acc[k];
  __syncthreads();
  if (tid < 28) {
    double s = 0.0;
    for (int i = 0; i < 128; ++i) s += part[tid][i];
    if (tid < 21) B[c * 21 + tid] = s; else if (tid < 27) gc[c * 6 + tid - 21] = s; else ccost[c] = s;
  }
}

// ---- Schur complement of one trust-region step ----------------------------------------------------------------------------------------
// scaled, damped point blocks C' = S C S + clamp(diag)/radius, their inverses, the scaled coupling blocks E' and T_o = E'_o C'^-1
__global__ void __launch_bounds__(128) k_schur_points(const double* __restrict__ C, const double* __restrict__ gp, const double* __restrict__ E, const int* __restrict__ pt_off,
                                                      const int* __restrict__ cam, const double* __restrict__ scc, const double* __restrict__ scp,
                                                      const unsigned char* __restrict__ pt_const, long n_pts, double radius, double* __restrict__ Cinv, double* __restrict__ gsp,
                                                      double* __restrict__ T, double* __restrict__ Es) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  const bool fixed = pt_const[p] != 0;
  const double s0 = scp[p * 3], s1 = scp[p * 3 + 1], s2 = scp[p * 3 + 2];
  double a00 = C[p * 6] * s0 * s0, a01 = C[p * 6 + 1] * s0 * s1, a02 = C[p * 6 + 2] * s0 * s2, a11 = C[p * 6 + 3] * s1 * s1, a12 = C[p * 6 + 4] * s1 * s2, a22 = C[p * 6 + 5] * s2 * s2;
  a00 += fmin(fmax(a00, 1e-6), 1e32) / radius; a11 += fmin(fmax(a11, 1e-6), 1e32) / radius; a22 += fmin(fmax(a22, 1e-6), 1e32) / radius;
  // symmetric inverse through the adjugate (C' is positive definite thanks to the damping)
  const double m00 = a11 * a22 - a12 * a12, m01 = a02 * a12 - a01 * a22, m02 = a01 * a12 - a02 * a11;
  const double det = a00 * m00 + a01 * m01 + a02 * m02;
  const double id = fixed ? 0.0 : 1.0 / det;
  const double i00 = m00 * id, i01 = m01 * id, i02 = m02 * id, i11 = (a00 * a22 - a02 * a02) * id, i12 = (a01 * a02 - a00 * a12) * id, i22 = (a00 * a11 - a01 * a01) * id;
  double* ci = Cinv + p * 6;
  ci[0] = i00; ci[1] = i01; ci[2] = i02; ci[3] = i11; ci[4] = i12; ci[5] = i22;
  gsp[p * 3] = gp[p * 3] * s0; gsp[p * 3 + 1] = gp[p * 3 + 1] * s1; gsp[p * 3 + 2] = gp[p * 3 + 2] * s2;
  for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) {
    const double* e = E + (size_t)o * 18;
    const double* sc = scc + 6 * (size_t)cam[o];
    double* es = Es + (size_t)o * 18;
    double* t = T + (size_t)o * 18;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double e0 = e[i * 3] * sc[i] * s0, e1 = e[i * 3 + 1] * sc[i] * s1, e2 = e[i * 3 + 2] * sc[i] * s2;
      es[i * 3] = e0; es[i * 3 + 1] = e1; es[i * 3 + 2] = e2;
      t[i * 3] = e0 * i00 + e1 * i01 + e2 * i02;
      t[i * 3 + 1] = e0 * i01 + e1 * i11 + e2 * i12;
      t[i * 3 + 2] = e0 * i02 + e1 * i12 + e2 * i22;
    }
  }
}

// one 6x6 block (c1 <= c2) of the reduced camera system, added to the lower triangle of A:
// [own_diag && c1 == c2] (S B S + damping) - sum_entries T_o1 E'_o2^T, entries in a fixed order
__global__ void __launch_bounds__(64) k_schur_assemble(const PairDest* __restrict__ dest, const PairEntry* __restrict__ entries, const double* __restrict__ T,
                                                       const double* __restrict__ Es, const double* __restrict__ B, const double* __restrict__ scc, const int* __restrict__ fidx,
                                                       double radius, int N, int own_diag, double* __restrict__ A) {
  const PairDest d = dest[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid >= 36) return;
  const int i = tid / 6, j = tid - i * 6;
  const int fi = fidx[6 * d.c1 + i], fj = fidx[6 * d.c2 + j];
  if (fi < 0 || fj < 0) return;
  if (d.c1 == d.c2 && fi < fj) return;                      // the diagonal block is symmetric: lower half only
  double acc = 0.0;
  for (int e = d.begin; e < d.end; ++e) {
    const PairEntry pe = entries[e];
    const double* t = T + (size_t)pe.o1 * 18 + i * 3;
    const double* es = Es + (size_t)pe.o2 * 18 + j * 3;
    acc += t[0] * es[0] + t[1] * es[1] + t[2] * es[2];
  }
  double v = -acc;
  if (own_diag && d.c1 == d.c2) {
    const int a = i < j ? i : j, b = i < j ? j : i;
    const double bs = B[(size_t)d.c1 * 21 + (a * 6 - a * (a - 1) / 2 + (b - a))] * scc[6 * (size_t)d.c1 + i] * scc[6 * (size_t)d.c1 + j];
    v += bs;
    if (i == j) v += fmin(fmax(bs, 1e-6), 1e32) / radius;
  }
  const int r = fi > fj ? fi : fj, c = fi > fj ? fj : fi;
  A[(size_t)r * N + c] += v;
}

// right-hand side of the reduced system, added to rhs: [own_gradient] -S g_c + sum_{o of c} T_o g'_p(o), fixed order per camera
__global__ void __launch_bounds__(128) k_schur_rhs(const int* __restrict__ cam_off, const int* __restrict__ cam_obs, const int* __restrict__ pt, const double* __restrict__ T,
                                                   const double* __restrict__ gsp, const double* __restrict__ gc, const double* __restrict__ scc, const int* __restrict__ fidx,
                                                   int own_gradient, double* __restrict__ rhs) {
  __shared__ double part[6][129];
  const int c = blockIdx.x, tid = threadIdx.x;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = cam_off[c] + tid; i < cam_off[c + 1]; i += 128) {
    const int o = cam_obs[i];
    const double* t = T + (size_t)o * 18;
    const double* g = gsp + 3 * (size_t)pt[o];
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] += t[k * 3] * g[0] + t[k * 3 + 1] * g[1] + t[k * 3 + 2] * g[2];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) part[k][tid] = acc[k];
  __syncthreads();
  if (tid < 6) {
    const int f = fidx[6 * c + tid];
    if (f < 0) return;
    double s = 0.0;
    for (int i = 0; i < 128; ++i) s += part[tid][i];
    if (own_gradient) s -= gc[6 * (size_t)c + tid] * scc[6 * (size_t)c + tid];
    rhs[f] += s;
  }
}

// joint solve: the cameras' own blocks J_c^T J_c and gradients join the dense pose-graph system (blk[c] = first unknown of camera c's block or -1)
__global__ void __launch_bounds__(64) k_add_cam_blocks(const double* __restrict__ B, const double* __restrict__ gc, const int* __restrict__ blk, int N, double* __restrict__ H,
                                                       double* __restrict__ g) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const int f = blk[c];
  if (f < 0 || tid >= 36) return;
  const int i = tid / 6, j = tid - i * 6;
  const int a = i < j ? i : j, b = i < j ? j : i;
  H[(size_t)(f + i) * N + f + j] += B[(size_t)c * 21 + (a * 6 - a * (a - 1) / 2 + (b - a))];
  if (j == 0) g[f + i] += gc[6 * (size_t)c + i];
}

// joint solve: a constant parameter inside a free block keeps its row, turned into the identity (step 0)
__global__ void __launch_bounds__(128) k_pin_params(const int* __restrict__ pinned, int N, double* __restrict__ A, double* __restrict__ rhs) {
  const int f = pinned[blockIdx.x];
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    if (j < f) A[(size_t)f * N + j] = 0.0;
    else if (j > f) A[(size_t)j * N + f] = 0.0;
    else { A[(size_t)f * N + f] = 1.0; rhs[f] = 0.0; }
  }
}

__global__ void k_pad_identity(double* __restrict__ A, double* __restrict__ rhs, int n, int N) {
  const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { A[(size_t)i * N + i] = 1.0; rhs[i] = 0.0; }
}

// y_p = -C'^-1 g'_p - sum_{o of p} T_o^T y_c(o)
__global__ void __launch_bounds__(128) k_schur_backsub(const double* __restrict__ Cinv, const double* __restrict__ gsp, const double* __restrict__ T, const int* __restrict__ pt_off,
                                                       const int* __restrict__ cam, const int* __restrict__ fidx, const double* __restrict__ y, long n_pts, double* __restrict__ yp) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  const double* ci = Cinv + p * 6;
  const double g0 = gsp[p * 3], g1 = gsp[p * 3 + 1], g2 = gsp[p * 3 + 2];
  double y0 = -(ci[0] * g0 + ci[1] * g1 + ci[2] * g2), y1 = -(ci[1] * g0 + ci[3] * g1 + ci[4] * g2), y2 = -(ci[2] * g0 + ci[4] * g1 + ci[5] * g2);
  for (int o = pt_off[p]; o < pt_off[p + 1]; ++o) {
    const double* t = T + (size_t)o * 18;
    const int* f = fidx + 6 * (size_t)cam[o];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (f[i] < 0) continue;
      const double yc = y[f[i]];
      y0 -= t[i * 3] * yc; y1 -= t[i * 3 + 1] * yc; y2 -= t[i * 3 + 2] * yc;
    }
  }
  yp[p * 3] = y0; yp[p * 3 + 1] = y1; yp[p * 3 + 2] = y2;
}

int upload_state(pvb_ctx* ctx, BAState* S, const double* cams, const double* points) {
  const size_t pb = (size_t)S->n_cam * sizeof(PosePrep), xb = (size_t)S->n_pts * 24;
  CK(S->h_stage.ensure(pb + xb));
  CK(S->d_prep.ensure(pb)); CK(S->d_X.ensure(std::max<size_t>(xb, 8)));
  CK(cudaStreamSynchronize(ctx->stream));
  PosePrep* hp = S->h_stage.as<PosePrep>();
  for (int c = 0; c < S->n_cam; ++c) prepare_pose(cams + 6 * c, hp[c]);
  memcpy(reinterpret_cast<char*>(S->h_stage.p) + pb, points, xb);
  CK(cudaMemcpyAsync(S->d_prep.p, hp, pb, cudaMemcpyHostToDevice, ctx->stream));
  if (xb) CK(cudaMemcpyAsync(S->d_X.p, reinterpret_cast<char*>(S->h_stage.p) + pb, xb, cudaMemcpyHostToDevice, ctx->stream));
  return PVB_OK;
}

// contribution lists of the Schur complement, sorted by destination block, then by (point, observation pair)
int build_pairs(pvb_ctx* ctx, BAState* S) {
  if (S->pairs_built) return PVB_OK;
  struct Item { int c1, c2, o1, o2; };
  std::vector<Item> items;
  for (long p = 0; p < S->n_pts; ++p)
    for (int a = S->pt_off[p]; a < S->pt_off[p + 1]; ++a)
      for (int b = S->pt_off[p]; b < S->pt_off[p + 1]; ++b) {
        const int ca = S->cam[a], cb = S->cam[b];
        if (ca < cb || (ca == cb)) items.push_back(Item{ca, cb, a, b});
      }
  for (int c = 0; c < S->n_cam; ++c) items.push_back(Item{c, c, -1, -1});            // every camera owns its diagonal block even without observations
  std::sort(items.begin(), items.end(), [](const Item& x, const Item& y) {
    if (x.c1 != y.c1) return x.c1 < y.c1;
    if (x.c2 != y.c2) return x.c2 < y.c2;
    if (x.o1 != y.o1) return x.o1 < y.o1;
    return x.o2 < y.o2;
  });
  std::vector<PairEntry> entries; entries.reserve(items.size());
  std::vector<PairDest> dest;
  for (size_t i = 0; i < items.size(); ++i) {
    if (i == 0 || items[i].c1 != items[i - 1].c1 || items[i].c2 != items[i - 1].c2) {
      if (!dest.empty()) dest.back().end = (int)entries.size();
      dest.push_back(PairDest{items[i].c1, items[i].c2, (int)entries.size(), 0});
    }
    if (items[i].o1 >= 0) entries.push_back(PairEntry{items[i].o1, items[i].o2});
  }
  if (!dest.empty()) dest.back().end = (int)entries.size();
  S->n_dest = (int)dest.size();
  CK(S->d_entries.ensure(std::max<size_t>(16, entries.size() * sizeof(PairEntry)))); CK(S->d_dest.ensure(std::max<size_t>(16, dest.size() * sizeof(PairDest))));
  if (!entries.empty()) CK(cudaMemcpyAsync(S->d_entries.p, entries.data(), entries.size() * sizeof(PairEntry), cudaMemcpyHostToDevice, ctx->stream));
  if (!dest.empty()) CK(cudaMemcpyAsync(S->d_dest.p, dest.data(), dest.size() * sizeof(PairDest), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  S->pairs_built = true;
  return PVB_OK;
}

}  // namespace

extern "C" {

int pvb_reproj_set(pvb_ctx* ctx, long n_obs, const int* cam, const int* point, const double* bearing3, double weight, double huber, int n_cams, long n_points) {
  if (!ctx || n_obs < 0 || n_cams <= 0 || n_points < 0 || (n_obs > 0 && (!cam || !point || !bearing3))) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_reproj_set: bad arguments") : PVB_ERR_ARG;
  if (n_obs > 0x7fffffffL / kW) return ctx->fail(PVB_ERR_ARG, "pvb_reproj_set: too many observations");
  for (long i = 0; i < n_obs; ++i)
    if (cam[i] < 0 || cam[i] >= n_cams || point[i] < 0 || point[i] >= n_points) return ctx->fail(PVB_ERR_ARG, "pvb_reproj_set: observation %ld refers to camera %d / point %d", i, cam[i], point[i]);
  CK(cudaSetDevice(ctx->device));
  BAState* S = ba_get(ctx, true);
  S->n_obs = n_obs; S->n_cam = n_cams; S->n_pts = n_points; S->weight = weight; S->huber = huber;
  S->has_rows = S->has_sys = S->pairs_built = false;
  if (!S->ev0) { CK(cudaEventCreate(&S->ev0)); CK(cudaEventCreate(&S->ev1)); }
  // point-major order, the caller's order kept inside a track (stable)
  std::vector<int> order(n_obs);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return point[a] < point[b]; });
  S->cam.resize(n_obs); S->pt.resize(n_obs); S->orig = order;
  std::vector<double> bear((size_t)n_obs * 3);
  S->pt_off.assign(n_points + 1, 0); S->cam_off.assign(n_cams + 1, 0);
  for (long i = 0; i < n_obs; ++i) {
    const int s = order[i];
    S->cam[i] = cam[s]; S->pt[i] = point[s];
    const double* b = bearing3 + 3 * (size_t)s;
    const double nrm = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);      // point_sphere.normalize() (CostFunction.h:225)
    for (int k = 0; k < 3; ++k) bear[3 * (size_t)i + k] = b[k] / nrm;
    S->pt_off[point[s] + 1]++; S->cam_off[cam[s] + 1]++;
  }
  for (long p = 0; p < n_points; ++p) S->pt_off[p + 1] += S->pt_off[p];
  for (int c = 0; c < n_cams; ++c) S->cam_off[c + 1] += S->cam_off[c];
  S->cam_obs.resize(n_obs);
  { std::vector<int> fill(S->cam_off.begin(), S->cam_off.end() - 1); for (long i = 0; i < n_obs; ++i) S->cam_obs[fill[S->cam[i]]++] = (int)i; }
  const size_t n = std::max<long>(n_obs, 1);
  CK(S->d_cam.ensure(n * 4)); CK(S->d_pt.ensure(n * 4)); CK(S->d_orig.ensure(n * 4)); CK(S->d_bearing.ensure(n * 24)); CK(S->d_cam_obs.ensure(n * 4));
  CK(S->d_pt_off.ensure((size_t)(n_points + 1) * 4)); CK(S->d_cam_off.ensure((size_t)(n_cams + 1) * 4));
  CK(S->d_W.ensure(n * kW * 8)); CK(S->d_r.ensure(n * 8)); CK(S->d_J.ensure(n * 72)); CK(S->d_E.ensure(n * 144));
  CK(S->d_C.ensure(std::max<size_t>(8, (size_t)n_points * 48))); CK(S->d_gp.ensure(std::max<size_t>(8, (size_t)n_points * 24)));
  CK(S->d_B.ensure((size_t)n_cams * 168)); CK(S->d_gc.ensure((size_t)n_cams * 48)); CK(S->d_ccost.ensure((size_t)n_cams * 8));
  CK(S->h_r.ensure(n * 8)); CK(S->h_J.ensure(n * 72));
  CK(S->h_blocks.ensure((size_t)n_cams * 224 + (size_t)n_points * 72 + 64));
  if (n_obs) {
    CK(cudaMemcpyAsync(S->d_cam.p, S->cam.data(), n_obs * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S->d_pt.p, S->pt.data(), n_obs * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S->d_orig.p, S->orig.data(), n_obs * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S->d_bearing.p, bear.data(), (size_t)n_obs * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S->d_cam_obs.p, S->cam_obs.data(), n_obs * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaMemcpyAsync(S->d_pt_off.p, S->pt_off.data(), (size_t)(n_points + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(S->d_cam_off.p, S->cam_off.data(), (size_t)(n_cams + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));      // `bear` is a local
  return PVB_OK;
}

int pvb_reproj_evaluate(pvb_ctx* ctx, const double* cams6, const double* points3, int want_rows, int want_system) {
  BAState* S = ctx ? ba_get(ctx, false) : nullptr;
  if (!ctx || !cams6 || (!points3 && S && S->n_pts > 0)) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_reproj_evaluate: bad arguments") : PVB_ERR_ARG;
  if (!S) return ctx->fail(PVB_ERR_STATE, "pvb_reproj_set has not been called");
  if (want_rows == 2 && want_system) return ctx->fail(PVB_ERR_ARG, "pvb_reproj_evaluate: raw rows (want_rows = 2) and the reduced system (always loss-corrected) need two calls");
  CK(cudaSetDevice(ctx->device));
  int rc = upload_state(ctx, S, cams6, points3); if (rc) return rc;
  ReprojArgs a{};
  a.cam = S->d_cam.as<int>(); a.pt = S->d_pt.as<int>(); a.orig = S->d_orig.as<int>(); a.bearing = S->d_bearing.as<double>();
  a.prep = S->d_prep.as<PosePrep>(); a.X = S->d_X.as<double>(); a.weight = S->weight; a.huber = S->huber; a.n = S->n_obs;
  if (want_rows == 2) a.huber = 0.0;              // raw rows: the caller registers the loss with its solver (ceres::HuberLoss)
  a.W = want_system ? S->d_W.as<double>() : nullptr;
  a.r_rows = want_rows ? S->d_r.as<double>() : nullptr; a.J_rows = want_rows ? S->d_J.as<double>() : nullptr;
  if (!want_system && !want_rows) a.W = S->d_W.as<double>();
  CK(cudaEventRecord(S->ev0, ctx->stream));
  if (S->n_obs) { k_reproj_rows<<<(unsigned)((S->n_obs + 127) / 128), 128, 0, ctx->stream>>>(a); CKL(); }
  CK(cudaEventRecord(S->ev1, ctx->stream));
  S->ev_valid = true;
  S->has_rows = want_rows != 0; S->has_sys = false;
  if (want_rows && S->n_obs) {
    CK(cudaMemcpyAsync(S->h_r.p, S->d_r.p, (size_t)S->n_obs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(S->h_J.p, S->d_J.p, (size_t)S->n_obs * 72, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (want_system) {
    if (S->n_pts) { k_reproj_point_blocks<<<(unsigned)((S->n_pts + 127) / 128), 128, 0, ctx->stream>>>(S->d_W.as<double>(), S->d_pt_off.as<int>(), S->n_pts, S->d_C.as<double>(), S->d_gp.as<double>(), S->d_E.as<double>()); CKL(); }
    k_reproj_cam_blocks<<<S->n_cam, 128, 0, ctx->stream>>>(S->d_W.as<double>(), S->d_cam_off.as<int>(), S->d_cam_obs.as<int>(), S->d_B.as<double>(), S->d_gc.as<double>(), S->d_ccost.as<double>());
    CKL();
    // host mirrors of the small blocks: [B 21 nc | gc 6 nc | cost nc | C 6 np | gp 3 np]
    double* h = S->h_blocks.as<double>();
    const size_t nc = S->n_cam, np = S->n_pts;
    CK(cudaMemcpyAsync(h, S->d_B.p, nc * 168, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h + 21 * nc, S->d_gc.p, nc * 48, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h + 27 * nc, S->d_ccost.p, nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (np) {
      CK(cudaMemcpyAsync(h + 28 * nc, S->d_C.p, np * 48, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaMemcpyAsync(h + 28 * nc + 6 * np, S->d_gp.p, np * 24, cudaMemcpyDeviceToHost, ctx->stream));
    }
    S->has_sys = true;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

const double* pvb_reproj_residuals(const pvb_ctx* ctx) { const BAState* S = ctx ? static_cast<const BAState*>(ctx->ba_state) : nullptr; return S && S->has_rows ? S->h_r.as<double>() : nullptr; }
const double* pvb_reproj_jacobians(const pvb_ctx* ctx) { const BAState* S = ctx ? static_cast<const BAState*>(ctx->ba_state) : nullptr; return S && S->has_rows ? S->h_J.as<double>() : nullptr; }

int pvb_reproj_cost(const pvb_ctx* ctx, double* cost) {
  const BAState* S = ctx ? static_cast<const BAState*>(ctx->ba_state) : nullptr;
  if (!S || !S->has_sys || !cost) return PVB_ERR_STATE;
  const double* h = S->h_blocks.as<double>() + 27 * (size_t)S->n_cam;
  double c = 0.0;
  for (int i = 0; i < S->n_cam; ++i) c += h[i];
  *cost = c;
  return PVB_OK;
}

int pvb_reproj_blocks(pvb_ctx* ctx, double* cam_H21, double* cam_g6, double* pt_H6, double* pt_g3, double* obs_E18) {
  BAState* S = ctx ? ba_get(ctx, false) : nullptr;
  if (!S || !S->has_sys) return ctx ? ctx->fail(PVB_ERR_STATE, "pvb_reproj_blocks: no evaluated system") : PVB_ERR_ARG;
  const double* h = S->h_blocks.as<double>();
  const size_t nc = S->n_cam, np = S->n_pts;
  if (cam_H21) memcpy(cam_H21, h, nc * 168);
  if (cam_g6) memcpy(cam_g6, h + 21 * nc, nc * 48);
  if (pt_H6 && np) memcpy(pt_H6, h + 28 * nc, np * 48);
  if (pt_g3 && np) memcpy(pt_g3, h + 28 * nc + 6 * np, np * 24);
  if (obs_E18 && S->n_obs) {
    CK(cudaSetDevice(ctx->device));
    std::vector<double> tmp((size_t)S->n_obs * 18);
    CK(cudaMemcpyAsync(tmp.data(), S->d_E.p, tmp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (long i = 0; i < S->n_obs; ++i) memcpy(obs_E18 + 18 * (size_t)S->orig[i], tmp.data() + 18 * (size_t)i, 144);
  }
  return PVB_OK;
}

int pvb_reproj_kernel_time_ms(pvb_ctx* ctx, float* ms) {
  BAState* S = ctx ? ba_get(ctx, false) : nullptr;
  if (!S || !S->ev_valid || !ms) return ctx ? ctx->fail(PVB_ERR_STATE, "no reprojection evaluate to time") : PVB_ERR_ARG;
  CK(cudaEventElapsedTime(ms, S->ev0, S->ev1));
  return PVB_OK;
}

// SfMGlobalBA's ceres::Solve (util/Optimization.cpp:59-61): trust-region LM with Ceres' defaults (the loop of pvb::solve_lm, pvb_host.hpp);
// the linear step eliminates the points first, like the DENSE_SCHUR / SPARSE_SCHUR solvers SetOptionsSfM selects (:611-636).
int pvb_reproj_solve_lm(pvb_ctx* ctx, double* cams6, double* points3, const unsigned char* cam_param_const, const unsigned char* point_const, int max_iterations,
                        double* summary6) {
  BAState* S = ctx ? ba_get(ctx, false) : nullptr;
  if (!ctx || !cams6) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_reproj_solve_lm: bad arguments") : PVB_ERR_ARG;
  if (!S) return ctx->fail(PVB_ERR_STATE, "pvb_reproj_set has not been called");
  if (S->n_pts > 0 && !points3) return ctx->fail(PVB_ERR_ARG, "pvb_reproj_solve_lm: null points");
  CK(cudaSetDevice(ctx->device));
  const int nc = S->n_cam; const long np = S->n_pts;
  int rc = build_pairs(ctx, S); if (rc) return rc;
  // free-parameter map of the reduced camera system
  std::vector<int> fidx(6 * (size_t)nc, -1);
  int n = 0;
  for (int i = 0; i < 6 * nc; ++i) if (!cam_param_const || !cam_param_const[i]) fidx[i] = n++;
  std::vector<unsigned char> ptc(std::max<long>(np, 1), 0);
  long n_free_pts = 0;
  for (long p = 0; p < np; ++p) { ptc[p] = point_const && point_const[p] ? 1 : 0; if (!ptc[p]) ++n_free_pts; }
  const int N = std::max(kPad, ((n + kPad - 1) / kPad) * kPad);
  CK(S->d_fidx.ensure(fidx.size() * 4)); CK(S->d_scc.ensure((size_t)nc * 48)); CK(S->d_scp.ensure(std::max<size_t>(8, (size_t)np * 24))); CK(S->d_ptconst.ensure(ptc.size()));
  CK(S->d_Cinv.ensure(std::max<size_t>(8, (size_t)np * 48))); CK(S->d_gsp.ensure(std::max<size_t>(8, (size_t)np * 24))); CK(S->d_yp.ensure(std::max<size_t>(8, (size_t)np * 24)));
  CK(S->d_T.ensure(std::max<size_t>(8, (size_t)S->n_obs * 144))); CK(S->d_Es.ensure(std::max<size_t>(8, (size_t)S->n_obs * 144)));
  CK(ctx->s_A.ensure((size_t)N * N * 8)); CK(ctx->s_rhs.ensure((size_t)N * 8)); CK(ctx->s_fail.ensure(16));
  CK(cudaMemcpyAsync(S->d_fidx.p, fidx.data(), fidx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(S->d_ptconst.p, ptc.data(), ptc.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  LMOptions opt; opt.max_iterations = max_iterations;
  LMSummary R;
  double cost = 0.0;
  auto evaluate = [&](const double* xc, const double* xp, double* out_cost) -> int {
    const int r = pvb_reproj_evaluate(ctx, xc, xp, 0, 1); if (r) return r;
    return pvb_reproj_cost(ctx, out_cost);
  };
  rc = evaluate(cams6, points3, &cost); if (rc) return rc;
  R.initial_cost = cost;
  const double* hb = S->h_blocks.as<double>();
  auto Bdiag = [&](int c, int k) { return hb[21 * (size_t)c + (k * 6 - k * (k - 1) / 2)]; };
  auto Cdiag = [&](long p, int k) { static const int di[3] = {0, 3, 5}; return hb[28 * (size_t)nc + 6 * (size_t)p + di[k]]; };
  auto gcam = [&](int i) { return hb[21 * (size_t)nc + i]; };
  auto gpt = [&](long i) { return hb[28 * (size_t)nc + 6 * (size_t)np + i]; };
  // Jacobi scaling, fixed from the first Jacobian
  std::vector<double> scc(6 * (size_t)nc, 1.0), scp(3 * (size_t)std::max<long>(np, 1), 1.0);
  for (int c = 0; c < nc; ++c) for (int k = 0; k < 6; ++k) scc[6 * c + k] = 1.0 / (1.0 + std::sqrt(Bdiag(c, k)));
  for (long p = 0; p < np; ++p) for (int k = 0; k < 3; ++k) scp[3 * p + k] = 1.0 / (1.0 + std::sqrt(Cdiag(p, k)));
  CK(cudaMemcpyAsync(S->d_scc.p, scc.data(), (size_t)nc * 48, cudaMemcpyHostToDevice, ctx->stream));
  if (np) CK(cudaMemcpyAsync(S->d_scp.p, scp.data(), (size_t)np * 24, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  auto gmax = [&]() {
    double m = 0;
    for (int i = 0; i < 6 * nc; ++i) if (fidx[i] >= 0) m = std::max(m, std::fabs(gcam(i)));
    for (long p = 0; p < np; ++p) if (!ptc[p]) for (int k = 0; k < 3; ++k) m = std::max(m, std::fabs(gpt(3 * p + k)));
    return m;
  };
  auto finish = [&]() {
    R.final_cost = cost;
    if (summary6) { summary6[0] = R.initial_cost; summary6[1] = R.final_cost; summary6[2] = R.iterations; summary6[3] = R.successful; summary6[4] = R.unsuccessful; summary6[5] = R.termination; }
    return PVB_OK;
  };
  if ((n == 0 && n_free_pts == 0) || gmax() <= opt.gradient_tolerance) { R.termination = 2; return finish(); }
  std::vector<double> y(N), yp(3 * (size_t)std::max<long>(np, 1), 0.0), cand_c(6 * (size_t)nc), cand_p(3 * (size_t)std::max<long>(np, 1));
  double radius = 1e4, decrease = 2.0;
  int invalid = 0;
  for (int it = 1; it <= opt.max_iterations; ++it) {
    R.iterations = it;
    if (np) {
      k_schur_points<<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>(S->d_C.as<double>(), S->d_gp.as<double>(), S->d_E.as<double>(), S->d_pt_off.as<int>(), S->d_cam.as<int>(),
                                                                            S->d_scc.as<double>(), S->d_scp.as<double>(), S->d_ptconst.as<unsigned char>(), np, radius,
                                                                            S->d_Cinv.as<double>(), S->d_gsp.as<double>(), S->d_T.as<double>(), S->d_Es.as<double>());
      CKL();
    }
    CK(cudaMemsetAsync(ctx->s_A.p, 0, (size_t)N * N * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->s_rhs.p, 0, (size_t)N * 8, ctx->stream));
    k_schur_assemble<<<S->n_dest, 64, 0, ctx->stream>>>(S->d_dest.as<PairDest>(), S->d_entries.as<PairEntry>(), S->d_T.as<double>(), S->d_Es.as<double>(), S->d_B.as<double>(),
                                                       S->d_scc.as<double>(), S->d_fidx.as<int>(), radius, N, 1, ctx->s_A.as<double>());
    CKL();
    k_schur_rhs<<<nc, 128, 0, ctx->stream>>>(S->d_cam_off.as<int>(), S->d_cam_obs.as<int>(), S->d_pt.as<int>(), S->d_T.as<double>(), S->d_gsp.as<double>(), S->d_gc.as<double>(),
                                             S->d_scc.as<double>(), S->d_fidx.as<int>(), 1, ctx->s_rhs.as<double>());
    CKL();
    if (N > n) { k_pad_identity<<<(N - n + 63) / 64, 64, 0, ctx->stream>>>(ctx->s_A.as<double>(), ctx->s_rhs.as<double>(), n, N); CKL(); }
    bool ok = false;
    rc = pvb_internal_factor_solve(ctx, N, &ok); if (rc) return rc;
    double model = 0.0;
    if (ok) {
      if (np) {
        k_schur_backsub<<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>(S->d_Cinv.as<double>(), S->d_gsp.as<double>(), S->d_T.as<double>(), S->d_pt_off.as<int>(), S->d_cam.as<int>(),
                                                                               S->d_fidx.as<int>(), ctx->s_rhs.as<double>(), np, S->d_yp.as<double>());
        CKL();
        CK(cudaMemcpyAsync(yp.data(), S->d_yp.p, (size_t)np * 24, cudaMemcpyDeviceToHost, ctx->stream));
      }
      CK(cudaMemcpyAsync(y.data(), ctx->s_rhs.p, (size_t)N * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      // model cost change -y^T (g' + 0.5 H' y) with H' y = -g' - D y at the solution of the damped system: 0.5 (-y.g' + y^T D y), fixed order
      for (int i = 0; i < 6 * nc; ++i) {
        if (fidx[i] < 0) continue;
        const double yi = y[fidx[i]], s = scc[i], hd = Bdiag(i / 6, i % 6) * s * s;
        model += 0.5 * (-yi * gcam(i) * s + std::min(std::max(hd, 1e-6), 1e32) / radius * yi * yi);
      }
      for (long p = 0; p < np; ++p) {
        if (ptc[p]) continue;
        for (int k = 0; k < 3; ++k) {
          const double yi = yp[3 * p + k], s = scp[3 * p + k], hd = Cdiag(p, k) * s * s;
          model += 0.5 * (-yi * gpt(3 * p + k) * s + std::min(std::max(hd, 1e-6), 1e32) / radius * yi * yi);
        }
      }
      ok = model > 0.0;
    }
    if (!ok) { radius *= 0.5; R.unsuccessful++; if (++invalid >= 5 || radius < 1e-32) { R.termination = 4; break; } continue; }
    invalid = 0;
    double sn = 0, xn = 0;
    std::copy(cams6, cams6 + 6 * (size_t)nc, cand_c.begin());
    if (np) std::copy(points3, points3 + 3 * (size_t)np, cand_p.begin());
    for (int i = 0; i < 6 * nc; ++i) { if (fidx[i] < 0) continue; const double d = y[fidx[i]] * scc[i]; cand_c[i] += d; sn += d * d; xn += cams6[i] * cams6[i]; }
    for (long p = 0; p < np; ++p) { if (ptc[p]) continue; for (int k = 0; k < 3; ++k) { const double d = yp[3 * p + k] * scp[3 * p + k]; cand_p[3 * p + k] += d; sn += d * d; xn += points3[3 * p + k] * points3[3 * p + k]; } }
    sn = std::sqrt(sn); xn = std::sqrt(xn);
    double new_cost = 0.0;
    rc = evaluate(cand_c.data(), cand_p.data(), &new_cost); if (rc) return rc;
    bool accepted = false, stop = false;
    if (sn <= opt.parameter_tolerance * (xn + opt.parameter_tolerance)) { R.termination = 3; stop = true; }
    else {
      const double change = cost - new_cost;
      if (std::fabs(change) <= opt.function_tolerance * cost) { R.termination = 1; stop = true; }
      else {
        const double rho = change / model;
        if (rho > 1e-3) {
          accepted = true;
          std::copy(cand_c.begin(), cand_c.end(), cams6);
          if (np) std::copy(cand_p.begin(), cand_p.begin() + 3 * (size_t)np, points3);
          cost = new_cost;
          R.successful++;
          radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
          decrease = 2.0;
          if (gmax() <= opt.gradient_tolerance) { R.termination = 2; stop = true; }
        } else {
          radius /= decrease; decrease *= 2.0; R.unsuccessful++;
          if (radius < 1e-32) { R.termination = 4; stop = true; }
        }
      }
    }
    if (stop) break;
    if (!accepted) {   // the candidate evaluation overwrote the blocks of the current point (device and host mirror): bring them back
      rc = evaluate(cams6, points3, &new_cost); if (rc) return rc;
    }
  }
  return finish();
}

// CameraLidarOptimizer::Optimize's ceres::Solve (joint_optimization/CameraLidarOptimizer.cpp:387-548): camera-camera reprojection blocks
// (AddCameraResidual, :431), LiDAR-LiDAR blocks and camera-LiDAR blocks (pvb_blocks_set: AddLidar*Residual / AddCameraLidarResidual,
// :437-460) in ONE trust-region problem over the pose blocks [cameras | LiDARs] and the structure points.  The pose-graph part is reduced
// per edge and assembled into the dense system by the LiDAR LM's kernels; the cameras' own blocks are added, the points are eliminated
// (Schur complement) and the Cholesky of the LM loop solves for the poses.  Control flow = pvb::solve_lm (Ceres' defaults).
int pvb_joint_solve_lm(pvb_ctx* ctx, double* poses6, double* points3, const unsigned char* pose_param_const, const unsigned char* point_const, int max_iterations,
                       double* summary6) {
  BAState* S = ctx ? ba_get(ctx, false) : nullptr;
  if (!ctx || !poses6) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_joint_solve_lm: bad arguments") : PVB_ERR_ARG;
  if (!S || ctx->nb <= 0) return ctx->fail(PVB_ERR_STATE, "pvb_joint_solve_lm needs pvb_blocks_set and pvb_reproj_set");
  if (S->n_cam > ctx->nb) return ctx->fail(PVB_ERR_ARG, "pvb_joint_solve_lm: %d cameras but only %d pose blocks", S->n_cam, ctx->nb);
  if (S->n_pts > 0 && !points3) return ctx->fail(PVB_ERR_ARG, "pvb_joint_solve_lm: null points");
  CK(cudaSetDevice(ctx->device));
  const int nb = ctx->nb, nc = S->n_cam; const long np = S->n_pts;
  int rc = build_pairs(ctx, S); if (rc) return rc;
  // a block whose six parameters are all constant leaves the system; a partly constant block stays and its constant rows are pinned
  std::vector<unsigned char> blk_const(nb, 0);
  for (int b = 0; b < nb; ++b) { bool all = pose_param_const != nullptr; for (int k = 0; k < 6 && all; ++k) all = pose_param_const[6 * b + k] != 0; blk_const[b] = all ? 1 : 0; }
  rc = pvb_internal_solver_prepare(ctx, blk_const.data()); if (rc) return rc;
  const int n = ctx->s_n, N = std::max(ctx->s_N, kPad);
  std::vector<int> slot(6 * (size_t)nb, -1), fidx(6 * (size_t)nb, -1), cam_blk(nc, -1), pinned;
  { int f = 0;
    for (int b = 0; b < nb; ++b) {
      if (blk_const[b]) continue;
      for (int k = 0; k < 6; ++k) {
        slot[6 * b + k] = f + k;
        if (pose_param_const && pose_param_const[6 * b + k]) pinned.push_back(f + k); else fidx[6 * b + k] = f + k;
      }
      if (b < nc) cam_blk[b] = f;
      f += 6;
    } }
  std::vector<unsigned char> ptc(std::max<long>(np, 1), 0);
  long n_free_pts = 0;
  for (long p = 0; p < np; ++p) { ptc[p] = point_const && point_const[p] ? 1 : 0; if (!ptc[p]) ++n_free_pts; }
  DevBuf& d_camblk = S->d_camblk; DevBuf& d_pinned = S->d_pinned;
  CK(d_camblk.ensure((size_t)nc * 4)); CK(d_pinned.ensure(std::max<size_t>(4, pinned.size() * 4)));
  CK(S->d_fidx.ensure(fidx.size() * 4)); CK(S->d_scc.ensure((size_t)nc * 48)); CK(S->d_scp.ensure(std::max<size_t>(8, (size_t)np * 24))); CK(S->d_ptconst.ensure(ptc.size()));
  CK(S->d_Cinv.ensure(std::max<size_t>(8, (size_t)np * 48))); CK(S->d_gsp.ensure(std::max<size_t>(8, (size_t)np * 24))); CK(S->d_yp.ensure(std::max<size_t>(8, (size_t)np * 24)));
  CK(S->d_T.ensure(std::max<size_t>(8, (size_t)S->n_obs * 144))); CK(S->d_Es.ensure(std::max<size_t>(8, (size_t)S->n_obs * 144)));
  CK(cudaMemcpyAsync(S->d_fidx.p, fidx.data(), fidx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(S->d_ptconst.p, ptc.data(), ptc.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_camblk.p, cam_blk.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!pinned.empty()) CK(cudaMemcpyAsync(d_pinned.p, pinned.data(), pinned.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));

  LMOptions opt; opt.max_iterations = max_iterations;
  LMSummary R;
  std::vector<double> g(std::max(n, 1)), hdiag(std::max(n, 1)), sc(std::max(n, 1), 1.0);
  // evaluation of both parts at (poses, points) + the dense pose system of the free blocks with the cameras' own blocks added
  auto evaluate = [&](const double* xp, const double* xs, double* out_cost) -> int {
    int r = pvb_blocks_evaluate(ctx, xp, 0, 1); if (r) return r;
    r = pvb_reproj_evaluate(ctx, xp, xs, 0, 1); if (r) return r;
    double c1 = 0, c2 = 0;
    r = pvb_blocks_cost(ctx, &c1, nullptr); if (r) return r;
    r = pvb_reproj_cost(ctx, &c2); if (r) return r;
    *out_cost = c1 + c2;
    return PVB_OK;
  };
  auto assemble = [&]() -> int {
    if (n == 0) return PVB_OK;
    int r = pvb_internal_solver_assemble(ctx, g.data()); if (r) return r;
    k_add_cam_blocks<<<nc, 64, 0, ctx->stream>>>(S->d_B.as<double>(), S->d_gc.as<double>(), d_camblk.as<int>(), ctx->s_N, ctx->s_H.as<double>(), ctx->s_g.as<double>());
    CKL();
    CK(cudaMemcpyAsync(g.data(), ctx->s_g.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpy2DAsync(hdiag.data(), 8, ctx->s_H.p, (size_t)(ctx->s_N + 1) * 8, 8, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PVB_OK;
  };
  double cost = 0.0;
  rc = evaluate(poses6, points3, &cost); if (rc) return rc;
  rc = assemble(); if (rc) return rc;
  R.initial_cost = cost;
  const double* hb = S->h_blocks.as<double>();
  auto Cdiag = [&](long p, int k) { static const int di[3] = {0, 3, 5}; return hb[28 * (size_t)nc + 6 * (size_t)p + di[k]]; };
  auto gpt = [&](long i) { return hb[28 * (size_t)nc + 6 * (size_t)np + i]; };
  // Jacobi scaling, fixed from the first Jacobian: poses from the diagonal of the total dense system, points from their own blocks
  std::vector<double> scc(6 * (size_t)nc, 1.0), scp(3 * (size_t)std::max<long>(np, 1), 1.0);
  if (n) {
    rc = pvb_internal_jacobi_scale(ctx); if (rc) return rc;
    CK(cudaMemcpyAsync(sc.data(), ctx->s_sc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  for (int i = 0; i < 6 * nc; ++i) if (slot[i] >= 0) scc[i] = sc[slot[i]];
  for (long p = 0; p < np; ++p) for (int k = 0; k < 3; ++k) scp[3 * p + k] = 1.0 / (1.0 + std::sqrt(Cdiag(p, k)));
  CK(cudaMemcpyAsync(S->d_scc.p, scc.data(), (size_t)nc * 48, cudaMemcpyHostToDevice, ctx->stream));
  if (np) CK(cudaMemcpyAsync(S->d_scp.p, scp.data(), (size_t)np * 24, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  auto gmax = [&]() {
    double m = 0;
    for (int i = 0; i < 6 * nb; ++i) if (fidx[i] >= 0) m = std::max(m, std::fabs(g[fidx[i]]));
    for (long p = 0; p < np; ++p) if (!ptc[p]) for (int k = 0; k < 3; ++k) m = std::max(m, std::fabs(gpt(3 * p + k)));
    return m;
  };
  auto finish = [&]() {
    R.final_cost = cost;
    if (summary6) { summary6[0] = R.initial_cost; summary6[1] = R.final_cost; summary6[2] = R.iterations; summary6[3] = R.successful; summary6[4] = R.unsuccessful; summary6[5] = R.termination; }
    return PVB_OK;
  };
  const bool any_pose = (int)pinned.size() < n;
  if ((!any_pose && n_free_pts == 0) || gmax() <= opt.gradient_tolerance) { R.termination = 2; return finish(); }
  if (n == 0) return ctx->fail(PVB_ERR_ARG, "pvb_joint_solve_lm: every pose block is constant (use pvb_reproj_solve_lm for a structure-only refinement)");
  std::vector<double> y(N), yp(3 * (size_t)std::max<long>(np, 1), 0.0), cand_x(6 * (size_t)nb), cand_p(3 * (size_t)std::max<long>(np, 1));
  double radius = 1e4, decrease = 2.0;
  int invalid = 0;
  for (int it = 1; it <= opt.max_iterations; ++it) {
    R.iterations = it;
    rc = pvb_internal_build_damped(ctx, radius); if (rc) return rc;
    if (np) {
      k_schur_points<<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>(S->d_C.as<double>(), S->d_gp.as<double>(), S->d_E.as<double>(), S->d_pt_off.as<int>(), S->d_cam.as<int>(),
                                                                            S->d_scc.as<double>(), S->d_scp.as<double>(), S->d_ptconst.as<unsigned char>(), np, radius,
                                                                            S->d_Cinv.as<double>(), S->d_gsp.as<double>(), S->d_T.as<double>(), S->d_Es.as<double>());
      CKL();
      k_schur_assemble<<<S->n_dest, 64, 0, ctx->stream>>>(S->d_dest.as<PairDest>(), S->d_entries.as<PairEntry>(), S->d_T.as<double>(), S->d_Es.as<double>(), S->d_B.as<double>(),
                                                         S->d_scc.as<double>(), S->d_fidx.as<int>(), radius, ctx->s_N, 0, ctx->s_A.as<double>());
      CKL();
      k_schur_rhs<<<nc, 128, 0, ctx->stream>>>(S->d_cam_off.as<int>(), S->d_cam_obs.as<int>(), S->d_pt.as<int>(), S->d_T.as<double>(), S->d_gsp.as<double>(), S->d_gc.as<double>(),
                                               S->d_scc.as<double>(), S->d_fidx.as<int>(), 0, ctx->s_rhs.as<double>());
      CKL();
    }
    if (!pinned.empty()) { k_pin_params<<<(unsigned)pinned.size(), 128, 0, ctx->stream>>>(d_pinned.as<int>(), ctx->s_N, ctx->s_A.as<double>(), ctx->s_rhs.as<double>()); CKL(); }
    bool ok = false;
    rc = pvb_internal_factor_solve(ctx, ctx->s_N, &ok); if (rc) return rc;
    double model = 0.0;
    if (ok) {
      if (np) {
        k_schur_backsub<<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>(S->d_Cinv.as<double>(), S->d_gsp.as<double>(), S->d_T.as<double>(), S->d_pt_off.as<int>(), S->d_cam.as<int>(),
                                                                               S->d_fidx.as<int>(), ctx->s_rhs.as<double>(), np, S->d_yp.as<double>());
        CKL();
        CK(cudaMemcpyAsync(yp.data(), S->d_yp.p, (size_t)np * 24, cudaMemcpyDeviceToHost, ctx->stream));
      }
      CK(cudaMemcpyAsync(y.data(), ctx->s_rhs.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      // model cost change 0.5 (-y.g' + y^T D y) (see pvb_reproj_solve_lm), fixed order
      for (int i = 0; i < 6 * nb; ++i) {
        const int f = fidx[i];
        if (f < 0) continue;
        const double yi = y[f], s_ = sc[f], hd = hdiag[f] * s_ * s_;
        model += 0.5 * (-yi * g[f] * s_ + std::min(std::max(hd, 1e-6), 1e32) / radius * yi * yi);
      }
      for (long p = 0; p < np; ++p) {
        if (ptc[p]) continue;
        for (int k = 0; k < 3; ++k) {
          const double yi = yp[3 * p + k], s_ = scp[3 * p + k], hd = Cdiag(p, k) * s_ * s_;
          model += 0.5 * (-yi * gpt(3 * p + k) * s_ + std::min(std::max(hd, 1e-6), 1e32) / radius * yi * yi);
        }
      }
      ok = model > 0.0;
    }
    if (!ok) { radius *= 0.5; R.unsuccessful++; if (++invalid >= 5 || radius < 1e-32) { R.termination = 4; break; } continue; }
    invalid = 0;
    double sn = 0, xn = 0;
    std::copy(poses6, poses6 + 6 * (size_t)nb, cand_x.begin());
    if (np) std::copy(points3, points3 + 3 * (size_t)np, cand_p.begin());
    for (int i = 0; i < 6 * nb; ++i) { const int f = fidx[i]; if (f < 0) continue; const double d = y[f] * sc[f]; cand_x[i] += d; sn += d * d; xn += poses6[i] * poses6[i]; }
    for (long p = 0; p < np; ++p) { if (ptc[p]) continue; for (int k = 0; k < 3; ++k) { const double d = yp[3 * p + k] * scp[3 * p + k]; cand_p[3 * p + k] += d; sn += d * d; xn += points3[3 * p + k] * points3[3 * p + k]; } }
    sn = std::sqrt(sn); xn = std::sqrt(xn);
    double new_cost = 0.0;
    rc = evaluate(cand_x.data(), cand_p.data(), &new_cost); if (rc) return rc;
    bool accepted = false, stop = false;
    if (sn <= opt.parameter_tolerance * (xn + opt.parameter_tolerance)) { R.termination = 3; stop = true; }
    else {
      const double change = cost - new_cost;
      if (std::fabs(change) <= opt.function_tolerance * cost) { R.termination = 1; stop = true; }
      else {
        const double rho = change / model;
        if (rho > 1e-3) {
          accepted = true;
          std::copy(cand_x.begin(), cand_x.end(), poses6);
          if (np) std::copy(cand_p.begin(), cand_p.begin() + 3 * (size_t)np, points3);
          cost = new_cost;
          rc = assemble(); if (rc) return rc;
          R.successful++;
          radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
          decrease = 2.0;
          if (gmax() <= opt.gradient_tolerance) { R.termination = 2; stop = true; }
        } else {
          radius /= decrease; decrease *= 2.0; R.unsuccessful++;
          if (radius < 1e-32) { R.termination = 4; stop = true; }
        }
      }
    }
    if (stop) break;
    if (!accepted) {   // the candidate evaluation overwrote the reduced blocks of the current point: bring them back (the dense pose system is untouched)
      rc = evaluate(poses6, points3, &new_cost); if (rc) return rc;
    }
  }
  return finish();
}

/* AddCameraResidual's observation loop (util/Optimization.cpp:187-219, ANGLE_RESIDUAL_1): features of frames without a valid pose are skipped (:193);
 * the key point goes through Equirectangular::ImageToCam(cv::Point2i) — the cv::Point2f is rounded to the pixel grid by the implicit conversion
 * (cvRound, half to even) and mapped in float32 (sensors/Equirectangular.h:98-105, 118-131, 155-164). */
int pvb_build_reproj_observations(int rows, int cols, long n_tracks, const int* track_off, const int* feat_frame, const float* feat_xy, const unsigned char* pose_valid,
                                  long cap, int* cam, int* point, double* bearing3) {
  if (rows <= 0 || cols <= 0 || n_tracks < 0 || (n_tracks > 0 && (!track_off || !feat_frame || !feat_xy)) || (cap > 0 && (!cam || !point || !bearing3))) return PVB_ERR_ARG;
  long m = 0;
  for (long t = 0; t < n_tracks; ++t)
    for (int f = track_off[t]; f < track_off[t + 1]; ++f) {
      const int fr = feat_frame[f];
      if (pose_valid && !pose_valid[fr]) continue;
      if (m >= cap) return PVB_ERR_NOMEM;
      const float px = (float)std::nearbyintf(feat_xy[2 * (size_t)f]), py = (float)std::nearbyintf(feat_xy[2 * (size_t)f + 1]);
      const float lon = (float)((double)(2 * px / cols - 1) * M_PI);
      const float lat = (float)((0.5 - (double)(py / rows)) * M_PI);
      const float cy = std::cos(lat);
      cam[m] = fr; point[m] = (int)t;
      bearing3[3 * m] = (double)(1.f * cy * std::sin(lon));
      bearing3[3 * m + 1] = (double)(-1.f * std::sin(lat));
      bearing3[3 * m + 2] = (double)(1.f * cy * std::cos(lon));
      ++m;
    }
  return (int)m;
}

}  // extern "C"
