// panovlm_b200 — pixel-space camera-LiDAR association, device stage.
//
// First stage of CameraLidarLineAssociate::Associate (joint_optimization/CameraLidarLineAssociate.cpp:22-91), the fallback for frames without
// LiDAR line segments (CameraLidarOptimizer.cpp:360-367): every LiDAR point -> camera frame (pcl::transformPointCloud: double math, float32
// store) -> pixel (Equirectangular::CamToImage in float with FastAtan2) -> its 3 nearest image sub-line mid points (cv::flann exact search on
// float squared L2).  The mid points (a few thousand per image) are staged tile by tile in shared memory and every thread scans them for its
// own point: ~28.8 k points x ~3 k mid points per frame.  No tensor cores: there is no contraction, only a 3-best selection.
#include "pvb_ctx.hpp"
#include "pvb_math.cuh"

using namespace pvb;

namespace {

constexpr int kMidTile = 2048;       // mid points staged per pass (16 KB of shared memory)

struct Pose34 { double R[9]; double t[3]; };

__global__ void __launch_bounds__(128) k_pixel_knn3(const F4* __restrict__ cloud, int n, Pose34 T, int rows, int cols, const float2* __restrict__ mid, int n_mid,
                                                    int* __restrict__ idx3, float* __restrict__ d2_3, float* __restrict__ pixel2) {
  __shared__ float2 tile[kMidTile];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < n;
  float u = 0.f, v = 0.f;
  if (act) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(cloud) + i);
    float x, y, z;
    transform_point_f32(T.R, T.t, p.x, p.y, p.z, x, y, z);
    cam_to_image_f32(x, y, z, rows, cols, u, v);
  }
  float b0 = INFINITY, b1 = INFINITY, b2 = INFINITY;
  int i0 = -1, i1 = -1, i2 = -1;
  for (int base = 0; base < n_mid; base += kMidTile) {
    const int m = min(kMidTile, n_mid - base);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) tile[k] = mid[base + k];
    __syncthreads();
    if (!act) continue;
#pragma unroll 4
    for (int k = 0; k < m; ++k) {
      const float2 c = tile[k];
      const float dx = fsub(u, c.x), dy = fsub(v, c.y);
      const float d = fadd(fmul(dx, dx), fmul(dy, dy));           // FLANN L2<float>: un-fused, x then y
      if (d < b2) {                                               // strict: the earlier mid point wins a tie
        const int id = base + k;
        if (d < b1) {
          b2 = b1; i2 = i1;
          if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = id; } else { b1 = d; i1 = id; }
        } else { b2 = d; i2 = id; }
      }
    }
  }
  if (!act) return;
  idx3[3 * i] = i0; idx3[3 * i + 1] = i1; idx3[3 * i + 2] = i2;
  if (d2_3) { d2_3[3 * i] = b0; d2_3[3 * i + 1] = b1; d2_3[3 * i + 2] = b2; }
  if (pixel2) { pixel2[2 * i] = u; pixel2[2 * i + 1] = v; }
}

}  // namespace

extern "C" int pvb_pixel_knn3(pvb_ctx* ctx, int rows, int cols, const float* mid2, int n_mid, const float* cloud_local, int n_points, const double* T_cl16, int* idx3,
                              float* d2_3, float* pixel2) {
  if (!ctx || rows <= 0 || cols <= 0 || n_mid < 0 || n_points < 0 || (n_mid > 0 && !mid2) || (n_points > 0 && (!cloud_local || !idx3)) || !T_cl16)
    return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_pixel_knn3: bad arguments") : PVB_ERR_ARG;
  if (n_points == 0) return PVB_OK;
  CK(cudaSetDevice(ctx->device));
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));      // scratch buffers are shared with the source-upload pipeline
  if (ctx->sort_stream) CK(cudaStreamSynchronize(ctx->sort_stream));
  CK(ctx->m_a.ensure((size_t)n_points * 16)); CK(ctx->m_b.ensure(std::max<size_t>(8, (size_t)n_mid * 8))); CK(ctx->m_c.ensure((size_t)n_points * 12));
  CK(ctx->m_d.ensure((size_t)n_points * 12)); CK(ctx->m_e.ensure((size_t)n_points * 8));
  CK(cudaMemcpyAsync(ctx->m_a.p, cloud_local, (size_t)n_points * 16, cudaMemcpyHostToDevice, ctx->stream));
  if (n_mid) CK(cudaMemcpyAsync(ctx->m_b.p, mid2, (size_t)n_mid * 8, cudaMemcpyHostToDevice, ctx->stream));
  Pose34 T;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.R[r * 3 + c] = T_cl16[r * 4 + c]; T.t[r] = T_cl16[r * 4 + 3]; }
  k_pixel_knn3<<<(n_points + 127) / 128, 128, 0, ctx->stream>>>(ctx->m_a.as<F4>(), n_points, T, rows, cols, ctx->m_b.as<float2>(), n_mid, ctx->m_c.as<int>(),
                                                              ctx->m_d.as<float>(), ctx->m_e.as<float>());
  CKL();
  CK(cudaMemcpyAsync(idx3, ctx->m_c.p, (size_t)n_points * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (d2_3) CK(cudaMemcpyAsync(d2_3, ctx->m_d.p, (size_t)n_points * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (pixel2) CK(cudaMemcpyAsync(pixel2, ctx->m_e.p, (size_t)n_points * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}
