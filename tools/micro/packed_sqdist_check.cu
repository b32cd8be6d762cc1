#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__global__ void k(const float4* __restrict__ rec, int ngroups, float qx, float qy, float qz, float one, float* out_packed, float* out_scalar) {
  const unsigned long long QX = f2_pack(qx, qx), QY = f2_pack(qy, qy), QZ = f2_pack(qz, qz), ONE = f2_pack(one, one);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += gridDim.x * blockDim.x) {
    const float4 X = rec[3 * i], Y = rec[3 * i + 1], Z = rec[3 * i + 2];
    unsigned long long dx0 = f2_sub(f2_pack(X.x, X.y), QX), dx1 = f2_sub(f2_pack(X.z, X.w), QX);
    unsigned long long dy0 = f2_sub(f2_pack(Y.x, Y.y), QY), dy1 = f2_sub(f2_pack(Y.z, Y.w), QY);
    unsigned long long dz0 = f2_sub(f2_pack(Z.x, Z.y), QZ), dz1 = f2_sub(f2_pack(Z.z, Z.w), QZ);
    unsigned long long s0 = f2_fma(f2_fma(f2_mul(dx0, dx0), ONE, f2_mul(dy0, dy0)), ONE, f2_mul(dz0, dz0));
    unsigned long long s1 = f2_fma(f2_fma(f2_mul(dx1, dx1), ONE, f2_mul(dy1, dy1)), ONE, f2_mul(dz1, dz1));
    float d[4]; f2_unpack(s0, d[0], d[1]); f2_unpack(s1, d[2], d[3]);
    const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
    for (int k = 0; k < 4; ++k) {
      out_packed[4 * i + k] = d[k];
      const float dx = __fsub_rn(xs[k], qx), dy = __fsub_rn(ys[k], qy), dz = __fsub_rn(zs[k], qz);     // argument order as sqdist_f32(q, c)? see below
      out_scalar[4 * i + k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
  }
}
int main() {
  const int ng = 1 << 20;
  std::vector<float> h((size_t)ng * 12);
  srand(1);
  for (auto& v : h) v = (float)rand() / RAND_MAX * 200.f - 100.f + (float)rand() / RAND_MAX * 1e-3f;
  float4* d; float *a, *b;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&a, (size_t)ng * 16); cudaMalloc(&b, (size_t)ng * 16);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  k<<<592, 256>>>(d, ng, 1.2345f, -7.891f, 0.333f, 1.0f, a, b);
  std::vector<float> ha((size_t)ng * 4), hb((size_t)ng * 4);
  cudaMemcpy(ha.data(), a, ha.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), b, hb.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (size_t i = 0; i < ha.size(); ++i) bad += ha[i] != hb[i];
  printf("packed vs scalar: %ld of %zu differ (%s)\n", bad, ha.size(), cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
