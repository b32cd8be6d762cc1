// panovlm_b200 — sm_100a kernels of the hot path.  HBM-bound integer/float/double work: no tensor cores.
//   k_transform_world      T1  Velodyne::Transform2LidarWorld per cloud (float32 store) + per-cloud AABB
//   k_cell_keys (+ cub sort / scan), k_gather_f4   cell-sorted target layout (uniform grid, x fastest)
//   k_superrow_counts / k_superrow_fill / k_target_rk   merged super-rows + static search bounds of a static target (dense mode)
//   k_target_cell_keys     query order by target cell (dense mode)
//   k_associate<K,...>     K2p fused: query -> world, exact k-NN on the grid, class test, plane fit, collinearity,
//                              [emit correspondence] [residual + Jacobian rows] [per-tile 6x6 normal-equation partial]
//   k_eval_blocks          K3  residual + analytic Jacobian over a correspondence list (+ per-tile 12x12 partial)
//   k_sum_partials         K4  deterministic (fixed order) assembly of the per-edge / per-frame normal equations
//   k_project / k_splat    K1  SE(3) + equirectangular projection, sparse depth image
//   k_line_votes           K2l AssociateLine2Line vote matrix
//   k_angle_votes          K2c AssociateByAngle vote counts
#pragma once
#include <cuda_runtime.h>
#include "pvb_knn.cuh"

namespace pvb {

constexpr int kTile = 128;   // queries / residual rows per thread block
#ifndef PVB_K3_MINB
#define PVB_K3_MINB 6        // resident blocks per SM the residual kernel is compiled for: measured 4: 0.176 ms, 5: 0.163, 6: 0.153, 8: 0.156 (reduced mode, Room-shaped list)
#endif

struct WorldPose { double R[9]; double t[3]; };   // T_wl of a pose block (row-major R)

struct CloudTile { int cloud; int start; int count; int pad; };          // points [start, start+count) of `cloud`
struct Pair { int target_cloud; int query_cloud; int ref_block; int nei_block; };
struct QueryTile { int pair; int start; int count; int out_base; };       // queries [start, start+count) (global index); output slot of the first

PVB_HD uint32_t ordered_f32(float f) { const uint32_t u = f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
PVB_HD float unordered_f32(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

#ifdef __CUDACC__

__device__ __forceinline__ F4 ldg_f4(const F4* p) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}

// two small host-mapped arrays -> device (pose staging, see upload_poses)
__global__ void __launch_bounds__(256) k_copy_words(const unsigned long long* __restrict__ a, long long na, unsigned long long* __restrict__ da,
                                                    const unsigned long long* __restrict__ b, long long nb, unsigned long long* __restrict__ db) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) da[i] = a[i];
  else if (i < na + nb) db[i - na] = b[i - na];
}

// ---- T1: local -> world (float32) for every point of every cloud; w = (index in cloud << 5) | class ----------
__global__ void __launch_bounds__(256) k_transform_world(const F4* __restrict__ local, const CloudTile* __restrict__ tiles,
                                                         const int* __restrict__ cloud_block, const WorldPose* __restrict__ wpose,
                                                         const int* __restrict__ cloud_off, F4* __restrict__ world, uint32_t* __restrict__ aabb /*[cloud][6]*/) {
  const CloudTile t = tiles[blockIdx.x];
  const WorldPose& wp = wpose[cloud_block[t.cloud]];
  const int i = threadIdx.x;
  float x = 0, y = 0, z = 0;
  const bool act = i < t.count;
  if (act) {
    const F4 p = ldg_f4(local + t.start + i);
    transform_point_f32(wp.R, wp.t, p.x, p.y, p.z, x, y, z);
    F4 o; o.x = x; o.y = y; o.z = z;
    o.w = u2f(((uint32_t)(t.start + i - cloud_off[t.cloud]) << 5) | ((uint32_t)p.w & 31u));
    reinterpret_cast<float4*>(world)[t.start + i] = make_float4(o.x, o.y, o.z, o.w);
  }
  if (aabb) {
    uint32_t lo[3] = {act ? ordered_f32(x) : 0xFFFFFFFFu, act ? ordered_f32(y) : 0xFFFFFFFFu, act ? ordered_f32(z) : 0xFFFFFFFFu};
    uint32_t hi[3] = {act ? ordered_f32(x) : 0u, act ? ordered_f32(y) : 0u, act ? ordered_f32(z) : 0u};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o)); hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o)); }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { atomicMin(aabb + t.cloud * 6 + c, lo[c]); atomicMax(aabb + t.cloud * 6 + 3 + c, hi[c]); }
    }
  }
}

__global__ void __launch_bounds__(256) k_transform_simple(const F4* __restrict__ in, long long n, WorldPose wp, F4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const F4 p = ldg_f4(in + i);
  float x, y, z;
  transform_point_f32(wp.R, wp.t, p.x, p.y, p.z, x, y, z);
  reinterpret_cast<float4*>(out)[i] = make_float4(x, y, z, p.w);
}

// ---- sweep undistortion of a batch of frames (Velodyne::UndistortCloud): 16 B in, 16 B out per point, two FP64 sines --------------
__global__ void __launch_bounds__(256) k_undistort(const F4* __restrict__ in, const CloudTile* __restrict__ tiles, const UndistortPrep* __restrict__ prep,
                                                   const int* __restrict__ cloud_off, F4* __restrict__ out) {
  const CloudTile t = tiles[blockIdx.x];
  const int i = threadIdx.x;
  if (i >= t.count) return;
  const F4 p = ldg_f4(in + t.start + i);
  const UndistortPrep& u = prep[t.cloud];
  float x = p.x, y = p.y, z = p.z;
  if (u.enabled) undistort_point_f32(u, (long long)(t.start + i - cloud_off[t.cloud]), (long long)(cloud_off[t.cloud + 1] - cloud_off[t.cloud]), p.x, p.y, p.z, x, y, z);
  reinterpret_cast<float4*>(out)[t.start + i] = make_float4(x, y, z, p.w);
}

// ---- grid build -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_keys(const F4* __restrict__ world, const CloudTile* __restrict__ tiles, const GridDesc* __restrict__ grids,
                                                   unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist) {
  const CloudTile t = tiles[blockIdx.x];
  const int i = threadIdx.x;
  if (i >= t.count) return;
  const GridDesc& g = grids[t.cloud];
  const F4 p = ldg_f4(world + t.start + i);
  const int cx = cell_coord((double)p.x, g.origin[0], g.inv_h, g.dims[0]);
  const int cy = cell_coord((double)p.y, g.origin[1], g.inv_h, g.dims[1]);
  const int cz = cell_coord((double)p.z, g.origin[2], g.inv_h, g.dims[2]);
  const unsigned long long cell = (unsigned long long)g.cell_base + ((unsigned long long)cz * g.dims[1] + cy) * g.dims[0] + cx;
  keys[t.start + i] = cell;
  vals[t.start + i] = (uint32_t)(t.start + i);
  atomicAdd(hist + cell, 1u);
}

__global__ void __launch_bounds__(256) k_gather_f4(const F4* __restrict__ src, const uint32_t* __restrict__ idx, long long n, F4* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + idx[i]);
}

// squared distance of every target record to its K-th nearest target record (itself included) within sqrt(sq_thr), +inf when there are fewer: the static
// search bound of MODE 4 (knn_select_superrow)
template <int K>
__global__ void __launch_bounds__(128) k_target_rk(GridDesc g, const uint32_t* __restrict__ cell_start, const F4* __restrict__ sorted, long long n, float sq_thr, int rmax, float* __restrict__ rk) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* cs = cell_start + g.cell_base;
  auto cells = [cs](long long c) { return (long long)__ldg(cs + c); };
  auto load = [sorted](long long p) { return ldg_f4(sorted + p); };
  const F4 q = ldg_f4(sorted + i);
  uint32_t tau = 0x7F800000u;
  const int found = knn_select_pruned<K>(g, cells, load, q.x, q.y, q.z, sq_thr, 1, rmax, [](int, uint32_t, uint32_t) {}, &tau);
  rk[i] = found >= K ? __uint_as_float(tau) : INFINITY;
}

// ---- merged super-rows of a static target (MODE 4): segment (row = (y, z), x) = records of the 9 cells (x, y + dy, z + dz), dz outer, dy inner ----------
__global__ void __launch_bounds__(256) k_superrow_counts(GridDesc g, const uint32_t* __restrict__ cell_start, long long ncells, uint32_t* __restrict__ cnt) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncells) return;
  if (c == ncells) { cnt[c] = 0u; return; }                     // the slot after the last cell (end of the last segment)
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int x = (int)(c % nx); const long long row = c / nx; const int y = (int)(row % ny), z = (int)(row / ny);
  const uint32_t* cs = cell_start + g.cell_base;
  uint32_t sum = 0u;
  for (int dz = -1; dz <= 1; ++dz) {
    const int zz = z + dz; if (zz < 0 || zz >= nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy; if (yy < 0 || yy >= ny) continue;
      const long long cc = ((long long)zz * ny + yy) * nx + x;
      sum += __ldg(cs + cc + 1) - __ldg(cs + cc);
    }
  }
  cnt[c] = (sum + 3u) & ~3u;                                     // segments are padded to whole groups of 4 records
}
// one thread per destination segment (build time only): records as groups of 4 in {x0..3}, {y0..3}, {z0..3} float4 triples, the tail of the last
// group filled with points at +infinity (never below any bound)
__global__ void __launch_bounds__(256) k_superrow_fill(GridDesc g, const uint32_t* __restrict__ cell_start, const F4* __restrict__ sorted, long long ncells,
                                                       const uint32_t* __restrict__ sstart, float* __restrict__ quads, uint32_t* __restrict__ srw, const float* __restrict__ rk,
                                                       float* __restrict__ srk) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int x = (int)(c % nx); const long long row = c / nx; const int y = (int)(row % ny), z = (int)(row / ny);
  const uint32_t* cs = cell_start + g.cell_base;
  uint32_t dst = sstart[c];
  const uint32_t end = sstart[c + 1];
  if (end == dst) return;
  auto put = [&](uint32_t r, float px, float py, float pz, uint32_t w, float k) {
    float* q = quads + (size_t)(r >> 2) * 12 + (r & 3u);
    q[0] = px; q[4] = py; q[8] = pz; srw[r] = w; srk[r] = k;
  };
  for (int dz = -1; dz <= 1; ++dz) {
    const int zz = z + dz; if (zz < 0 || zz >= nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy; if (yy < 0 || yy >= ny) continue;
      const long long cc = ((long long)zz * ny + yy) * nx + x;
      const uint32_t lo = __ldg(cs + cc), hi = __ldg(cs + cc + 1);
      for (uint32_t i = lo; i < hi; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(sorted) + i);          // cell_start holds positions in the sorted array of the whole cloud set
        put(dst++, v.x, v.y, v.z, (i << 5) | (__float_as_uint(v.w) & 31u), rk[i]);
      }
    }
  }
  for (; dst < end; ++dst) put(dst, INFINITY, INFINITY, INFINITY, 0u, INFINITY);
}

// Morton key of a source point in its own sensor frame (cells of 1 / inv_cell metres inside +-512 m, `bits` bits per axis), frame id in the high bits.
__device__ __forceinline__ unsigned long long spread3(uint32_t v) {       // <= 21 bits -> every third bit
  unsigned long long x = v & 0x1FFFFFull;
  x = (x | (x << 32)) & 0x001F00000000FFFFull;
  x = (x | (x << 16)) & 0x001F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}
__global__ void __launch_bounds__(256) k_morton_keys(const F4* __restrict__ local, const CloudTile* __restrict__ tiles, int first_cloud, float inv_cell, int bits,
                                                     unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const CloudTile t = tiles[blockIdx.x];
  const int i = threadIdx.x;
  if (i >= t.count) return;
  const F4 p = ldg_f4(local + t.start + i);
  const float top = (float)((1u << bits) - 1u);
  auto q = [&](float v) { const float f = floorf((v + 512.f) * inv_cell); return (uint32_t)(f < 0.f ? 0.f : (f > top ? top : f)); };
  const unsigned long long m = spread3(q(p.x)) | (spread3(q(p.y)) << 1) | (spread3(q(p.z)) << 2);
  keys[t.start + i] = ((unsigned long long)(t.cloud - first_cloud) << (3 * bits)) | m;
  vals[t.start + i] = (uint32_t)(t.start + i);
}

// Target-cell key of a source point under the CURRENT pose of its frame (dense mode, MODE 4): queries sorted by (frame, cell of the target grid, x fastest)
// put the lanes of a warp into the same few cells, so their super-row ranges coincide (same trip counts, broadcast loads).  Also the largest squared
// distance of a frame's points from its sensor (float bits): the host bounds how far a pose update can move a point without looking at the cloud again.
template <typename KeyT>
__global__ void __launch_bounds__(256) k_target_cell_keys(const F4* __restrict__ local, const CloudTile* __restrict__ tiles, int first_cloud, const WorldPose* __restrict__ wpose,
                                                          GridDesc g, int cellbits, KeyT* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ rmax2) {
  const CloudTile t = tiles[blockIdx.x];
  const int i = threadIdx.x;
  float r2 = 0.f;
  if (i < t.count) {
    const F4 p = ldg_f4(local + t.start + i);
    const WorldPose& wp = wpose[t.cloud + 1];                    // block 0 is the target's identity pose
    float x, y, z;
    transform_point_f32(wp.R, wp.t, p.x, p.y, p.z, x, y, z);
    const int cx = cell_coord((double)x, g.origin[0], g.inv_h, g.dims[0]);
    const int cy = cell_coord((double)y, g.origin[1], g.inv_h, g.dims[1]);
    const int cz = cell_coord((double)z, g.origin[2], g.inv_h, g.dims[2]);
    const unsigned long long cell = ((unsigned long long)cz * g.dims[1] + cy) * g.dims[0] + cx;
    keys[t.start + i] = (KeyT)(((unsigned long long)(t.cloud - first_cloud) << cellbits) | cell);
    vals[t.start + i] = (uint32_t)(t.start + i);
    r2 = p.x * p.x + p.y * p.y + p.z * p.z;
  }
  if (rmax2) {                                                   // only after an upload (the cloud does not change with the poses); one atomic per block:
    __shared__ uint32_t s_m[8];                                  // same-address atomics of every warp made this kernel 4x slower than its memory traffic
    uint32_t m = __float_as_uint(r2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < 8; ++w) m = max(m, s_m[w]);
      if (m) atomicMax(rmax2 + t.cloud, m);
    }
  }
}

// Locality of a query order (dense mode): number of adjacent pairs (j - 1, j) of the same frame whose target cells are NOT neighbours (differ by more than one
// cell along an axis).  keys[] are the target-cell keys by ORIGINAL index (k_target_cell_keys), order[] the permutation in use.  After a sort the count is
// the number of jumps of the occupied-cell sequence; a re-used permutation whose count stays near that value still groups the lanes of a warp spatially.
template <typename KeyT>
__global__ void __launch_bounds__(256) k_order_locality(const KeyT* __restrict__ keys, const uint32_t* __restrict__ order, long long cnt, GridDesc g, int cellbits,
                                                        unsigned int* __restrict__ counter) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool far = false;
  if (j > 0 && j < cnt) {
    const unsigned long long a = (unsigned long long)keys[order[j]], b = (unsigned long long)keys[order[j - 1]];
    if ((a >> cellbits) == (b >> cellbits)) {
      const unsigned long long mask = (1ull << cellbits) - 1ull;
      const long long ca = (long long)(a & mask), cb = (long long)(b & mask);
      const int nx = g.dims[0], ny = g.dims[1];
      const int xa = (int)(ca % nx), xb = (int)(cb % nx);
      const long long ra = ca / nx, rb = cb / nx;
      const int ya = (int)(ra % ny), yb = (int)(rb % ny), za = (int)(ra / ny), zb = (int)(rb / ny);
      far = abs(xa - xb) > 1 || abs(ya - yb) > 1 || abs(za - zb) > 1;
    }
  }
  const int n = __syncthreads_count(far ? 1 : 0);
  if (threadIdx.x == 0 && n) atomicAdd(counter, (unsigned int)n);
}

// ---- K2p: fused associate (+ residual + reduce) --------------------------------------------------------------------
struct AssocArgs {
  const F4* q_local;            // query records, local frame: x,y,z,intensity(class)  (dense mode: Morton order)
  const uint32_t* q_orig;       // original index of each (re-ordered) query, or null (identity)
  const QueryTile* tiles;
  const Pair* pairs;
  const GridDesc* grids;
  const uint32_t* cell_start;
  const F4* sorted;             // cell-sorted world target records
  const WorldPose* wpose;       // per pose block
  const PosePrep* prep;         // per pose block
  AssocParams prm;              // rmax is derived per target grid in the kernel
  double thr;                   // distance threshold (m)
  // residual
  int residual_type, normalize; double huber, weight;
  // outputs (any may be null)
  unsigned char* out_valid; double* out_point; double* out_plane;   // indexed by output slot: q_orig[query] (dense) or tile.out_base + lane
  double* out_res; double* out_jac6;                                 // idem
  int* out_nn_idx; float* out_nn_d2;                                 // idem, K per query (debug / parity)
  double* partials;                                                  // [tile][warp][29]
  unsigned long long* stats;                                         // optional: [0] tiles staged through TMA, [1] tiles on the global path
  // search-radius hints of the buffered single-pass search (MODE 2 / 4), one record per query by OUTPUT SLOT (the query's original index in dense mode,
  // so the hints survive a re-ordering of the queries): {x, y, z of the query in the world
  // frame at the last evaluation, its K-th squared distance then (float; not finite = no hint)}.  Read and rewritten in place; may be null.
  F4* hint;
  int use_hint;                                                      // 0: ignore the stored hints (they are still rewritten)
  int flat_walk;                                                     // 1: the hinted walk runs over the flattened row ranges (walk_block_collect_flat)
  double tight_frac;                                                 // MODE 4: a bound counts as tight when the predicted survivors stay below this share of the list
  int use_static;                                                    // MODE 4: 1 = queries without a usable hint take the static bound of the target (srk)
  // MODE 4: merged super-rows of a static target (knn_select_superrow): records with w = (position in `sorted` << 5) | class, and the start of
  // every (row, x cell) segment, indexed like cell_start (GridDesc::cell_base applies)
  const float4* srow; const uint32_t* sstart; const uint32_t* srw; const float* srk;      // srow: groups of 4 records as {x0..3},{y0..3},{z0..3}; srw[r] = (position << 5) | class;
  float one;                                                         // srk[r]: squared distance of record r to its 10th nearest target point (static bound); one = 1.0f (see DevSuperRow::sqdist4)
};

// ---- TMA staging helpers (sm_90+/sm_100a): 1-D bulk copies global -> shared completing on an mbarrier ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  }
}

constexpr int kStageRows = 64;     // (y,z) rows of a tile's cell box that can be staged
constexpr int kStageCap = 768;     // staged records per tile (12 KB of shared memory)

constexpr int kListCap = 24;       // per-query candidate list of the buffered single-pass search (8 B per entry)
#ifndef PVB_LC4
#define PVB_LC4 24
#endif
constexpr int kListCap4 = PVB_LC4;      // the same for MODE 4 (room for the ~1.7 K survivors of the static bound before a compaction)

// MODE: 0 = TMA-staged tile + exhaustive walk, 1 = pruned two-pass walk, 2 = buffered single pass with search-radius hints (frames mode default),
//       4 = buffered single pass over the merged super-rows of a static target (dense mode default)
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct DevSuperRow {
  const uint32_t* st; const float4* quads; const uint32_t* wp; const float* rkp; float one;
  __device__ __forceinline__ float rk(uint32_t r) const { return __ldg(rkp + r); }
  __device__ __forceinline__ uint32_t w(uint32_t r) const { return __ldg(wp + r); }
  __device__ __forceinline__ uint32_t start(long long i) const { return __ldg(st + i); }
  // squared distances of the 4 records of group G to q: packed FP32x2 arithmetic, two candidates per instruction.  Every operation rounds like the scalar
  // one (sub.rn / mul.rn per half); the additions are fma(a, one, b) with `one` = 1.0f read from the kernel arguments: a * 1 is exact, so the result is
  // the correctly rounded a + b, and ptxas cannot contract it with the multiplication that produced `a` (it does contract mul.rn.f32x2 + add.rn.f32x2
  // into FFMA2, which would change the last bit; tools/micro/packed_sqdist_check.cu compares 4 M distances with the scalar __fmul_rn / __fadd_rn path).
  struct Quad { float4 X, Y, Z; };
  __device__ __forceinline__ Quad load3(uint32_t G) const { Quad q; q.X = __ldg(quads + 3u * G); q.Y = __ldg(quads + 3u * G + 1u); q.Z = __ldg(quads + 3u * G + 2u); return q; }
  __device__ __forceinline__ void sqdist4(const Quad& c, float qx, float qy, float qz, uint32_t (&kb)[4]) const {
    const float4 X = c.X, Y = c.Y, Z = c.Z;
    const unsigned long long QX = f2_pack(qx, qx), QY = f2_pack(qy, qy), QZ = f2_pack(qz, qz), ONE = f2_pack(one, one);
    const unsigned long long dx0 = f2_sub(f2_pack(X.x, X.y), QX), dx1 = f2_sub(f2_pack(X.z, X.w), QX);
    const unsigned long long dy0 = f2_sub(f2_pack(Y.x, Y.y), QY), dy1 = f2_sub(f2_pack(Y.z, Y.w), QY);
    const unsigned long long dz0 = f2_sub(f2_pack(Z.x, Z.y), QZ), dz1 = f2_sub(f2_pack(Z.z, Z.w), QZ);
    const unsigned long long s0 = f2_fma(f2_fma(f2_mul(dx0, dx0), ONE, f2_mul(dy0, dy0)), ONE, f2_mul(dz0, dz0));
    const unsigned long long s1 = f2_fma(f2_fma(f2_mul(dx1, dx1), ONE, f2_mul(dy1, dy1)), ONE, f2_mul(dz1, dz1));
    float d0, d1, d2, d3;
    f2_unpack(s0, d0, d1); f2_unpack(s1, d2, d3);
    kb[0] = __float_as_uint(d0); kb[1] = __float_as_uint(d1); kb[2] = __float_as_uint(d2); kb[3] = __float_as_uint(d3);
  }
};
template <int K, bool REDUCE, int MINB, bool DEBUG_NN, int MODE, bool REF_ID>
__global__ void __launch_bounds__(kTile, MINB) k_associate(const AssocArgs a) {
  constexpr bool STAGE = MODE == 0;
  constexpr bool LIST = MODE == 2 || MODE == 4;      // per-query candidate list in shared memory
  constexpr int LC = MODE == 4 ? kListCap4 : kListCap;
  __shared__ double sJ_own[(REDUCE && !LIST) ? kTile : 1][8];   // per row: J6 (nei pose) | r | cost   (row stride 8 doubles = 64 B); MODE 2: aliases the warp's list
  __shared__ uint32_t s_win[LIST ? 1 : K][kTile];           // record positions of each query's K neighbours (MODE 2: slots 0..K-1 of the list)
  // MODE 2: per warp (d2 bits, record position) of the candidates below the hinted bound, [warp][entry][lane]; its last slots double as the row table of the
  // flattened walk and, once the warp's searches and plane fits are done, its first 2 KB as the staging rows of the warp's reduction
  __shared__ __align__(16) U2 s_list[LIST ? kTile / 32 : 1][LIST ? LC : 1][32];
  static_assert(!LIST || LC * 8 >= K * 16, "the neighbour records alias the warp's list");
  static_assert(!LIST || sizeof(U2) * LC * 32 >= sizeof(double) * 32 * 8, "the reduction rows alias the warp's list");
  double (*sJ)[8] = LIST ? reinterpret_cast<double (*)[8]>(&s_list[0][0][0]) : sJ_own;      // MODE 2: re-pointed per warp below
  __shared__ uint32_t s_rng[18][(MODE == 0) ? kTile : 1];          // the <= 9 (lo, hi) row ranges of each query's 3x3x3 cell block
  __shared__ __align__(16) F4 s_pts[STAGE ? kStageCap : 1];          // the tile's candidate rows, copied by TMA
  __shared__ uint32_t s_row_lo[STAGE ? kStageRows : 1], s_row_base[STAGE ? kStageRows : 1];
  __shared__ int s_box[6];
  __shared__ int s_staged;
  __shared__ __align__(8) uint64_t s_bar;
  const QueryTile t = a.tiles[blockIdx.x];
  const Pair pr = a.pairs[t.pair];
  const int i = threadIdx.x;
  const bool act = i < t.count;
  bool valid = false, dyn_tight = false;
  int found = 0;
  uint32_t lim_hint = 0u, tau = 0x7F800000u;
  double p_local[3] = {0, 0, 0}, plane[4] = {0, 0, 0, 0};
  double r = 0.0, cost = 0.0, J[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) J[k] = 0.0;
  uint32_t qi = 0;
  const GridDesc& g = a.grids[pr.target_cloud];
  const uint32_t* cs = a.cell_start + g.cell_base;
  const F4* srt = a.sorted;
  const WorldPose& wn = a.wpose[pr.nei_block];
  const WorldPose& wr = a.wpose[pr.ref_block];
  F4 q; q.x = q.y = q.z = q.w = 0.f;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (act) {
    q = ldg_f4(a.q_local + t.start + i);
    transform_point_f32(wn.R, wn.t, q.x, q.y, q.z, qx, qy, qz);
  }
  int nyb = 1, by0 = 0, bz0 = 0;
  bool staged = false;
  if (STAGE) {
    // ---- stage the rows of the tile's cell box (its queries' cells +- 1) in shared memory with TMA bulk copies ------------
    if (i == 0) { s_box[0] = s_box[1] = s_box[2] = 0x7fffffff; s_box[3] = s_box[4] = s_box[5] = -1; s_staged = 0; }
    __syncthreads();
    int c3[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, d3[3] = {-1, -1, -1};
    if (act) {
      c3[0] = d3[0] = cell_coord((double)qx, g.origin[0], g.inv_h, g.dims[0]);
      c3[1] = d3[1] = cell_coord((double)qy, g.origin[1], g.inv_h, g.dims[1]);
      c3[2] = d3[2] = cell_coord((double)qz, g.origin[2], g.inv_h, g.dims[2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { c3[k] = min(c3[k], __shfl_xor_sync(0xffffffffu, c3[k], o)); d3[k] = max(d3[k], __shfl_xor_sync(0xffffffffu, d3[k], o)); }
    }
    if ((i & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { atomicMin(&s_box[k], c3[k]); atomicMax(&s_box[3 + k], d3[k]); }
    }
    __syncthreads();
    const int bx0 = max(0, s_box[0] - 1), bx1 = min(g.dims[0] - 1, s_box[3] + 1);
    by0 = max(0, s_box[1] - 1); const int by1 = min(g.dims[1] - 1, s_box[4] + 1);
    bz0 = max(0, s_box[2] - 1); const int bz1 = min(g.dims[2] - 1, s_box[5] + 1);
    nyb = by1 - by0 + 1;
    const int rows = nyb * (bz1 - bz0 + 1);
    if (s_box[3] >= 0 && rows <= kStageRows) {
      if (i < rows) {
        const int y = by0 + i % nyb, z = bz0 + i / nyb;
        const long long row = ((long long)z * g.dims[1] + y) * g.dims[0];
        const uint32_t lo = __ldg(cs + row + bx0), hi = __ldg(cs + row + bx1 + 1);
        s_row_lo[i] = lo; s_row_base[i] = hi - lo;            // length for now
      }
      __syncthreads();
      if (i == 0) {
        uint32_t total = 0;
        for (int k = 0; k < rows; ++k) { const uint32_t len = s_row_base[k]; s_row_base[k] = total; total += len; }
        if (total > 0 && total <= (uint32_t)kStageCap) {
          mbar_init(&s_bar, 1);
          mbar_expect_tx(&s_bar, total * 16u);
          for (int k = 0; k < rows; ++k) {
            const uint32_t len = (k + 1 < rows ? s_row_base[k + 1] : total) - s_row_base[k];
            if (len) tma_bulk_g2s(&s_pts[s_row_base[k]], srt + s_row_lo[k], len * 16u, &s_bar);
          }
          s_staged = 1;
        }
      }
      __syncthreads();
      staged = s_staged != 0;
      if (staged) mbar_wait(&s_bar, 0);
    }
    if (a.stats && i == 0) atomicAdd(a.stats + (staged ? 0 : 1), 1ull);
  }
  if (act) {
    const int gq = t.start + i;
    qi = a.q_orig ? a.q_orig[gq] : (uint32_t)(t.out_base + i);
    auto cells = [cs](long long c) { return (long long)__ldg(cs + c); };
    auto loadg = [srt](long long p) { return ldg_f4(srt + p); };
    auto load1 = [&](long long p) { return (STAGE && staged) ? s_pts[p] : ldg_f4(srt + p); };
    auto row_map = [&](int y, int z, uint32_t& lo, uint32_t& hi) {
      if (STAGE && staged) {
        const int rl = (z - bz0) * nyb + (y - by0);
        const uint32_t nlo = s_row_base[rl] + (lo - s_row_lo[rl]);
        hi = nlo + (hi - lo); lo = nlo;
      }
    };
    U2* const my_list = &s_list[LIST ? (i >> 5) : 0][0][LIST ? (i & 31) : 0];          // entry e of this query: my_list[e * 32]
    auto win = [&](int j) { return LIST ? my_list[j * 32].y : s_win[j][i]; };
    auto set_win = [&](int j, uint32_t pos) { if (LIST) my_list[j * 32].y = pos; else s_win[j][i] = pos; };
    auto range_set = [&](int k, uint32_t lo, uint32_t hi) { if (MODE == 0) { s_rng[2 * k][i] = lo; s_rng[2 * k + 1][i] = hi; } };
    auto range_get = [&](int k, uint32_t& lo, uint32_t& hi) { if (MODE == 0) { lo = s_rng[2 * k][i]; hi = s_rng[2 * k + 1][i]; } else { lo = hi = 0u; } };
    AssocParams prm = a.prm;
    prm.rmax = (int)ceil(a.thr / g.h);
    if (DEBUG_NN && a.out_nn_idx) {
#pragma unroll
      for (int j = 0; j < K; ++j) set_win(j, 0xFFFFFFFFu);
    }
    // search-radius hint: K target points lay within sqrt(tau_old) of q_old => within sqrt(tau_old) + |q - q_old| of q
    if (LIST && a.hint && a.use_hint) {
      const F4 hq = ldg_f4(a.hint + qi);
      if (hq.w >= 0.f && hq.w < 3.0e38f) {
        const double dx = (double)qx - (double)hq.x, dy = (double)qy - (double)hq.y, dz = (double)qz - (double)hq.z;
        const double rad = (sqrt((double)hq.w) + sqrt(dx * dx + dy * dy + dz * dz)) * (1.0 + 1e-5) + 1e-9;
        const double lim2 = rad * rad;
        // the bound is "tight" when the list can be expected to hold what lies below it (points on a surface: count ~ K * lim2 / tau_old); after a large
        // pose change it is loose: MODE 2 then takes the two-pass search, MODE 4 also looks for the static bound of the target and keeps the smaller one
        if (lim2 < (double)prm.sq_thr) {
          dyn_tight = lim2 * (double)K < (double)hq.w * ((MODE == 4 ? a.tight_frac : 0.8) * LC);      // MODE 4: a looser bound is still used, together with the static one
          if (dyn_tight || MODE == 4) lim_hint = f2u((float)lim2) + 2u;      // +1 ulp for the float rounding, +1 to make the bound exclusive
        }
      }
    }
    if (MODE == 2) {
      // a warp whose lanes took different search paths pays for both: the bounded single pass is only taken when every active lane of the warp has a usable
      // bound (after a large pose change most warps fall back as a whole and cost what the two-pass search costs, not the sum).  MODE 4 has the static
      // bound of the target for the lanes without one, so every lane takes the single pass.
      const unsigned am = __activemask();
      if (__ballot_sync(am, lim_hint == 0u)) lim_hint = 0u;
    }
    if (!LIST) {
      DevSuperRow sr{};
      valid = associate_point2plane<K, REF_ID, MODE, LC>(g, cells, load1, loadg, row_map, prm, qx, qy, qz, (uint32_t)q.w & 31u, wr.R, wr.t, wn.R, wn.t, p_local, plane, win, set_win, range_set,
                                       range_get, lim_hint, &tau, my_list, 32, a.flat_walk != 0, sr);
    } else if (MODE == 2) {
      DevSuperRow sr{};
      found = search_list_modes<K, MODE, LC, SingleLane>(g, cells, loadg, prm, true, qx, qy, qz, set_win, lim_hint, dyn_tight, &tau, my_list, 32, a.flat_walk != 0, sr, false);
    }
  }
  if (MODE == 4) {      // warp-synchronous search: every lane of the warp takes part (inactive lanes with an empty range)
    auto cells = [cs](long long c) { return (long long)__ldg(cs + c); };
    auto loadg = [srt](long long p) { return ldg_f4(srt + p); };
    U2* const my_list = &s_list[LIST ? (i >> 5) : 0][0][LIST ? (i & 31) : 0];
    auto set_win = [&](int j, uint32_t pos) { my_list[j * 32].y = pos; };
    AssocParams prm = a.prm;
    prm.rmax = (int)ceil(a.thr / g.h);
    DevSuperRow sr; sr.st = a.sstart + g.cell_base; sr.quads = a.srow; sr.wp = a.srw; sr.rkp = a.srk; sr.one = a.one;
    found = search_list_modes<K, MODE, LC, WarpLanes>(g, cells, loadg, prm, act, qx, qy, qz, set_win, lim_hint, dyn_tight, &tau, my_list, 32, false, sr, a.use_static != 0);
  }
  if (LIST && act && a.hint) reinterpret_cast<float4*>(a.hint)[qi] = make_float4(qx, qy, qz, u2f(tau));
  // list modes: every lane of the warp leaves its search here; the K positions move to registers and the warp's list region becomes the store of the
  // neighbours' records ([j][lane], 16 B each: K * 16 <= LC * 8), fetched once per query and read by the three passes of the plane fit
  uint32_t wp[LIST ? K : 1];
  if (LIST) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; ++j) wp[j] = (act && found >= K) ? s_list[i >> 5][j][i & 31].y : 0xFFFFFFFFu;
    __syncwarp();
  }
  if (act) {
    auto loadg = [srt](long long p) { return ldg_f4(srt + p); };
    AssocParams prm = a.prm;
    auto win = [&](int j) { return LIST ? wp[LIST ? j : 0] : s_win[j][i]; };
    if (LIST && found >= K) {
      struct SmemNb {
        F4* base;
        __device__ __forceinline__ void sync() {}
        __device__ __forceinline__ void put(int j, const F4& v) { reinterpret_cast<float4*>(base)[j * 32] = make_float4(v.x, v.y, v.z, v.w); }
        __device__ __forceinline__ F4 get(int j) const { const float4 v = reinterpret_cast<const float4*>(base)[j * 32]; F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
      } nb;
      nb.base = reinterpret_cast<F4*>(&s_list[i >> 5][0][0]) + (i & 31);
      auto set_wp = [&](int j, uint32_t pos) { wp[LIST ? j : 0] = pos; };
      valid = plane_from_window<K, REF_ID>(loadg, prm, qx, qy, qz, (uint32_t)q.w & 31u, wr.R, wr.t, wn.R, wn.t, p_local, plane, win, set_wp, nb);
    }
    auto load = loadg;     // the debug view below is only built without staging
    if (DEBUG_NN && a.out_nn_idx) {     // debug / parity view: neighbours ordered by (d2, record position)
      for (int j = 0; j < K; ++j) {
        const uint32_t pj = win(j);
        if (pj == 0xFFFFFFFFu) { a.out_nn_idx[(size_t)qi * K + (K - 1 - j)] = -1; a.out_nn_d2[(size_t)qi * K + (K - 1 - j)] = INFINITY; continue; }
        const F4 rj = load((long long)pj);
        const float dj = sqdist_f32(qx, qy, qz, rj.x, rj.y, rj.z);
        int rank = 0;
        for (int m = 0; m < K; ++m) {
          const uint32_t pm = win(m);
          if (pm == 0xFFFFFFFFu || m == j) continue;
          const F4 rm = load((long long)pm);
          const float dm = sqdist_f32(qx, qy, qz, rm.x, rm.y, rm.z);
          rank += (dm < dj || (dm == dj && pm < pj)) ? 1 : 0;
        }
        a.out_nn_idx[(size_t)qi * K + rank] = (int)(f2u(rj.w) >> 5);
        a.out_nn_d2[(size_t)qi * K + rank] = dj;
      }
    }
    if (valid && (REDUCE || a.out_res)) {
      double c[8] = {p_local[0], p_local[1], p_local[2], plane[0], plane[1], plane[2], plane[3], a.weight};
      double q[3], P[3], g[3];
      transform_nei_to_ref(a.prep[pr.ref_block], a.prep[pr.nei_block], c, q, P);
      r = tail_point_plane(a.residual_type, a.normalize != 0, c, P, g);
      accumulate_row(a.prep[pr.ref_block], a.prep[pr.nei_block], q, P, g, J);
      cost = huber_correct(a.huber, r, J, 12);
    }
    if (a.out_valid) a.out_valid[qi] = valid ? 1 : 0;
    if (a.out_point && valid) {
#pragma unroll
      for (int k = 0; k < 3; ++k) a.out_point[(size_t)qi * 3 + k] = p_local[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) a.out_plane[(size_t)qi * 4 + k] = plane[k];
    }
    if (a.out_res) {
      a.out_res[qi] = r;
#pragma unroll
      for (int k = 0; k < 6; ++k) a.out_jac6[(size_t)qi * 6 + k] = J[6 + k];
    }
  }
  if (REDUCE) {
    // Per-WARP reduction (no block barrier: warps of a tile finish their searches at different times).  Rows are
    // staged in shared memory, then 28 lanes each own one entry of (H upper 21 | g 6 | cost); fixed summation
    // order => run-to-run identical results.  partials: [tile][warp][29].
    const int w = i >> 5, lane = i & 31;
    if (LIST) { __syncwarp(); sJ = reinterpret_cast<double (*)[8]>(&s_list[w][0][0]) - w * 32; }   // every lane of the warp is done with its list: sJ[i] = row (i & 31) of the warp's region
#pragma unroll
    for (int k = 0; k < 6; ++k) sJ[i][k] = valid ? J[6 + k] : 0.0;
    sJ[i][6] = valid ? r : 0.0;
    sJ[i][7] = valid ? cost : 0.0;
    const unsigned b = __ballot_sync(0xffffffffu, valid);
    __syncwarp();
    double* out = a.partials + ((size_t)blockIdx.x * (kTile / 32) + w) * 29;
    if (lane < 28) {
      int ia = 0, ib = 0; double acc = 0.0;
      if (lane < 21) { int o = lane; ia = 0; while (o >= 6 - ia) { o -= 6 - ia; ++ia; } ib = ia + o; }
      else if (lane < 27) { ia = lane - 21; ib = 6; }
      const double (*rows)[8] = &sJ[w * 32];
      if (lane < 27) { for (int row = 0; row < 32; ++row) acc += rows[row][ia] * rows[row][ib]; }
      else { for (int row = 0; row < 32; ++row) acc += rows[row][7]; }
      out[lane] = acc;
    } else if (lane == 28) {
      out[28] = (double)__popc(b);
    }
  }
}

// ---- A3: point-to-line association (5-NN + PCA line) --------------------------------------------------------------------
struct LineAssocArgs {
  const F4* q_local; const QueryTile* tiles; const Pair* pairs; const GridDesc* grids; const uint32_t* cell_start; const F4* sorted;
  const WorldPose* wpose; float sq_thr; double thr;
  unsigned char* out_valid; double* out_point; double* out_a; double* out_b;     // indexed by tile.out_base + lane
  int* out_nn;                                                                   // when set: neighbour indices only (K per slot, -1 = rejected), no line fit
};
template <int K>
__global__ void __launch_bounds__(kTile) k_associate_line(const LineAssocArgs a) {
  __shared__ uint32_t s_win[K][kTile];
  __shared__ uint32_t s_rng[18][kTile];
  const QueryTile t = a.tiles[blockIdx.x];
  const Pair pr = a.pairs[t.pair];
  const int i = threadIdx.x;
  if (i >= t.count) return;
  const int gq = t.start + i;
  const uint32_t slot = (uint32_t)(t.out_base + i);
  const F4 q = ldg_f4(a.q_local + gq);
  const WorldPose& wn = a.wpose[pr.nei_block];
  const WorldPose& wr = a.wpose[pr.ref_block];
  float qx, qy, qz;
  transform_point_f32(wn.R, wn.t, q.x, q.y, q.z, qx, qy, qz);
  const GridDesc& g = a.grids[pr.target_cloud];
  const uint32_t* cs = a.cell_start + g.cell_base;
  const F4* srt = a.sorted;
  auto cells = [cs](long long c) { return (long long)__ldg(cs + c); };
  auto load = [srt](long long p) { return ldg_f4(srt + p); };
  auto win = [&](int j) { return s_win[j][i]; };
  auto set_win = [&](int j, uint32_t pos) { s_win[j][i] = pos; };
  auto range_set = [&](int k, uint32_t lo, uint32_t hi) { s_rng[2 * k][i] = lo; s_rng[2 * k + 1][i] = hi; };
  auto range_get = [&](int k, uint32_t& lo, uint32_t& hi) { lo = s_rng[2 * k][i]; hi = s_rng[2 * k + 1][i]; };
  if (a.out_nn) {
    int idx[K];
    const bool ok = knn_indices<K>(g, cells, load, a.sq_thr, (int)ceil(a.thr / g.h), qx, qy, qz, idx, win, set_win, range_set, range_get);
#pragma unroll
    for (int j = 0; j < K; ++j) a.out_nn[(size_t)slot * K + j] = ok ? idx[j] : -1;
    return;
  }
  double p_local[3], pa[3], pb[3];
  const bool valid = associate_point2line<K>(g, cells, load, a.sq_thr, (int)ceil(a.thr / g.h), qx, qy, qz, wr.R, wr.t, wn.R, wn.t, p_local, pa, pb, win, set_win,
                                             range_set, range_get);
  a.out_valid[slot] = valid ? 1 : 0;
  if (valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { a.out_point[(size_t)slot * 3 + k] = p_local[k]; a.out_a[(size_t)slot * 3 + k] = pa[k]; a.out_b[(size_t)slot * 3 + k] = pb[k]; }
  }
}

// ---- K3: residual + analytic Jacobian over a correspondence list sorted by pose-graph edge --------------------------
struct BlockTile { int edge; int start; int count; int pad; };
struct EvalArgs {
  const BlockTile* tiles;
  const int* edge_ref; const int* edge_nei;
  const int* type; const int* normalize; const double* huber;
  const double* consts;        // SoA: consts[k * n + row]
  const uint32_t* orig;        // original block index of each sorted row
  long long n;
  const PosePrep* prep;
  double* out_r; double* out_J;   // original order (may be null)
  double* partials;               // [tile][92] (may be null)
  int raw_rows;                   // 1: out_r / out_J hold the rows BEFORE the robust loss (the caller registers the loss with its solver); partials are always corrected
};

__global__ void __launch_bounds__(kTile, PVB_K3_MINB) k_eval_blocks(const EvalArgs a) {
  __shared__ double sJ[kTile][14];   // J12 | r | cost  (stride 14 doubles)
  __shared__ PosePrep sPrep[2];      // the tile's two pose blocks (all rows of a tile belong to one edge)
  const BlockTile t = a.tiles[blockIdx.x];
  const int i = threadIdx.x;
  const bool act = i < t.count;
  {
    constexpr int kD = (int)(sizeof(PosePrep) / sizeof(double));
    if (i < 2 * kD) {
      const int which = i / kD, k = i - which * kD;
      const int blk = which == 0 ? a.edge_ref[t.edge] : a.edge_nei[t.edge];
      reinterpret_cast<double*>(&sPrep[which])[k] = __ldg(reinterpret_cast<const double*>(a.prep + blk) + k);
    }
    __syncthreads();
  }
  double J[12], r = 0.0, cost = 0.0;
#pragma unroll
  for (int k = 0; k < 12; ++k) J[k] = 0.0;
  if (act) {
    const long long row = (long long)t.start + i;
    double c[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) c[k] = __ldg(a.consts + (size_t)k * a.n + row);
    r = eval_block(a.type[row], a.normalize[row] != 0, c, sPrep[0], sPrep[1], J);
    if (!a.raw_rows) cost = huber_correct(a.huber[row], r, J, 12);
    if (a.out_r) {
      const uint32_t o = a.orig[row];
      a.out_r[o] = r;
      double2* dst = reinterpret_cast<double2*>(a.out_J + (size_t)o * 12);   // 96-byte rows, 16-byte aligned: 128-bit stores
#pragma unroll
      for (int k = 0; k < 6; ++k) dst[k] = make_double2(J[2 * k], J[2 * k + 1]);
    }
    if (a.raw_rows) cost = huber_correct(a.huber[row], r, J, 12);
  }
  if (a.partials) {
    // every warp sums its own 32 rows (3 of the 92 entries per lane: three independent accumulation chains of 32 instead of one of 128 per thread, and
    // no idle warp), then 92 threads add the four warp sums in warp order: the order of every sum is fixed => run-to-run identical results
    __shared__ double sW[kTile / 32][92];
#pragma unroll
    for (int k = 0; k < 12; ++k) sJ[i][k] = J[k];
    sJ[i][12] = r; sJ[i][13] = cost;
    __syncwarp();
    const int w = i >> 5, lane = i & 31;
    const double (*rows)[14] = &sJ[w * 32];
    int ia[3], ib[3];
#pragma unroll
    for (int e3 = 0; e3 < 3; ++e3) {
      const int e = lane + 32 * e3;
      ia[e3] = 13; ib[e3] = 13;                                  // e >= 90: the cost column (entry 90; entries 91.. are not stored)
      if (e < 78) { int o = e, a0 = 0; while (o >= 12 - a0) { o -= 12 - a0; ++a0; } ia[e3] = a0; ib[e3] = a0 + o; }
      else if (e < 90) { ia[e3] = e - 78; ib[e3] = 12; }
    }
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll 4
    for (int row = 0; row < 32; ++row) {
#pragma unroll
      for (int e3 = 0; e3 < 3; ++e3) {
        const double va = rows[row][ia[e3]], vb = lane + 32 * e3 < 90 ? rows[row][ib[e3]] : 1.0;
        acc[e3] += va * vb;
      }
    }
#pragma unroll
    for (int e3 = 0; e3 < 3; ++e3) if (lane + 32 * e3 < 91) sW[w][lane + 32 * e3] = acc[e3];
    __syncthreads();
    if (i < 91) a.partials[(size_t)blockIdx.x * 92 + i] = ((sW[0][i] + sW[1][i]) + sW[2][i]) + sW[3][i];
    else if (i == 91) a.partials[(size_t)blockIdx.x * 92 + 91] = (double)t.count;
  }
}

// ---- B1 on the device: the accepted correspondences of an association become residual blocks without leaving HBM ------------------------
// flags (one byte per query slot) -> 32-bit flags for the prefix sum that compacts them
__global__ void __launch_bounds__(256) k_flags_to_u32(const unsigned char* __restrict__ valid, long long n, uint32_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) out[i] = i < n ? (valid[i] ? 1u : 0u) : 0u;          // one extra element: its exclusive sum is the total
}
__global__ void __launch_bounds__(256) k_gather_u32(const uint32_t* __restrict__ src, const int* __restrict__ idx, int n, uint32_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// slot s (valid) -> block row pos[s]: Point2Plane_Angle / _Meter with the reference's Huber width (util/Optimization.cpp:513-517, 541-557),
// constants SoA consts[k * n_total + row] = point (0..2) | plane (3..6) | weight (7)
__global__ void __launch_bounds__(256) k_blocks_from_point2plane(const unsigned char* __restrict__ valid, const uint32_t* __restrict__ pos, long long n_slots,
                                                                 const double* __restrict__ point, const double* __restrict__ plane, int type, int normalize, double huber,
                                                                 double weight, long long n_total, int* __restrict__ b_type, int* __restrict__ b_norm,
                                                                 double* __restrict__ b_huber, double* __restrict__ b_consts, uint32_t* __restrict__ b_orig) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots || !valid[s]) return;
  const long long r = pos[s];
  b_type[r] = type; b_norm[r] = normalize; b_huber[r] = huber; b_orig[r] = (uint32_t)r;
#pragma unroll
  for (int k = 0; k < 3; ++k) b_consts[(size_t)k * n_total + r] = point[s * 3 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k) b_consts[(size_t)(3 + k) * n_total + r] = plane[s * 4 + k];
  b_consts[(size_t)7 * n_total + r] = weight;
#pragma unroll
  for (int k = 8; k < 12; ++k) b_consts[(size_t)k * n_total + r] = 0.0;
}

// ---- K4: sum the per-tile partials of each edge / frame in tile order --------------------------------------------------
// Stage 1: block (group, chunk) sums its share of the group's partial rows (8 slices x 32 value lanes, slices added in
// order) into chunk_out[group][chunk][NV]; stage 2 (k_sum_chunks) adds the chunks in order.  Every order is fixed =>
// deterministic.  NV <= 96.
constexpr int kSumChunks = 16;
template <int NV>
__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partials, const int* __restrict__ tile_begin /*[n_groups+1]*/, double* __restrict__ chunk_out) {
  __shared__ double sm[8][96];
  const int gidx = blockIdx.x, chunk = blockIdx.y, lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int g0 = tile_begin[gidx], g1 = tile_begin[gidx + 1];
  const int per = (g1 - g0 + kSumChunks - 1) / kSumChunks;
  const int t0 = g0 + chunk * per, t1 = min(g1, t0 + per);
  for (int v = lane; v < NV; v += 32) {
    double acc = 0.0;
    for (int t = t0 + slice; t < t1; t += 8) acc += partials[(size_t)t * NV + v];
    sm[slice][v] = acc;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double acc = 0.0;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) acc += sm[sl][threadIdx.x];
    chunk_out[((size_t)gidx * kSumChunks + chunk) * NV + threadIdx.x] = acc;
  }
}
template <int NV>
__global__ void k_sum_chunks(const double* __restrict__ chunk_out, double* __restrict__ out) {
  const int gidx = blockIdx.x, v = threadIdx.x;
  if (v >= NV) return;
  double acc = 0.0;
#pragma unroll
  for (int c = 0; c < kSumChunks; ++c) acc += chunk_out[((size_t)gidx * kSumChunks + c) * NV + v];
  out[(size_t)gidx * NV + v] = acc;
}

// ---- K1: SE(3) + equirectangular projection (float32 FastAtan2 path) --------------------------------------------------
__global__ void __launch_bounds__(256) k_project(const F4* __restrict__ pts, long long n, WorldPose T, int rows, int cols, float* __restrict__ uvd,
                                                 unsigned long long* __restrict__ splat, int size) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const F4 p = ldg_f4(pts + i);
  float x, y, z, u, v;
  transform_point_f32(T.R, T.t, p.x, p.y, p.z, x, y, z);
  cam_to_image_f32(x, y, z, rows, cols, u, v);
  const float depth = sqrtf(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
  if (uvd) { uvd[i * 3] = u; uvd[i * 3 + 1] = v; uvd[i * 3 + 2] = depth; }
  if (splat) {
    const int h = size / 2;
    const int rbx = (int)ceilf(u) + h, rby = (int)ceilf(v) + h, ltx = (int)floorf(u) - h, lty = (int)floorf(v) - h;
    const bool in_rb = rbx >= 0 && rby >= 0 && rbx + 1 <= cols && rby + 1 <= rows;
    const bool in_lt = ltx >= 0 && lty >= 0 && ltx + 1 <= cols && lty + 1 <= rows;
    if (in_rb && in_lt) {
      const uint32_t rel = (uint32_t)((int)((double)depth * 256.0)) & 0xFFFFu;
      const unsigned long long key = ((unsigned long long)(i + 1) << 16) | rel;     // "last writer wins" == largest point index
      for (int yy = lty; yy <= rby; ++yy) for (int xx = ltx; xx <= rbx; ++xx) atomicMax(splat + (size_t)yy * cols + xx, key);
    }
  }
}
__global__ void __launch_bounds__(256) k_splat_finalize(const unsigned long long* __restrict__ splat, long long n, uint16_t* __restrict__ img) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) img[i] = (uint16_t)(splat[i] & 0xFFFFull);
}

// ---- non-fused double helpers for threshold-critical vote kernels (bit-identical to the -ffp-contract=off CPU) ----------
__device__ __forceinline__ double dot3_nf(const double* a, const double* b) { return dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2])); }
__device__ __forceinline__ double vector_angle_nf(const double* a, const double* b) {   // Geometry.hpp:450-466
  double c = dot3_nf(a, b);
  const double n1 = sqrt(dot3_nf(a, a)), n2 = sqrt(dot3_nf(b, b));
  c = c / dmul(n1, n2);
  if (c >= 1.0) return 0.0;
  if (c <= -1.0) return M_PI;
  return acos(c);
}

// ---- K2l: line-to-line vote matrix ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_line_votes(const double* __restrict__ ref_lines, int S_ref, const F4* __restrict__ pts, int n_pts,
                                                    const int* __restrict__ p2s_off, const int* __restrict__ p2s_ids, double thr, int* __restrict__ M) {
  extern __shared__ double s_lines[];
  for (int k = threadIdx.x; k < S_ref * 6; k += blockDim.x) s_lines[k] = ref_lines[k];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const F4 p = ldg_f4(pts + i);
  const double P[3] = {(double)p.x, (double)p.y, (double)p.z};
  const int e0 = p2s_off[i], e1 = p2s_off[i + 1];
  if (e0 == e1) return;
  for (int s = 0; s < S_ref; ++s) {
    const double* l = s_lines + s * 6;   // PointToLineDistance3D (Geometry.hpp:198-211)
    const double d0 = dsub(P[0], l[0]), d1 = dsub(P[1], l[1]), d2 = dsub(P[2], l[2]);
    const double k = dadd(dadd(dmul(l[3], d0), dmul(l[4], d1)), dmul(l[5], d2)) / dadd(dadd(dmul(l[3], l[3]), dmul(l[4], l[4])), dmul(l[5], l[5]));
    const double e[3] = {dsub(dadd(dmul(k, l[3]), l[0]), P[0]), dsub(dadd(dmul(k, l[4]), l[1]), P[1]), dsub(dadd(dmul(k, l[5]), l[2]), P[2])};
    const double dist = sqrt(dadd(dadd(dmul(e[0], e[0]), dmul(e[1], e[1])), dmul(e[2], e[2])));
    if (dist > thr) continue;
    for (int q = e0; q < e1; ++q) atomicAdd(M + (size_t)p2s_ids[q] * S_ref + s, 1);
  }
}

// ---- K2l batched: the vote matrices of ALL frame pairs of a pose graph in one launch ----------------------------------------------------------
// Same arithmetic per (point, line) as k_line_votes.  Inputs are the concatenation over frames of the world-frame corner clouds, their point -> segment
// CSR lists and the world-frame segment lines; a tile = <= 128 consecutive corner points of the NEIGHBOUR frame of one pair.
struct VotePair { int ref, nei; long long m_off; };                  // frames of the pair, offset of its S_nei x S_ref matrix in M
struct VoteTile { int pair, start, count, pad; };                    // corner points [start, start + count) of the pair's neighbour frame (frame-relative)
__global__ void __launch_bounds__(128) k_line_votes_batch(const VoteTile* __restrict__ tiles, const VotePair* __restrict__ pairs, const F4* __restrict__ world,
                                                          const int* __restrict__ corner_off /*[n_frames+1]*/, const int* __restrict__ p2s_off /*[sum corners + n_frames]: per frame n+1 entries*/,
                                                          const int* __restrict__ p2s_base /*[n_frames]: offset of the frame's ids in p2s_ids*/, const int* __restrict__ p2s_ids,
                                                          const double* __restrict__ lines /*[sum segments][6]*/, const int* __restrict__ seg_off /*[n_frames+1]*/, double thr,
                                                          int* __restrict__ M) {
  extern __shared__ double s_lines[];
  const VoteTile t = tiles[blockIdx.x];
  const VotePair pr = pairs[t.pair];
  const int S_ref = seg_off[pr.ref + 1] - seg_off[pr.ref];
  const double* rl = lines + (size_t)seg_off[pr.ref] * 6;
  for (int k = threadIdx.x; k < S_ref * 6; k += blockDim.x) s_lines[k] = rl[k];
  __syncthreads();
  if ((int)threadIdx.x >= t.count) return;
  const int li = t.start + threadIdx.x;                              // point index inside the neighbour frame
  const int gi = corner_off[pr.nei] + li;
  const int* po = p2s_off + corner_off[pr.nei] + pr.nei;             // the frame's n + 1 CSR offsets
  const int e0 = po[li], e1 = po[li + 1];
  if (e0 == e1) return;
  const F4 p = ldg_f4(world + gi);
  const double P[3] = {(double)p.x, (double)p.y, (double)p.z};
  const int* ids = p2s_ids + p2s_base[pr.nei];
  int* Mp = M + pr.m_off;
  for (int s = 0; s < S_ref; ++s) {
    const double* l = s_lines + s * 6;   // PointToLineDistance3D (Geometry.hpp:198-211)
    const double d0 = dsub(P[0], l[0]), d1 = dsub(P[1], l[1]), d2 = dsub(P[2], l[2]);
    const double k = dadd(dadd(dmul(l[3], d0), dmul(l[4], d1)), dmul(l[5], d2)) / dadd(dadd(dmul(l[3], l[3]), dmul(l[4], l[4])), dmul(l[5], l[5]));
    const double e[3] = {dsub(dadd(dmul(k, l[3]), l[0]), P[0]), dsub(dadd(dmul(k, l[4]), l[1]), P[1]), dsub(dadd(dmul(k, l[5]), l[2]), P[2])};
    const double dist = sqrt(dadd(dadd(dmul(e[0], e[0]), dmul(e[1], e[1])), dmul(e[2], e[2])));
    if (dist > thr) continue;
    for (int q = e0; q < e1; ++q) atomicAdd(Mp + (size_t)ids[q] * S_ref + s, 1);
  }
}

// ---- AssociatePoint2LineSegment (LidarFeatureAssociate.cpp:319-383): nearest infinite line per point, first minimum wins --------
__global__ void __launch_bounds__(128) k_nearest_line(const double* __restrict__ ref_lines, int S_ref, const F4* __restrict__ pts, int n_pts,
                                                      int* __restrict__ out_line, double* __restrict__ out_dist) {
  extern __shared__ double s_lines[];
  for (int k = threadIdx.x; k < S_ref * 6; k += blockDim.x) s_lines[k] = ref_lines[k];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  const F4 p = ldg_f4(pts + i);
  const double P[3] = {(double)p.x, (double)p.y, (double)p.z};
  double best = DBL_MAX; int arg = -1;
  for (int s = 0; s < S_ref; ++s) {
    const double* l = s_lines + s * 6;   // PointToLineDistance3D (Geometry.hpp:198-211)
    const double d0 = dsub(P[0], l[0]), d1 = dsub(P[1], l[1]), d2 = dsub(P[2], l[2]);
    const double k = dadd(dadd(dmul(l[3], d0), dmul(l[4], d1)), dmul(l[5], d2)) / dadd(dadd(dmul(l[3], l[3]), dmul(l[4], l[4])), dmul(l[5], l[5]));
    const double e[3] = {dsub(dadd(dmul(k, l[3]), l[0]), P[0]), dsub(dadd(dmul(k, l[4]), l[1]), P[1]), dsub(dadd(dmul(k, l[5]), l[2]), P[2])};
    const double dist = sqrt(dadd(dadd(dmul(e[0], e[0]), dmul(e[1], e[1])), dmul(e[2], e[2])));
    if (dist < best) { best = dist; arg = s; }
  }
  out_line[i] = arg; out_dist[i] = best;
}

// ---- K2c: camera-LiDAR AssociateByAngle vote counts ------------------------------------------------------------------------
struct ImageLinePlane { double n[4]; double p4[3]; double scope; };   // unit plane through the origin, arc midpoint, half-arc angle
__global__ void __launch_bounds__(128) k_angle_votes(const ImageLinePlane* __restrict__ lines, int L, const F4* __restrict__ cloud_local, int P, WorldPose T,
                                                     const int* __restrict__ p2s_off, const int* __restrict__ p2s_ids, int S, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (i >= P) return;
  const int e0 = p2s_off[i], e1 = p2s_off[i + 1];
  if (e0 == e1) return;
  const F4 pl = ldg_f4(cloud_local + i);
  const float range = fadd(fadd(fmul(pl.x, pl.x), fmul(pl.y, pl.y)), fmul(pl.z, pl.z));
  if (range > 15 * 15) return;                                           // CameraLidarLineAssociate.cpp:395
  float x, y, z;
  transform_point_f32(T.R, T.t, pl.x, pl.y, pl.z, x, y, z);
  const double p[3] = {(double)x, (double)y, (double)z};
  const ImageLinePlane ln = lines[l];
  // ProjectPointToPlane(p, plane, normalized = true)  (Geometry.hpp:301-316)
  const double s = dadd(dadd(dadd(dmul(ln.n[0], p[0]), dmul(ln.n[1], p[1])), dmul(ln.n[2], p[2])), ln.n[3]);
  const double dis = fabs(s);
  double pp[3] = {dsub(p[0], dmul(dis, ln.n[0])), dsub(p[1], dmul(dis, ln.n[1])), dsub(p[2], dmul(dis, ln.n[2]))};
  if (fabs(dadd(dadd(dadd(dmul(ln.n[0], pp[0]), dmul(ln.n[1], pp[1])), dmul(ln.n[2], pp[2])), ln.n[3])) > 1e-4) {
    pp[0] = dadd(p[0], dmul(dis, ln.n[0])); pp[1] = dadd(p[1], dmul(dis, ln.n[1])); pp[2] = dadd(p[2], dmul(dis, ln.n[2]));
  }
  const double thr = 3.0 / 180.0 * M_PI;
  if (vector_angle_nf(p, pp) >= thr) return;                             // :402
  if (vector_angle_nf(ln.p4, pp) >= ln.scope + thr) return;              // :405
  for (int q = e0; q < e1; ++q) atomicAdd(counts + (size_t)l * S + p2s_ids[q], 1);
}

#endif  // __CUDACC__
}  // namespace pvb
