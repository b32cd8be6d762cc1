// panovlm_b200 — exact k-nearest-neighbour association on a cell-sorted target (per-query logic).
//
// Replaces pcl::KdTreeFLANN::nearestKSearch + the per-query body of AssociatePoint2Plane
// (lidar_mapping/LidarFeatureAssociate.cpp:570-599): k nearest on float32 squared L2, k-th <= thr^2,
// same-class test, neighbours -> reference sensor frame, LSQ plane + tolerance, collinearity reject.
//
// Data layout: the target cloud (world frame, float32) is bucketed into a uniform grid (cell size h, x fastest)
// and stored sorted by cell as 16-byte records {x, y, z, (orig_idx << 5) | class}; cell_start[c] is the offset
// of the first record of cell c.  All cells of one (y,z) row that a query needs are one contiguous range.
// Exactness: ring r covers every point within r*h of the query (clamping to the grid box is monotone and
// 1-Lipschitz per axis); the search stops when the k-th best is closer than r*h or r*h >= thr.
#pragma once
#include "pvb_math.cuh"

namespace pvb {

struct alignas(16) F4 { float x, y, z, w; };

struct GridDesc {
  double origin[3];
  double inv_h, h;
  int dims[3];            // nx, ny, nz
  int n_points;
  long long cell_base;    // offset of this cloud's cells in the global cell_start array (ncells + 1 entries)
  long long point_base;   // offset of this cloud's records in the sorted array
};

PVB_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PVB_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

PVB_HD int cell_coord(double v, double origin, double inv_h, int n) {
  const double f = floor((v - origin) * inv_h);
  int c = f < 0.0 ? 0 : (f > (double)(n - 1) ? n - 1 : (int)f);
  return c;
}

constexpr unsigned long long kKeyEmptyLow = 0xFFFFFFFFull;

template <int K>
PVB_HD void topk_insert(unsigned long long (&keys)[K], unsigned long long key) {
  unsigned long long k = key;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const unsigned long long a = keys[j];
    const bool lt = k < a;
    keys[j] = lt ? k : a;
    k = lt ? a : k;
  }
}

// scan one contiguous record range, inserting candidates (d2 <= thr^2 is encoded in the initial keys)
template <int K, typename PointLoader>
PVB_HD void scan_range(const PointLoader& load, long long lo, long long hi, float qx, float qy, float qz, unsigned long long (&keys)[K]) {
  for (long long i = lo; i < hi; ++i) {
    const F4 p = load(i);
    const float d2 = sqdist_f32(qx, qy, qz, p.x, p.y, p.z);
    const unsigned long long key = ((unsigned long long)f2u(d2) << 32) | (unsigned long long)(uint32_t)i;
    if (key < keys[K - 1]) topk_insert<K>(keys, key);
  }
}

// Exact K-NN of (qx,qy,qz) within sqrt(sq_thr).  keys[] come back ascending by (d2, record position);
// returns the number of valid neighbours (== K when the K-th is within the threshold).
// `pos` in the keys is relative to g.point_base.  CellLoader: cell_start value; PointLoader: record.
template <int K, typename CellLoader, typename PointLoader>
PVB_HD int knn_search(const GridDesc& g, const CellLoader& cells, const PointLoader& load, float qx, float qy, float qz, float sq_thr, int rmax,
                      unsigned long long (&keys)[K]) {
  const unsigned long long init = ((unsigned long long)f2u(sq_thr) << 32) | kKeyEmptyLow;
#pragma unroll
  for (int j = 0; j < K; ++j) keys[j] = init;
  const int cx = cell_coord((double)qx, g.origin[0], g.inv_h, g.dims[0]);
  const int cy = cell_coord((double)qy, g.origin[1], g.inv_h, g.dims[1]);
  const int cz = cell_coord((double)qz, g.origin[2], g.inv_h, g.dims[2]);
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  for (int r = 1; r <= rmax; ++r) {
    const int z0 = cz - r < 0 ? 0 : cz - r, z1 = cz + r > nz - 1 ? nz - 1 : cz + r;
    const int y0 = cy - r < 0 ? 0 : cy - r, y1 = cy + r > ny - 1 ? ny - 1 : cy + r;
    const int x0 = cx - r < 0 ? 0 : cx - r, x1 = cx + r > nx - 1 ? nx - 1 : cx + r;
    for (int z = z0; z <= z1; ++z) {
      const bool zshell = (z == cz - r) || (z == cz + r);
      for (int y = y0; y <= y1; ++y) {
        const long long row = ((long long)z * ny + y) * nx;
        const bool full = (r == 1) || zshell || (y == cy - r) || (y == cy + r);
        if (full) {
          scan_range<K>(load, cells(row + x0), cells(row + x1 + 1), qx, qy, qz, keys);
        } else {
          if (cx - r >= 0) scan_range<K>(load, cells(row + cx - r), cells(row + cx - r + 1), qx, qy, qz, keys);
          if (cx + r <= nx - 1) scan_range<K>(load, cells(row + cx + r), cells(row + cx + r + 1), qx, qy, qz, keys);
        }
      }
    }
    // all points closer than r*h have been seen
    if ((keys[K - 1] & kKeyEmptyLow) != kKeyEmptyLow) {
      const double reach = (double)r * g.h;
      if ((double)u2f((uint32_t)(keys[K - 1] >> 32)) < reach * reach * (1.0 - 1e-6)) break;
    }
  }
  int n = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) n += ((keys[j] & kKeyEmptyLow) != kKeyEmptyLow) ? 1 : 0;
  return n;
}

struct AssocParams {
  float sq_thr;           // point_to_plane_dis_threshold^2 computed in float (LidarFeatureAssociate.cpp:557)
  int rmax;               // ceil(thr / h)
  double plane_tol;       // lidar_plane_tolerance
  double collinear_tol;   // 3.0 (LidarFeatureAssociate.cpp:594)
};

// Per-query body of AssociatePoint2Plane.  qw = query in world (float32), qcls = class label.
// R_ref/t_ref, R_nei/t_nei = R_wl, t_wl of the two frames.  On success: p_local (query in the neighbour's
// sensor frame, double) and plane (n, d) in the reference sensor frame.
template <int K, typename CellLoader, typename PointLoader>
PVB_HD bool associate_point2plane(const GridDesc& g, const CellLoader& cells, const PointLoader& load, const AssocParams& prm,
                                  float qx, float qy, float qz, uint32_t qcls,
                                  const double* R_ref, const double* t_ref, const double* R_nei, const double* t_nei,
                                  double p_local[3], double plane[4], unsigned long long (&keys)[K]) {
  const int found = knn_search<K>(g, cells, load, qx, qy, qz, prm.sq_thr, prm.rmax, keys);
  if (found < K) return false;                                   // :578 (k-th beyond the threshold) + quirk C.6 guard
  double pts[K][3];
  int same = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const F4 p = load((long long)(uint32_t)(keys[j] & kKeyEmptyLow));
    same += ((f2u(p.w) & 31u) == qcls) ? 1 : 0;                  // :586 pt.intensity == point.intensity
    const double pw[3] = {(double)p.x, (double)p.y, (double)p.z};
    world2local(R_ref, t_ref, pw, pts[j]);                       // :587
  }
  if (same < K) return false;                                    // :590
  if (!form_plane_lsq<K>(pts, prm.plane_tol, plane)) return false;   // :593
  if (points_collinear<K>(pts, prm.collinear_tol)) return false;     // :594-596
  const double qw[3] = {(double)qx, (double)qy, (double)qz};
  world2local(R_nei, t_nei, qw, p_local);                        // :598-599
  return true;
}

}  // namespace pvb
