// panovlm_b200 — exact k-nearest-neighbour association on a cell-sorted target (per-query logic).
//
// Replaces pcl::KdTreeFLANN::nearestKSearch + the per-query body of AssociatePoint2Plane
// (lidar_mapping/LidarFeatureAssociate.cpp:570-599): k nearest on float32 squared L2, k-th <= thr^2,
// same-class test, neighbours -> reference sensor frame, LSQ plane + tolerance, collinearity reject.
//
// (Staging: a thread block may copy the rows its queries need into shared memory with TMA bulk copies; ring-1 ranges are
// then translated into that staging area by `row_map` — see k_associate in pvb_kernels.cuh.)
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
struct alignas(8) U2 { uint32_t x, y; };          // candidate-list entry of the hinted search: (d2 bits, record position)

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

// ---- exact k-NN selection, two passes over the candidate cells ------------------------------------------------
// Pass 1 keeps only the K smallest squared distances: float32 bit patterns of non-negative floats order like
// unsigned integers, so the sorted list is maintained with a branch-free min/max chain (2 ALU ops per slot, no
// payload, no divergence).  Pass 2 re-scans the same cells and hands every candidate that belongs to the K smallest
// (d2 < tau, plus as many d2 == tau as needed, in scan order) to `sink(j, position)`.
template <int K>
PVB_HD void topk_values_insert(uint32_t (&keys)[K], uint32_t key) {
  // sorted ascending; inserting `key` and dropping the largest: new[j] = median(old[j-1], key, old[j])
  //                                                                   = min(old[j], max(old[j-1], key))
  // every slot is computed from the OLD values only: 2 independent min/max per slot, dependency depth 2.
  uint32_t prev = 0u;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const uint32_t a = keys[j];
    const uint32_t m = prev > key ? prev : key;
    keys[j] = a < m ? a : m;
    prev = a;
  }
}

template <int K, typename PointLoader>
PVB_HD void scan_values(const PointLoader& load, long long lo, long long hi, float qx, float qy, float qz, uint32_t (&keys)[K]) {
#pragma unroll 2
  for (long long i = lo; i < hi; ++i) {
    const F4 c = load(i);
    topk_values_insert<K>(keys, f2u(sqdist_f32(qx, qy, qz, c.x, c.y, c.z)));
  }
}

template <typename PointLoader, typename Sink>
PVB_HD void scan_collect(const PointLoader& load, long long lo, long long hi, float qx, float qy, float qz, uint32_t tau, int eq_needed, int& eq_taken, int& n_out,
                         const Sink& sink) {
#pragma unroll 2
  for (long long i = lo; i < hi; ++i) {
    const F4 c = load(i);
    const uint32_t kb = f2u(sqdist_f32(qx, qy, qz, c.x, c.y, c.z));
    bool take = kb < tau;
    if (kb == tau && eq_taken < eq_needed) { take = true; ++eq_taken; }
    if (take) { sink(n_out, (uint32_t)i, kb); ++n_out; }
  }
}

// visits the cells of ring `r` (r == 0: the whole (2R+1)^3 block of radius R) row by row; f(range_lo, range_hi)
template <typename CellLoader, typename F>
PVB_HD void for_each_range(const GridDesc& g, const CellLoader& cells, int cx, int cy, int cz, int r, bool whole_block, const F& f) {
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int z0 = cz - r < 0 ? 0 : cz - r, z1 = cz + r > nz - 1 ? nz - 1 : cz + r;
  const int y0 = cy - r < 0 ? 0 : cy - r, y1 = cy + r > ny - 1 ? ny - 1 : cy + r;
  const int x0 = cx - r < 0 ? 0 : cx - r, x1 = cx + r > nx - 1 ? nx - 1 : cx + r;
  for (int z = z0; z <= z1; ++z) {
    const bool zshell = (z == cz - r) || (z == cz + r);
    for (int y = y0; y <= y1; ++y) {
      const long long row = ((long long)z * ny + y) * nx;
      const bool full = whole_block || (r == 1) || zshell || (y == cy - r) || (y == cy + r);
      // one call site of f (keeps the inlined scan loop single: instruction-cache footprint)
      const int nparts = full ? 1 : 2;
      for (int part = 0; part < nparts; ++part) {
        long long c0, c1;
        if (full) { c0 = row + x0; c1 = row + x1 + 1; }
        else if (part == 0) { if (cx - r < 0) continue; c0 = row + cx - r; c1 = c0 + 1; }
        else { if (cx + r > nx - 1) continue; c0 = row + cx + r; c1 = c0 + 1; }
        f(cells(c0), cells(c1));
      }
    }
  }
}

// Walk over a list of record ranges (per-row loops).
template <typename RangeGet, typename Body>
PVB_HD void walk_ranges_nested(int n_ranges, const RangeGet& range, const Body& body) {
  for (int row = 0; row < n_ranges; ++row) {
    uint32_t lo, hi;
    range(row, lo, hi);
#pragma unroll 2
    for (uint32_t i = lo; i < hi; ++i) body((long long)i);
  }
}

// Exact K-NN of (qx,qy,qz) within sqrt(sq_thr).  Returns the number of neighbours handed to sink (K, or 0 when the
// K-th nearest is beyond the threshold / fewer than K points are in reach).  sink(j, record position, d2 bits).
//  * ring 1 (the 3x3x3 cell block = up to 9 contiguous row ranges) is looked up once and walked twice through `load1`;
//    row_map(y, z, lo, hi) may translate a row range into another index space (the tile's shared-memory staging area);
//  * wider rings (rare) go through `loadg` with positions in the global sorted array.
//  ring_out tells the caller which space the neighbour positions are in (1: load1's, > 1: loadg's).
//  range_set(idx, lo, hi) / range_get(idx, lo&, hi&): caller-provided storage for the <= 9 row ranges.
template <int K, typename CellLoader, typename Load1, typename LoadG, typename RowMap, typename Sink, typename RangeSet, typename RangeGet>
PVB_HD int knn_select(const GridDesc& g, const CellLoader& cells, const Load1& load1, const LoadG& loadg, const RowMap& row_map, float qx, float qy, float qz,
                      float sq_thr, int rmax, const Sink& sink, const RangeSet& range_set, const RangeGet& range_get, int& ring_out) {
  const uint32_t init = f2u(sq_thr) + 1u;          // every d2 <= sq_thr is below it
  uint32_t keys[K];
#pragma unroll
  for (int j = 0; j < K; ++j) keys[j] = init;
  const double fx = ((double)qx - g.origin[0]) * g.inv_h, fy = ((double)qy - g.origin[1]) * g.inv_h, fz = ((double)qz - g.origin[2]) * g.inv_h;
  const int cx = cell_coord((double)qx, g.origin[0], g.inv_h, g.dims[0]);
  const int cy = cell_coord((double)qy, g.origin[1], g.inv_h, g.dims[1]);
  const int cz = cell_coord((double)qz, g.origin[2], g.inv_h, g.dims[2]);
  // distance from the query to the nearest face of its own cell (0 when the query was clamped into the grid)
  double slack = 0.5;
  {
    const double f[3] = {fx - cx, fy - cy, fz - cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      double m = f[a] < 1.0 - f[a] ? f[a] : 1.0 - f[a];
      m = m < 0.0 ? 0.0 : m;
      slack = m < slack ? m : slack;
    }
  }
  // ---- ring 1
  int n_ranges = 0;
  {
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const int z0 = cz - 1 < 0 ? 0 : cz - 1, z1 = cz + 1 > nz - 1 ? nz - 1 : cz + 1;
    const int y0 = cy - 1 < 0 ? 0 : cy - 1, y1 = cy + 1 > ny - 1 ? ny - 1 : cy + 1;
    const int x0 = cx - 1 < 0 ? 0 : cx - 1, x1 = cx + 1 > nx - 1 ? nx - 1 : cx + 1;
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const long long row = ((long long)z * ny + y) * nx;
        uint32_t lo = (uint32_t)cells(row + x0), hi = (uint32_t)cells(row + x1 + 1);
        if (hi > lo) { row_map(y, z, lo, hi); range_set(n_ranges, lo, hi); ++n_ranges; }
      }
  }
  walk_ranges_nested(n_ranges, range_get, [&](long long i) {
    const F4 c = load1(i);
    topk_values_insert<K>(keys, f2u(sqdist_f32(qx, qy, qz, c.x, c.y, c.z)));
  });
  int r = 1;
  bool done = false;
  if (keys[K - 1] != init) {
    const double reach = (1.0 + slack) * g.h;
    done = (double)u2f(keys[K - 1]) < reach * reach * (1.0 - 1e-6);
  }
  if (!done) {        // rare: widen ring by ring (generic nested loops, global positions)
    for (r = 2; r <= rmax; ++r) {
      for_each_range(g, cells, cx, cy, cz, r, false, [&](long long lo, long long hi) { scan_values<K>(loadg, lo, hi, qx, qy, qz, keys); });
      if (keys[K - 1] != init) {      // ring r covers every point closer than (r + slack) * h
        const double reach = ((double)r + slack) * g.h;
        if ((double)u2f(keys[K - 1]) < reach * reach * (1.0 - 1e-6)) break;
      }
    }
    if (r > rmax) r = rmax;
  }
  ring_out = r;
  if (keys[K - 1] == init) return 0;
  const uint32_t tau = keys[K - 1];
  int n_lt = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) n_lt += keys[j] < tau ? 1 : 0;
  int eq_taken = 0, n_out = 0;
  const int eq_needed = K - n_lt;
  if (r == 1) {
    walk_ranges_nested(n_ranges, range_get, [&](long long i) {
      const F4 c = load1(i);
      const uint32_t kb = f2u(sqdist_f32(qx, qy, qz, c.x, c.y, c.z));
      bool take = kb < tau;
      if (kb == tau && eq_taken < eq_needed) { take = true; ++eq_taken; }
      if (take) { sink(n_out, (uint32_t)i, kb); ++n_out; }
    });
  } else {
    for_each_range(g, cells, cx, cy, cz, r, true, [&](long long lo, long long hi) { scan_collect(loadg, lo, hi, qx, qy, qz, tau, eq_needed, eq_taken, n_out, sink); });
  }
  return n_out;
}

// ---- pruned variant (default path): the 3x3x3 block is visited row by row, nearest rows first; a row - and, inside a row, each of
// the two outer cells - is skipped when even its nearest face is farther than the current K-th distance (pass 1) / than tau (pass 2).
// A skipped cell cannot hold one of the K nearest points (strict comparison with a 1e-5 relative safety margin on the bound for the
// float rounding of d2), so the selected set is the same as that of the exhaustive walk; on surface-like clouds about half of the
// candidates are never touched.  Row ranges are looked up when needed (no per-thread range storage).
// g2[a][0] / g2[a][1]: squared distance (float, already shrunk by the safety margin) from the query to the lower / upper face of its
// own cell along axis a = lower bound of the squared axis distance to any point of the neighbouring cell on that side.
struct FaceGaps { float lo[3], hi[3]; };

template <int K, typename CellLoader, typename F>
PVB_HD void walk_block_pruned(const GridDesc& g, const CellLoader& cells, int cx, int cy, int cz, const FaceGaps& fg, const double gap_lo[3], const double gap_hi[3], int R,
                              const uint32_t& limit_key, const F& body) {
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int W = 2 * R + 1;
  // |d| >= 2 (only with the 5x5x5 start block of finer grids): bound through the generic double path
  auto far2 = [&](int a, int d) { const double v = d < 0 ? gap_lo[a] + (double)(-d - 1) * g.h : gap_hi[a] + (double)(d - 1) * g.h; return (float)(v * v * (1.0 - 1e-5)); };
  auto bound2 = [&](int a, int d) { return d == 0 ? 0.0f : (d == -1 ? fg.lo[a] : (d == 1 ? fg.hi[a] : far2(a, d))); };
#pragma unroll 1
  for (int o = 0; o < W * W; ++o) {
    const int iz = o / W, iy = o - iz * W;
    const int dz = (iz & 1) ? -((iz + 1) >> 1) : (iz >> 1);      // 0, -1, +1, -2, +2: nearest rows first
    const int dy = (iy & 1) ? -((iy + 1) >> 1) : (iy >> 1);
    const int y = cy + dy, z = cz + dz;
    if (y < 0 || y >= ny || z < 0 || z >= nz) continue;
    const float b2 = bound2(1, dy) + bound2(2, dz);              // float sum of two margin-shrunk terms: still below the true bound
    if (f2u(b2) > limit_key) continue;
    int xa = cx, xb = cx;
    for (int d = 1; d <= R; ++d) {
      if (cx - d < 0 || f2u(b2 + bound2(0, -d)) > limit_key) break;
      xa = cx - d;
    }
    for (int d = 1; d <= R; ++d) {
      if (cx + d > nx - 1 || f2u(b2 + bound2(0, d)) > limit_key) break;
      xb = cx + d;
    }
    const long long row = ((long long)z * ny + y) * nx;
    const uint32_t lo = (uint32_t)cells(row + xa), hi = (uint32_t)cells(row + xb + 1);
#pragma unroll 2
    for (uint32_t i = lo; i < hi; ++i) body((long long)i);
  }
}

template <int K, typename CellLoader, typename Load, typename Sink>
PVB_HD int knn_select_pruned(const GridDesc& g, const CellLoader& cells, const Load& load, float qx, float qy, float qz, float sq_thr, int r0, int rmax, const Sink& sink,
                             uint32_t* tau_out = nullptr) {
  const uint32_t init = f2u(sq_thr) + 1u;          // every d2 <= sq_thr is below it
  uint32_t keys[K];
#pragma unroll
  for (int j = 0; j < K; ++j) keys[j] = init;
  const double fx = ((double)qx - g.origin[0]) * g.inv_h, fy = ((double)qy - g.origin[1]) * g.inv_h, fz = ((double)qz - g.origin[2]) * g.inv_h;
  const int cx = cell_coord((double)qx, g.origin[0], g.inv_h, g.dims[0]);
  const int cy = cell_coord((double)qy, g.origin[1], g.inv_h, g.dims[1]);
  const int cz = cell_coord((double)qz, g.origin[2], g.inv_h, g.dims[2]);
  // distances (m) from the query to the faces of its own cell: lower bounds for every point of the neighbouring cells on that side
  // (0 when the query was clamped into the grid on that side)
  double gap_lo[3], gap_hi[3], slack = 0.5;
  {
    const double f[3] = {fx - cx, fy - cy, fz - cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double lo = f[a] < 0.0 ? 0.0 : f[a], hi = 1.0 - f[a] < 0.0 ? 0.0 : 1.0 - f[a];
      gap_lo[a] = lo * g.h; gap_hi[a] = hi * g.h;
      const double m = lo < hi ? lo : hi;
      slack = m < slack ? m : slack;
    }
  }
  FaceGaps fg;
#pragma unroll
  for (int a = 0; a < 3; ++a) { fg.lo[a] = (float)(gap_lo[a] * gap_lo[a] * (1.0 - 1e-5)); fg.hi[a] = (float)(gap_hi[a] * gap_hi[a] * (1.0 - 1e-5)); }
  walk_block_pruned<K>(g, cells, cx, cy, cz, fg, gap_lo, gap_hi, r0, keys[K - 1], [&](long long i) {
    const F4 c = load(i);
    topk_values_insert<K>(keys, f2u(sqdist_f32(qx, qy, qz, c.x, c.y, c.z)));
  });
  int r = r0;
  bool done = false;
  if (keys[K - 1] != init) {
    const double reach = ((double)r0 + slack) * g.h;
    done = (double)u2f(keys[K - 1]) < reach * reach * (1.0 - 1e-6);
  }
  if (rmax < r0) rmax = r0;
  if (!done) {        // rare: widen ring by ring (generic nested loops)
    for (r = r0 + 1; r <= rmax; ++r) {
      for_each_range(g, cells, cx, cy, cz, r, false, [&](long long lo, long long hi) { scan_values<K>(load, lo, hi, qx, qy, qz, keys); });
      if (keys[K - 1] != init) {      // ring r covers every point closer than (r + slack) * h
        const double reach = ((double)r + slack) * g.h;
        if ((double)u2f(keys[K - 1]) < reach * reach * (1.0 - 1e-6)) break;
      }
    }
    if (r > rmax) r = rmax;
  }
  if (keys[K - 1] == init) return 0;
  const uint32_t tau = keys[K - 1];
  if (tau_out) *tau_out = tau;
  int n_lt = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) n_lt += keys[j] < tau ? 1 : 0;
  int eq_taken = 0, n_out = 0;
  const int eq_needed = K - n_lt;
  if (r == r0) {
    walk_block_pruned<K>(g, cells, cx, cy, cz, fg, gap_lo, gap_hi, r0, tau, [&](long long i) {
      const F4 c = load(i);
      const uint32_t kb = f2u(sqdist_f32(qx, qy, qz, c.x, c.y, c.z));
      bool take = kb < tau;
      if (kb == tau && eq_taken < eq_needed) { take = true; ++eq_taken; }
      if (take) { sink(n_out, (uint32_t)i, kb); ++n_out; }
    });
  } else {
    for_each_range(g, cells, cx, cy, cz, r, true, [&](long long lo, long long hi) { scan_collect(load, lo, hi, qx, qy, qz, tau, eq_needed, eq_taken, n_out, sink); });
  }
  return n_out;
}

// ---- hinted single-pass variant (default path since round 2) ---------------------------------------------------------------------
// Search-radius hint: `lim_hint` (0 = none) is an exclusive upper bound (d2 bit pattern) that the caller claims holds for the K-th
// squared distance, e.g. from the previous evaluation of the same query: K target points lay within sqrt(tau_old) of q_old, so they
// lie within sqrt(tau_old) + |q - q_old| of q.  With a bound known up front ONE walk over the block is enough: every candidate below
// the bound goes to a small per-query list (key = d2 bits, payload = record position; LC entries, caller-provided storage) with a
// branch-free body, rows and cells beyond the bound are never touched, and the list (typically K .. K+3 entries once the poses settle)
// is cut to its K smallest entries afterwards (value-only min/max network over the list + in-place compaction; nothing to do when it
// holds exactly K).  The hint only shortens the walk: when fewer than K candidates are found below it (stale hint) or more than LC
// (bound far too wide), or when there is no hint, the pruned two-pass search above runs instead, so the result never depends on the
// hint being right.  Ties at the K-th distance are resolved by visiting order like in the two-pass search (same row order, a later
// candidate must be strictly closer to displace an earlier one).
//   The K nearest end up in list slots 0..K-1 (entry e at lbase[e * lstride]).
template <int K, typename LKey, typename LMove>
PVB_HD uint32_t list_cut_to_k(int& n, const LKey& lkey, const LMove& lmove) {
  uint32_t keys[K];
#pragma unroll
  for (int j = 0; j < K; ++j) keys[j] = 0xFFFFFFFFu;
#pragma unroll 2
  for (int i = 0; i < n; ++i) topk_values_insert<K>(keys, lkey(i));
  const uint32_t tau = keys[K - 1];
  int n_lt = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) n_lt += keys[j] < tau ? 1 : 0;
  const int eq_needed = K - n_lt;
  int eq_taken = 0, out = 0;
  for (int i = 0; i < n; ++i) {
    const uint32_t kb = lkey(i);
    bool take = kb < tau;
    if (kb == tau && eq_taken < eq_needed) { take = true; ++eq_taken; }
    if (take) { if (out != i) lmove(out, i); ++out; }
  }
  n = out;
  return tau;
}

// One walk over the (2 rb + 1)^3 block in the row order of walk_block_pruned (nearest rows first) with a FIXED exclusive limit `lim`:
// appends every candidate with d2 bits < lim (list entries beyond LC are dropped, the count keeps running).  32-bit cell arithmetic
// (the cell table has < 2^32 entries), four candidates in flight per iteration.
// List storage: entry e of this query lives at lbase[e * lstride] (shared memory on the device: stride = threads per block).
template <int LC, typename CellLoader, typename Load>
PVB_HD int walk_block_collect(const GridDesc& g, const CellLoader& cells, const Load& load, int cx, int cy, int cz, const FaceGaps& fg, const double gap_lo[3], const double gap_hi[3],
                              int rb, uint32_t lim, float qx, float qy, float qz, U2* lbase, int lstride) {
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  const int W = 2 * rb + 1;
  auto far2 = [&](int a, int d) { const double v = d < 0 ? gap_lo[a] + (double)(-d - 1) * g.h : gap_hi[a] + (double)(d - 1) * g.h; return (float)(v * v * (1.0 - 1e-5)); };
  auto bound2 = [&](int a, int d) { return d == 0 ? 0.0f : (d == -1 ? fg.lo[a] : (d == 1 ? fg.hi[a] : far2(a, d))); };
  U2* lp = lbase;
  U2* const lend = lbase + (long long)LC * lstride;
#pragma unroll 1
  for (int iz = 0; iz < W; ++iz) {
    const int dz = (iz & 1) ? -((iz + 1) >> 1) : (iz >> 1);      // 0, -1, +1, -2, +2: nearest rows first
    const int z = cz + dz;
    if (z < 0 || z >= nz) continue;
    const float bz = bound2(2, dz);
    if (f2u(bz) >= lim) continue;
#pragma unroll 1
    for (int iy = 0; iy < W; ++iy) {
      const int dy = (iy & 1) ? -((iy + 1) >> 1) : (iy >> 1);
      const int y = cy + dy;
      if (y < 0 || y >= ny) continue;
      const float b2 = bound2(1, dy) + bz;                       // float sum of two margin-shrunk terms: still below the true bound
      if (f2u(b2) >= lim) continue;
      int xa = cx, xb = cx;
      for (int d = 1; d <= rb; ++d) {
        if (cx - d < 0 || f2u(b2 + bound2(0, -d)) >= lim) break;
        xa = cx - d;
      }
      for (int d = 1; d <= rb; ++d) {
        if (cx + d > nx - 1 || f2u(b2 + bound2(0, d)) >= lim) break;
        xb = cx + d;
      }
      const uint32_t row = ((uint32_t)z * (uint32_t)ny + (uint32_t)y) * (uint32_t)nx;
      const uint32_t lo = (uint32_t)cells((long long)(row + (uint32_t)xa)), hi = (uint32_t)cells((long long)(row + (uint32_t)xb + 1u));
#pragma unroll 1
      for (uint32_t i = lo; i < hi; i += 4u) {
        F4 c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint32_t ik = i + k < hi ? i + k : hi - 1u; c[k] = load((long long)ik); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t kb = f2u(sqdist_f32(qx, qy, qz, c[k].x, c[k].y, c[k].z));
          const bool p = kb < lim && i + k < hi;
          if (p && lp < lend) { U2 e; e.x = kb; e.y = i + k; *lp = e; }
          lp += p ? lstride : 0;                                 // keeps counting past the capacity: the caller sees the overflow
        }
      }
    }
  }
  return (int)((lp - lbase) / lstride);
}

// The same walk for the 3x3x3 block with the candidates of all rows FLATTENED into one loop: the <= 9 surviving row ranges go to a small per-query table
// first, then one loop runs over the concatenation, two candidates per trip.  The lanes of a warp then iterate max(total candidates) times instead of
// sum over rows of max(row length): with rows of 0..15 records that is ~1.6x fewer trips.  Same visiting order as walk_block_collect, so the same ties
// are resolved the same way.
// The table lives in the LAST kRowTab slots of the candidate list itself (no extra shared memory): table entry r is read into registers when row r starts,
// so list slot LC - kRowTab + r may be overwritten from then on; the list end moves up by one slot with every row start.  A query whose survivors outrun its
// rows simply "overflows" earlier (count > capacity => the caller falls back to the two-pass search); once the poses settle a query has K .. K+3 survivors.
constexpr int kRowTab = 10;          // 9 rows + a spare slot that is read (never used) after the last row
template <int LC, typename CellLoader, typename Load>
PVB_HD int walk_block_collect_flat(const GridDesc& g, const CellLoader& cells, const Load& load, int cx, int cy, int cz, const FaceGaps& fg, uint32_t lim, float qx, float qy, float qz,
                                   U2* lbase, int lstride) {
  static_assert(LC > kRowTab + 4, "list too short to hold the row table");
  const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
  U2* const rtab = lbase + (long long)(LC - kRowTab) * lstride;
  int nr = 0;
  uint32_t rem = 0;
  // the 18 cell-table look-ups of the 9 rows are independent: issue them all before the first one is needed (one memory round trip instead of nine)
  uint32_t c_lo[9], c_hi[9];
#pragma unroll
  for (int o = 0; o < 9; ++o) {
    const int iz = o / 3, iy = o - 3 * iz;
    const int dz = (iz & 1) ? -1 : (iz >> 1), dy = (iy & 1) ? -1 : (iy >> 1);      // 0, -1, +1: nearest rows first
    const int z = cz + dz, y = cy + dy;
    c_lo[o] = c_hi[o] = 0u;
    if (z < 0 || z >= nz || y < 0 || y >= ny) continue;
    const float b2 = (dy == 0 ? 0.0f : (dy < 0 ? fg.lo[1] : fg.hi[1])) + (dz == 0 ? 0.0f : (dz < 0 ? fg.lo[2] : fg.hi[2]));
    if (f2u(b2) >= lim) continue;
    const int xa = (cx - 1 < 0 || f2u(b2 + fg.lo[0]) >= lim) ? cx : cx - 1;
    const int xb = (cx + 1 > nx - 1 || f2u(b2 + fg.hi[0]) >= lim) ? cx : cx + 1;
    const uint32_t row = ((uint32_t)z * (uint32_t)ny + (uint32_t)y) * (uint32_t)nx;
    c_lo[o] = (uint32_t)cells((long long)(row + (uint32_t)xa)); c_hi[o] = (uint32_t)cells((long long)(row + (uint32_t)xb + 1u));
  }
#pragma unroll
  for (int o = 0; o < 9; ++o)
    if (c_hi[o] > c_lo[o]) { U2 e; e.x = c_lo[o]; e.y = c_hi[o]; rtab[(long long)nr * lstride] = e; ++nr; rem += c_hi[o] - c_lo[o]; }
  if (rem == 0u) return 0;
  U2* lp = lbase;
  const U2* rt = rtab;
  uint32_t i = rt->x, hi = rt->y;
  U2* lend = lbase + (long long)(LC - kRowTab + 1) * lstride;      // table entry 0 is in registers: its slot is free
  int over = 0;
#pragma unroll 1
  while (rem > 0u) {
    const bool two = rem >= 2u;
    const uint32_t ia = i;
    ++i;
    bool adv = i == hi;
    if (adv) { rt += lstride; i = rt->x; hi = rt->y; }              // the slot after the last row is never used as a range (rem runs out first)
    const uint32_t ib = two ? i : ia;
    bool adv2 = false;
    if (two) { ++i; adv2 = i == hi; if (adv2) { rt += lstride; i = rt->x; hi = rt->y; } }
    const F4 ca = load((long long)ia), cb = load((long long)ib);
    const uint32_t ka = f2u(sqdist_f32(qx, qy, qz, ca.x, ca.y, ca.z)), kb = f2u(sqdist_f32(qx, qy, qz, cb.x, cb.y, cb.z));
    const bool pa = ka < lim, pb = two && kb < lim;
    // the entries are written AFTER the table reads of this trip; the list end follows the rows already in registers
    lend += (adv ? lstride : 0) + (adv2 ? lstride : 0);
    if (pa) { if (lp < lend) { U2 e; e.x = ka; e.y = ia; *lp = e; lp += lstride; } else over = 1; }
    if (pb) { if (lp < lend) { U2 e; e.x = kb; e.y = ib; *lp = e; lp += lstride; } else over = 1; }
    rem -= two ? 2u : 1u;
  }
  return over ? LC + 1 : (int)((lp - lbase) / lstride);
}

template <int K, int LC, typename CellLoader, typename Load, typename WinSet>
PVB_HD int knn_select_hinted(const GridDesc& g, const CellLoader& cells, const Load& load, float qx, float qy, float qz, float sq_thr, int r0, int rmax, uint32_t lim_hint,
                             U2* lbase, int lstride, const WinSet& set_win, uint32_t* tau_out, bool flat = true) {
  auto lkey = [&](int e) { return lbase[(long long)e * lstride].x; };
  auto lmove = [&](int dst, int src) { lbase[(long long)dst * lstride] = lbase[(long long)src * lstride]; };
  static_assert(LC >= K + 2, "list capacity");
  const uint32_t init = f2u(sq_thr) + 1u;          // every d2 <= sq_thr is below it
  if (lim_hint != 0u && lim_hint < init) {
    const double fx = ((double)qx - g.origin[0]) * g.inv_h, fy = ((double)qy - g.origin[1]) * g.inv_h, fz = ((double)qz - g.origin[2]) * g.inv_h;
    const int cx = cell_coord((double)qx, g.origin[0], g.inv_h, g.dims[0]);
    const int cy = cell_coord((double)qy, g.origin[1], g.inv_h, g.dims[1]);
    const int cz = cell_coord((double)qz, g.origin[2], g.inv_h, g.dims[2]);
    double gap_lo[3], gap_hi[3], slack = 0.5;
    {
      const double f[3] = {fx - cx, fy - cy, fz - cz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double lo = f[a] < 0.0 ? 0.0 : f[a], hi = 1.0 - f[a] < 0.0 ? 0.0 : 1.0 - f[a];
        gap_lo[a] = lo * g.h; gap_hi[a] = hi * g.h;
        const double m = lo < hi ? lo : hi;
        slack = m < slack ? m : slack;
      }
    }
    FaceGaps fg;
#pragma unroll
    for (int a = 0; a < 3; ++a) { fg.lo[a] = (float)(gap_lo[a] * gap_lo[a] * (1.0 - 1e-5)); fg.hi[a] = (float)(gap_hi[a] * gap_hi[a] * (1.0 - 1e-5)); }
    // block radius (cells) that covers the hinted search radius: every point closer than (r + slack) * h lies in the (2r+1)^3 block
    const double lim = (double)u2f(lim_hint);
    const double reach1 = (1.0 + slack) * g.h, reach2 = (2.0 + slack) * g.h;
    int rb = 0;
    if (lim < reach1 * reach1 * (1.0 - 1e-6)) rb = 1;
    else if (lim < reach2 * reach2 * (1.0 - 1e-6)) rb = 2;
    if (rb != 0) {
      int n = (rb == 1 && flat) ? walk_block_collect_flat<LC>(g, cells, load, cx, cy, cz, fg, lim_hint, qx, qy, qz, lbase, lstride)
                                           : walk_block_collect<LC>(g, cells, load, cx, cy, cz, fg, gap_lo, gap_hi, rb, lim_hint, qx, qy, qz, lbase, lstride);
      if (n >= K && n <= LC) {
        uint32_t tau;
        if (n > K) tau = list_cut_to_k<K>(n, lkey, lmove);
        else {
          tau = 0u;
#pragma unroll
          for (int j = 0; j < K; ++j) { const uint32_t v = lkey(j); tau = v > tau ? v : tau; }
        }
        if (tau_out) *tau_out = tau;
        return K;
      }
    }
  }
  // no usable hint: the pruned two-pass search
  return knn_select_pruned<K>(g, cells, load, qx, qy, qz, sq_thr, r0, rmax, [&](int j, uint32_t pos, uint32_t) { set_win(j, pos); }, tau_out);
}

// ---- super-row variant (MODE 4, dense mode default since round 2) ------------------------------------------------------------------
// The 3x3x3 block walk above touches up to 9 row ranges per query (18 cell-table look-ups, per-lane row bookkeeping: the integer overhead was
// ~2/3 of the walk's instructions).  For a STATIC target the 9 rows (y +- 1, z +- 1) around a row can be merged once, at index-build time:
// super-row (y, z) holds, for every x cell, the records of the 9 cells (x, y + dy, z + dz) back to back (dz outer, dy inner), x ascending.
// The 27-cell neighbourhood of a query in cell (cx, cy, cz) is then ONE contiguous range of the super-row array,
//     [sstart[row * nx + cx - 1], sstart[row * nx + cx + 2])        with row = cz * ny + cy,
// found with a few table look-ups and walked by one loop with no bookkeeping.  The array is ~9x the target (1.9 GB for 10 M points): HBM
// capacity and bandwidth traded for instruction issue slots, which is what limits this kernel.
// Layout (SR): every (row, x) segment is padded to a multiple of 4 records with points at +infinity; group G = records 4G .. 4G+3 is stored as
// three float4 {x0..x3}, {y0..y3}, {z0..z3} (48 B), so one 128-bit load feeds two packed FP32x2 operations (two candidates per instruction);
// w(r) = (position in the cell-sorted array << 5) | class and rk(r) (below) live in parallel arrays that are only read for the few survivors.
//   sr.start(i): sstart[i] of this cloud's table (record units, multiples of 4);  sr.load3(G) / sr.sqdist4(quad, q, kb[4]): float32 bit patterns of the squared
//   distances of group G's records to q (flann::L2_Simple: non-fused, x then y then z);  sr.w(r), sr.rk(r).
// The function is WARP-SYNCHRONOUS on the device (W = WarpLanes): all 32 lanes call it (inactive lanes with active = false), loops run to the longest
// range of the warp, and the list is compacted for every lane at once when any lane's list is about to overflow (cut to the K smallest, the
// bound tightens to the K-th) - so ANY bound works, even none, with one walk.  On the host (W = SingleLane) the same code runs per query.
// Bounds on the K-th squared distance (exclusive, float bits):
//   * lim_dyn: from the previous evaluation of the query (see k_associate), 0 = none; dyn_tight: the caller expects few survivors below it;
//   * static (use_static, taken when the dynamic one is absent or loose): sr.rk(p) is the squared distance from target point p to ITS K_build-th
//     nearest target point (p included, K_build >= K, computed once when the index is built), so K target points lie within sqrt(rk(p)) + |q - p| of
//     the query for ANY p; p = the nearest record of the query's own x segment (one short scan).  On surface-like clouds ~1.7 K survivors.
// Exactness: the range covers every point within (1 + slack) * h of the query (slack = distance to the nearest face of the query's own
// cell in cells); the x - 1 / x + 1 cells are skipped only when their nearest face is beyond the bound.  When fewer than K candidates are found
// or the K-th distance is not below that reach (sparse neighbourhood, wrong hint) `fallback` is set and the caller runs the generic ring search.
struct SingleLane {
  PVB_HD static bool any(bool p) { return p; }
  PVB_HD static uint32_t max_u32(uint32_t v) { return v; }
};
#ifdef __CUDACC__
struct WarpLanes {
  __device__ __forceinline__ static bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ static uint32_t max_u32(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
};
#endif
struct NoSuperRow {
  PVB_HD float rk(uint32_t) const { return 0.f; }
  PVB_HD uint32_t w(uint32_t) const { return 0u; }
  PVB_HD uint32_t start(long long) const { return 0u; }
  struct Quad { int none; };
  PVB_HD Quad load3(uint32_t) const { return Quad{0}; }
  PVB_HD void sqdist4(const Quad&, float, float, float, uint32_t (&kb)[4]) const { kb[0] = kb[1] = kb[2] = kb[3] = 0x7F800000u; }
};

template <int K, int LC, typename W, typename SR>
PVB_HD int knn_select_superrow(const GridDesc& g, const SR& sr, bool active, float qx, float qy, float qz, float sq_thr, uint32_t lim_dyn, bool dyn_tight, bool use_static,
                               U2* lbase, int lstride, uint32_t* tau_out, bool& fallback) {
  static_assert(LC >= K + 4, "list capacity");
  fallback = false;
  auto lkey = [&](int e) { return lbase[(long long)e * lstride].x; };
  auto lmove = [&](int dst, int src) { lbase[(long long)dst * lstride] = lbase[(long long)src * lstride]; };
  const uint32_t init = f2u(sq_thr) + 1u;          // every d2 <= sq_thr is below it
  const int nx = g.dims[0], ny = g.dims[1];
  const double fx = ((double)qx - g.origin[0]) * g.inv_h, fy = ((double)qy - g.origin[1]) * g.inv_h, fz = ((double)qz - g.origin[2]) * g.inv_h;
  const int cx = cell_coord((double)qx, g.origin[0], g.inv_h, g.dims[0]);
  const int cy = cell_coord((double)qy, g.origin[1], g.inv_h, g.dims[1]);
  const int cz = cell_coord((double)qz, g.origin[2], g.inv_h, g.dims[2]);
  double slack = 0.5, gx_lo = 0.0, gx_hi = 0.0;
  {
    const double f[3] = {fx - cx, fy - cy, fz - cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double lo = f[a] < 0.0 ? 0.0 : f[a], hi = 1.0 - f[a] < 0.0 ? 0.0 : 1.0 - f[a];
      if (a == 0) { gx_lo = lo * g.h; gx_hi = hi * g.h; }
      const double m = lo < hi ? lo : hi;
      slack = m < slack ? m : slack;
    }
  }
  // squared distance (float, shrunk by a safety margin) to the faces of the query's cell along x = lower bound for the cells beyond them
  const uint32_t b_lo = f2u((float)(gx_lo * gx_lo * (1.0 - 1e-5))), b_hi = f2u((float)(gx_hi * gx_hi * (1.0 - 1e-5)));
  const double reach1 = (1.0 + slack) * g.h;
  const double r1sq = reach1 * reach1 * (1.0 - 1e-6);
  const uint32_t base = ((uint32_t)cz * (uint32_t)ny + (uint32_t)cy) * (uint32_t)nx + (uint32_t)cx;
  const bool has_l = cx > 0, has_r = cx < nx - 1;
  uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;
  if (active) {
    s1 = sr.start((long long)base); s2 = sr.start((long long)(base + 1u));
    s0 = has_l ? sr.start((long long)(base - 1u)) : s1; s3 = has_r ? sr.start((long long)(base + 2u)) : s2;
  }
  uint32_t lim = (lim_dyn != 0u && lim_dyn < init) ? lim_dyn : init;
  // ---- static bound: nearest record of the query's own x segment + that record's K-th neighbour distance
  const bool need_static = active && use_static && !(lim < init && dyn_tight);
  if (W::any(need_static)) {
    const uint32_t a0 = need_static ? s1 : 0u, a1 = need_static ? s2 : 0u;
    const uint32_t it_n = W::max_u32((a1 - a0) >> 2);
    uint32_t best = 0x7F800000u, bestp = 0u, i = a0;
    typename SR::Quad cur = sr.load3(i < a1 ? (i >> 2) : 0u);
#pragma unroll 1
    for (uint32_t it = 0; it < it_n; ++it, i += 4u) {
      const bool in = i < a1;
      const typename SR::Quad nxt = sr.load3(i + 4u < a1 ? ((i + 4u) >> 2) : 0u);      // the next group is in flight while this one is evaluated
      uint32_t kb[4];
      sr.sqdist4(cur, qx, qy, qz, kb);
      cur = nxt;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const bool b = in && kb[k] < best; bestp = b ? i + (uint32_t)k : bestp; best = b ? kb[k] : best; }
    }
    if (need_static && best < 0x7F800000u) {
      const float rk = sr.rk(bestp);
      if (rk >= 0.f && rk < 3.0e38f) {
        const double rad = (sqrt((double)u2f(best)) + sqrt((double)rk)) * (1.0 + 1e-5) + 1e-9;
        const double lim2 = rad * rad;
        if (lim2 < (double)sq_thr) { const uint32_t ls = f2u((float)lim2) + 2u; lim = ls < lim ? ls : lim; }      // +1 ulp for the float rounding, +1 to make the bound exclusive
      }
    }
  }
  // ---- one walk over the range: everything below the bound goes to the list; the warp compacts its lists together when one is about to overflow
  uint32_t lo = (has_l && b_lo < lim) ? s0 : s1, hi = (has_r && b_hi < lim) ? s3 : s2;
  if (!active) lo = hi = 0u;
  const uint32_t it_n = W::max_u32((hi - lo) >> 2);
  int n = 0;
  {
    uint32_t i = lo;
    typename SR::Quad cur = sr.load3(i < hi ? (i >> 2) : 0u);
#pragma unroll 1
    for (uint32_t it = 0; it < it_n; ++it, i += 4u) {
      const bool in = i < hi;
      const typename SR::Quad nxt = sr.load3(i + 4u < hi ? ((i + 4u) >> 2) : 0u);      // the next group is in flight while this one is evaluated
      uint32_t kb[4];
      sr.sqdist4(cur, qx, qy, qz, kb);
      cur = nxt;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool p = in && kb[k] < lim;
        if (p) { U2 e; e.x = kb[k]; e.y = i + (uint32_t)k; lbase[(long long)n * lstride] = e; }
        n += p ? 1 : 0;
      }
      if (W::any(n > LC - 4)) {
        if (n > K) { const uint32_t t = list_cut_to_k<K>(n, lkey, lmove); lim = t < lim ? t : lim; }
      }
    }
  }
  if (!active) return 0;
  if (n < K) { fallback = true; return 0; }
  uint32_t tau;
  if (n > K) tau = list_cut_to_k<K>(n, lkey, lmove);
  else {
    tau = 0u;
#pragma unroll
    for (int j = 0; j < K; ++j) { const uint32_t v = lkey(j); tau = v > tau ? v : tau; }
  }
  if (!((double)u2f(tau) < r1sq)) { fallback = true; return 0; }
#pragma unroll
  for (int j = 0; j < K; ++j) lbase[(long long)j * lstride].y = sr.w(lbase[(long long)j * lstride].y) >> 5;      // record -> position in the cell-sorted array
  if (tau_out) *tau_out = tau;
  return K;
}

struct AssocParams {
  float sq_thr;           // point_to_plane_dis_threshold^2 computed in float (LidarFeatureAssociate.cpp:557)
  int rmax;               // ceil(thr / h)
  int r0;                 // radius (cells) of the block the pruned walk starts with: 1 = 3x3x3, 2 = 5x5x5 (finer grids)
  double plane_tol;       // lidar_plane_tolerance
  double collinear_tol;   // 3.0 (LidarFeatureAssociate.cpp:594)
};

// Per-query body of AssociatePoint2Plane.  qw = query in world (float32), qcls = class label.
// R_ref/t_ref, R_nei/t_nei = R_wl, t_wl of the two frames.  On success: p_local (query in the neighbour's
// sensor frame, double) and plane (n, d) in the reference sensor frame.  win(j) / set_win(j, pos) access the
// caller's per-query neighbour slots (shared memory on the device).
// REF_ID: the reference frame's pose is exactly the identity (rigid target map): World2Local of a neighbour is then the
// neighbour itself bit for bit (x*1 + y*0 + z*0 - 0), so the 3 x K matrix-vector products per query are skipped.
// Everything of AssociatePoint2Plane after the search (LidarFeatureAssociate.cpp:583-599): win(0..K-1) hold the record positions of the K nearest
// neighbours (any order), `load` reads a record.  Same-class test, neighbours -> reference sensor frame, LSQ plane + tolerance, collinearity reject.
// The K neighbour records are fetched ONCE (K independent gathers in flight) into a small store and read from there by the three passes of the fit
// (Gram matrix, refinement step, tolerance test); on the device the store is the warp's shared-memory list region (see k_associate), here a local array.
template <int K>
struct LocalNbStore {
  F4 r[K];
  PVB_HD void sync() {}
  PVB_HD void put(int j, const F4& v) { r[j] = v; }
  PVB_HD F4 get(int j) const { return r[j]; }
};

template <int K, bool REF_ID, typename Load, typename WinGet, typename WinSet, typename NB>
PVB_HD bool plane_from_window(const Load& load, const AssocParams& prm, float qx, float qy, float qz, uint32_t qcls,
                              const double* R_ref, const double* t_ref, const double* R_nei, const double* t_nei,
                              double p_local[3], double plane[4], const WinGet& win, const WinSet& set_win, NB& nb) {
  {
    // canonical neighbour order = ascending record position: the plane fit below sums over the neighbours, and the order the search
    // found them in depends on the walk (block radius, ring expansion, hints); sorting makes the result independent of all of that
    uint32_t wp[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wp[j] = win(j);
#pragma unroll
    for (int a = 1; a < K; ++a) {
#pragma unroll
      for (int b = a; b >= 1; --b) { const uint32_t lo = wp[b - 1] < wp[b] ? wp[b - 1] : wp[b], hi = wp[b - 1] < wp[b] ? wp[b] : wp[b - 1]; wp[b - 1] = lo; wp[b] = hi; }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) set_win(j, wp[j]);
    nb.sync();
#pragma unroll
    for (int j = 0; j < K; ++j) nb.put(j, load((long long)wp[j]));
  }
  // neighbours -> reference sensor frame (:587), streamed: Gram matrix for the LSQ plane and the scatter matrix
  PlaneAcc acc; plane_acc_clear(acc);
  int same = 0;
#pragma unroll 1
  for (int j = 0; j < K; ++j) {
    const F4 p = nb.get(j);
    same += ((f2u(p.w) & 31u) == qcls) ? 1 : 0;                  // :586 pt.intensity == point.intensity
    const double pw[3] = {(double)p.x, (double)p.y, (double)p.z};
    double pl[3];
    if (REF_ID) { pl[0] = pw[0]; pl[1] = pw[1]; pl[2] = pw[2]; } else world2local(R_ref, t_ref, pw, pl);
    plane_acc_add(acc, pl);
  }
  if (same < K) return false;                                    // :590
  if (collinear_from_gram(acc, K, prm.collinear_tol)) return false;   // :594-596
  Chol3 L;
  double x[3];
  if (chol3_factor(acc, L)) {
    { const double rhs[3] = {-acc.h0, -acc.h1, -acc.h2}; chol3_solve(L, rhs, x); }   // :593 FormPlane: A x = -1
    // one refinement step: t = A^T (b - A x)
    double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
      const F4 p = nb.get(j);
      const double pw[3] = {(double)p.x, (double)p.y, (double)p.z};
      double pl[3];
      if (REF_ID) { pl[0] = pw[0]; pl[1] = pw[1]; pl[2] = pw[2]; } else world2local(R_ref, t_ref, pw, pl);
      const double rj = -1.0 - (pl[0] * x[0] + pl[1] * x[1] + pl[2] * x[2]);
      t[0] += pl[0] * rj; t[1] += pl[1] * rj; t[2] += pl[2] * rj;
    }
    double dx[3];
    chol3_solve(L, t, dx);
    x[0] += dx[0]; x[1] += dx[1]; x[2] += dx[2];
  } else {                                                       // rank-deficient neighbour set: Eigen's basic solution
    double A[K * 3];
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
      const F4 p = nb.get(j);
      const double pw[3] = {(double)p.x, (double)p.y, (double)p.z};
      if (REF_ID) { A[j * 3] = pw[0]; A[j * 3 + 1] = pw[1]; A[j * 3 + 2] = pw[2]; } else world2local(R_ref, t_ref, pw, &A[j * 3]);
    }
    lstsq_minus_one_rolled(K, A, x);
  }
  const double nrm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const double d = 1.0 / nrm;                                     // Geometry.hpp:362-363: n = x / |x|, d = 1 / |x|
  const double n0 = x[0] * d, n1 = x[1] * d, n2 = x[2] * d;
  if (prm.plane_tol > 0) {                                       // Geometry.hpp:364-371
    bool ok = true;
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
      const F4 p = nb.get(j);
      const double pw[3] = {(double)p.x, (double)p.y, (double)p.z};
      double pl[3];
      if (REF_ID) { pl[0] = pw[0]; pl[1] = pw[1]; pl[2] = pw[2]; } else world2local(R_ref, t_ref, pw, pl);
      ok = ok && !(fabs(n0 * pl[0] + n1 * pl[1] + n2 * pl[2] + d) > prm.plane_tol);
    }
    if (!ok) return false;
  }
  plane[0] = n0; plane[1] = n1; plane[2] = n2; plane[3] = d;
  const double qw[3] = {(double)qx, (double)qy, (double)qz};
  world2local(R_nei, t_nei, qw, p_local);                        // :598-599
  return true;
}


// The search of the list modes alone (MODE 2: per-row walk, MODE 4: merged super-rows): the K nearest end up in list slots 0..K-1 (.y = position in the
// cell-sorted array); returns K, or 0 when the K-th nearest is beyond the threshold.  MODE 4 on the device: called by ALL lanes of the warp (W = WarpLanes).
template <int K, int MODE, int LC, typename W, typename CellLoader, typename LoadG, typename WinSet, typename SR>
PVB_HD int search_list_modes(const GridDesc& g, const CellLoader& cells, const LoadG& loadg, const AssocParams& prm, bool active, float qx, float qy, float qz, const WinSet& set_win,
                             uint32_t lim_hint, bool dyn_tight, uint32_t* tau_out, U2* lbase, int lstride, bool flat, const SR& sr, bool use_static = true) {
  if (MODE == 4) {
    bool fb = false;
    const int found = knn_select_superrow<K, LC, W>(g, sr, active, qx, qy, qz, prm.sq_thr, lim_hint, dyn_tight, use_static, lbase, lstride, tau_out, fb);
    if (!fb) return found;
    return knn_select_hinted<K, LC>(g, cells, loadg, qx, qy, qz, prm.sq_thr, prm.r0 < 1 ? 1 : prm.r0, prm.rmax, dyn_tight ? lim_hint : 0u, lbase, lstride, set_win, tau_out, false);
  }
  if (!active) return 0;
  return knn_select_hinted<K, LC>(g, cells, loadg, qx, qy, qz, prm.sq_thr, prm.r0 < 1 ? 1 : prm.r0, prm.rmax, lim_hint, lbase, lstride, set_win, tau_out, flat);
}

// MODE: 0 = exhaustive walk over stored row ranges (TMA-staged variant), 1 = pruned two-pass walk, 2 = hinted single pass with the two-pass walk as
// fallback (uses the list accessors and the search-radius hint, see knn_select_hinted; the K nearest end up in list slots 0..K-1 = win(0..K-1)),
// 4 = the same over the merged super-rows of a static target (knn_select_superrow; dense mode default).
template <int K, bool REF_ID, int MODE, int LC = K + 2, typename CellLoader, typename Load1, typename LoadG, typename RowMap, typename WinGet, typename WinSet, typename RangeSet, typename RangeGet,
          typename SR = NoSuperRow>
PVB_HD bool associate_point2plane(const GridDesc& g, const CellLoader& cells, const Load1& load1, const LoadG& loadg, const RowMap& row_map, const AssocParams& prm,
                                  float qx, float qy, float qz, uint32_t qcls,
                                  const double* R_ref, const double* t_ref, const double* R_nei, const double* t_nei,
                                  double p_local[3], double plane[4], const WinGet& win, const WinSet& set_win, const RangeSet& range_set, const RangeGet& range_get,
                                  uint32_t lim_hint, uint32_t* tau_out, U2* lbase, int lstride, bool flat = true, const SR& sr = SR(), bool use_static = true) {
  int ring = 1;
  int found;
  if (MODE == 4 || MODE == 2) { found = search_list_modes<K, MODE, LC, SingleLane>(g, cells, loadg, prm, true, qx, qy, qz, set_win, lim_hint, lim_hint != 0u, tau_out, lbase, lstride, flat, sr, use_static); ring = 2; }
  else if (MODE == 1) { found = knn_select_pruned<K>(g, cells, loadg, qx, qy, qz, prm.sq_thr, prm.r0 < 1 ? 1 : prm.r0, prm.rmax, [&](int j, uint32_t pos, uint32_t) { set_win(j, pos); }); ring = 2; }
  else found = knn_select<K>(g, cells, load1, loadg, row_map, qx, qy, qz, prm.sq_thr, prm.rmax, [&](int j, uint32_t pos, uint32_t) { set_win(j, pos); }, range_set, range_get, ring);
  if (found < K) return false;                                   // :578
  auto load = [&](long long pos) { return ring == 1 ? load1(pos) : loadg(pos); };   // neighbour positions live in the space they were found in (k-th beyond the threshold) + quirk C.6 guard
  LocalNbStore<K> nb;
  return plane_from_window<K, REF_ID>(load, prm, qx, qy, qz, qcls, R_ref, t_ref, R_nei, t_nei, p_local, plane, win, set_win, nb);
}

// Per-query body of AssociatePoint2Line (lidar_mapping/LidarFeatureAssociate.cpp:478-548): 5 nearest corner points of
// the reference frame (world, float32), PCA line test in the WORLD frame, synthetic end points c +- 0.1 d moved to the
// reference sensor frame, query moved to the neighbour's sensor frame.
template <int K, typename CellLoader, typename PointLoader, typename WinGet, typename WinSet, typename RangeSet, typename RangeGet>
PVB_HD bool associate_point2line(const GridDesc& g, const CellLoader& cells, const PointLoader& load, float sq_thr, int rmax, float qx, float qy, float qz,
                                 const double* R_ref, const double* t_ref, const double* R_nei, const double* t_nei,
                                 double p_local[3], double a_local[3], double b_local[3], const WinGet& win, const WinSet& set_win, const RangeSet& range_set,
                                 const RangeGet& range_get) {
  int ring = 1;
  const int found = knn_select<K>(g, cells, load, load, [](int, int, uint32_t&, uint32_t&) {}, qx, qy, qz, sq_thr, rmax, [&](int j, uint32_t pos, uint32_t) { set_win(j, pos); }, range_set, range_get, ring);
  if (found < K) return false;                                   // :497 + quirk C.6 guard
  double pts[K][3];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const F4 p = load((long long)win(j));
    pts[j][0] = (double)p.x; pts[j][1] = (double)p.y; pts[j][2] = (double)p.z;   // :503 (world coordinates)
  }
  double line[6];
  if (!form_line_pca<K>(pts, 10.0, 0.05, line)) return false;    // :506-509
  const double aw[3] = {0.1 * line[3] + line[0], 0.1 * line[4] + line[1], 0.1 * line[5] + line[2]};      // :514-515
  const double bw[3] = {-0.1 * line[3] + line[0], -0.1 * line[4] + line[1], -0.1 * line[5] + line[2]};
  const double qw[3] = {(double)qx, (double)qy, (double)qz};
  world2local(R_nei, t_nei, qw, p_local);                        // :517
  world2local(R_ref, t_ref, aw, a_local);                        // :518-519
  world2local(R_ref, t_ref, bw, b_local);
  return true;
}

// nearestKSearch alone (the segment-based variants, LidarFeatureAssociate.cpp:238-317, 385-440, only need the neighbour set):
// indices into the reference cloud of the K nearest points, or false when the K-th is beyond the threshold / the cloud is smaller than K.
template <int K, typename CellLoader, typename PointLoader, typename WinGet, typename WinSet, typename RangeSet, typename RangeGet>
PVB_HD bool knn_indices(const GridDesc& g, const CellLoader& cells, const PointLoader& load, float sq_thr, int rmax, float qx, float qy, float qz, int out_idx[K],
                        const WinGet& win, const WinSet& set_win, const RangeSet& range_set, const RangeGet& range_get) {
  int ring = 1;
  const int found = knn_select<K>(g, cells, load, load, [](int, int, uint32_t&, uint32_t&) {}, qx, qy, qz, sq_thr, rmax, [&](int j, uint32_t pos, uint32_t) { set_win(j, pos); }, range_set, range_get, ring);
  if (found < K) return false;                                   // :261 / :416 + quirk C.6 guard
#pragma unroll
  for (int j = 0; j < K; ++j) out_idx[j] = (int)(f2u(load((long long)win(j)).w) >> 5);
  return true;
}

}  // namespace pvb
