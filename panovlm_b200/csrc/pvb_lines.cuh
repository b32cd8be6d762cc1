// panovlm_b200 — device tails of AddLidarLineToLineResidual2 (util/Optimization.cpp:329-441): after the batched vote pass (k_line_votes_batch) the
// FindAssociations tail (lidar_mapping/LidarFeatureAssociate.cpp:120-197), the line-track gate (Optimization.cpp:383-400) and the Point2Line blocks
// (:402-435) are produced on the device, straight into the block arrays of the context - no vote matrix, world cloud or correspondence goes to the host.
//   k_line_assoc      : one thread per pose-graph edge (the tail is a sequential fold over <= ~50 neighbour segments with a map keyed by reference line)
//   k_line_blocks_emit: one CTA per (edge, reference line) slot that kept a neighbour segment; ordered compaction of the segment's member points
// The arithmetic is written with the uncontracted dadd / dmul / dsub of pvb_math.cuh in the operation order of the host tail (pvb_builders.cpp:
// find_associations, pvb_frames_line2line_blocks), so the constants of the blocks are bit-identical to the host path's.
#pragma once
#include "pvb_kernels.cuh"

namespace pvb {

__device__ inline double lt_sq(double x) { return dmul(x, x); }
// PointToLineDistance3D (Geometry.hpp:198-211) of the line's point p against line l
__device__ inline double lt_point_to_line(const double* p, const double* l) {
  const double k = dadd(dadd(dmul(l[3], dsub(p[0], l[0])), dmul(l[4], dsub(p[1], l[1]))), dmul(l[5], dsub(p[2], l[2]))) / dadd(dadd(lt_sq(l[3]), lt_sq(l[4])), lt_sq(l[5]));
  const double q0 = dadd(dmul(k, l[3]), l[0]), q1 = dadd(dmul(k, l[4]), l[1]), q2 = dadd(dmul(k, l[5]), l[2]);
  return sqrt(dadd(dadd(lt_sq(dsub(q0, p[0])), lt_sq(dsub(q1, p[1]))), lt_sq(dsub(q2, p[2]))));
}
// PlaneAngle (Geometry.hpp:471-485) of two direction vectors
__device__ inline double lt_plane_angle(const double* a, const double* b) {
  double c = fabs(dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2])));
  c = c / dmul(sqrt(dadd(dadd(lt_sq(a[0]), lt_sq(a[1])), lt_sq(a[2]))), sqrt(dadd(dadd(lt_sq(b[0]), lt_sq(b[1])), lt_sq(b[2]))));
  if (c >= 1.0) return 0.0;
  return acos(c);
}

// holder[h_off[e] + c] = neighbour segment kept for reference line c of edge e (-1: none); kbase = first block of that slot inside the edge; cnt[e] = blocks of the edge
__global__ void __launch_bounds__(128) k_line_assoc(const VotePair* __restrict__ pairs, int n_edges, const int* __restrict__ M, const int* __restrict__ seg_off,
                                                    const double* __restrict__ lines /*[sum segments][6], world*/, const int* __restrict__ seg_size /*p2s entries per segment*/,
                                                    const int* __restrict__ member_cnt /*points per segment*/, const int* __restrict__ track_of /*NULL: no gate*/,
                                                    const int* __restrict__ h_off, int* __restrict__ holder, int* __restrict__ kbase, int* __restrict__ cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const VotePair pr = pairs[e];
  const int so_r = seg_off[pr.ref], so_n = seg_off[pr.nei];
  const int Sr = seg_off[pr.ref + 1] - so_r, Sn = seg_off[pr.nei + 1] - so_n;
  int* hold = holder + h_off[e];
  int* kb = kbase + h_off[e];
  for (int c = 0; c < Sr; ++c) { hold[c] = -1; kb[c] = 0; }
  if (Sr == 0 || Sn == 0) { cnt[e] = 0; return; }                      // CheckLidarSegment (:208-216)
  const double* rw = lines + (size_t)so_r * 6;
  const double* nw = lines + (size_t)so_n * 6;
  const int* Mp = M + pr.m_off;
  for (int s = 0; s < Sn; ++s) {
    const int* row = Mp + (size_t)s * Sr;
    int max_col = 0, max_count = row[0];
    for (int c = 1; c < Sr; ++c) { const int v = row[c]; if (v > max_count) { max_count = v; max_col = c; } }
    if ((unsigned long long)max_count < (unsigned long long)seg_size[so_n + s] / 2) continue;           // :132
    if (lt_plane_angle(rw + 6 * max_col + 3, nw + 6 * s + 3) * 180.0 / M_PI > 7) continue;              // :138
    const int h = hold[max_col];
    if (h < 0) hold[max_col] = s;
    else {
      const double d1 = lt_point_to_line(nw + 6 * h, rw + 6 * max_col);
      const double d2 = lt_point_to_line(nw + 6 * s, rw + 6 * max_col);
      if (d2 < d1) hold[max_col] = s;
    }
  }
  int total = 0;
  for (int c = 0; c < Sr; ++c) {                                       // ascending reference line: the iteration order of the reference's map
    const int s = hold[c];
    if (s < 0) continue;
    if (track_of) {                                                    // Optimization.cpp:383-400: both lines in one track
      const int ta = track_of[so_r + c], tb = track_of[so_n + s];
      if (ta < 0 || ta != tb) { hold[c] = -1; continue; }
    }
    kb[c] = total;
    total += member_cnt[so_n + s];
  }
  cnt[e] = total;
}

__global__ void __launch_bounds__(128) k_line_blocks_emit(const int* __restrict__ slot_edge, const int* __restrict__ h_off, const int* __restrict__ holder, const int* __restrict__ kbase,
                                                          const long long* __restrict__ edge_base, const VotePair* __restrict__ pairs, const int* __restrict__ seg_off,
                                                          const double* __restrict__ coeffs_local /*[sum segments][6]*/, const int* __restrict__ corner_off, const int* __restrict__ p2s_off,
                                                          const int* __restrict__ p2s_base, const int* __restrict__ p2s_ids, const F4* __restrict__ world, const WorldPose* __restrict__ wp,
                                                          int type, int normalize, double huber, double weight, long long n /*rows of the block arrays*/, long long at /*first line block*/,
                                                          int* __restrict__ b_type, int* __restrict__ b_norm, double* __restrict__ b_huber, double* __restrict__ b_consts, uint32_t* __restrict__ b_orig) {
  __shared__ int s_warp[4];
  const int slot = blockIdx.x;
  const int s = holder[slot];
  if (s < 0) return;
  const int e = slot_edge[slot];
  const int c = slot - h_off[e];
  const VotePair pr = pairs[e];
  const double* cl = coeffs_local + (size_t)(seg_off[pr.ref] + c) * 6;                                  // the reference line in its own frame (:144-145)
  double a[3], d[3];
  for (int k = 0; k < 3; ++k) { a[k] = dadd(dmul(0.1, cl[3 + k]), cl[k]); d[k] = dsub(a[k], dadd(dmul(-0.1, cl[3 + k]), cl[k])); }
  const double nn = sqrt(dadd(dadd(lt_sq(d[0]), lt_sq(d[1])), lt_sq(d[2])));                            // Point2Line_*: line_direction = (a - b).normalized()
  for (int k = 0; k < 3; ++k) d[k] = d[k] / nn;
  const WorldPose W = wp[pr.nei];
  const int c0 = corner_off[pr.nei], n_pts = corner_off[pr.nei + 1] - c0;
  const int* po = p2s_off + c0 + pr.nei;                                                                // the frame's n + 1 CSR offsets
  const int* ids = p2s_ids + p2s_base[pr.nei];
  const long long row0 = at + edge_base[e] + kbase[slot];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int run = 0;
  for (int i0 = 0; i0 < n_pts; i0 += 128) {
    const int i = i0 + (int)threadIdx.x;
    bool member = false;
    if (i < n_pts) for (int q = po[i]; q < po[i + 1]; ++q) member = member || ids[q] == s;
    const unsigned bal = __ballot_sync(0xffffffffu, member);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int before = 0, tile = 0;
    for (int k = 0; k < 4; ++k) { const int v = s_warp[k]; tile += v; if (k < w) before += v; }
    if (member) {
      const long long r = row0 + run + before + __popc(bal & ((1u << lane) - 1u));
      const F4 p = ldg_f4(world + c0 + i);
      const double pw[3] = {(double)p.x, (double)p.y, (double)p.z};
      double pl[3];
      world2local(W.R, W.t, pw, pl);                                                                    // Optimization.cpp:407 / :422
      b_type[r] = type; b_norm[r] = normalize; b_huber[r] = huber; b_orig[r] = (uint32_t)r;
      b_consts[0 * n + r] = pl[0]; b_consts[1 * n + r] = pl[1]; b_consts[2 * n + r] = pl[2];
      b_consts[3 * n + r] = a[0]; b_consts[4 * n + r] = a[1]; b_consts[5 * n + r] = a[2];
      b_consts[6 * n + r] = d[0]; b_consts[7 * n + r] = d[1]; b_consts[8 * n + r] = d[2];
      b_consts[9 * n + r] = weight; b_consts[10 * n + r] = 0.0; b_consts[11 * n + r] = 0.0;
    }
    __syncthreads();
    run += tile;
  }
}

}  // namespace pvb
