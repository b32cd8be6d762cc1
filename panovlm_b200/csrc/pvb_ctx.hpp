// panovlm_b200 — the context behind the C ABI: device / pinned buffer holders, the per-mode state and the error macros.
// Shared by the translation units of the library (pvb_lib.cu: association / residual / solver modes, pvb_ba.cu: reprojection mode).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/panovlm_b200.h"
#include "pvb_knn.cuh"

namespace pvb {

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct WorldPoseHost { double R[9]; double t[3]; };

struct CloudSet {
  int n_clouds = 0; long long n_points = 0; int n_tiles = 0;
  std::vector<int> off;           // n_clouds + 1
  DevBuf local, d_off, tiles, cloud_block;
  void release() { local.release(); d_off.release(); tiles.release(); cloud_block.release(); }
};

struct TargetIndex {
  DevBuf world, sorted, keys, keys_alt, vals, vals_alt, hist, cell_start, grids, aabb, tmp;
  DevBuf srow, sstart, srk, srw; bool has_superrows = false;      // merged super-rows of a static single-cloud target (MODE 4, dense mode)
  std::vector<GridDesc> h_grids;
  long long total_cells = 0;
  bool built = false;
  void release() { world.release(); sorted.release(); keys.release(); keys_alt.release(); vals.release(); vals_alt.release(); hist.release(); cell_start.release(); grids.release(); aabb.release(); tmp.release(); srow.release(); sstart.release(); srk.release(); srw.release(); has_superrows = false; }
};

}  // namespace pvb

struct pvb_ctx {
  using DevBuf = pvb::DevBuf; using PinBuf = pvb::PinBuf; using CloudSet = pvb::CloudSet; using TargetIndex = pvb::TargetIndex; using GridDesc = pvb::GridDesc;
  int device = 0;
  cudaStream_t stream = nullptr; bool own_stream = true;
  std::string err;
  long launches = 0;
  int tune_minb = 6, tune_stage = 0, tune_r0 = 1, tune_mode = 2, tune_dense_mode = 4, tune_static = 1, tune_key64 = 0, tune_chunks = 8, tune_chunk_min = 1000000, tune_hints = 1, tune_morton_bits = 12, tune_flat = 1; double tune_cellcap = 4.0, tune_hscale = 1.0, tune_dense_hscale = 1.0, tune_reorder = 1.0, tune_tight = 0.8; DevBuf d_stats;   // fastest measured on B200 (tools/sweep_variants.py, profiles/r1f_sweep.log)
  // pose staging
  PinBuf h_pose; DevBuf d_prep, d_wpose;
  // ---- blocks mode
  long long bn = 0; int nb = 0; int b_tiles = 0;
  std::vector<int> edge_ref, edge_nei, edge_tile_begin;
  std::vector<uint32_t> b_orig;
  DevBuf b_chunk, d_chunk;
  DevBuf b_tile, b_eref, b_enei, b_type, b_norm, b_huber, b_consts, b_orig_d, b_r, b_J, b_part, b_esys, b_tbegin;
  PinBuf h_r, h_J, h_esys;
  bool b_has_rows = false, b_has_sys = false;
  // multi-GPU: the pose graph's GLOBAL edge list (every rank reduces into the same layout) and the exchange step of an evaluation
  std::vector<int> g_edge_ref, g_edge_nei;
  pvb_reduce_hook reduce_hook = nullptr; void* reduce_user = nullptr;
  // ---- frames mode
  CloudSet f_tgt, f_qry; TargetIndex f_index; int n_frames = 0;
  CloudSet f_corner; TargetIndex f_cindex; int n_corner_frames = 0;
  CloudSet p_cs; TargetIndex p_index;                                  // pair-level k-NN (pvb_pair_knn5)
  DevBuf f_la, f_lb; PinBuf fh_la, fh_lb;
  std::vector<int> l_edge, l_query; std::vector<double> l_point, l_a, l_b;
  DevBuf f_pairs, f_qtiles, f_valid, f_point, f_plane, f_nn_idx, f_nn_d2;
  PinBuf fh_valid, fh_point, fh_plane;
  std::vector<int> a_edge, a_query; std::vector<double> a_point, a_plane;
  // ---- dense mode
  CloudSet d_tgt, d_src; TargetIndex d_index; int d_frames = 0;
  DevBuf d_hint;                                                       // search-radius hints of the dense queries (launch order), see AssocArgs::hint
  DevBuf d_q_sorted, d_q_orig, d_pairs, d_qtiles, d_part, d_sys, d_tbegin, d_valid, d_point, d_plane, d_res, d_jac;
  PinBuf dh_sys;
  int d_ntiles = 0; double d_cell = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool ev_valid = false;
  cudaEvent_t bev0 = nullptr, bev1 = nullptr; bool bev_valid = false;     // brackets k_eval_blocks of the last blocks evaluate
  cudaStream_t copy_stream = nullptr, sort_stream = nullptr; cudaEvent_t eval_done = nullptr; std::vector<cudaEvent_t> chunk_ev, h2d_ev;
  std::vector<int> d_chunk_frame, d_chunk_ctile, d_chunk_qtile; bool d_chunks_pending = false;
  // MODE 4: the queries are ordered by target cell under the poses of an evaluation (k_target_cell_keys); the order is redone after an upload and when a
  // pose update may have moved a point by more than half a cell (bound from the pose change and the frame's largest sensor distance)
  bool d_cell_order = false, d_order_pending = false, d_order_valid = false; std::vector<pvb::WorldPoseHost> d_order_wpose; DevBuf d_rmax2; PinBuf dh_rmax2; cudaEvent_t rmax_ev = nullptr, pose_ev = nullptr; bool rmax_inflight = false;
  long d_reorders = 0, d_order_reuses = 0;
  // a fresh upload with the layout of the previous one re-uses the previous permutation (gather only, no keys / sort) while the poses stay close to those of the
  // last sort and the measured locality of the order (k_order_locality, read back one evaluation later) stays near the value seen right after that sort
  bool d_perm_valid = false, loc_inflight = false, loc_was_sort = false, loc_known = false; double loc_sorted_frac = 0.0, loc_last_frac = 1.0; int reuse_holdoff = 0, reuse_backoff = 1;
  DevBuf d_loc; PinBuf dh_loc; cudaEvent_t loc_ev = nullptr; int tune_reuse = 1;   // brackets the fused associate kernel of the last dense evaluate
  // ---- device linear solver of the LM loop (pvb_solver.cuh)
  int solver_kind = 0;                                                  // PVB_SOLVER_AUTO
  bool solver_attr_set = false;                                         // dynamic shared memory opt-in of the solver kernels done on this context's device
  DevBuf s_H, s_A, s_g, s_sc, s_rhs, s_y, s_term, s_con, s_seg, s_gcon, s_gseg, s_fail; PinBuf sh_vec;
  // block-sparse PCG form of the step (pvb_solver.cuh): BSR structure + values, CG vectors, per-CTA partial sums, device scalars
  DevBuf s_bsr_rowptr, s_bsr_col, s_bsr_diag, s_bsr_val, s_px, s_pr, s_pz, s_pp, s_pq, s_dmp, s_minv, s_part, s_scal; int s_nfb = 0, s_nblk = 0;
  bool pcg_active = false; double pcg_tol = 1e-12; int pcg_max_it = 4000; long pcg_iterations = 0, pcg_solves = 0;
  int s_n = 0, s_N = 0, s_ndest = 0, s_ngdest = 0; float s_last_factor_ms = 0.f;
  // ---- misc
  DevBuf m_a, m_b, m_c, m_d, m_e;
  DevBuf v_local, v_world, v_misc, v_M;                                 // batched line votes (pvb_line_votes_batch)
  // device tails of the line-to-line family (pvb_frames_line2line_blocks_device): where the tables of the last vote pass lie in v_misc, the tail's own tables
  // (lt_misc) and scratch, and the blocks that wait for the next pvb_frames_point2plane_blocks to place them in the block arrays
  struct LineLayout { bool valid = false; int n_frames = 0, n_pairs = 0; size_t o_coff = 0, o_soff = 0, o_pbase = 0, o_poff = 0, o_pids = 0, o_wp = 0, o_lines = 0, o_vp = 0; } line_layout;
  struct LinePending {
    bool pending = false; long long total = 0, slots = 0;
    std::vector<int> ref, nei, cnt;                                     // the edges (frame indices) and their block counts
    int type = 0, normalize = 0; double huber = 0.0, weight = 1.0;
    size_t o_hoff = 0, o_slot = 0, o_coeffs = 0;                        // offsets inside lt_misc
  } line_pending;
  DevBuf lt_misc, lt_hold, lt_kbase, lt_cnt, lt_base;
  PinBuf mh_a;
  // ---- reprojection / bundle-adjustment mode (pvb_ba.cu owns the state; released through ba_free by pvb_destroy)
  void* ba_state = nullptr; void (*ba_free)(void*) = nullptr;

  int fail(int code, const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf; return code;
  }
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->fail(PVB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CKL() do { ctx->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return ctx->fail(PVB_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

// ---- internal entry points shared between the translation units (not part of the ABI) ---------------------------------------------
// blocked FP64 Cholesky + substitution of the LM step: factor the N x N matrix in ctx->s_A (lower triangle, N a multiple of 64) in place
// and solve with the right-hand side in ctx->s_rhs; *ok = false when a pivot fails
extern "C" __attribute__((visibility("hidden"))) int pvb_internal_factor_solve(pvb_ctx* ctx, int N, bool* ok);
// pieces of the device LM step of the pose-graph blocks (pvb_lib.cu), reused by the joint camera-LiDAR solve:
//   prepare : free-block map and contribution lists for the given constant blocks (sets ctx->s_n, ctx->s_N)
//   assemble: dense J^T J (ctx->s_H, N x N) and gradient (ctx->s_g) of the free blocks from the edge systems of the last pvb_blocks_evaluate;
//             the gradient is also copied to h_g (s_n doubles)
//   jacobi_scale: ctx->s_sc = 1 / (1 + sqrt(diag H));  build_damped: ctx->s_A (lower triangle) / ctx->s_rhs from s_H, s_g, s_sc and the radius
extern "C" __attribute__((visibility("hidden"))) int pvb_internal_solver_prepare(pvb_ctx* ctx, const unsigned char* is_const_block);
extern "C" __attribute__((visibility("hidden"))) int pvb_internal_solver_assemble(pvb_ctx* ctx, double* h_g);
extern "C" __attribute__((visibility("hidden"))) int pvb_internal_jacobi_scale(pvb_ctx* ctx);
extern "C" __attribute__((visibility("hidden"))) int pvb_internal_build_damped(pvb_ctx* ctx, double radius);
