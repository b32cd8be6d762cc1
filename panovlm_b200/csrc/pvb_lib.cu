// panovlm_b200 — context, device-memory management and the extern "C" ABI (include/panovlm_b200.h).
// One context = one GPU + one stream.  Host code is C++ (the reference's host is C++); no torch types here.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "../../include/panovlm_b200.h"
#include "pvb_ctx.hpp"
#include "pvb_host.hpp"
#include "pvb_kernels.cuh"
#include "pvb_solver.cuh"
#include "pvb_lines.cuh"

using namespace pvb;

namespace {

// scratch buffers m_a..m_e are shared with the source-upload pipeline: wait for it before reusing them elsewhere
int quiesce_copy(pvb_ctx* ctx) {
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  if (ctx->sort_stream) CK(cudaStreamSynchronize(ctx->sort_stream));
  return PVB_OK;
}

int upload_poses(pvb_ctx* ctx, const double* poses, int nb, bool block0_identity_prefix) {
  // layout in pinned staging: PosePrep[nb'] then WorldPose[nb'] where nb' = nb (+1 if an identity block is prefixed)
  const int n = nb + (block0_identity_prefix ? 1 : 0);
  const size_t bytes = (size_t)n * (sizeof(PosePrep) + sizeof(WorldPose));
  CK(ctx->h_pose.ensure(bytes));
  CK(ctx->d_prep.ensure((size_t)n * sizeof(PosePrep)));
  CK(ctx->d_wpose.ensure((size_t)n * sizeof(WorldPose)));
  CK(cudaStreamSynchronize(ctx->stream));   // staging buffer may still be in flight from the previous call
  PosePrep* hp = ctx->h_pose.as<PosePrep>();
  WorldPose* hw = reinterpret_cast<WorldPose*>(hp + n);
  int o = 0;
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  if (block0_identity_prefix) { prepare_pose(zero6, hp[0]); world_pose(hp[0], hw[0].R, hw[0].t); o = 1; }
  for (int b = 0; b < nb; ++b) { prepare_pose(poses + 6 * b, hp[o + b]); world_pose(hp[o + b], hw[o + b].R, hw[o + b].t); }
  // The prepared poses travel by a KERNEL that reads the pinned staging buffer (mapped under unified addressing), not by the copy engine: a cudaMemcpyAsync
  // would queue behind the source-cloud chunks of pvb_dense_set_sources on the host-to-device engine, and every kernel of the step needs the poses first
  // (measured: upload 2.9 ms + evaluate 2.2 ms took 6.3 ms instead of overlapping).
  static_assert(sizeof(PosePrep) % 8 == 0 && sizeof(WorldPose) % 8 == 0, "copied as 64-bit words");
  const long long w_prep = (long long)n * (long long)(sizeof(PosePrep) / 8), w_world = (long long)n * (long long)(sizeof(WorldPose) / 8);
  k_copy_words<<<(unsigned)((w_prep + w_world + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long*>(hp), w_prep, ctx->d_prep.as<unsigned long long>(),
                                                                                     reinterpret_cast<const unsigned long long*>(hw), w_world, ctx->d_wpose.as<unsigned long long>());
  CKL();
  return PVB_OK;
}

int set_cloudset(pvb_ctx* ctx, CloudSet& cs, const std::vector<const float*>& ptrs, const std::vector<int>& counts, const std::vector<int>& blocks, int tile) {
  cs.n_clouds = (int)counts.size();
  cs.off.assign(cs.n_clouds + 1, 0);
  for (int c = 0; c < cs.n_clouds; ++c) cs.off[c + 1] = cs.off[c] + counts[c];
  cs.n_points = cs.off[cs.n_clouds];
  CK(cs.local.ensure(std::max<size_t>(16, (size_t)cs.n_points * sizeof(F4))));
  for (int c = 0; c < cs.n_clouds; ++c)
    if (counts[c] > 0) CK(cudaMemcpyAsync(cs.local.as<F4>() + cs.off[c], ptrs[c], (size_t)counts[c] * sizeof(F4), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<CloudTile> tiles;
  for (int c = 0; c < cs.n_clouds; ++c)
    for (int s = 0; s < counts[c]; s += tile) tiles.push_back(CloudTile{c, cs.off[c] + s, std::min(tile, counts[c] - s), 0});
  cs.n_tiles = (int)tiles.size();
  CK(cs.tiles.ensure(std::max<size_t>(16, tiles.size() * sizeof(CloudTile))));
  CK(cs.d_off.ensure((size_t)(cs.n_clouds + 1) * sizeof(int)));
  CK(cs.cloud_block.ensure(std::max<size_t>(16, (size_t)cs.n_clouds * sizeof(int))));
  if (!tiles.empty()) CK(cudaMemcpyAsync(cs.tiles.p, tiles.data(), tiles.size() * sizeof(CloudTile), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(cs.d_off.p, cs.off.data(), (size_t)(cs.n_clouds + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (cs.n_clouds) CK(cudaMemcpyAsync(cs.cloud_block.p, blocks.data(), (size_t)cs.n_clouds * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
  return PVB_OK;
}

// world transform + cell-sorted layout of every cloud of `cs` (poses already uploaded to ctx->d_wpose)
int build_target_index(pvb_ctx* ctx, CloudSet& cs, TargetIndex& ti, double cell_hint, double hscale = 0.0) {
  if (hscale <= 0.0) hscale = ctx->tune_hscale;
  const long long n = cs.n_points;
  ti.built = false;
  if (n == 0 || cs.n_clouds == 0) { ti.h_grids.assign(cs.n_clouds, GridDesc{}); ti.built = true; return PVB_OK; }
  CK(ti.world.ensure((size_t)n * sizeof(F4)));
  CK(ti.sorted.ensure((size_t)n * sizeof(F4)));
  CK(ti.aabb.ensure((size_t)cs.n_clouds * 6 * sizeof(uint32_t)));
  std::vector<uint32_t> init((size_t)cs.n_clouds * 6);
  for (int c = 0; c < cs.n_clouds; ++c) for (int k = 0; k < 6; ++k) init[(size_t)c * 6 + k] = k < 3 ? 0xFFFFFFFFu : 0u;
  CK(cudaMemcpyAsync(ti.aabb.p, init.data(), init.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  k_transform_world<<<cs.n_tiles, 256, 0, ctx->stream>>>(cs.local.as<F4>(), cs.tiles.as<CloudTile>(), cs.cloud_block.as<int>(), ctx->d_wpose.as<WorldPose>(),
                                                         cs.d_off.as<int>(), ti.world.as<F4>(), ti.aabb.as<uint32_t>());
  CKL();
  CK(cudaMemcpyAsync(init.data(), ti.aabb.p, init.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ti.h_grids.assign(cs.n_clouds, GridDesc{});
  long long cells = 0;
  for (int c = 0; c < cs.n_clouds; ++c) {
    GridDesc& g = ti.h_grids[c];
    const int cnt = cs.off[c + 1] - cs.off[c];
    g.n_points = cnt; g.point_base = cs.off[c]; g.cell_base = cells;
    if (cnt == 0) { g.dims[0] = g.dims[1] = g.dims[2] = 1; g.h = 1.0; g.inv_h = 1.0; g.origin[0] = g.origin[1] = g.origin[2] = 0; cells += 2; continue; }
    double lo[3], ext[3];
    for (int k = 0; k < 3; ++k) { lo[k] = unordered_f32(init[(size_t)c * 6 + k]); ext[k] = std::max(1e-3, (double)unordered_f32(init[(size_t)c * 6 + 3 + k]) - lo[k]); }
    double h = cell_hint > 0 ? cell_hint : hscale * std::cbrt(ext[0] * ext[1] * ext[2] / (double)cnt);
    h = std::max(h, 0.02);
    for (;;) {
      double nc = 1; for (int k = 0; k < 3; ++k) nc *= std::floor(ext[k] / h) + 1;
      if (nc <= ctx->tune_cellcap * cnt + 4096.0 && nc < 1.5e9) break;
      h *= 1.25;
    }
    g.h = h; g.inv_h = 1.0 / h;
    for (int k = 0; k < 3; ++k) { g.origin[k] = lo[k]; g.dims[k] = (int)std::floor(ext[k] / h) + 1; }
    cells += (long long)g.dims[0] * g.dims[1] * g.dims[2] + 1;   // +1: the row-end lookup cells(row + x1 + 1) of the last row
  }
  ti.total_cells = cells;
  if (cells >= (1ll << 32)) return ctx->fail(PVB_ERR_ARG, "grid too large (%lld cells)", cells);
  CK(ti.grids.ensure((size_t)cs.n_clouds * sizeof(GridDesc)));
  CK(cudaMemcpyAsync(ti.grids.p, ti.h_grids.data(), (size_t)cs.n_clouds * sizeof(GridDesc), cudaMemcpyHostToDevice, ctx->stream));
  CK(ti.hist.ensure((size_t)(cells + 1) * 4));
  CK(ti.cell_start.ensure((size_t)(cells + 1) * 4));
  CK(ti.keys.ensure((size_t)n * 8)); CK(ti.keys_alt.ensure((size_t)n * 8));
  CK(ti.vals.ensure((size_t)n * 4)); CK(ti.vals_alt.ensure((size_t)n * 4));
  CK(cudaMemsetAsync(ti.hist.p, 0, (size_t)(cells + 1) * 4, ctx->stream));
  k_cell_keys<<<cs.n_tiles, 256, 0, ctx->stream>>>(ti.world.as<F4>(), cs.tiles.as<CloudTile>(), ti.grids.as<GridDesc>(), ti.keys.as<unsigned long long>(),
                                                   ti.vals.as<uint32_t>(), ti.hist.as<uint32_t>());
  CKL();
  int end_bit = 1; while ((1ll << end_bit) < cells + 1 && end_bit < 63) ++end_bit;
  size_t tmp1 = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp1, ti.keys.as<unsigned long long>(), ti.keys_alt.as<unsigned long long>(), ti.vals.as<uint32_t>(), ti.vals_alt.as<uint32_t>(),
                                  (int)n, 0, end_bit, ctx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp2, ti.hist.as<uint32_t>(), ti.cell_start.as<uint32_t>(), (int)(cells + 1), ctx->stream);
  CK(ti.tmp.ensure(std::max(tmp1, tmp2)));
  size_t tb = ti.tmp.cap;
  CK(cub::DeviceRadixSort::SortPairs(ti.tmp.p, tb, ti.keys.as<unsigned long long>(), ti.keys_alt.as<unsigned long long>(), ti.vals.as<uint32_t>(), ti.vals_alt.as<uint32_t>(),
                                     (int)n, 0, end_bit, ctx->stream));
  tb = ti.tmp.cap;
  CK(cub::DeviceScan::ExclusiveSum(ti.tmp.p, tb, ti.hist.as<uint32_t>(), ti.cell_start.as<uint32_t>(), (int)(cells + 1), ctx->stream));
  k_gather_f4<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ti.world.as<F4>(), ti.vals_alt.as<uint32_t>(), n, ti.sorted.as<F4>());
  CKL();
  ti.built = true;
  return PVB_OK;
}

// merged super-rows of a single-cloud index (MODE 4, see knn_select_superrow): 9x the records, one contiguous range per 27-cell neighbourhood
int build_superrows(pvb_ctx* ctx, TargetIndex& ti) {
  ti.has_superrows = false;
  if (!ti.built || ti.h_grids.size() != 1 || ti.h_grids[0].n_points <= 0) return PVB_OK;
  const GridDesc g = ti.h_grids[0];
  const long long ncells = (long long)g.dims[0] * g.dims[1] * g.dims[2];
  if (10ll * g.n_points >= (1ll << 32)) return PVB_OK;                         // positions are 32-bit
  CK(ti.hist.ensure((size_t)(ncells + 1) * 4));
  CK(ti.sstart.ensure((size_t)(ncells + 1) * 4));
  const unsigned nb = (unsigned)((ncells + 1 + 255) / 256);
  k_superrow_counts<<<nb, 256, 0, ctx->stream>>>(g, ti.cell_start.as<uint32_t>(), ncells, ti.hist.as<uint32_t>());
  CKL();
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, ti.hist.as<uint32_t>(), ti.sstart.as<uint32_t>(), (int)(ncells + 1), ctx->stream);
  CK(ti.tmp.ensure(tb));
  tb = ti.tmp.cap;
  CK(cub::DeviceScan::ExclusiveSum(ti.tmp.p, tb, ti.hist.as<uint32_t>(), ti.sstart.as<uint32_t>(), (int)(ncells + 1), ctx->stream));
  uint32_t total = 0;
  CK(cudaMemcpyAsync(&total, ti.sstart.as<uint32_t>() + ncells, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(ti.srow.ensure(std::max<size_t>(64, (size_t)total * 12)));
  CK(ti.srk.ensure(std::max<size_t>(16, (size_t)total * 4))); CK(ti.srw.ensure(std::max<size_t>(16, (size_t)total * 4)));
  // static search bound: every target point's distance to its 10th nearest target point (within 1.5 cells: a looser bound could not be used anyway)
  CK(ti.vals.ensure((size_t)g.n_points * 4));
  const float rk_thr = (float)(1.5 * g.h);
  k_target_rk<10><<<(unsigned)((g.n_points + 127) / 128), 128, 0, ctx->stream>>>(g, ti.cell_start.as<uint32_t>(), ti.sorted.as<F4>(), (long long)g.n_points, rk_thr * rk_thr, 2, ti.vals.as<float>());
  CKL();
  k_superrow_fill<<<(unsigned)((ncells + 255) / 256), 256, 0, ctx->stream>>>(g, ti.cell_start.as<uint32_t>(), ti.sorted.as<F4>(), ncells, ti.sstart.as<uint32_t>(), ti.srow.as<float>(),
                                                                            ti.srw.as<uint32_t>(), ti.vals.as<float>(), ti.srk.as<float>());
  CKL();
  ti.has_superrows = true;
  return PVB_OK;
}

template <bool REDUCE>
int launch_associate(pvb_ctx* ctx, int k, int n_tiles, AssocArgs a, bool ref_identity = false) {
  if (n_tiles == 0) return PVB_OK;
  // tuning knobs (PVB_MINB: resident blocks per SM the register allocation targets; PVB_MODE: 2 = buffered single-pass search with search-radius
  // hints (default), 1 = pruned two-pass walk, 0 = stage each tile's candidate rows in shared memory with TMA bulk copies + exhaustive walk;
  // PVB_STAGE=1 is the old spelling of PVB_MODE=0).  Defaults: fastest measured on B200.
  const int minb = ctx->tune_minb;
  int mode = ctx->tune_stage ? 0 : ctx->tune_mode;
  if (a.srow && a.sstart && a.out_nn_idx == nullptr) mode = 4;      // the caller's index has merged super-rows (dense mode)
  else if (mode == 4) mode = 2;
  const bool dbg = a.out_nn_idx != nullptr;
  if (k != 5 && k != 10) return ctx->fail(PVB_ERR_ARG, "k must be 5 or 10 (got %d)", k);
  a.stats = nullptr;
  a.prm.r0 = ctx->tune_r0;
  a.use_hint = ctx->tune_hints; a.flat_walk = ctx->tune_flat; a.use_static = ctx->tune_static; a.tight_frac = ctx->tune_tight;
  if (mode == 0 && !dbg) { CK(ctx->d_stats.ensure(16)); a.stats = ctx->d_stats.as<unsigned long long>(); }
  if (mode == 3) mode = 2;            // (the warp-cooperative experiment of round 2 was 2x slower than MODE 2 and has been removed)
#define PVB_LAUNCH(KK, MB, DBG, MD) do { if (ref_identity) k_associate<KK, REDUCE, MB, DBG, MD, true><<<n_tiles, kTile, 0, ctx->stream>>>(a); else k_associate<KK, REDUCE, MB, DBG, MD, false><<<n_tiles, kTile, 0, ctx->stream>>>(a); } while (0)
#define PVB_MINB_SWITCH(KK, MD) do { (void)minb; PVB_LAUNCH(KK, 6, false, MD); } while (0)      // 6 resident blocks per SM measured fastest (4 / 5 were compiled in round 1: profiles/r1f_sweep.log)
#define PVB_DISPATCH(KK)                                                                                   \
  if (dbg) { if (mode == 1) PVB_LAUNCH(KK, 4, true, 1); else PVB_LAUNCH(KK, 4, true, 2); }                 \
  else if (mode == 0) PVB_LAUNCH(KK, 6, false, 0);                                                         \
  else if (mode == 1) PVB_LAUNCH(KK, 6, false, 1);                                                         \
  else if (mode == 4) PVB_MINB_SWITCH(KK, 4);                                                              \
  else PVB_MINB_SWITCH(KK, 2);
  if (k == 10) { PVB_DISPATCH(10) } else { PVB_DISPATCH(5) }
#undef PVB_DISPATCH
#undef PVB_MINB_SWITCH
#undef PVB_LAUNCH
  CKL();
  return PVB_OK;
}

}  // namespace

extern "C" {

// ================================================================ lifecycle
int pvb_create(int device, pvb_ctx** out) {
  if (!out) return PVB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return PVB_ERR_CUDA;   // no CPU fallback by design
  if (device < 0 || device >= n) return PVB_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return PVB_ERR_CUDA;
  pvb_ctx* ctx = new pvb_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PVB_ERR_CUDA; }
  cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->bev0); cudaEventCreate(&ctx->bev1);
  if (const char* e = getenv("PVB_MINB")) ctx->tune_minb = atoi(e);
  if (const char* e = getenv("PVB_STAGE")) ctx->tune_stage = atoi(e) != 0;
  if (const char* e = getenv("PVB_MODE")) { ctx->tune_dense_mode = std::min(4, std::max(0, atoi(e))); ctx->tune_mode = ctx->tune_dense_mode == 4 ? 2 : ctx->tune_dense_mode; }
  if (const char* e = getenv("PVB_HINTS")) ctx->tune_hints = atoi(e) != 0;
  if (const char* e = getenv("PVB_FLAT")) ctx->tune_flat = atoi(e) != 0;
  if (const char* e = getenv("PVB_MORTON_BITS")) ctx->tune_morton_bits = std::min(16, std::max(10, atoi(e)));   // bits per axis of the source re-ordering: cells of 1024 m / 2^bits
  if (const char* e = getenv("PVB_R0")) ctx->tune_r0 = atoi(e) >= 2 ? 2 : 1;
  if (const char* e = getenv("PVB_CELLCAP")) ctx->tune_cellcap = std::max(1.0, atof(e));
  if (const char* e = getenv("PVB_HSCALE")) ctx->tune_hscale = ctx->tune_dense_hscale = std::max(0.1, atof(e));
  if (const char* e = getenv("PVB_REUSE")) ctx->tune_reuse = atoi(e) != 0;      // fresh uploads may keep the previous query permutation
  if (const char* e = getenv("PVB_CHUNKS")) ctx->tune_chunks = std::max(1, atoi(e));
  if (const char* e = getenv("PVB_CHUNK_MIN")) ctx->tune_chunk_min = std::max(1, atoi(e));
  if (const char* e = getenv("PVB_KEY64")) ctx->tune_key64 = atoi(e) != 0;      // force the 64-bit cell keys of very large grids (test hook)
  if (const char* e = getenv("PVB_TIGHT")) ctx->tune_tight = atof(e);
  if (const char* e = getenv("PVB_STATIC")) ctx->tune_static = atoi(e) != 0;
  if (const char* e = getenv("PVB_REORDER")) ctx->tune_reorder = std::max(0.0, atof(e));      // re-order the dense queries when a pose update may move a point by more than this many cells
  if (const char* e = getenv("PVB_DENSE_HSCALE")) ctx->tune_dense_hscale = std::max(0.1, atof(e));
  if (ctx->d_stats.ensure(16) == cudaSuccess) cudaMemset(ctx->d_stats.p, 0, 16);
  *out = ctx;
  return PVB_OK;
}

void pvb_destroy(pvb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf* dbs[] = {&ctx->d_prep, &ctx->d_wpose, &ctx->b_tile, &ctx->b_eref, &ctx->b_enei, &ctx->b_type, &ctx->b_norm, &ctx->b_huber, &ctx->b_consts, &ctx->b_orig_d, &ctx->b_r, &ctx->b_J,
                   &ctx->b_part, &ctx->b_esys, &ctx->b_tbegin, &ctx->f_pairs, &ctx->f_qtiles, &ctx->f_valid, &ctx->f_point, &ctx->f_plane, &ctx->f_nn_idx, &ctx->f_nn_d2,
                   &ctx->d_q_sorted, &ctx->d_q_orig, &ctx->d_pairs, &ctx->d_qtiles, &ctx->d_part, &ctx->d_sys, &ctx->d_tbegin, &ctx->d_valid, &ctx->d_point, &ctx->d_plane,
                   &ctx->d_res, &ctx->d_jac, &ctx->m_a, &ctx->m_b, &ctx->m_c, &ctx->m_d, &ctx->m_e, &ctx->v_local, &ctx->v_world, &ctx->v_misc, &ctx->v_M, &ctx->lt_misc, &ctx->lt_hold, &ctx->lt_kbase, &ctx->lt_cnt, &ctx->lt_base, &ctx->d_hint, &ctx->b_chunk, &ctx->d_chunk, &ctx->d_stats};
  for (DevBuf* b : dbs) b->release();
  PinBuf* pbs[] = {&ctx->h_pose, &ctx->h_r, &ctx->h_J, &ctx->h_esys, &ctx->fh_valid, &ctx->fh_point, &ctx->fh_plane, &ctx->dh_sys, &ctx->mh_a};
  for (PinBuf* b : pbs) b->release();
  ctx->f_tgt.release(); ctx->f_qry.release(); ctx->f_index.release(); ctx->f_corner.release(); ctx->f_cindex.release(); ctx->f_la.release(); ctx->f_lb.release();
  ctx->fh_la.release(); ctx->fh_lb.release(); ctx->d_tgt.release(); ctx->d_src.release(); ctx->d_index.release();
  for (cudaEvent_t e : ctx->chunk_ev) cudaEventDestroy(e);
  if (ctx->eval_done) cudaEventDestroy(ctx->eval_done);
  if (ctx->rmax_ev) cudaEventDestroy(ctx->rmax_ev);
  if (ctx->loc_ev) cudaEventDestroy(ctx->loc_ev);
  ctx->d_loc.release(); ctx->dh_loc.release();
  if (ctx->pose_ev) cudaEventDestroy(ctx->pose_ev);
  ctx->d_rmax2.release(); ctx->dh_rmax2.release();
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->sort_stream) { cudaStreamSynchronize(ctx->sort_stream); cudaStreamDestroy(ctx->sort_stream); }
  for (cudaEvent_t e : ctx->h2d_ev) cudaEventDestroy(e);
  if (ctx->bev0) cudaEventDestroy(ctx->bev0);
  if (ctx->bev1) cudaEventDestroy(ctx->bev1);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  ctx->p_cs.release(); ctx->p_index.release();
  DevBuf* sbs[] = {&ctx->s_H, &ctx->s_A, &ctx->s_g, &ctx->s_sc, &ctx->s_rhs, &ctx->s_y, &ctx->s_term, &ctx->s_con, &ctx->s_seg, &ctx->s_gcon, &ctx->s_gseg, &ctx->s_fail,
                   &ctx->s_bsr_rowptr, &ctx->s_bsr_col, &ctx->s_bsr_diag, &ctx->s_bsr_val, &ctx->s_px, &ctx->s_pr, &ctx->s_pz, &ctx->s_pp, &ctx->s_pq, &ctx->s_dmp, &ctx->s_minv, &ctx->s_part, &ctx->s_scal};
  for (DevBuf* b : sbs) b->release();
  ctx->sh_vec.release();
  if (ctx->ba_state && ctx->ba_free) ctx->ba_free(ctx->ba_state);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* pvb_last_error(const pvb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pvb_set_stream(pvb_ctx* ctx, void* s) {
  if (!ctx) return PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (s == nullptr) {
    if (!ctx->own_stream) { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return PVB_OK;
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s; ctx->own_stream = false;
  return PVB_OK;
}

int pvb_synchronize(pvb_ctx* ctx) { if (!ctx) return PVB_ERR_ARG; CK(cudaStreamSynchronize(ctx->stream)); return PVB_OK; }
long pvb_kernel_launches(const pvb_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* pvb_stream(const pvb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ================================================================ A. correspondence-list mode
int pvb_blocks_set(pvb_ctx* ctx, long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts, int nb) {
  if (!ctx) return PVB_ERR_ARG;
  if (n < 0 || nb <= 0 || (n > 0 && (!type || !ref || !nei || !normalize || !huber || !consts))) return ctx->fail(PVB_ERR_ARG, "pvb_blocks_set: bad arguments");
  CK(cudaSetDevice(ctx->device));
  for (long i = 0; i < n; ++i) {
    if (ref[i] < 0 || ref[i] >= nb || nei[i] < 0 || nei[i] >= nb) return ctx->fail(PVB_ERR_ARG, "block %ld: pose index out of range", i);
    if (type[i] < 0 || type[i] > PVB_LINE2LINE_ANGLE) return ctx->fail(PVB_ERR_ARG, "block %ld: unknown residual type %d", i, type[i]);
  }
  ctx->bn = n; ctx->nb = nb; ctx->b_has_rows = ctx->b_has_sys = false;
  // group rows by pose-graph edge (stable: rows of an edge keep their registration order)
  std::vector<uint32_t> order(n);
  for (long i = 0; i < n; ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    const long long ka = (long long)ref[a] * nb + nei[a], kb = (long long)ref[b] * nb + nei[b];
    return ka < kb;
  });
  ctx->b_orig = order;
  ctx->edge_ref.clear(); ctx->edge_nei.clear(); ctx->edge_tile_begin.clear();
  std::vector<BlockTile> tiles;
  std::vector<int> s_type(n), s_norm(n); std::vector<double> s_huber(n), s_consts((size_t)n * 12);
  if (ctx->g_edge_ref.empty()) {
    long i = 0;
    while (i < n) {
      long j = i;
      const int er = ref[order[i]], en = nei[order[i]];
      while (j < n && ref[order[j]] == er && nei[order[j]] == en) ++j;
      const int e = (int)ctx->edge_ref.size();
      ctx->edge_ref.push_back(er); ctx->edge_nei.push_back(en); ctx->edge_tile_begin.push_back((int)tiles.size());
      for (long s = i; s < j; s += kTile) tiles.push_back(BlockTile{e, (int)s, (int)std::min<long>(kTile, j - s), 0});
      i = j;
    }
  } else {
    // multi-GPU: the global edge list gives the layout; edges without local blocks keep an empty tile range (their systems reduce to zero)
    long i = 0;
    for (size_t e = 0; e < ctx->g_edge_ref.size(); ++e) {
      const int er = ctx->g_edge_ref[e], en = ctx->g_edge_nei[e];
      if (er >= nb || en >= nb) return ctx->fail(PVB_ERR_ARG, "edge list entry %zu refers to a pose block beyond %d", e, nb);
      long j = i;
      while (j < n && ref[order[j]] == er && nei[order[j]] == en) ++j;
      ctx->edge_ref.push_back(er); ctx->edge_nei.push_back(en); ctx->edge_tile_begin.push_back((int)tiles.size());
      for (long s = i; s < j; s += kTile) tiles.push_back(BlockTile{(int)e, (int)s, (int)std::min<long>(kTile, j - s), 0});
      i = j;
    }
    if (i != n) return ctx->fail(PVB_ERR_ARG, "block %u (edge %d -> %d) is not in the edge list set by pvb_blocks_set_edge_list", order[i], ref[order[i]], nei[order[i]]);
  }
  ctx->edge_tile_begin.push_back((int)tiles.size());
  ctx->b_tiles = (int)tiles.size();
  for (long r = 0; r < n; ++r) {
    const uint32_t o = order[r];
    s_type[r] = type[o]; s_norm[r] = normalize[o]; s_huber[r] = huber[o];
    for (int k = 0; k < 12; ++k) s_consts[(size_t)k * n + r] = consts[(size_t)o * 12 + k];
  }
  const int ne = (int)ctx->edge_ref.size();
  CK(ctx->b_tile.ensure(std::max<size_t>(16, tiles.size() * sizeof(BlockTile))));
  CK(ctx->b_eref.ensure(std::max<size_t>(16, (size_t)ne * 4))); CK(ctx->b_enei.ensure(std::max<size_t>(16, (size_t)ne * 4)));
  CK(ctx->b_tbegin.ensure((size_t)(ne + 1) * 4));
  CK(ctx->b_type.ensure(std::max<size_t>(16, (size_t)n * 4))); CK(ctx->b_norm.ensure(std::max<size_t>(16, (size_t)n * 4)));
  CK(ctx->b_huber.ensure(std::max<size_t>(16, (size_t)n * 8))); CK(ctx->b_consts.ensure(std::max<size_t>(16, (size_t)n * 96)));
  CK(ctx->b_orig_d.ensure(std::max<size_t>(16, (size_t)n * 4)));
  CK(ctx->b_r.ensure(std::max<size_t>(16, (size_t)n * 8))); CK(ctx->b_J.ensure(std::max<size_t>(16, (size_t)n * 96)));
  CK(ctx->b_part.ensure(std::max<size_t>(16, tiles.size() * 92 * 8))); CK(ctx->b_esys.ensure(std::max<size_t>(16, (size_t)ne * 92 * 8)));
  CK(ctx->h_esys.ensure(std::max<size_t>(16, (size_t)ne * 92 * 8)));
  if (n > 0) {
    CK(cudaMemcpyAsync(ctx->b_tile.p, tiles.data(), tiles.size() * sizeof(BlockTile), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_eref.p, ctx->edge_ref.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_enei.p, ctx->edge_nei.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_type.p, s_type.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_norm.p, s_norm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_huber.p, s_huber.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_consts.p, s_consts.data(), (size_t)n * 96, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_orig_d.p, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->b_tbegin.p, ctx->edge_tile_begin.data(), (size_t)(ne + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

int pvb_blocks_evaluate(pvb_ctx* ctx, const double* poses, int want_rows, int want_system) {
  if (!ctx || !poses) return PVB_ERR_ARG;
  if (ctx->nb <= 0) return ctx->fail(PVB_ERR_STATE, "pvb_blocks_set has not been called");
  CK(cudaSetDevice(ctx->device));
  int rc = upload_poses(ctx, poses, ctx->nb, false);
  if (rc) return rc;
  const long long n = ctx->bn;
  const int ne = (int)ctx->edge_ref.size();
  ctx->b_has_rows = ctx->b_has_sys = false;
  if (n > 0) {
    EvalArgs a;
    a.tiles = ctx->b_tile.as<BlockTile>(); a.edge_ref = ctx->b_eref.as<int>(); a.edge_nei = ctx->b_enei.as<int>();
    a.type = ctx->b_type.as<int>(); a.normalize = ctx->b_norm.as<int>(); a.huber = ctx->b_huber.as<double>(); a.consts = ctx->b_consts.as<double>();
    a.orig = ctx->b_orig_d.as<uint32_t>(); a.n = n; a.prep = ctx->d_prep.as<PosePrep>();
    a.out_r = want_rows ? ctx->b_r.as<double>() : nullptr; a.out_J = want_rows ? ctx->b_J.as<double>() : nullptr;
    a.partials = want_system ? ctx->b_part.as<double>() : nullptr;
    a.raw_rows = want_rows == 2 ? 1 : 0;
    CK(cudaEventRecord(ctx->bev0, ctx->stream));
    k_eval_blocks<<<ctx->b_tiles, kTile, 0, ctx->stream>>>(a);
    CKL();
    CK(cudaEventRecord(ctx->bev1, ctx->stream));
    ctx->bev_valid = true;
    if (want_rows) {
      CK(ctx->h_r.ensure((size_t)n * 8)); CK(ctx->h_J.ensure((size_t)n * 96));
      CK(cudaMemcpyAsync(ctx->h_r.p, ctx->b_r.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaMemcpyAsync(ctx->h_J.p, ctx->b_J.p, (size_t)n * 96, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  if (want_system && ne > 0) {       // also with no local blocks (a rank of a sharded pose graph that owns no residual): its systems are zero
    CK(ctx->b_chunk.ensure((size_t)ne * kSumChunks * 92 * 8));
    k_sum_partials<92><<<dim3(ne, kSumChunks), 256, 0, ctx->stream>>>(ctx->b_part.as<double>(), ctx->b_tbegin.as<int>(), ctx->b_chunk.as<double>());
    CKL();
    k_sum_chunks<92><<<ne, 96, 0, ctx->stream>>>(ctx->b_chunk.as<double>(), ctx->b_esys.as<double>());
    CKL();
    if (ctx->reduce_hook) ctx->reduce_hook(ctx->reduce_user, ctx->b_esys.as<double>(), (long)ne * 92, (void*)ctx->stream);   // the ONE exchange step (allreduce) of an evaluation
    CK(cudaMemcpyAsync(ctx->h_esys.p, ctx->b_esys.p, (size_t)ne * 92 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->b_has_rows = want_rows != 0; ctx->b_has_sys = want_system != 0;
  return PVB_OK;
}

int pvb_blocks_set_edge_list(pvb_ctx* ctx, int n_edges, const int* ref, const int* nei) {
  if (!ctx || n_edges < 0 || (n_edges > 0 && (!ref || !nei))) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_blocks_set_edge_list: bad arguments") : PVB_ERR_ARG;
  for (int e = 0; e < n_edges; ++e) {
    if (ref[e] < 0 || nei[e] < 0) return ctx->fail(PVB_ERR_ARG, "edge %d: negative pose index", e);
    if (e > 0 && !(ref[e - 1] < ref[e] || (ref[e - 1] == ref[e] && nei[e - 1] < nei[e]))) return ctx->fail(PVB_ERR_ARG, "edge list must be sorted by (ref, nei) and unique (entry %d)", e);
  }
  ctx->g_edge_ref.assign(ref, ref + n_edges); ctx->g_edge_nei.assign(nei, nei + n_edges);
  return PVB_OK;
}

int pvb_blocks_set_reduce_hook(pvb_ctx* ctx, pvb_reduce_hook hook, void* user) {
  if (!ctx) return PVB_ERR_ARG;
  ctx->reduce_hook = hook; ctx->reduce_user = user;
  return PVB_OK;
}

int pvb_blocks_kernel_time_ms(pvb_ctx* ctx, float* ms) {
  if (!ctx || !ms) return PVB_ERR_ARG;
  if (!ctx->bev_valid) return ctx->fail(PVB_ERR_STATE, "no blocks evaluate has run");
  CK(cudaEventSynchronize(ctx->bev1));
  CK(cudaEventElapsedTime(ms, ctx->bev0, ctx->bev1));
  return PVB_OK;
}

const double* pvb_blocks_residuals(const pvb_ctx* ctx) { return (ctx && ctx->b_has_rows) ? ctx->h_r.as<double>() : nullptr; }
const double* pvb_blocks_jacobians(const pvb_ctx* ctx) { return (ctx && ctx->b_has_rows) ? ctx->h_J.as<double>() : nullptr; }

int pvb_blocks_cost(const pvb_ctx* ctx, double* cost, long* nres) {
  if (!ctx || !ctx->b_has_sys) return PVB_ERR_STATE;
  double c = 0, n = 0;
  const double* s = ctx->h_esys.as<double>();
  for (size_t e = 0; e < ctx->edge_ref.size(); ++e) { c += s[e * 92 + 90]; n += s[e * 92 + 91]; }
  if (cost) *cost = c;
  if (nres) *nres = (long)n;
  return PVB_OK;
}

int pvb_blocks_num_edges(const pvb_ctx* ctx) { return ctx ? (int)ctx->edge_ref.size() : PVB_ERR_ARG; }
int pvb_blocks_edges(const pvb_ctx* ctx, int* ref, int* nei) {
  if (!ctx) return PVB_ERR_ARG;
  for (size_t e = 0; e < ctx->edge_ref.size(); ++e) { ref[e] = ctx->edge_ref[e]; nei[e] = ctx->edge_nei[e]; }
  return PVB_OK;
}
const double* pvb_blocks_edge_systems_ptr(const pvb_ctx* ctx) { return (ctx && ctx->b_has_sys) ? ctx->h_esys.as<double>() : nullptr; }
int pvb_blocks_edge_systems(const pvb_ctx* ctx, double* out) {
  if (!ctx || !ctx->b_has_sys) return PVB_ERR_STATE;
  memcpy(out, ctx->h_esys.p, ctx->edge_ref.size() * 92 * 8);
  return PVB_OK;
}

static double assemble_dense(const pvb_ctx* ctx, double* H, double* g) {
  const int D = 6 * ctx->nb;
  if (H) std::fill(H, H + (size_t)D * D, 0.0);
  if (g) std::fill(g, g + D, 0.0);
  double cost = 0;
  const double* s = ctx->h_esys.as<double>();
  for (size_t e = 0; e < ctx->edge_ref.size(); ++e) {
    const double* S = s + e * 92;
    cost += S[90];
    if (!H && !g) continue;
    const int o[2] = {6 * ctx->edge_ref[e], 6 * ctx->edge_nei[e]};
    int q = 0;
    for (int a = 0; a < 12; ++a)
      for (int b = a; b < 12; ++b, ++q) {
        const int ia = o[a / 6] + a % 6, ib = o[b / 6] + b % 6;
        if (H) {
          if (ia == ib) H[(size_t)ia * D + ib] += S[q] * (a == b ? 1.0 : 2.0);   // ref == nei edge: both (a,b) and (b,a) land on the diagonal
          else { H[(size_t)ia * D + ib] += S[q]; H[(size_t)ib * D + ia] += S[q]; }
        }
      }
    if (g) for (int a = 0; a < 12; ++a) g[o[a / 6] + a % 6] += S[78 + a];
  }
  return cost;
}

int pvb_blocks_dense_system(const pvb_ctx* ctx, double* H, double* g, double* cost) {
  if (!ctx || !ctx->b_has_sys) return PVB_ERR_STATE;
  const double c = assemble_dense(ctx, H, g);
  if (cost) *cost = c;
  return PVB_OK;
}

// ---- device linear algebra of the trust-region step --------------------------------------------------------------------------------
// factor the N x N matrix in ctx->s_A (lower triangle, in place) and solve with the right-hand side in ctx->s_rhs; *ok = 0 when a pivot fails
constexpr size_t kTrsmSmem = 2 * kNB * (kNB + 1) * sizeof(double);
int pvb_internal_factor_solve(pvb_ctx* ctx, int N, bool* ok) {
  if (!ctx->solver_attr_set) {       // function attributes are per device: one opt-in per context (= per device), not per process
    CK(cudaFuncSetAttribute(k_trsm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
    CK(cudaFuncSetAttribute(k_syrk_update_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
    ctx->solver_attr_set = true;
  }
  CK(cudaMemsetAsync(ctx->s_fail.p, 0, 4, ctx->stream));
  CK(ctx->s_y.ensure((size_t)N * 8));
  for (int k0 = 0; k0 < N; k0 += kNB) {
    k_potrf_diag<<<1, 256, 0, ctx->stream>>>(ctx->s_A.as<double>(), N, k0, ctx->s_fail.as<int>());
    CKL();
    const int T = (N - k0 - kNB) / kNB;
    if (T > 0) {
      k_trsm_panel<<<T, 256, kTrsmSmem, ctx->stream>>>(ctx->s_A.as<double>(), N, k0);
      CKL();
    }
    k_fwd_step<<<std::max(T, 1), 256, 0, ctx->stream>>>(ctx->s_A.as<double>(), N, k0, ctx->s_rhs.as<double>(), ctx->s_y.as<double>());   // L y = rhs, block by block
    CKL();
    if (T > 0) {
      const int Ti = (T + 1) / 2;                                   // 128-row tiles; row tile ti owns 2 ti + 2 column tiles of 64
      k_syrk_update_mma<<<Ti * (Ti + 1), 256, kSyrkSmem, ctx->stream>>>(ctx->s_A.as<double>(), N, k0);
      CKL();
    }
  }
  for (int k0 = N - kNB; k0 >= 0; k0 -= kNB) {                     // L^T x = y; x lands in s_rhs
    k_bwd_step<<<std::max(1, (k0 + 255) / 256), 256, 0, ctx->stream>>>(ctx->s_A.as<double>(), N, k0, ctx->s_y.as<double>(), ctx->s_rhs.as<double>());
    CKL();
  }
  int fail = 0;
  CK(cudaMemcpyAsync(&fail, ctx->s_fail.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *ok = fail == 0;
  return PVB_OK;
}

// free-block index map + contribution lists of every edge system to the dense matrix / gradient (sorted by destination, then edge)
static int solver_prepare(pvb_ctx* ctx, const unsigned char* is_const) {
  const int nb = ctx->nb, ne = (int)ctx->edge_ref.size();
  std::vector<int> fidx(nb, -1);
  int nf = 0;
  for (int b = 0; b < nb; ++b) if (!is_const || !is_const[b]) fidx[b] = nf++;
  ctx->s_n = 6 * nf; ctx->s_N = ((ctx->s_n + kNB - 1) / kNB) * kNB;
  if (ctx->s_n == 0) return PVB_OK;
  std::vector<HContrib> con, gcon;
  for (int e = 0; e < ne; ++e) {
    const int blk[2] = {fidx[ctx->edge_ref[e]], fidx[ctx->edge_nei[e]]};
    for (int kr = 0; kr < 2; ++kr) {
      if (blk[kr] < 0) continue;
      gcon.push_back(HContrib{blk[kr], 0, e, kr});
      for (int kc = 0; kc < 2; ++kc) if (blk[kc] >= 0) con.push_back(HContrib{blk[kr], blk[kc], e, kr * 2 + kc});
    }
  }
  auto less = [](const HContrib& a, const HContrib& b) { if (a.dest_r != b.dest_r) return a.dest_r < b.dest_r; if (a.dest_c != b.dest_c) return a.dest_c < b.dest_c;
                                                          if (a.edge != b.edge) return a.edge < b.edge; return a.kind < b.kind; };
  std::sort(con.begin(), con.end(), less); std::sort(gcon.begin(), gcon.end(), less);
  std::vector<int> seg, gseg;
  for (size_t i = 0; i < con.size(); ++i) if (i == 0 || con[i].dest_r != con[i - 1].dest_r || con[i].dest_c != con[i - 1].dest_c) seg.push_back((int)i);
  seg.push_back((int)con.size());
  for (size_t i = 0; i < gcon.size(); ++i) if (i == 0 || gcon[i].dest_r != gcon[i - 1].dest_r) gseg.push_back((int)i);
  gseg.push_back((int)gcon.size());
  ctx->s_ndest = (int)seg.size() - 1; ctx->s_ngdest = (int)gseg.size() - 1;
  const size_t N = (size_t)ctx->s_N;
  if (ctx->pcg_active) {
    // a free pose block without any residual has no diagonal block to precondition with: such (degenerate) graphs take the dense path, whose damping keeps them solvable
    std::vector<char> has_diag(nf, 0);
    for (int d = 0; d < ctx->s_ndest; ++d) { const HContrib& c = con[seg[d]]; if (c.dest_r == c.dest_c) has_diag[c.dest_r] = 1; }
    for (int b = 0; b < nf; ++b) if (!has_diag[b]) { ctx->pcg_active = false; break; }
  }
  if (!ctx->pcg_active) { CK(ctx->s_H.ensure(N * N * 8)); CK(ctx->s_A.ensure(N * N * 8)); }
  else {
    // block-sparse form: one 6x6 block per destination (row block, column block), rows sorted (the order of `con`)
    const int nd = ctx->s_ndest;
    std::vector<int> rowptr(nf + 1, 0), col(nd), diag(nf, -1);
    for (int d = 0; d < nd; ++d) { const HContrib& c = con[seg[d]]; rowptr[c.dest_r + 1]++; col[d] = c.dest_c; if (c.dest_r == c.dest_c) diag[c.dest_r] = d; }
    for (int b = 0; b < nf; ++b) rowptr[b + 1] += rowptr[b];
    ctx->s_nfb = nf; ctx->s_nblk = nd;
    const int ncta = (ctx->s_n + kPcgThreads - 1) / kPcgThreads;
    CK(ctx->s_bsr_rowptr.ensure((size_t)(nf + 1) * 4)); CK(ctx->s_bsr_col.ensure(std::max<size_t>(16, (size_t)nd * 4))); CK(ctx->s_bsr_diag.ensure((size_t)nf * 4));
    CK(ctx->s_bsr_val.ensure(std::max<size_t>(16, (size_t)nd * 36 * 8))); CK(ctx->s_minv.ensure((size_t)nf * 36 * 8));
    for (DevBuf* b : {&ctx->s_px, &ctx->s_pr, &ctx->s_pz, &ctx->s_pq, &ctx->s_dmp}) CK(b->ensure(N * 8));
    CK(ctx->s_pp.ensure(2 * N * 8));
    CK(ctx->s_part.ensure((size_t)(2 * ncta + (nf + kSpmvRows - 1) / kSpmvRows) * 8)); CK(ctx->s_scal.ensure(64));
    CK(cudaMemcpyAsync(ctx->s_bsr_rowptr.p, rowptr.data(), rowptr.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (nd) CK(cudaMemcpyAsync(ctx->s_bsr_col.p, col.data(), col.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->s_bsr_diag.p, diag.data(), diag.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(ctx->s_g.ensure(N * 8)); CK(ctx->s_sc.ensure(N * 8)); CK(ctx->s_rhs.ensure(N * 8)); CK(ctx->s_term.ensure(N * 8)); CK(ctx->s_fail.ensure(16));
  CK(ctx->s_con.ensure(std::max<size_t>(16, con.size() * sizeof(HContrib)))); CK(ctx->s_seg.ensure(seg.size() * 4));
  CK(ctx->s_gcon.ensure(std::max<size_t>(16, gcon.size() * sizeof(HContrib)))); CK(ctx->s_gseg.ensure(gseg.size() * 4));
  CK(ctx->sh_vec.ensure(4 * N * 8));
  if (!con.empty()) CK(cudaMemcpyAsync(ctx->s_con.p, con.data(), con.size() * sizeof(HContrib), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->s_seg.p, seg.data(), seg.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!gcon.empty()) CK(cudaMemcpyAsync(ctx->s_gcon.p, gcon.data(), gcon.size() * sizeof(HContrib), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->s_gseg.p, gseg.data(), gseg.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

// dense H (free unknowns) and gradient from the edge systems of the last evaluate, on the device; g also copied to the host (h_g)
static int solver_assemble(pvb_ctx* ctx, double* h_g) {
  const size_t N = (size_t)ctx->s_N;
  CK(cudaMemsetAsync(ctx->s_g.p, 0, N * 8, ctx->stream));
  if (ctx->pcg_active) {
    if (ctx->s_ndest) { k_assemble_bsr<<<ctx->s_ndest, 64, 0, ctx->stream>>>(ctx->s_con.as<HContrib>(), ctx->s_seg.as<int>(), ctx->b_esys.as<double>(), ctx->s_bsr_val.as<double>()); CKL(); }
  } else {
  CK(cudaMemsetAsync(ctx->s_H.p, 0, N * N * 8, ctx->stream));
  if (ctx->s_ndest) { k_assemble_H<<<ctx->s_ndest, 64, 0, ctx->stream>>>(ctx->s_con.as<HContrib>(), ctx->s_seg.as<int>(), ctx->b_esys.as<double>(), ctx->s_N, ctx->s_H.as<double>()); CKL(); }
  }
  if (ctx->s_ngdest) { k_assemble_g<<<ctx->s_ngdest, 32, 0, ctx->stream>>>(ctx->s_gcon.as<HContrib>(), ctx->s_gseg.as<int>(), ctx->b_esys.as<double>(), ctx->s_g.as<double>()); CKL(); }
  CK(cudaMemcpyAsync(h_g, ctx->s_g.p, (size_t)ctx->s_n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

// One trust-region step by preconditioned conjugate gradients on the block-sparse matrix: (S H S + D) y = -S g, y -> ctx->s_rhs; the model terms -> ctx->s_term.
// The host reads the residual norm every 16 iterations (one small D2H); the CG scalars themselves stay on the device.
static int pcg_solve(pvb_ctx* ctx, double radius, bool* ok) {
  const int n = ctx->s_n, nfb = ctx->s_nfb, ncu = (n + kPcgThreads - 1) / kPcgThreads, ncs = (nfb + kSpmvRows - 1) / kSpmvRows;
  double* part = ctx->s_part.as<double>();
  double *part_pq = part, *part_rz = part + ncs, *part_rr = part + ncs + ncu;
  double* scal = ctx->s_scal.as<double>();                       // 4 doubles, then the two arrival counters of grid_total
  unsigned* counters = reinterpret_cast<unsigned*>(scal + 4);
  const int* rowptr = ctx->s_bsr_rowptr.as<int>(); const int* col = ctx->s_bsr_col.as<int>(); const double* val = ctx->s_bsr_val.as<double>();
  double *x = ctx->s_px.as<double>(), *r = ctx->s_pr.as<double>(), *z = ctx->s_pz.as<double>(), *q = ctx->s_pq.as<double>();
  double* pbuf[2] = {ctx->s_pp.as<double>(), ctx->s_pp.as<double>() + ctx->s_N};      // the direction ping-pongs between two buffers
  CK(cudaMemsetAsync(ctx->s_fail.p, 0, 4, ctx->stream));
  CK(cudaMemsetAsync(scal, 0, 48, ctx->stream));
  k_bsr_prepare_step<<<(nfb + 127) / 128, 128, 0, ctx->stream>>>(val, ctx->s_bsr_diag.as<int>(), ctx->s_g.as<double>(), ctx->s_sc.as<double>(), nfb, radius, ctx->s_rhs.as<double>(),
                                                               ctx->s_dmp.as<double>(), ctx->s_minv.as<double>(), ctx->s_fail.as<int>());
  CKL();
  k_pcg_update<<<ncu, kPcgThreads, 0, ctx->stream>>>(1, 0, n, ctx->s_rhs.as<double>(), ctx->s_minv.as<double>(), scal, x, r, z, pbuf[0], q, part_rz, part_rr, counters + 1);
  CKL();
  double h_scal[4] = {0, 0, 0, 0};
  int fail = 0;
  CK(cudaMemcpyAsync(h_scal, scal, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&fail, ctx->s_fail.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (fail) { *ok = false; return PVB_OK; }
  const double rr0 = h_scal[2];
  int it = 0;
  ctx->pcg_solves++;
  while (rr0 > 0.0 && it < ctx->pcg_max_it) {
    for (int k = 0; k < 16; ++k, ++it) {
      k_pcg_spmv<<<ncs, kSpmvRows * 32, 0, ctx->stream>>>(0, it, nfb, rowptr, col, val, ctx->s_sc.as<double>(), ctx->s_dmp.as<double>(), scal, z, pbuf[it & 1], pbuf[(it + 1) & 1], q,
                                                            part_pq, counters);
      CKL();
      k_pcg_update<<<ncu, kPcgThreads, 0, ctx->stream>>>(0, it, n, ctx->s_rhs.as<double>(), ctx->s_minv.as<double>(), scal, x, r, z, pbuf[(it + 1) & 1], q, part_rz, part_rr, counters + 1);
      CKL();
    }
    ctx->pcg_iterations += 16;
    CK(cudaMemcpyAsync(h_scal, scal, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!(h_scal[2] == h_scal[2])) { *ok = false; return PVB_OK; }      // NaN: breakdown
    if (h_scal[2] <= ctx->pcg_tol * ctx->pcg_tol * rr0) break;
  }
  // y = x; model terms need S H S y (no damping)
  CK(cudaMemcpyAsync(ctx->s_rhs.p, x, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  k_pcg_spmv<<<ncs, kSpmvRows * 32, 0, ctx->stream>>>(1, 0, nfb, rowptr, col, val, ctx->s_sc.as<double>(), nullptr, scal, x, nullptr, nullptr, q, nullptr, nullptr);
  CKL();
  k_bsr_model_terms<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->s_g.as<double>(), ctx->s_sc.as<double>(), x, q, n, ctx->s_term.as<double>());
  CKL();
  *ok = true;
  return PVB_OK;
}

// The loop of pvb::solve_lm (pvb_host.hpp) with the linear algebra on the device.  Same control flow and constants; the candidate
// evaluation already leaves the edge systems of the accepted point on the device, so an accepted step costs one evaluation, not two.
static int solve_lm_device(pvb_ctx* ctx, double* poses, const unsigned char* is_const, const LMOptions& opt, LMSummary& S) {
  const int nb = ctx->nb, D = 6 * nb;
  int rc = solver_prepare(ctx, is_const); if (rc) return rc;
  const int n = ctx->s_n, N = ctx->s_N;
  std::vector<int> fi;
  for (int b = 0; b < nb; ++b) if (!is_const || !is_const[b]) for (int k = 0; k < 6; ++k) fi.push_back(6 * b + k);
  auto evaluate = [&](const double* x, double* cost) -> int {
    const int r = pvb_blocks_evaluate(ctx, x, 0, 1); if (r) return r;
    return pvb_blocks_cost(ctx, cost, nullptr);
  };
  double cost = 0;
  rc = evaluate(poses, &cost); if (rc) return rc;
  bool stale = false;
  S = LMSummary(); S.initial_cost = cost;
  if (n == 0) { S.final_cost = cost; S.termination = 2; return PVB_OK; }
  double* hv = ctx->sh_vec.as<double>();
  double *gs = hv, *sc = hv + N, *y = hv + 2 * N, *term = hv + 3 * N;
  rc = solver_assemble(ctx, gs); if (rc) return rc;
  if (ctx->pcg_active) k_bsr_jacobi_scale<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->s_bsr_val.as<double>(), ctx->s_bsr_diag.as<int>(), n, ctx->s_sc.as<double>());
  else k_jacobi_scale<<<(N + 255) / 256, 256, 0, ctx->stream>>>(ctx->s_H.as<double>(), n, N, ctx->s_sc.as<double>());
  CKL();
  CK(cudaMemcpyAsync(sc, ctx->s_sc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  auto gmax = [&]() { double m = 0; for (int i = 0; i < n; ++i) m = std::max(m, std::fabs(gs[i])); return m; };
  double radius = 1e4, decrease = 2.0;
  int invalid = 0;
  std::vector<double> cand(D);
  if (gmax() <= opt.gradient_tolerance) { S.final_cost = cost; S.termination = 2; return PVB_OK; }
  for (int it = 1; it <= opt.max_iterations; ++it) {
    S.iterations = it;
    bool ok = false;
    if (ctx->pcg_active) { rc = pcg_solve(ctx, radius, &ok); if (rc) return rc; }
    else {
      k_build_damped<<<dim3((N + 255) / 256, N), 256, 0, ctx->stream>>>(ctx->s_H.as<double>(), ctx->s_g.as<double>(), ctx->s_sc.as<double>(), n, N, radius, ctx->s_A.as<double>(),
                                                                       ctx->s_rhs.as<double>());
      CKL();
      rc = pvb_internal_factor_solve(ctx, N, &ok); if (rc) return rc;
    }
    double model = 0;
    if (ok) {
      if (!ctx->pcg_active) {
        k_model_terms<<<(n + 7) / 8, 256, 0, ctx->stream>>>(ctx->s_H.as<double>(), ctx->s_g.as<double>(), ctx->s_sc.as<double>(), ctx->s_rhs.as<double>(), n, N, ctx->s_term.as<double>());
        CKL();
      }
      CK(cudaMemcpyAsync(y, ctx->s_rhs.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaMemcpyAsync(term, ctx->s_term.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (int i = 0; i < n; ++i) model -= term[i];             // fixed summation order
      ok = model > 0.0;
    }
    if (!ok) { radius *= 0.5; S.unsuccessful++; if (++invalid >= 5 || radius < 1e-32) { S.termination = 4; break; } continue; }
    invalid = 0;
    double sn = 0, xn = 0;
    std::copy(poses, poses + D, cand.begin());
    for (int i = 0; i < n; ++i) { const double d = y[i] * sc[i]; cand[fi[i]] += d; sn += d * d; xn += poses[fi[i]] * poses[fi[i]]; }
    sn = std::sqrt(sn); xn = std::sqrt(xn);
    double new_cost = 0;
    rc = evaluate(cand.data(), &new_cost); if (rc) return rc;
    stale = true;                                               // the device now holds the CANDIDATE's edge systems
    if (sn <= opt.parameter_tolerance * (xn + opt.parameter_tolerance)) { S.termination = 3; break; }
    const double change = cost - new_cost;
    if (std::fabs(change) <= opt.function_tolerance * cost) { S.termination = 1; break; }
    const double rho = change / model;
    if (rho > 1e-3) {
      std::copy(cand.begin(), cand.end(), poses);
      cost = new_cost;
      rc = solver_assemble(ctx, gs); if (rc) return rc;        // edge systems of the accepted point are already on the device
      stale = false;
      S.successful++;
      radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease = 2.0;
      if (gmax() <= opt.gradient_tolerance) { S.termination = 2; break; }
    } else {
      radius /= decrease; decrease *= 2.0; S.unsuccessful++;
      if (radius < 1e-32) { S.termination = 4; break; }
    }
  }
  S.final_cost = cost;
  // a loop that ends on a rejected / terminating candidate leaves that candidate's systems behind: they do not describe the returned poses, so
  // pvb_blocks_cost / pvb_blocks_edge_systems / pvb_blocks_dense_system report "no evaluation" until the caller evaluates again
  if (stale) ctx->b_has_sys = ctx->b_has_rows = false;
  return PVB_OK;
}

// ---- internal entry points for the joint camera-LiDAR solve (pvb_ba.cu); declared in pvb_ctx.hpp -------------------------------------
int pvb_internal_solver_prepare(pvb_ctx* ctx, const unsigned char* is_const_block) { return solver_prepare(ctx, is_const_block); }
int pvb_internal_solver_assemble(pvb_ctx* ctx, double* h_g) { return solver_assemble(ctx, h_g); }
int pvb_internal_jacobi_scale(pvb_ctx* ctx) {
  k_jacobi_scale<<<(ctx->s_N + 255) / 256, 256, 0, ctx->stream>>>(ctx->s_H.as<double>(), ctx->s_n, ctx->s_N, ctx->s_sc.as<double>());
  CKL();
  return PVB_OK;
}
int pvb_internal_build_damped(pvb_ctx* ctx, double radius) {
  const int N = ctx->s_N;
  k_build_damped<<<dim3((N + 255) / 256, N), 256, 0, ctx->stream>>>(ctx->s_H.as<double>(), ctx->s_g.as<double>(), ctx->s_sc.as<double>(), ctx->s_n, N, radius, ctx->s_A.as<double>(),
                                                                   ctx->s_rhs.as<double>());
  CKL();
  return PVB_OK;
}

int pvb_blocks_set_linear_solver(pvb_ctx* ctx, int kind) {
  if (!ctx || kind < PVB_SOLVER_AUTO || kind > PVB_SOLVER_PCG) return ctx ? ctx->fail(PVB_ERR_ARG, "unknown linear solver %d", kind) : PVB_ERR_ARG;
  ctx->solver_kind = kind;
  return PVB_OK;
}

// x = A^-1 b for a symmetric positive definite row-major A (n x n) with the device Cholesky of the LM loop (parity / benchmark entry)
int pvb_blocks_pcg_stats(const pvb_ctx* ctx, long* solves, long* iterations) {
  if (!ctx) return PVB_ERR_ARG;
  if (solves) *solves = ctx->pcg_solves;
  if (iterations) *iterations = ctx->pcg_iterations;
  return PVB_OK;
}

int pvb_cholesky_solve(pvb_ctx* ctx, const double* A, int n, const double* b, double* x, float* factor_ms) {
  if (!ctx || !A || !b || !x || n <= 0) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_cholesky_solve: bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const int N = ((n + kNB - 1) / kNB) * kNB;
  CK(ctx->s_A.ensure((size_t)N * N * 8)); CK(ctx->s_rhs.ensure((size_t)N * 8)); CK(ctx->s_fail.ensure(16));
  std::vector<double> pad((size_t)N * N, 0.0), rhs(N, 0.0);
  for (int i = 0; i < N; ++i) { if (i < n) { memcpy(&pad[(size_t)i * N], A + (size_t)i * n, (size_t)n * 8); rhs[i] = b[i]; } else pad[(size_t)i * N + i] = 1.0; }
  CK(cudaMemcpyAsync(ctx->s_A.p, pad.data(), pad.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->s_rhs.p, rhs.data(), rhs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->bev0, ctx->stream));
  bool ok = false;
  int rc = pvb_internal_factor_solve(ctx, N, &ok); if (rc) return rc;
  CK(cudaEventRecord(ctx->bev1, ctx->stream));
  CK(cudaMemcpyAsync(rhs.data(), ctx->s_rhs.p, (size_t)N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->bev_valid = false;
  if (factor_ms) CK(cudaEventElapsedTime(factor_ms, ctx->bev0, ctx->bev1));
  if (!ok) return ctx->fail(PVB_ERR_ARG, "pvb_cholesky_solve: matrix is not positive definite");
  memcpy(x, rhs.data(), (size_t)n * 8);
  return PVB_OK;
}

int pvb_blocks_solve_lm(pvb_ctx* ctx, double* poses, const unsigned char* is_const, int max_iterations, double* summary6) {
  if (!ctx || !poses) return PVB_ERR_ARG;
  if (ctx->nb <= 0) return ctx->fail(PVB_ERR_STATE, "pvb_blocks_set has not been called");
  int rc_inner = PVB_OK;
  EvalFn eval = [&](const double* x, double* H, double* g) -> double {
    const int rc = pvb_blocks_evaluate(ctx, x, 0, 1);
    if (rc) { rc_inner = rc; return 0.0; }
    return assemble_dense(ctx, H, g);
  };
  LMOptions opt; opt.max_iterations = max_iterations;
  int n_free = 0;
  for (int b = 0; b < ctx->nb; ++b) if (!is_const || !is_const[b]) n_free += 6;
  // SetOptionsLidar picks the linear solver by problem size (Optimization.cpp:647-662); here: device Cholesky once the dense host
  // algebra would dominate (a 6x6 .. ~40-block system is faster on the host than ~3 launches per 64 columns)
  // SetOptionsLidar leaves the dense solver at 50 frames (sparse, exact) and goes iterative above 2000 (util/Optimization.cpp:647-662); here the block-sparse PCG,
  // converged to rounding, takes over from 500 pose blocks: measured on Floor (1593 frames) same LM steps, poses equal to 1e-14, 0.14 s instead of 0.37 s
  const bool pcg = ctx->solver_kind == PVB_SOLVER_PCG || (ctx->solver_kind == PVB_SOLVER_AUTO && ctx->nb > 500);
  const bool on_device = pcg || ctx->solver_kind == PVB_SOLVER_DEVICE || (ctx->solver_kind == PVB_SOLVER_AUTO && n_free >= 256);
  LMSummary S;
  if (on_device) {
    CK(cudaSetDevice(ctx->device));
    ctx->pcg_active = pcg && n_free > 0;
    const int rc = solve_lm_device(ctx, poses, is_const, opt, S);
    ctx->pcg_active = false;
    if (rc) return rc;
  }
  else S = solve_lm(eval, poses, ctx->nb, is_const, opt);
  if (rc_inner) return rc_inner;
  if (summary6) { summary6[0] = S.initial_cost; summary6[1] = S.final_cost; summary6[2] = S.iterations; summary6[3] = S.successful; summary6[4] = S.unsuccessful; summary6[5] = S.termination; }
  return PVB_OK;
}

// ================================================================ B. frames
int pvb_frames_set(pvb_ctx* ctx, int n_frames, const pvb_frame* frames) {
  if (!ctx || n_frames <= 0 || !frames) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_frames_set: bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  std::vector<const float*> tp(n_frames), qp(n_frames); std::vector<int> tc(n_frames), qc(n_frames), blocks(n_frames);
  for (int f = 0; f < n_frames; ++f) { tp[f] = frames[f].surf_target; tc[f] = frames[f].n_target; qp[f] = frames[f].surf_query; qc[f] = frames[f].n_query; blocks[f] = f; }
  int rc = set_cloudset(ctx, ctx->f_tgt, tp, tc, blocks, 256); if (rc) return rc;
  rc = set_cloudset(ctx, ctx->f_qry, qp, qc, blocks, 256); if (rc) return rc;
  ctx->n_frames = n_frames; ctx->f_index.built = false;
  return PVB_OK;
}

static int frames_run(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei, const pvb_assoc_params* prm, bool want_nn, long long* total_slots,
                      std::vector<int>* slot_edge_begin) {
  if (ctx->n_frames <= 0) return ctx->fail(PVB_ERR_STATE, "pvb_frames_set has not been called");
  if (!prm || (prm->k != 5 && prm->k != 10)) return ctx->fail(PVB_ERR_ARG, "k must be 5 or 10");
  for (int e = 0; e < n_edges; ++e)
    if (ref[e] < 0 || ref[e] >= ctx->n_frames || nei[e] < 0 || nei[e] >= ctx->n_frames) return ctx->fail(PVB_ERR_ARG, "edge %d out of range", e);
  int rc = upload_poses(ctx, poses, ctx->n_frames, false); if (rc) return rc;
  rc = build_target_index(ctx, ctx->f_tgt, ctx->f_index, prm->cell_size); if (rc) return rc;
  std::vector<Pair> pairs(n_edges); std::vector<QueryTile> tiles; slot_edge_begin->assign(n_edges + 1, 0);
  long long slots = 0;
  for (int e = 0; e < n_edges; ++e) {
    pairs[e] = Pair{ref[e], nei[e], ref[e], nei[e]};
    (*slot_edge_begin)[e] = (int)slots;
    const int q0 = ctx->f_qry.off[nei[e]], qn = ctx->f_qry.off[nei[e] + 1] - q0;
    const bool has_target = ctx->f_tgt.off[ref[e] + 1] > ctx->f_tgt.off[ref[e]];
    if (has_target) for (int s = 0; s < qn; s += kTile) tiles.push_back(QueryTile{e, q0 + s, std::min(kTile, qn - s), (int)(slots + s)});
    slots += qn;
  }
  (*slot_edge_begin)[n_edges] = (int)slots;
  *total_slots = slots;
  CK(ctx->f_pairs.ensure(std::max<size_t>(16, pairs.size() * sizeof(Pair)))); CK(ctx->f_qtiles.ensure(std::max<size_t>(16, tiles.size() * sizeof(QueryTile))));
  CK(ctx->f_valid.ensure(std::max<size_t>(16, (size_t)slots))); CK(ctx->f_point.ensure(std::max<size_t>(16, (size_t)slots * 24))); CK(ctx->f_plane.ensure(std::max<size_t>(16, (size_t)slots * 32)));
  if (want_nn) { CK(ctx->f_nn_idx.ensure(std::max<size_t>(16, (size_t)slots * prm->k * 4))); CK(ctx->f_nn_d2.ensure(std::max<size_t>(16, (size_t)slots * prm->k * 4))); }
  if (!pairs.empty()) CK(cudaMemcpyAsync(ctx->f_pairs.p, pairs.data(), pairs.size() * sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  if (!tiles.empty()) CK(cudaMemcpyAsync(ctx->f_qtiles.p, tiles.data(), tiles.size() * sizeof(QueryTile), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->f_valid.p, 0, std::max<size_t>(16, (size_t)slots), ctx->stream));
  AssocArgs a{};
  a.q_local = ctx->f_qry.local.as<F4>(); a.q_orig = nullptr; a.tiles = ctx->f_qtiles.as<QueryTile>(); a.pairs = ctx->f_pairs.as<Pair>();
  a.grids = ctx->f_index.grids.as<GridDesc>(); a.cell_start = ctx->f_index.cell_start.as<uint32_t>(); a.sorted = ctx->f_index.sorted.as<F4>();
  a.wpose = ctx->d_wpose.as<WorldPose>(); a.prep = ctx->d_prep.as<PosePrep>();
  a.hint = nullptr;      // pose-graph pairs: every query set is searched once per outer iteration, no hints
  a.prm.sq_thr = prm->dist_threshold * prm->dist_threshold; a.prm.rmax = 1; a.prm.plane_tol = prm->plane_tolerance; a.prm.collinear_tol = 3.0;
  a.thr = (double)prm->dist_threshold;
  a.residual_type = PVB_P2PLANE_METER; a.normalize = 0; a.huber = 0; a.weight = 1;
  a.out_valid = ctx->f_valid.as<unsigned char>(); a.out_point = ctx->f_point.as<double>(); a.out_plane = ctx->f_plane.as<double>();
  a.out_nn_idx = want_nn ? ctx->f_nn_idx.as<int>() : nullptr; a.out_nn_d2 = want_nn ? ctx->f_nn_d2.as<float>() : nullptr;
  rc = launch_associate<false>(ctx, prm->k, (int)tiles.size(), a); if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));   // pairs/tiles vectors go out of scope
  return PVB_OK;
}

int pvb_frames_associate_point2plane(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei, const pvb_assoc_params* prm, long* n_assoc) {
  if (!ctx || !poses || n_edges < 0 || (n_edges > 0 && (!ref || !nei))) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  long long slots = 0; std::vector<int> sb;
  int rc = frames_run(ctx, poses, n_edges, ref, nei, prm, false, &slots, &sb); if (rc) return rc;
  CK(ctx->fh_valid.ensure(std::max<size_t>(16, (size_t)slots))); CK(ctx->fh_point.ensure(std::max<size_t>(16, (size_t)slots * 24))); CK(ctx->fh_plane.ensure(std::max<size_t>(16, (size_t)slots * 32)));
  if (slots > 0) {
    CK(cudaMemcpyAsync(ctx->fh_valid.p, ctx->f_valid.p, (size_t)slots, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->fh_point.p, ctx->f_point.p, (size_t)slots * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->fh_plane.p, ctx->f_plane.p, (size_t)slots * 32, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->a_edge.clear(); ctx->a_query.clear(); ctx->a_point.clear(); ctx->a_plane.clear();
  const unsigned char* v = ctx->fh_valid.as<unsigned char>(); const double* pt = ctx->fh_point.as<double>(); const double* pl = ctx->fh_plane.as<double>();
  for (int e = 0; e < n_edges; ++e)
    for (int s = sb[e]; s < sb[e + 1]; ++s)
      if (v[s]) {
        ctx->a_edge.push_back(e); ctx->a_query.push_back(s - sb[e]);
        ctx->a_point.insert(ctx->a_point.end(), pt + (size_t)s * 3, pt + (size_t)s * 3 + 3);
        ctx->a_plane.insert(ctx->a_plane.end(), pl + (size_t)s * 4, pl + (size_t)s * 4 + 4);
      }
  if (n_assoc) *n_assoc = (long)ctx->a_edge.size();
  return PVB_OK;
}

// AddLidarPointToPlaneResidual without the host round trip: associate, compact the accepted correspondences with a prefix sum and write the
// residual blocks (edge-major, query order inside an edge = the reference's push_back order) straight into the blocks-mode buffers; host-built
// extra blocks (line-to-line, camera-LiDAR, ...) are appended behind them as further edges.
int pvb_frames_point2plane_blocks(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei, const pvb_assoc_params* prm, int angle_residual,
                                  int normalize_distance, double weight, int block_offset, int n_pose_blocks, long n_extra, const int* x_type, const int* x_ref,
                                  const int* x_nei, const int* x_normalize, const double* x_huber, const double* x_consts, long* n_blocks) {
  if (!ctx || !poses || n_edges < 0 || (n_edges > 0 && (!ref || !nei)) || n_extra < 0 || (n_extra > 0 && (!x_type || !x_ref || !x_nei || !x_normalize || !x_huber || !x_consts)))
    return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_frames_point2plane_blocks: bad arguments") : PVB_ERR_ARG;
  if (block_offset < 0 || n_pose_blocks < block_offset + ctx->n_frames) return ctx->fail(PVB_ERR_ARG, "pvb_frames_point2plane_blocks: %d pose blocks do not hold %d frames at offset %d", n_pose_blocks, ctx->n_frames, block_offset);
  const int nb = n_pose_blocks;
  for (long i = 0; i < n_extra; ++i) {
    if (x_ref[i] < 0 || x_ref[i] >= nb || x_nei[i] < 0 || x_nei[i] >= nb) return ctx->fail(PVB_ERR_ARG, "extra block %ld: pose index out of range", i);
    if (x_type[i] < 0 || x_type[i] > PVB_LINE2LINE_ANGLE) return ctx->fail(PVB_ERR_ARG, "extra block %ld: unknown residual type %d", i, x_type[i]);
  }
  CK(cudaSetDevice(ctx->device));
  long long slots = 0; std::vector<int> sb;
  int rc = frames_run(ctx, poses, n_edges, ref, nei, prm, false, &slots, &sb); if (rc) return rc;
  // ---- compaction of the accepted slots: pos[s] = row of slot s, pos[slots] = their number
  CK(ctx->m_a.ensure((size_t)(slots + 1) * 4)); CK(ctx->m_b.ensure((size_t)(slots + 1) * 4)); CK(ctx->m_c.ensure((size_t)(n_edges + 1) * 4)); CK(ctx->m_d.ensure((size_t)(n_edges + 1) * 4));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  k_flags_to_u32<<<(unsigned)((slots + 256) / 256), 256, 0, ctx->stream>>>(ctx->f_valid.as<unsigned char>(), slots, ctx->m_a.as<uint32_t>());
  CKL();
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->m_a.as<uint32_t>(), ctx->m_b.as<uint32_t>(), (int)(slots + 1), ctx->stream);
  CK(ctx->m_e.ensure(tmp));
  size_t tb = ctx->m_e.cap;
  CK(cub::DeviceScan::ExclusiveSum(ctx->m_e.p, tb, ctx->m_a.as<uint32_t>(), ctx->m_b.as<uint32_t>(), (int)(slots + 1), ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_c.p, sb.data(), (size_t)(n_edges + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  k_gather_u32<<<(n_edges + 256) / 256, 256, 0, ctx->stream>>>(ctx->m_b.as<uint32_t>(), ctx->m_c.as<int>(), n_edges + 1, ctx->m_d.as<uint32_t>());
  CKL();
  std::vector<uint32_t> row_begin(n_edges + 1);                      // first row of every association edge
  CK(cudaMemcpyAsync(row_begin.data(), ctx->m_d.p, (size_t)(n_edges + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // line-to-line blocks left on the device by pvb_frames_line2line_blocks_device: placed after the host extras, one reduction edge per pose-graph edge that has any
  pvb_ctx::LinePending& LP = ctx->line_pending;
  const bool lx = LP.pending && ctx->line_layout.valid;
  const long long n_lx = lx ? LP.total : 0;
  std::vector<long long> lx_base;                                    // first line block of every line edge (relative to the first line block)
  if (lx) {
    lx_base.assign(LP.cnt.size(), 0);
    long long run = 0;
    for (size_t e = 0; e < LP.cnt.size(); ++e) {
      if (block_offset + LP.ref[e] >= nb || block_offset + LP.nei[e] >= nb) return ctx->fail(PVB_ERR_ARG, "pvb_frames_point2plane_blocks: line edge %d -> %d exceeds %d pose blocks", LP.ref[e], LP.nei[e], nb);
      lx_base[e] = run; run += LP.cnt[e];
    }
  }
  const long long n_dev = row_begin[n_edges], n = n_dev + n_extra + n_lx, lx_at = n_dev + n_extra;
  ctx->a_edge.clear(); ctx->a_query.clear(); ctx->a_point.clear(); ctx->a_plane.clear();     // the correspondences stay on the device in this path
  // ---- edges and tiles: the association edges in their order, then the extra blocks grouped by (ref, nei) like pvb_blocks_set
  ctx->bn = n; ctx->nb = nb; ctx->b_has_rows = ctx->b_has_sys = false;
  ctx->edge_ref.clear(); ctx->edge_nei.clear(); ctx->edge_tile_begin.clear();
  std::vector<BlockTile> tiles;
  std::vector<uint32_t> order(n_extra);
  for (long i = 0; i < n_extra; ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (long long)x_ref[a] * nb + x_nei[a] < (long long)x_ref[b] * nb + x_nei[b]; });
  if (ctx->g_edge_ref.empty()) {
    for (int e = 0; e < n_edges; ++e) {
      ctx->edge_ref.push_back(block_offset + ref[e]); ctx->edge_nei.push_back(block_offset + nei[e]); ctx->edge_tile_begin.push_back((int)tiles.size());
      for (long long s0 = row_begin[e]; s0 < row_begin[e + 1]; s0 += kTile) tiles.push_back(BlockTile{e, (int)s0, (int)std::min<long long>(kTile, row_begin[e + 1] - s0), 0});
    }
    for (long i = 0; i < n_extra;) {
      long j = i;
      const int er = x_ref[order[i]], en = x_nei[order[i]];
      while (j < n_extra && x_ref[order[j]] == er && x_nei[order[j]] == en) ++j;
      const int e = (int)ctx->edge_ref.size();
      ctx->edge_ref.push_back(er); ctx->edge_nei.push_back(en); ctx->edge_tile_begin.push_back((int)tiles.size());
      for (long s0 = i; s0 < j; s0 += kTile) tiles.push_back(BlockTile{e, (int)(n_dev + s0), (int)std::min<long>(kTile, j - s0), 0});
      i = j;
    }
    if (lx)
      for (size_t le = 0; le < LP.cnt.size(); ++le) {
        if (LP.cnt[le] == 0) continue;
        const int e = (int)ctx->edge_ref.size();
        ctx->edge_ref.push_back(block_offset + LP.ref[le]); ctx->edge_nei.push_back(block_offset + LP.nei[le]); ctx->edge_tile_begin.push_back((int)tiles.size());
        for (long long s0 = 0; s0 < LP.cnt[le]; s0 += kTile) tiles.push_back(BlockTile{e, (int)(lx_at + lx_base[le] + s0), (int)std::min<long long>(kTile, LP.cnt[le] - s0), 0});
      }
  } else {
    // sharded pose graph: the GLOBAL edge list is the reduction layout (every rank sums into the same n_edges x 92 buffer); this rank's association
    // edges and extra blocks are filed under their global edge, edges of other ranks keep an empty tile range
    const size_t ng = ctx->g_edge_ref.size();
    auto find_edge = [&](int er, int en) -> long {
      size_t lo = 0, hi = ng;                                     // the list is sorted by (ref, nei) (pvb_blocks_set_edge_list)
      while (lo < hi) { const size_t mid = (lo + hi) / 2; if (ctx->g_edge_ref[mid] < er || (ctx->g_edge_ref[mid] == er && ctx->g_edge_nei[mid] < en)) lo = mid + 1; else hi = mid; }
      return (lo < ng && ctx->g_edge_ref[lo] == er && ctx->g_edge_nei[lo] == en) ? (long)lo : -1;
    };
    std::vector<std::pair<long, BlockTile>> filed;
    for (int e = 0; e < n_edges; ++e) {
      const long ge = find_edge(block_offset + ref[e], block_offset + nei[e]);
      if (ge < 0) return ctx->fail(PVB_ERR_ARG, "association edge %d -> %d is not in the edge list set by pvb_blocks_set_edge_list", ref[e], nei[e]);
      for (long long s0 = row_begin[e]; s0 < row_begin[e + 1]; s0 += kTile) filed.push_back({ge, BlockTile{(int)ge, (int)s0, (int)std::min<long long>(kTile, row_begin[e + 1] - s0), 0}});
    }
    for (long i = 0; i < n_extra;) {
      long j = i;
      const int er = x_ref[order[i]], en = x_nei[order[i]];
      while (j < n_extra && x_ref[order[j]] == er && x_nei[order[j]] == en) ++j;
      const long ge = find_edge(er, en);
      if (ge < 0) return ctx->fail(PVB_ERR_ARG, "extra block edge %d -> %d is not in the edge list set by pvb_blocks_set_edge_list", er, en);
      for (long s0 = i; s0 < j; s0 += kTile) filed.push_back({ge, BlockTile{(int)ge, (int)(n_dev + s0), (int)std::min<long>(kTile, j - s0), 0}});
      i = j;
    }
    if (lx)
      for (size_t le = 0; le < LP.cnt.size(); ++le) {
        if (LP.cnt[le] == 0) continue;
        const long ge = find_edge(block_offset + LP.ref[le], block_offset + LP.nei[le]);
        if (ge < 0) return ctx->fail(PVB_ERR_ARG, "line edge %d -> %d is not in the edge list set by pvb_blocks_set_edge_list", LP.ref[le], LP.nei[le]);
        for (long long s0 = 0; s0 < LP.cnt[le]; s0 += kTile) filed.push_back({ge, BlockTile{(int)ge, (int)(lx_at + lx_base[le] + s0), (int)std::min<long long>(kTile, LP.cnt[le] - s0), 0}});
      }
    std::stable_sort(filed.begin(), filed.end(), [](const std::pair<long, BlockTile>& a, const std::pair<long, BlockTile>& b) { return a.first < b.first; });
    size_t f = 0;
    for (size_t ge = 0; ge < ng; ++ge) {
      ctx->edge_ref.push_back(ctx->g_edge_ref[ge]); ctx->edge_nei.push_back(ctx->g_edge_nei[ge]); ctx->edge_tile_begin.push_back((int)tiles.size());
      while (f < filed.size() && filed[f].first == (long)ge) tiles.push_back(filed[f++].second);
    }
  }
  ctx->edge_tile_begin.push_back((int)tiles.size());
  ctx->b_tiles = (int)tiles.size();
  ctx->b_orig.clear();
  const int ne = (int)ctx->edge_ref.size();
  CK(ctx->b_tile.ensure(std::max<size_t>(16, tiles.size() * sizeof(BlockTile))));
  CK(ctx->b_eref.ensure(std::max<size_t>(16, (size_t)ne * 4))); CK(ctx->b_enei.ensure(std::max<size_t>(16, (size_t)ne * 4)));
  CK(ctx->b_tbegin.ensure((size_t)(ne + 1) * 4));
  CK(ctx->b_type.ensure(std::max<size_t>(16, (size_t)n * 4))); CK(ctx->b_norm.ensure(std::max<size_t>(16, (size_t)n * 4)));
  CK(ctx->b_huber.ensure(std::max<size_t>(16, (size_t)n * 8))); CK(ctx->b_consts.ensure(std::max<size_t>(16, (size_t)n * 96)));
  CK(ctx->b_orig_d.ensure(std::max<size_t>(16, (size_t)n * 4)));
  CK(ctx->b_r.ensure(std::max<size_t>(16, (size_t)n * 8))); CK(ctx->b_J.ensure(std::max<size_t>(16, (size_t)n * 96)));
  CK(ctx->b_part.ensure(std::max<size_t>(16, tiles.size() * 92 * 8))); CK(ctx->b_esys.ensure(std::max<size_t>(16, (size_t)ne * 92 * 8)));
  CK(ctx->h_esys.ensure(std::max<size_t>(16, (size_t)ne * 92 * 8)));
  if (n_dev > 0) {
    k_blocks_from_point2plane<<<(unsigned)((slots + 255) / 256), 256, 0, ctx->stream>>>(ctx->f_valid.as<unsigned char>(), ctx->m_b.as<uint32_t>(), slots, ctx->f_point.as<double>(),
        ctx->f_plane.as<double>(), angle_residual ? PVB_P2PLANE_ANGLE : PVB_P2PLANE_METER, normalize_distance, angle_residual ? 2 * M_PI / 180.0 : 0.2, weight, n,
        ctx->b_type.as<int>(), ctx->b_norm.as<int>(), ctx->b_huber.as<double>(), ctx->b_consts.as<double>(), ctx->b_orig_d.as<uint32_t>());
    CKL();
  }
  std::vector<int> s_type(n_extra), s_norm(n_extra); std::vector<double> s_huber(n_extra), s_consts((size_t)n_extra * 12); std::vector<uint32_t> s_orig(n_extra);
  for (long r = 0; r < n_extra; ++r) {
    const uint32_t o = order[r];
    s_type[r] = x_type[o]; s_norm[r] = x_normalize[o]; s_huber[r] = x_huber[o]; s_orig[r] = (uint32_t)(n_dev + o);
    for (int k = 0; k < 12; ++k) s_consts[(size_t)k * n_extra + r] = x_consts[(size_t)o * 12 + k];
  }
  if (n_extra > 0) {
    CK(cudaMemcpyAsync(ctx->b_type.as<int>() + n_dev, s_type.data(), (size_t)n_extra * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_norm.as<int>() + n_dev, s_norm.data(), (size_t)n_extra * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_huber.as<double>() + n_dev, s_huber.data(), (size_t)n_extra * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_orig_d.as<uint32_t>() + n_dev, s_orig.data(), (size_t)n_extra * 4, cudaMemcpyHostToDevice, ctx->stream));
    for (int k = 0; k < 12; ++k)
      CK(cudaMemcpyAsync(ctx->b_consts.as<double>() + (size_t)k * n + n_dev, s_consts.data() + (size_t)k * n_extra, (size_t)n_extra * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (n_lx > 0) {
    const pvb_ctx::LineLayout& L = ctx->line_layout;
    CK(ctx->lt_base.ensure(lx_base.size() * 8));
    CK(cudaMemcpyAsync(ctx->lt_base.p, lx_base.data(), lx_base.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned char* d = ctx->v_misc.as<unsigned char>();
    const unsigned char* t = ctx->lt_misc.as<unsigned char>();
    k_line_blocks_emit<<<(unsigned)LP.slots, 128, 0, ctx->stream>>>(reinterpret_cast<const int*>(t + LP.o_slot), reinterpret_cast<const int*>(t + LP.o_hoff), ctx->lt_hold.as<int>(), ctx->lt_kbase.as<int>(),
        ctx->lt_base.as<long long>(), reinterpret_cast<const VotePair*>(d + L.o_vp), reinterpret_cast<const int*>(d + L.o_soff), reinterpret_cast<const double*>(t + LP.o_coeffs),
        reinterpret_cast<const int*>(d + L.o_coff), reinterpret_cast<const int*>(d + L.o_poff), reinterpret_cast<const int*>(d + L.o_pbase), reinterpret_cast<const int*>(d + L.o_pids),
        ctx->v_world.as<F4>(), reinterpret_cast<const WorldPose*>(d + L.o_wp), LP.type, LP.normalize, LP.huber, LP.weight, n, lx_at,
        ctx->b_type.as<int>(), ctx->b_norm.as<int>(), ctx->b_huber.as<double>(), ctx->b_consts.as<double>(), ctx->b_orig_d.as<uint32_t>());
    CKL();
  }
  LP.pending = false;                                                // consumed (or dropped: a pending list never outlives the next block build)
  if (!tiles.empty()) CK(cudaMemcpyAsync(ctx->b_tile.p, tiles.data(), tiles.size() * sizeof(BlockTile), cudaMemcpyHostToDevice, ctx->stream));
  if (ne > 0) {
    CK(cudaMemcpyAsync(ctx->b_eref.p, ctx->edge_ref.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->b_enei.p, ctx->edge_nei.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->b_tbegin.p, ctx->edge_tile_begin.data(), (size_t)(ne + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (n_blocks) *n_blocks = (long)n;
  return PVB_OK;
}

int pvb_frames_get_point2plane(const pvb_ctx* ctx, long cap, int* edge, int* query, double* point3, double* plane4) {
  if (!ctx) return PVB_ERR_ARG;
  const long n = std::min<long>(cap, (long)ctx->a_edge.size());
  for (long i = 0; i < n; ++i) { if (edge) edge[i] = ctx->a_edge[i]; if (query) query[i] = ctx->a_query[i]; }
  if (point3) memcpy(point3, ctx->a_point.data(), (size_t)n * 24);
  if (plane4) memcpy(plane4, ctx->a_plane.data(), (size_t)n * 32);
  return PVB_OK;
}

int pvb_frames_knn(pvb_ctx* ctx, const double* poses, int ref, int nei, const pvb_assoc_params* prm, int* idx, float* d2) {
  if (!ctx || !poses || !idx || !d2) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  long long slots = 0; std::vector<int> sb;
  int rc = frames_run(ctx, poses, 1, &ref, &nei, prm, true, &slots, &sb); if (rc) return rc;
  if (slots > 0) {
    CK(cudaMemcpyAsync(idx, ctx->f_nn_idx.p, (size_t)slots * prm->k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(d2, ctx->f_nn_d2.p, (size_t)slots * prm->k * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

// ---- point-to-line association on the corner clouds (A3) ------------------------------------------------------------
int pvb_frames_set_corners(pvb_ctx* ctx, int n_frames, const float* const* corner, const int* n_corner) {
  if (!ctx || n_frames <= 0 || !corner || !n_corner) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_frames_set_corners: bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  std::vector<const float*> p(corner, corner + n_frames); std::vector<int> c(n_corner, n_corner + n_frames), blocks(n_frames);
  for (int f = 0; f < n_frames; ++f) blocks[f] = f;
  int rc = set_cloudset(ctx, ctx->f_corner, p, c, blocks, 256); if (rc) return rc;
  ctx->n_corner_frames = n_frames; ctx->f_cindex.built = false;
  return PVB_OK;
}

int pvb_frames_associate_point2line(pvb_ctx* ctx, const double* poses, int n_edges, const int* ref, const int* nei, float dist_threshold, double cell_size, long* n_assoc) {
  if (!ctx || !poses || n_edges < 0 || (n_edges > 0 && (!ref || !nei))) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  if (ctx->n_corner_frames <= 0) return ctx->fail(PVB_ERR_STATE, "pvb_frames_set_corners has not been called");
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  const int nfr = ctx->n_corner_frames;
  for (int e = 0; e < n_edges; ++e) if (ref[e] < 0 || ref[e] >= nfr || nei[e] < 0 || nei[e] >= nfr) return ctx->fail(PVB_ERR_ARG, "edge %d out of range", e);
  int rc = upload_poses(ctx, poses, nfr, false); if (rc) return rc;
  CloudSet& cs = ctx->f_corner;
  rc = build_target_index(ctx, cs, ctx->f_cindex, cell_size); if (rc) return rc;
  std::vector<Pair> pairs(n_edges); std::vector<QueryTile> tiles; std::vector<int> sb(n_edges + 1, 0);
  long long slots = 0;
  for (int e = 0; e < n_edges; ++e) {
    pairs[e] = Pair{ref[e], nei[e], ref[e], nei[e]};
    sb[e] = (int)slots;
    const int q0 = cs.off[nei[e]], qn = cs.off[nei[e] + 1] - q0;
    if (cs.off[ref[e] + 1] > cs.off[ref[e]]) for (int s0 = 0; s0 < qn; s0 += kTile) tiles.push_back(QueryTile{e, q0 + s0, std::min(kTile, qn - s0), (int)(slots + s0)});
    slots += qn;
  }
  sb[n_edges] = (int)slots;
  CK(ctx->f_pairs.ensure(std::max<size_t>(16, pairs.size() * sizeof(Pair)))); CK(ctx->f_qtiles.ensure(std::max<size_t>(16, tiles.size() * sizeof(QueryTile))));
  const size_t S = std::max<size_t>(16, (size_t)slots);
  CK(ctx->f_valid.ensure(S)); CK(ctx->f_point.ensure(S * 24)); CK(ctx->f_la.ensure(S * 24)); CK(ctx->f_lb.ensure(S * 24));
  CK(ctx->fh_valid.ensure(S)); CK(ctx->fh_point.ensure(S * 24)); CK(ctx->fh_la.ensure(S * 24)); CK(ctx->fh_lb.ensure(S * 24));
  if (!pairs.empty()) CK(cudaMemcpyAsync(ctx->f_pairs.p, pairs.data(), pairs.size() * sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  if (!tiles.empty()) CK(cudaMemcpyAsync(ctx->f_qtiles.p, tiles.data(), tiles.size() * sizeof(QueryTile), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->f_valid.p, 0, S, ctx->stream));
  if (!tiles.empty()) {
    LineAssocArgs a{};
    a.q_local = cs.local.as<F4>(); a.tiles = ctx->f_qtiles.as<QueryTile>(); a.pairs = ctx->f_pairs.as<Pair>(); a.grids = ctx->f_cindex.grids.as<GridDesc>();
    a.cell_start = ctx->f_cindex.cell_start.as<uint32_t>(); a.sorted = ctx->f_cindex.sorted.as<F4>(); a.wpose = ctx->d_wpose.as<WorldPose>();
    a.sq_thr = dist_threshold * dist_threshold; a.thr = (double)dist_threshold;
    a.out_valid = ctx->f_valid.as<unsigned char>(); a.out_point = ctx->f_point.as<double>(); a.out_a = ctx->f_la.as<double>(); a.out_b = ctx->f_lb.as<double>();
    k_associate_line<5><<<(int)tiles.size(), kTile, 0, ctx->stream>>>(a);
    CKL();
  }
  if (slots > 0) {
    CK(cudaMemcpyAsync(ctx->fh_valid.p, ctx->f_valid.p, (size_t)slots, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->fh_point.p, ctx->f_point.p, (size_t)slots * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->fh_la.p, ctx->f_la.p, (size_t)slots * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->fh_lb.p, ctx->f_lb.p, (size_t)slots * 24, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->l_edge.clear(); ctx->l_query.clear(); ctx->l_point.clear(); ctx->l_a.clear(); ctx->l_b.clear();
  const unsigned char* v = ctx->fh_valid.as<unsigned char>();
  const double *pt = ctx->fh_point.as<double>(), *pa = ctx->fh_la.as<double>(), *pb = ctx->fh_lb.as<double>();
  for (int e = 0; e < n_edges; ++e)
    for (int sl = sb[e]; sl < sb[e + 1]; ++sl)
      if (v[sl]) {
        ctx->l_edge.push_back(e); ctx->l_query.push_back(sl - sb[e]);
        ctx->l_point.insert(ctx->l_point.end(), pt + (size_t)sl * 3, pt + (size_t)sl * 3 + 3);
        ctx->l_a.insert(ctx->l_a.end(), pa + (size_t)sl * 3, pa + (size_t)sl * 3 + 3);
        ctx->l_b.insert(ctx->l_b.end(), pb + (size_t)sl * 3, pb + (size_t)sl * 3 + 3);
      }
  if (n_assoc) *n_assoc = (long)ctx->l_edge.size();
  return PVB_OK;
}

int pvb_frames_get_point2line(const pvb_ctx* ctx, long cap, int* edge, int* query, double* point3, double* a3, double* b3) {
  if (!ctx) return PVB_ERR_ARG;
  const long n = std::min<long>(cap, (long)ctx->l_edge.size());
  for (long i = 0; i < n; ++i) { if (edge) edge[i] = ctx->l_edge[i]; if (query) query[i] = ctx->l_query[i]; }
  if (point3) memcpy(point3, ctx->l_point.data(), (size_t)n * 24);
  if (a3) memcpy(a3, ctx->l_a.data(), (size_t)n * 24);
  if (b3) memcpy(b3, ctx->l_b.data(), (size_t)n * 24);
  return PVB_OK;
}

// ================================================================ C. dense ICP sweep
int pvb_dense_set_target(pvb_ctx* ctx, const float* xyzc, long n, double cell_size) {
  if (!ctx || !xyzc || n <= 0) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_dense_set_target: bad arguments") : PVB_ERR_ARG;
  if (n >= (1l << 27)) return ctx->fail(PVB_ERR_ARG, "target cloud too large (%ld >= 2^27 points)", n);
  CK(cudaSetDevice(ctx->device));
  int rc = set_cloudset(ctx, ctx->d_tgt, {xyzc}, {(int)n}, {0}, 256); if (rc) return rc;
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  rc = upload_poses(ctx, zero6, 1, false); if (rc) return rc;      // target frame == world
  ctx->d_cell = cell_size;
  const bool sr = ctx->tune_dense_mode == 4 && !ctx->tune_stage;
  rc = build_target_index(ctx, ctx->d_tgt, ctx->d_index, cell_size, sr ? ctx->tune_dense_hscale : ctx->tune_hscale); if (rc) return rc;
  ctx->d_index.srow.release(); ctx->d_index.sstart.release(); ctx->d_index.srk.release(); ctx->d_index.srw.release(); ctx->d_index.has_superrows = false;
  ctx->d_order_valid = false; ctx->d_perm_valid = false; ctx->loc_known = false;      // the query order belongs to the old grid
  if (sr) { rc = build_superrows(ctx, ctx->d_index); if (rc) return rc; }
  CK(cudaStreamSynchronize(ctx->stream));
  // the unsorted world copy and sort scratch are not needed after the build
  if (ctx->d_hint.p && ctx->d_src.n_points > 0) CK(cudaMemsetAsync(ctx->d_hint.p, 0x7f, (size_t)ctx->d_src.n_points * sizeof(F4), ctx->stream));   // hints belong to the old target
  ctx->d_index.world.release(); ctx->d_index.keys.release(); ctx->d_index.keys_alt.release(); ctx->d_index.vals.release(); ctx->d_index.vals_alt.release(); ctx->d_index.hist.release();
  return PVB_OK;
}

// Source upload is pipelined against the next evaluate: the frames are split into chunks; every chunk is copied,
// Morton-keyed, sorted and gathered on a copy stream and signals an event; pvb_dense_evaluate* then launches the fused
// kernel chunk by chunk, each launch waiting only for its own chunk, so the H2D copy and the re-ordering of chunk c+1
// overlap the association of chunk c.  (With a pinned host buffer the copies are truly asynchronous.)
static int dense_prepare_layout(pvb_ctx* ctx, const int* offsets, int n_frames) {
  CloudSet& cs = ctx->d_src;
  cs.n_clouds = n_frames;
  cs.off.assign(offsets, offsets + n_frames + 1);
  if (cs.off[0] != 0) return ctx->fail(PVB_ERR_ARG, "offsets must start at 0");
  for (int f = 0; f < n_frames; ++f) if (cs.off[f + 1] < cs.off[f]) return ctx->fail(PVB_ERR_ARG, "offsets not monotone");
  cs.n_points = cs.off[n_frames];
  const long long n = cs.n_points;
  std::vector<CloudTile> ctiles; std::vector<int> blocks(n_frames);
  std::vector<Pair> pairs(n_frames); std::vector<QueryTile> tiles; std::vector<int> tbegin(n_frames + 1, 0);
  // chunks of the upload pipeline: at most tune_chunks, and not smaller than ~tune_chunk_min points (a chunk's ordering is ~8 launches: below that size their fixed cost outweighs the overlap)
  const int n_chunks = (int)std::max<long long>(1, std::min<long long>(std::min(n_frames, ctx->tune_chunks), (long long)cs.off[n_frames] / ctx->tune_chunk_min));
  ctx->d_chunk_frame.assign(n_chunks + 1, 0); ctx->d_chunk_ctile.assign(n_chunks + 1, 0); ctx->d_chunk_qtile.assign(n_chunks + 1, 0);
  for (int c = 0; c <= n_chunks; ++c) ctx->d_chunk_frame[c] = (int)((long long)n_frames * c / n_chunks);
  int chunk = 0;
  for (int f = 0; f < n_frames; ++f) {
    while (chunk < n_chunks && ctx->d_chunk_frame[chunk] == f) { ctx->d_chunk_ctile[chunk] = (int)ctiles.size(); ctx->d_chunk_qtile[chunk] = (int)tiles.size(); ++chunk; }
    const int cnt = cs.off[f + 1] - cs.off[f];
    blocks[f] = f + 1;
    pairs[f] = Pair{0, f, 0, f + 1};
    tbegin[f] = (int)tiles.size() * (kTile / 32);      // partials are per warp: kTile/32 entries per tile
    for (int s0 = 0; s0 < cnt; s0 += 256) ctiles.push_back(CloudTile{f, cs.off[f] + s0, std::min(256, cnt - s0), 0});
    for (int s0 = 0; s0 < cnt; s0 += kTile) tiles.push_back(QueryTile{f, cs.off[f] + s0, std::min(kTile, cnt - s0), 0});
  }
  ctx->d_chunk_ctile[n_chunks] = (int)ctiles.size(); ctx->d_chunk_qtile[n_chunks] = (int)tiles.size();
  tbegin[n_frames] = (int)tiles.size() * (kTile / 32);
  cs.n_tiles = (int)ctiles.size();
  ctx->d_ntiles = (int)tiles.size();
  ctx->d_frames = n_frames;
  ctx->d_perm_valid = false; ctx->d_order_valid = false; ctx->loc_known = false; ctx->loc_inflight = false;      // a new layout: the old permutation means nothing
  CK(cs.local.ensure(std::max<size_t>(16, (size_t)n * sizeof(F4))));
  CK(cs.tiles.ensure(std::max<size_t>(16, ctiles.size() * sizeof(CloudTile))));
  CK(ctx->m_a.ensure(std::max<size_t>(16, (size_t)n * 8))); CK(ctx->m_b.ensure(std::max<size_t>(16, (size_t)n * 8)));
  CK(ctx->m_c.ensure(std::max<size_t>(16, (size_t)n * 4))); CK(ctx->m_d.ensure(std::max<size_t>(16, (size_t)n * 4)));
  CK(ctx->d_q_sorted.ensure(std::max<size_t>(16, (size_t)n * sizeof(F4)))); CK(ctx->d_q_orig.ensure(std::max<size_t>(16, (size_t)n * 4)));
  CK(ctx->d_hint.ensure(std::max<size_t>(16, (size_t)n * sizeof(F4))));
  CK(cudaMemsetAsync(ctx->d_hint.p, 0x7f, (size_t)n * sizeof(F4), ctx->stream));      // 0x7f7f7f7f = 3.39e38: "no hint"
  CK(ctx->d_pairs.ensure(pairs.size() * sizeof(Pair))); CK(ctx->d_qtiles.ensure(std::max<size_t>(16, tiles.size() * sizeof(QueryTile)))); CK(ctx->d_tbegin.ensure(tbegin.size() * 4));
  CK(ctx->d_part.ensure(std::max<size_t>(16, tiles.size() * (kTile / 32) * 29 * 8))); CK(ctx->d_sys.ensure((size_t)n_frames * 29 * 8)); CK(ctx->dh_sys.ensure((size_t)n_frames * 29 * 8));
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, ctx->m_a.as<unsigned long long>(), ctx->m_b.as<unsigned long long>(), ctx->m_c.as<uint32_t>(), ctx->m_d.as<uint32_t>(), (int)std::max<long long>(1, n), 0, 64, ctx->stream);
  CK(ctx->m_e.ensure(tb));
  if (!ctiles.empty()) CK(cudaMemcpyAsync(cs.tiles.p, ctiles.data(), ctiles.size() * sizeof(CloudTile), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_pairs.p, pairs.data(), pairs.size() * sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  if (!tiles.empty()) CK(cudaMemcpyAsync(ctx->d_qtiles.p, tiles.data(), tiles.size() * sizeof(QueryTile), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_tbegin.p, tbegin.data(), tbegin.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  while ((int)ctx->chunk_ev.size() < n_chunks) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
  return PVB_OK;
}

int pvb_dense_set_sources(pvb_ctx* ctx, const float* xyzc, const int* offsets, int n_frames) {
  if (!ctx || !xyzc || !offsets || n_frames <= 0) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_dense_set_sources: bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CloudSet& cs = ctx->d_src;
  const bool same = ctx->d_frames == n_frames && (int)cs.off.size() == n_frames + 1 && std::equal(offsets, offsets + n_frames + 1, cs.off.begin());
  if (!same) { int rc = dense_prepare_layout(ctx, offsets, n_frames); if (rc) return rc; }
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&ctx->sort_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->eval_done, cudaEventDisableTiming));
  }
  // two side streams: the copy stream only moves bytes (the chunks' H2D copies run back to back at PCIe rate), the sort stream re-orders chunk c
  // (Morton keys, radix sort, gather) as soon as its copy has landed, while chunk c + 1 is still on the wire; the evaluate waits per chunk
  cudaStream_t cp = ctx->copy_stream, st = ctx->sort_stream;
  // the previous evaluate may still be reading the query buffers
  CK(cudaEventRecord(ctx->eval_done, ctx->stream));
  CK(cudaStreamWaitEvent(cp, ctx->eval_done, 0));
  CK(cudaStreamWaitEvent(st, ctx->eval_done, 0));
  const int n_chunks = (int)ctx->d_chunk_frame.size() - 1;
  while ((int)ctx->h2d_ev.size() < n_chunks) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->h2d_ev.push_back(e); }
  // MODE 4 orders the queries by TARGET cell, which needs the poses: only the copies are queued here, the ordering runs at the next evaluate (dense_order_queries)
  ctx->d_cell_order = ctx->tune_dense_mode == 4 && !ctx->tune_stage && ctx->d_index.has_superrows;
  for (int c = 0; c < n_chunks; ++c) {
    const int f0 = ctx->d_chunk_frame[c], f1 = ctx->d_chunk_frame[c + 1];
    const long long p0 = cs.off[f0], cnt = (long long)cs.off[f1] - p0;
    if (cnt > 0 && ctx->d_cell_order) {
      CK(cudaMemcpyAsync(cs.local.as<F4>() + p0, xyzc + (size_t)p0 * 4, (size_t)cnt * sizeof(F4), cudaMemcpyHostToDevice, cp));
      CK(cudaEventRecord(ctx->h2d_ev[c], cp));
      continue;
    }
    if (cnt > 0) {
      CK(cudaMemcpyAsync(cs.local.as<F4>() + p0, xyzc + (size_t)p0 * 4, (size_t)cnt * sizeof(F4), cudaMemcpyHostToDevice, cp));
      CK(cudaEventRecord(ctx->h2d_ev[c], cp));
      CK(cudaStreamWaitEvent(st, ctx->h2d_ev[c], 0));
      const int t0 = ctx->d_chunk_ctile[c], t1 = ctx->d_chunk_ctile[c + 1];
      // Morton re-ordering per frame (keys = (frame - first frame of the chunk) << 3 bits | morton); vals are global point indices
      int frame_bits = 1; while ((1 << frame_bits) < f1 - f0) ++frame_bits;
      k_morton_keys<<<t1 - t0, 256, 0, st>>>(cs.local.as<F4>(), cs.tiles.as<CloudTile>() + t0, f0, (float)(1 << (ctx->tune_morton_bits - 10)), ctx->tune_morton_bits,
                                             ctx->m_a.as<unsigned long long>(), ctx->m_c.as<uint32_t>());
      CKL();
      size_t tb = ctx->m_e.cap;
      CK(cub::DeviceRadixSort::SortPairs(ctx->m_e.p, tb, ctx->m_a.as<unsigned long long>() + p0, ctx->m_b.as<unsigned long long>() + p0, ctx->m_c.as<uint32_t>() + p0,
                                         ctx->m_d.as<uint32_t>() + p0, (int)cnt, 0, 3 * ctx->tune_morton_bits + frame_bits, st));
      k_gather_f4<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(cs.local.as<F4>(), ctx->m_d.as<uint32_t>() + p0, cnt, ctx->d_q_sorted.as<F4>() + p0);
      CKL();
      CK(cudaMemcpyAsync(ctx->d_q_orig.as<uint32_t>() + p0, ctx->m_d.as<uint32_t>() + p0, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaEventRecord(ctx->chunk_ev[c], st));
  }
  ctx->d_chunks_pending = true;
  ctx->d_order_pending = ctx->d_cell_order;
  return PVB_OK;
}

// MODE 4: (re-)order the queries of every frame by target cell under the poses just uploaded (ctx->h_pose / d_wpose).  Runs chunk by chunk on the sort
// stream (after the chunk's H2D copy when the upload is fresh) and signals chunk_ev[c]; the evaluate then launches chunk by chunk as for a fresh upload.
// Returns *did = false when the present order is still good: no upload since, and no frame's pose moved a point by more than half a cell.
static int dense_order_queries(pvb_ctx* ctx, bool* did) {
  *did = false;
  if (!ctx->d_cell_order) return PVB_OK;
  CloudSet& cs = ctx->d_src;
  const int nf = ctx->d_frames;
  const GridDesc g = ctx->d_index.h_grids[0];
  const WorldPose* hw = reinterpret_cast<const WorldPose*>(ctx->h_pose.as<PosePrep>() + (nf + 1)) + 1;      // staging of upload_poses: block 0 = identity
  // have the poses moved a point of some frame by more than tune_reorder cells since the last SORT?  (bound from the pose change and the frame's largest sensor distance)
  auto poses_moved = [&]() -> bool {
    if (!ctx->d_order_valid || (int)ctx->d_order_wpose.size() != nf) return true;
    if (ctx->rmax_inflight) { if (cudaEventQuery(ctx->rmax_ev) == cudaSuccess) ctx->rmax_inflight = false; else return true; }
    const float* r2 = ctx->dh_rmax2.as<float>();
    for (int f = 0; f < nf; ++f) {
      double fr = 0, dt = 0;
      for (int k = 0; k < 9; ++k) { const double d = hw[f].R[k] - ctx->d_order_wpose[f].R[k]; fr += d * d; }
      for (int k = 0; k < 3; ++k) { const double d = hw[f].t[k] - ctx->d_order_wpose[f].t[k]; dt += d * d; }
      if (std::sqrt(fr) * std::sqrt((double)r2[f]) + std::sqrt(dt) > ctx->tune_reorder * g.h) {
        if (getenv("PVB_DEBUG_ORDER")) fprintf(stderr, "[pvb] re-order: frame %d dR %.3g rmax %.3g dt %.3g h %.3g\n", f, std::sqrt(fr), std::sqrt((double)r2[f]), std::sqrt(dt), g.h);
        return true;
      }
    }
    return false;
  };
  const bool moved = poses_moved();
  if (!ctx->d_order_pending && !moved) return PVB_OK;
  // locality of the order used by the previous fresh upload (measured on the device, read back here)
  if (ctx->loc_inflight && cudaEventQuery(ctx->loc_ev) == cudaSuccess) {
    ctx->loc_inflight = false; ctx->loc_known = true;
    const double frac = (double)*ctx->dh_loc.as<unsigned int>() / (double)std::max<long long>(1, cs.n_points);
    ctx->loc_last_frac = frac;
    if (ctx->loc_was_sort) { ctx->loc_sorted_frac = frac; }
    else if (frac > ctx->loc_sorted_frac + 0.05) { ctx->reuse_holdoff = ctx->reuse_backoff; ctx->reuse_backoff = std::min(64, 2 * ctx->reuse_backoff); }      // the data changed: sort, and wait longer before trusting an old order again
    else ctx->reuse_backoff = 1;
  }
  // a fresh upload may keep the previous permutation (gather only): same layout, poses near those of the last sort, and the order was still local when last measured
  bool reuse = ctx->d_order_pending && ctx->tune_reuse && ctx->d_perm_valid && !moved && ctx->loc_known && !ctx->loc_inflight &&
               ctx->loc_last_frac <= ctx->loc_sorted_frac + 0.05 && ctx->reuse_holdoff == 0;
  if (ctx->d_order_pending && !reuse && ctx->reuse_holdoff > 0) --ctx->reuse_holdoff;
  if (getenv("PVB_DEBUG_ORDER")) fprintf(stderr, "[pvb] re-order #%ld: pending %d valid %d moved %d reuse %d locality %.3f (sorted %.3f)\n", ctx->d_reorders, (int)ctx->d_order_pending, (int)ctx->d_order_valid, (int)moved, (int)reuse, ctx->loc_last_frac, ctx->loc_sorted_frac);
  if (!ctx->loc_ev) { CK(cudaEventCreateWithFlags(&ctx->loc_ev, cudaEventDisableTiming)); CK(ctx->d_loc.ensure(16)); CK(ctx->dh_loc.ensure(16)); }
  if (!ctx->rmax_ev) { CK(cudaEventCreateWithFlags(&ctx->rmax_ev, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->pose_ev, cudaEventDisableTiming)); }
  const bool fresh = ctx->d_order_pending;
  // fresh upload: chunk by chunk on the sort stream behind the chunks' H2D copies (the evaluate launches per chunk); pose update: ONE sort of all frames on the
  // main stream, one launch afterwards
  cudaStream_t st = fresh ? ctx->sort_stream : ctx->stream;
  CK(ctx->d_rmax2.ensure((size_t)nf * 4)); CK(ctx->dh_rmax2.ensure((size_t)nf * 4));
  if (fresh) {
    CK(cudaEventRecord(ctx->pose_ev, ctx->stream));                // the poses of this evaluate
    CK(cudaStreamWaitEvent(st, ctx->pose_ev, 0));
  }
  if (fresh) { CK(cudaMemsetAsync(ctx->d_rmax2.p, 0, (size_t)nf * 4, st)); CK(cudaMemsetAsync(ctx->d_loc.p, 0, 4, st)); }
  uint32_t* const rmax_dst = fresh ? ctx->d_rmax2.as<uint32_t>() : nullptr;
  long long ncells = (long long)g.dims[0] * g.dims[1] * g.dims[2];
  int cellbits = 1; while ((1ll << cellbits) < ncells) ++cellbits;
  const int n_chunks = fresh ? (int)ctx->d_chunk_frame.size() - 1 : 1;
  for (int c = 0; c < n_chunks; ++c) {
    const int f0 = fresh ? ctx->d_chunk_frame[c] : 0, f1 = fresh ? ctx->d_chunk_frame[c + 1] : nf;
    const long long p0 = cs.off[f0], cnt = (long long)cs.off[f1] - p0;
    if (cnt > 0) {
      if (fresh) CK(cudaStreamWaitEvent(st, ctx->h2d_ev[c], 0));
      const int t0 = fresh ? ctx->d_chunk_ctile[c] : 0, t1 = fresh ? ctx->d_chunk_ctile[c + 1] : cs.n_tiles;
      int frame_bits = 1; while ((1 << frame_bits) < f1 - f0) ++frame_bits;
      size_t tb = ctx->m_e.cap;
      const unsigned lb = (unsigned)((cnt + 255) / 256);
      if (cellbits + frame_bits <= 32 && !ctx->tune_key64) {
        k_target_cell_keys<uint32_t><<<t1 - t0, 256, 0, st>>>(cs.local.as<F4>(), cs.tiles.as<CloudTile>() + t0, f0, ctx->d_wpose.as<WorldPose>(), g, cellbits,
                                                              ctx->m_a.as<uint32_t>(), ctx->m_c.as<uint32_t>(), rmax_dst);
        CKL();
        if (!reuse) {
          CK(cub::DeviceRadixSort::SortPairs(ctx->m_e.p, tb, ctx->m_a.as<uint32_t>() + p0, ctx->m_b.as<uint32_t>() + p0, ctx->m_c.as<uint32_t>() + p0,
                                             ctx->m_d.as<uint32_t>() + p0, (int)cnt, 0, cellbits + frame_bits, st));
          CK(cudaMemcpyAsync(ctx->d_q_orig.as<uint32_t>() + p0, ctx->m_d.as<uint32_t>() + p0, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, st));
        }
        if (fresh) { k_order_locality<uint32_t><<<lb, 256, 0, st>>>(ctx->m_a.as<uint32_t>(), ctx->d_q_orig.as<uint32_t>() + p0, cnt, g, cellbits, ctx->d_loc.as<unsigned int>()); CKL(); }
      } else {
        k_target_cell_keys<unsigned long long><<<t1 - t0, 256, 0, st>>>(cs.local.as<F4>(), cs.tiles.as<CloudTile>() + t0, f0, ctx->d_wpose.as<WorldPose>(), g, cellbits,
                                                                        ctx->m_a.as<unsigned long long>(), ctx->m_c.as<uint32_t>(), rmax_dst);
        CKL();
        if (!reuse) {
          CK(cub::DeviceRadixSort::SortPairs(ctx->m_e.p, tb, ctx->m_a.as<unsigned long long>() + p0, ctx->m_b.as<unsigned long long>() + p0, ctx->m_c.as<uint32_t>() + p0,
                                             ctx->m_d.as<uint32_t>() + p0, (int)cnt, 0, cellbits + frame_bits, st));
          CK(cudaMemcpyAsync(ctx->d_q_orig.as<uint32_t>() + p0, ctx->m_d.as<uint32_t>() + p0, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, st));
        }
        if (fresh) { k_order_locality<unsigned long long><<<lb, 256, 0, st>>>(ctx->m_a.as<unsigned long long>(), ctx->d_q_orig.as<uint32_t>() + p0, cnt, g, cellbits, ctx->d_loc.as<unsigned int>()); CKL(); }
      }
      k_gather_f4<<<lb, 256, 0, st>>>(cs.local.as<F4>(), ctx->d_q_orig.as<uint32_t>() + p0, cnt, ctx->d_q_sorted.as<F4>() + p0);
      CKL();
    }
    if (fresh) CK(cudaEventRecord(ctx->chunk_ev[c], st));
  }
  if (fresh) {
    CK(cudaMemcpyAsync(ctx->dh_rmax2.p, ctx->d_rmax2.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->rmax_ev, st));
    ctx->rmax_inflight = true;
  }
  if (fresh) {
    CK(cudaMemcpyAsync(ctx->dh_loc.p, ctx->d_loc.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->loc_ev, st));
    ctx->loc_inflight = true; ctx->loc_was_sort = !reuse;
  }
  if (!reuse) {
    ctx->d_order_wpose.resize(nf);
    for (int f = 0; f < nf; ++f) { memcpy(ctx->d_order_wpose[f].R, hw[f].R, 72); memcpy(ctx->d_order_wpose[f].t, hw[f].t, 24); }
    ctx->d_reorders++;
  } else ctx->d_order_reuses++;
  ctx->d_order_pending = false; ctx->d_order_valid = true; ctx->d_perm_valid = true; ctx->d_chunks_pending = fresh;
  *did = true;
  return PVB_OK;
}

static int dense_args(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, AssocArgs& a) {
  if (!ctx->d_index.built) return ctx->fail(PVB_ERR_STATE, "pvb_dense_set_target has not been called");
  if (ctx->d_frames <= 0) return ctx->fail(PVB_ERR_STATE, "pvb_dense_set_sources has not been called");
  if (!prm || (prm->k != 5 && prm->k != 10)) return ctx->fail(PVB_ERR_ARG, "k must be 5 or 10");
  if (prm->residual_type != PVB_P2PLANE_METER && prm->residual_type != PVB_P2PLANE_ANGLE) return ctx->fail(PVB_ERR_ARG, "residual_type must be a point-to-plane type");
  int rc = upload_poses(ctx, poses_lw, ctx->d_frames, true); if (rc) return rc;
  a = AssocArgs{};
  a.q_local = ctx->d_q_sorted.as<F4>(); a.q_orig = ctx->d_q_orig.as<uint32_t>(); a.tiles = ctx->d_qtiles.as<QueryTile>(); a.pairs = ctx->d_pairs.as<Pair>();
  a.grids = ctx->d_index.grids.as<GridDesc>(); a.cell_start = ctx->d_index.cell_start.as<uint32_t>(); a.sorted = ctx->d_index.sorted.as<F4>();
  a.wpose = ctx->d_wpose.as<WorldPose>(); a.prep = ctx->d_prep.as<PosePrep>();
  a.hint = ctx->d_hint.as<F4>();
  if (ctx->d_index.has_superrows && ctx->tune_dense_mode == 4 && !ctx->tune_stage) { a.srow = ctx->d_index.srow.as<float4>(); a.sstart = ctx->d_index.sstart.as<uint32_t>(); a.srw = ctx->d_index.srw.as<uint32_t>(); a.srk = ctx->d_index.srk.as<float>(); }
  a.one = 1.0f;
  a.prm.sq_thr = prm->dist_threshold * prm->dist_threshold; a.prm.rmax = 1; a.prm.plane_tol = prm->plane_tolerance; a.prm.collinear_tol = 3.0;
  a.thr = (double)prm->dist_threshold;
  a.residual_type = prm->residual_type; a.normalize = prm->normalize; a.huber = prm->huber; a.weight = prm->weight;
  return PVB_OK;
}

int pvb_dense_evaluate_device(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, double** dev_sys) {
  if (!ctx || !poses_lw) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  AssocArgs a; int rc = dense_args(ctx, poses_lw, prm, a); if (rc) return rc;
  a.partials = ctx->d_part.as<double>();
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  { bool did = false; rc = dense_order_queries(ctx, &did); if (rc) return rc; }
  if (ctx->d_chunks_pending) {            // fresh upload: consume it chunk by chunk as the copy stream delivers
    const int n_chunks = (int)ctx->d_chunk_frame.size() - 1;
    for (int c = 0; c < n_chunks; ++c) {
      CK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
      const int t0 = ctx->d_chunk_qtile[c], t1 = ctx->d_chunk_qtile[c + 1];
      AssocArgs ac = a;
      ac.tiles = a.tiles + t0;
      ac.partials = a.partials + (size_t)t0 * (kTile / 32) * 29;
      rc = launch_associate<true>(ctx, prm->k, t1 - t0, ac, true); if (rc) return rc;
    }
    ctx->d_chunks_pending = false;
  } else {
    rc = launch_associate<true>(ctx, prm->k, ctx->d_ntiles, a, true); if (rc) return rc;   // the target frame is the world: identity reference pose
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->ev_valid = true;
  double* dst = (dev_sys && *dev_sys) ? *dev_sys : ctx->d_sys.as<double>();   // caller-provided device buffer (e.g. a slice of an allreduce buffer)
  CK(ctx->d_chunk.ensure((size_t)ctx->d_frames * kSumChunks * 29 * 8));
  k_sum_partials<29><<<dim3(ctx->d_frames, kSumChunks), 256, 0, ctx->stream>>>(ctx->d_part.as<double>(), ctx->d_tbegin.as<int>(), ctx->d_chunk.as<double>());
  CKL();
  k_sum_chunks<29><<<ctx->d_frames, 32, 0, ctx->stream>>>(ctx->d_chunk.as<double>(), dst);
  CKL();
  if (dev_sys) *dev_sys = dst;
  return PVB_OK;
}

int pvb_dense_set_hints(pvb_ctx* ctx, int enable) { if (!ctx) return PVB_ERR_ARG; ctx->tune_hints = enable != 0; return PVB_OK; }
int pvb_dense_reset_hints(pvb_ctx* ctx) {
  if (!ctx) return PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_hint.p && ctx->d_src.n_points > 0) CK(cudaMemsetAsync(ctx->d_hint.p, 0x7f, (size_t)ctx->d_src.n_points * sizeof(F4), ctx->stream));
  return PVB_OK;
}

int pvb_dense_order_stats(const pvb_ctx* ctx, long* sorts, long* reuses) {
  if (!ctx) return PVB_ERR_ARG;
  if (sorts) *sorts = ctx->d_reorders;
  if (reuses) *reuses = ctx->d_order_reuses;
  return PVB_OK;
}

int pvb_debug_counters(pvb_ctx* ctx, unsigned long long* out2) {
  if (!ctx || !out2) return PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out2, ctx->d_stats.p, 16, cudaMemcpyDeviceToHost));
  return PVB_OK;
}

int pvb_dense_kernel_time_ms(pvb_ctx* ctx, float* ms) {
  if (!ctx || !ms) return PVB_ERR_ARG;
  if (!ctx->ev_valid) return ctx->fail(PVB_ERR_STATE, "no dense evaluate has run");
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return PVB_OK;
}

int pvb_dense_evaluate(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, double* out_sys) {
  if (!ctx || !out_sys) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  double* none = nullptr;
  int rc = pvb_dense_evaluate_device(ctx, poses_lw, prm, &none); if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->dh_sys.p, ctx->d_sys.p, (size_t)ctx->d_frames * 29 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out_sys, ctx->dh_sys.p, (size_t)ctx->d_frames * 29 * 8);
  return PVB_OK;
}

int pvb_dense_get_rows(pvb_ctx* ctx, const double* poses_lw, const pvb_dense_params* prm, unsigned char* valid, double* point3, double* plane4, double* residual, double* jac6) {
  if (!ctx || !poses_lw) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  AssocArgs a; int rc = dense_args(ctx, poses_lw, prm, a); if (rc) return rc;
  const size_t n = (size_t)ctx->d_src.n_points;
  CK(ctx->d_valid.ensure(n)); CK(ctx->d_point.ensure(n * 24)); CK(ctx->d_plane.ensure(n * 32)); CK(ctx->d_res.ensure(n * 8)); CK(ctx->d_jac.ensure(n * 48));
  CK(cudaMemsetAsync(ctx->d_valid.p, 0, n, ctx->stream)); CK(cudaMemsetAsync(ctx->d_point.p, 0, n * 24, ctx->stream)); CK(cudaMemsetAsync(ctx->d_plane.p, 0, n * 32, ctx->stream));
  a.out_valid = ctx->d_valid.as<unsigned char>(); a.out_point = ctx->d_point.as<double>(); a.out_plane = ctx->d_plane.as<double>();
  a.out_res = ctx->d_res.as<double>(); a.out_jac6 = ctx->d_jac.as<double>();
  { bool did = false; rc = dense_order_queries(ctx, &did); if (rc) return rc; }
  if (ctx->d_chunks_pending) { CK(cudaStreamSynchronize(ctx->copy_stream)); CK(cudaStreamSynchronize(ctx->sort_stream)); ctx->d_chunks_pending = false; }
  rc = launch_associate<false>(ctx, prm->k, ctx->d_ntiles, a, true); if (rc) return rc;
  if (valid) CK(cudaMemcpyAsync(valid, ctx->d_valid.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (point3) CK(cudaMemcpyAsync(point3, ctx->d_point.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (plane4) CK(cudaMemcpyAsync(plane4, ctx->d_plane.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (residual) CK(cudaMemcpyAsync(residual, ctx->d_res.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (jac6) CK(cudaMemcpyAsync(jac6, ctx->d_jac.p, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

int pvb_dense_gauss_newton_step(const double* sys29, int n_frames, double lambda, double* poses_lw) {
  if (!sys29 || !poses_lw || n_frames <= 0) return PVB_ERR_ARG;
  for (int f = 0; f < n_frames; ++f) {
    const double* S = sys29 + (size_t)f * 29;
    if (S[28] < 6) continue;
    std::vector<double> A(36), b(6);
    int q = 0;
    for (int a = 0; a < 6; ++a) for (int c = a; c < 6; ++c, ++q) { A[a * 6 + c] = S[q]; A[c * 6 + a] = S[q]; }
    for (int a = 0; a < 6; ++a) { A[a * 6 + a] += lambda * std::max(A[a * 6 + a], 1e-6); b[a] = -S[21 + a]; }
    if (!cholesky_solve(A, 6, b)) continue;
    for (int a = 0; a < 6; ++a) poses_lw[6 * f + a] += b[a];
  }
  return PVB_OK;
}

// ================================================================ D. projection
static void T16_to_pose(const double* T, WorldPose& w) { for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) w.R[r * 3 + c] = T[r * 4 + c]; w.t[r] = T[r * 4 + 3]; } }

int pvb_project_equirect(pvb_ctx* ctx, const float* xyzi, long n, const double* T, int rows, int cols, float* uvd) {
  if (!ctx || !xyzi || !T || !uvd || n < 0) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  if (n == 0) return PVB_OK;
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  CK(ctx->m_a.ensure((size_t)n * 16)); CK(ctx->m_b.ensure((size_t)n * 12));
  CK(cudaMemcpyAsync(ctx->m_a.p, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  WorldPose w; T16_to_pose(T, w);
  k_project<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->m_a.as<F4>(), n, w, rows, cols, ctx->m_b.as<float>(), nullptr, 0);
  CKL();
  CK(cudaMemcpyAsync(uvd, ctx->m_b.p, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

int pvb_project_depth_image(pvb_ctx* ctx, const float* xyzi, long n, const double* T, int rows, int cols, int size, uint16_t* image) {
  if (!ctx || !xyzi || !T || !image || n < 0 || rows <= 0 || cols <= 0) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  const size_t px = (size_t)rows * cols;
  CK(ctx->m_a.ensure(std::max<size_t>(16, (size_t)n * 16))); CK(ctx->m_c.ensure(px * 8)); CK(ctx->m_d.ensure(px * 2));
  CK(cudaMemsetAsync(ctx->m_c.p, 0, px * 8, ctx->stream));
  if (n > 0) {
    CK(cudaMemcpyAsync(ctx->m_a.p, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    WorldPose w; T16_to_pose(T, w);
    k_project<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->m_a.as<F4>(), n, w, rows, cols, nullptr, ctx->m_c.as<unsigned long long>(), size);
    CKL();
  }
  k_splat_finalize<<<(unsigned)((px + 255) / 256), 256, 0, ctx->stream>>>(ctx->m_c.as<unsigned long long>(), (long long)px, ctx->m_d.as<uint16_t>());
  CKL();
  CK(cudaMemcpyAsync(image, ctx->m_d.p, px * 2, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

int pvb_transform_cloud(pvb_ctx* ctx, const float* xyzi, long n, const double* R, const double* t, float* out) {
  if (!ctx || n < 0 || (n > 0 && (!xyzi || !out)) || !R || !t) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  if (n == 0) return PVB_OK;
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  CK(ctx->m_a.ensure((size_t)n * 16)); CK(ctx->m_b.ensure((size_t)n * 16));
  CK(cudaMemcpyAsync(ctx->m_a.p, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  WorldPose w; for (int k = 0; k < 9; ++k) w.R[k] = R[k]; for (int k = 0; k < 3; ++k) w.t[k] = t[k];
  k_transform_simple<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->m_a.as<F4>(), n, w, ctx->m_b.as<F4>());
  CKL();
  CK(cudaMemcpyAsync(out, ctx->m_b.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

// ================================================================ sweep undistortion
int pvb_undistort_clouds(pvb_ctx* ctx, const float* xyzi, const int* offsets, int n_frames, const double* T_wl16, const double* T_we16, const unsigned char* has_end,
                         float* out) {
  if (!ctx || n_frames < 0 || (n_frames > 0 && (!offsets || !T_wl16 || !T_we16))) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_undistort_clouds: bad arguments") : PVB_ERR_ARG;
  if (n_frames == 0) return PVB_OK;
  const long long n = offsets[n_frames];
  for (int f = 0; f < n_frames; ++f) if (offsets[f + 1] < offsets[f]) return ctx->fail(PVB_ERR_ARG, "pvb_undistort_clouds: offsets must ascend");
  if (n == 0) return PVB_OK;
  if (!xyzi || !out) return ctx->fail(PVB_ERR_ARG, "pvb_undistort_clouds: null cloud");
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  std::vector<UndistortPrep> prep(n_frames);
  std::vector<CloudTile> tiles;
  for (int f = 0; f < n_frames; ++f) {
    UndistortPrep& u = prep[f];
    memset(&u, 0, sizeof(u));
    if (!has_end || has_end[f]) {
      const double* A = T_wl16 + (size_t)f * 16; const double* E = T_we16 + (size_t)f * 16;
      double R_se[9], t_se[3], q[4];
      for (int r = 0; r < 3; ++r) {                              // R_wl^T R_we, R_wl^T (t_we - t_wl)  (Velodyne.cpp:1647-1648)
        for (int c = 0; c < 3; ++c) { double acc = 0.0; for (int k = 0; k < 3; ++k) acc += A[k * 4 + r] * E[k * 4 + c]; R_se[r * 3 + c] = acc; }
        double acc = 0.0; for (int k = 0; k < 3; ++k) acc += A[k * 4 + r] * (E[k * 4 + 3] - A[k * 4 + 3]); t_se[r] = acc;
      }
      quat_from_matrix_eigen(R_se, q);
      undistort_prepare(q, t_se, u);
    }
    for (int s0 = offsets[f]; s0 < offsets[f + 1]; s0 += 256) tiles.push_back(CloudTile{f, s0, std::min(256, offsets[f + 1] - s0), 0});
  }
  CK(ctx->m_a.ensure((size_t)n * 16)); CK(ctx->m_b.ensure((size_t)n * 16)); CK(ctx->m_c.ensure(tiles.size() * sizeof(CloudTile)));
  CK(ctx->m_d.ensure(prep.size() * sizeof(UndistortPrep))); CK(ctx->m_e.ensure((size_t)(n_frames + 1) * 4));
  CK(cudaMemcpyAsync(ctx->m_a.p, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_c.p, tiles.data(), tiles.size() * sizeof(CloudTile), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_d.p, prep.data(), prep.size() * sizeof(UndistortPrep), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_e.p, offsets, (size_t)(n_frames + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  k_undistort<<<(unsigned)tiles.size(), 256, 0, ctx->stream>>>(ctx->m_a.as<F4>(), ctx->m_c.as<CloudTile>(), ctx->m_d.as<UndistortPrep>(), ctx->m_e.as<int>(), ctx->m_b.as<F4>());
  CKL();
  CK(cudaMemcpyAsync(out, ctx->m_b.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // tiles / prep are host vectors: the copies above must have left them
  return PVB_OK;
}

// ================================================================ pair-level 5-NN and nearest line (segment-based variants of A3)
int pvb_pair_knn5(pvb_ctx* ctx, const float* ref_local, int n_ref, const double* R_ref, const double* t_ref, const float* nei_local, int n_nei, const double* R_nei,
                  const double* t_nei, float dist_threshold, double cell_size, int* idx5) {
  if (!ctx || n_ref < 0 || n_nei < 0 || (n_nei > 0 && (!nei_local || !idx5)) || (n_ref > 0 && !ref_local) || !R_ref || !t_ref || !R_nei || !t_nei)
    return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_pair_knn5: bad arguments") : PVB_ERR_ARG;
  if (n_nei == 0) return PVB_OK;
  if (n_ref < 5) { for (long i = 0; i < (long)n_nei * 5; ++i) idx5[i] = -1; return PVB_OK; }   // nearestKSearch returns < 5 neighbours (quirk C.6 guard)
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  int rc = set_cloudset(ctx, ctx->p_cs, {ref_local, nei_local}, {n_ref, n_nei}, {0, 1}, 256); if (rc) return rc;
  // the two T_wl go to the device as given (no angle-axis round trip)
  CK(ctx->h_pose.ensure(2 * sizeof(WorldPose))); CK(ctx->d_wpose.ensure(2 * sizeof(WorldPose)));
  CK(cudaStreamSynchronize(ctx->stream));
  WorldPose* hw = ctx->h_pose.as<WorldPose>();
  for (int k = 0; k < 9; ++k) { hw[0].R[k] = R_ref[k]; hw[1].R[k] = R_nei[k]; }
  for (int k = 0; k < 3; ++k) { hw[0].t[k] = t_ref[k]; hw[1].t[k] = t_nei[k]; }
  CK(cudaMemcpyAsync(ctx->d_wpose.p, hw, 2 * sizeof(WorldPose), cudaMemcpyHostToDevice, ctx->stream));
  rc = build_target_index(ctx, ctx->p_cs, ctx->p_index, cell_size); if (rc) return rc;
  const Pair pair{0, 1, 0, 1};
  std::vector<QueryTile> tiles;
  for (int s0 = 0; s0 < n_nei; s0 += kTile) tiles.push_back(QueryTile{0, n_ref + s0, std::min(kTile, n_nei - s0), s0});
  CK(ctx->f_pairs.ensure(sizeof(Pair))); CK(ctx->f_qtiles.ensure(tiles.size() * sizeof(QueryTile))); CK(ctx->f_nn_idx.ensure((size_t)n_nei * 5 * 4));
  CK(cudaMemcpyAsync(ctx->f_pairs.p, &pair, sizeof(Pair), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->f_qtiles.p, tiles.data(), tiles.size() * sizeof(QueryTile), cudaMemcpyHostToDevice, ctx->stream));
  LineAssocArgs a{};
  a.q_local = ctx->p_cs.local.as<F4>(); a.tiles = ctx->f_qtiles.as<QueryTile>(); a.pairs = ctx->f_pairs.as<Pair>(); a.grids = ctx->p_index.grids.as<GridDesc>();
  a.cell_start = ctx->p_index.cell_start.as<uint32_t>(); a.sorted = ctx->p_index.sorted.as<F4>(); a.wpose = ctx->d_wpose.as<WorldPose>();
  a.sq_thr = dist_threshold * dist_threshold; a.thr = (double)dist_threshold;
  a.out_nn = ctx->f_nn_idx.as<int>();
  k_associate_line<5><<<(int)tiles.size(), kTile, 0, ctx->stream>>>(a);
  CKL();
  CK(cudaMemcpyAsync(idx5, ctx->f_nn_idx.p, (size_t)n_nei * 5 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

int pvb_nearest_line(pvb_ctx* ctx, const double* lines_world6, int n_lines, const float* points_world, int n_points, int* line, double* dist) {
  if (!ctx || n_lines < 0 || n_points < 0 || (n_points > 0 && (!points_world || !line || !dist)) || (n_lines > 0 && !lines_world6))
    return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_nearest_line: bad arguments") : PVB_ERR_ARG;
  if (n_points == 0) return PVB_OK;
  if ((size_t)n_lines * 48 > 200 * 1024) return ctx->fail(PVB_ERR_ARG, "pvb_nearest_line: %d lines do not fit in shared memory", n_lines);
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  CK(ctx->m_a.ensure(std::max<size_t>(16, (size_t)n_lines * 48))); CK(ctx->m_b.ensure((size_t)n_points * 16)); CK(ctx->m_c.ensure((size_t)n_points * 4)); CK(ctx->m_d.ensure((size_t)n_points * 8));
  if (n_lines) CK(cudaMemcpyAsync(ctx->m_a.p, lines_world6, (size_t)n_lines * 48, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_b.p, points_world, (size_t)n_points * 16, cudaMemcpyHostToDevice, ctx->stream));
  const size_t smem = (size_t)n_lines * 48;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_nearest_line, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_nearest_line<<<(n_points + 127) / 128, 128, smem, ctx->stream>>>(ctx->m_a.as<double>(), n_lines, ctx->m_b.as<F4>(), n_points, ctx->m_c.as<int>(), ctx->m_d.as<double>());
  CKL();
  CK(cudaMemcpyAsync(line, ctx->m_c.p, (size_t)n_points * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(dist, ctx->m_d.p, (size_t)n_points * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

// ================================================================ E. line votes
int pvb_line_votes(pvb_ctx* ctx, const double* ref_lines, int S_ref, const float* pts, int n_pts, const int* p2s_off, const int* p2s_ids, int S_nei, double thr, int* M) {
  if (!ctx || !M || S_ref < 0 || S_nei < 0 || n_pts < 0) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  const size_t msz = (size_t)S_ref * S_nei;
  if (msz == 0) return PVB_OK;
  if (n_pts == 0) { memset(M, 0, msz * 4); return PVB_OK; }
  if (!ref_lines || !pts || !p2s_off || !p2s_ids) return ctx->fail(PVB_ERR_ARG, "null input");
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  const int n_ids = p2s_off[n_pts];
  CK(ctx->m_a.ensure((size_t)S_ref * 48)); CK(ctx->m_b.ensure((size_t)n_pts * 16)); CK(ctx->m_c.ensure((size_t)(n_pts + 1) * 4)); CK(ctx->m_d.ensure(std::max<size_t>(16, (size_t)n_ids * 4))); CK(ctx->m_e.ensure(msz * 4));
  CK(cudaMemcpyAsync(ctx->m_a.p, ref_lines, (size_t)S_ref * 48, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_b.p, pts, (size_t)n_pts * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_c.p, p2s_off, (size_t)(n_pts + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (n_ids) CK(cudaMemcpyAsync(ctx->m_d.p, p2s_ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->m_e.p, 0, msz * 4, ctx->stream));
  k_line_votes<<<(n_pts + 127) / 128, 128, (size_t)S_ref * 48, ctx->stream>>>(ctx->m_a.as<double>(), S_ref, ctx->m_b.as<F4>(), n_pts, ctx->m_c.as<int>(), ctx->m_d.as<int>(), thr, ctx->m_e.as<int>());
  CKL();
  CK(cudaMemcpyAsync(M, ctx->m_e.p, msz * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

// The vote matrices of AssociateLine2Line (LidarFeatureAssociate.cpp:442-476, TransformLines :219-236) for many frame pairs at once: every frame's corner cloud
// goes to the world frame in one launch (float32 store, as Transform2LidarWorld), every pair's votes in a second one, one download.  lines_world: the frames'
// segment lines already in the world frame (host, concatenated in frame order).  M: pair p's S_nei x S_ref matrix at m_off[p].  world_out (may be NULL): the
// world-frame corner clouds, concatenated in frame order (the residual builders read the points from it).
int pvb_line_votes_batch(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, const double* lines_world, int n_pairs, const int* pair_ref, const int* pair_nei,
                         const long long* m_off, long long m_total, double thr, int* M, float* world_out) {
  if (!ctx || n_frames < 0 || n_pairs < 0 || (n_frames > 0 && !frames) || (n_pairs > 0 && (!pair_ref || !pair_nei || !m_off))) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_line_votes_batch: bad arguments") : PVB_ERR_ARG;
  ctx->line_layout.valid = false; ctx->line_pending.pending = false;   // the tables the device tail reads are about to be rewritten
  if (n_frames == 0) return PVB_OK;
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  std::vector<int> coff(n_frames + 1, 0), soff(n_frames + 1, 0), pbase(n_frames, 0);
  long long n_ids = 0;
  for (int f = 0; f < n_frames; ++f) {
    coff[f + 1] = coff[f] + frames[f].n_corner; soff[f + 1] = soff[f] + frames[f].n_segments;
    pbase[f] = (int)n_ids;
    n_ids += frames[f].n_corner > 0 ? frames[f].p2s_off[frames[f].n_corner] : 0;
  }
  const int nc = coff[n_frames], ns = soff[n_frames];
  if (ns > 0 && !lines_world) return ctx->fail(PVB_ERR_ARG, "pvb_line_votes_batch: lines_world is NULL");
  // host staging: local clouds, CSR offsets (n + 1 per frame), ids, poses
  std::vector<F4> local((size_t)std::max(1, nc));
  std::vector<int> poff((size_t)nc + n_frames, 0), pids((size_t)std::max<long long>(1, n_ids));
  std::vector<WorldPose> wp(n_frames);
  std::vector<CloudTile> ctiles; std::vector<int> cblock(n_frames);
  for (int f = 0; f < n_frames; ++f) {
    const pvb_line_frame& fr = frames[f];
    if (fr.n_corner > 0) {
      memcpy(&local[coff[f]], fr.corner_local, (size_t)fr.n_corner * 16);
      memcpy(&poff[(size_t)coff[f] + f], fr.p2s_off, (size_t)(fr.n_corner + 1) * 4);
      const int m = fr.p2s_off[fr.n_corner];
      if (m > 0) memcpy(&pids[pbase[f]], fr.p2s_ids, (size_t)m * 4);
    }
    for (int k = 0; k < 9; ++k) wp[f].R[k] = fr.R_wl[k];
    for (int k = 0; k < 3; ++k) wp[f].t[k] = fr.t_wl[k];
    cblock[f] = f;
    for (int s0 = 0; s0 < fr.n_corner; s0 += 256) ctiles.push_back(CloudTile{f, coff[f] + s0, std::min(256, fr.n_corner - s0), 0});
  }
  std::vector<VotePair> vp(std::max(1, n_pairs)); std::vector<VoteTile> vt;
  int smax = 0;
  for (int p = 0; p < n_pairs; ++p) {
    const int r = pair_ref[p], n = pair_nei[p];
    if (r < 0 || r >= n_frames || n < 0 || n >= n_frames) return ctx->fail(PVB_ERR_ARG, "pvb_line_votes_batch: pair %d out of range", p);
    vp[p] = VotePair{r, n, m_off[p]};
    if (frames[r].n_segments == 0 || frames[n].n_segments == 0) continue;      // CheckLidarSegment: nothing to vote on
    smax = std::max(smax, frames[r].n_segments);
    for (int s0 = 0; s0 < frames[n].n_corner; s0 += 128) vt.push_back(VoteTile{p, s0, std::min(128, frames[n].n_corner - s0), 0});
  }
  DevBuf& d_local = ctx->v_local; DevBuf& d_world = ctx->v_world; DevBuf& d_misc = ctx->v_misc; DevBuf& d_M = ctx->v_M;
  // one packed upload of the small integer / double tables
  const size_t o_coff = 0, o_soff = o_coff + (size_t)(n_frames + 1) * 4, o_pbase = o_soff + (size_t)(n_frames + 1) * 4, o_poff = o_pbase + (size_t)n_frames * 4,
               o_pids = o_poff + poff.size() * 4, o_cblock = o_pids + pids.size() * 4, o_ctiles = (o_cblock + (size_t)n_frames * 4 + 15) / 16 * 16,
               o_wp = o_ctiles + std::max<size_t>(1, ctiles.size()) * sizeof(CloudTile), o_lines = o_wp + (size_t)n_frames * sizeof(WorldPose),
               o_vp = o_lines + (size_t)std::max(1, ns) * 48, o_vt = o_vp + vp.size() * sizeof(VotePair), total = o_vt + std::max<size_t>(1, vt.size()) * sizeof(VoteTile);
  CK(ctx->mh_a.ensure(total)); CK(d_misc.ensure(total));
  unsigned char* h = ctx->mh_a.as<unsigned char>();
  memcpy(h + o_coff, coff.data(), (size_t)(n_frames + 1) * 4); memcpy(h + o_soff, soff.data(), (size_t)(n_frames + 1) * 4); memcpy(h + o_pbase, pbase.data(), (size_t)n_frames * 4);
  memcpy(h + o_poff, poff.data(), poff.size() * 4); memcpy(h + o_pids, pids.data(), pids.size() * 4); memcpy(h + o_cblock, cblock.data(), (size_t)n_frames * 4);
  if (!ctiles.empty()) memcpy(h + o_ctiles, ctiles.data(), ctiles.size() * sizeof(CloudTile));
  memcpy(h + o_wp, wp.data(), (size_t)n_frames * sizeof(WorldPose));
  if (ns > 0) memcpy(h + o_lines, lines_world, (size_t)ns * 48);
  memcpy(h + o_vp, vp.data(), vp.size() * sizeof(VotePair));
  if (!vt.empty()) memcpy(h + o_vt, vt.data(), vt.size() * sizeof(VoteTile));
  CK(d_local.ensure(local.size() * 16)); CK(d_world.ensure(local.size() * 16)); CK(d_M.ensure(std::max<size_t>(16, (size_t)m_total * 4)));
  CK(cudaMemcpyAsync(d_misc.p, h, total, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_local.p, local.data(), local.size() * 16, cudaMemcpyHostToDevice, ctx->stream));
  unsigned char* d = d_misc.as<unsigned char>();
  if (!ctiles.empty()) {
    k_transform_world<<<(unsigned)ctiles.size(), 256, 0, ctx->stream>>>(d_local.as<F4>(), reinterpret_cast<const CloudTile*>(d + o_ctiles), reinterpret_cast<const int*>(d + o_cblock),
                                                                        reinterpret_cast<const WorldPose*>(d + o_wp), reinterpret_cast<const int*>(d + o_coff), d_world.as<F4>(), nullptr);
    CKL();
  }
  if (m_total > 0) CK(cudaMemsetAsync(d_M.p, 0, (size_t)m_total * 4, ctx->stream));
  if (!vt.empty()) {
    const size_t smem = (size_t)smax * 48;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_line_votes_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_line_votes_batch<<<(unsigned)vt.size(), 128, smem, ctx->stream>>>(reinterpret_cast<const VoteTile*>(d + o_vt), reinterpret_cast<const VotePair*>(d + o_vp), d_world.as<F4>(),
                                                                        reinterpret_cast<const int*>(d + o_coff), reinterpret_cast<const int*>(d + o_poff), reinterpret_cast<const int*>(d + o_pbase),
                                                                        reinterpret_cast<const int*>(d + o_pids), reinterpret_cast<const double*>(d + o_lines), reinterpret_cast<const int*>(d + o_soff),
                                                                        thr, d_M.as<int>());
    CKL();
  }
  if (M && m_total > 0) CK(cudaMemcpyAsync(M, d_M.p, (size_t)m_total * 4, cudaMemcpyDeviceToHost, ctx->stream));   // M == NULL: the matrices stay on the device (device tail)
  if (world_out && nc > 0) CK(cudaMemcpyAsync(world_out, d_world.p, (size_t)nc * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->line_layout = pvb_ctx::LineLayout{true, n_frames, n_pairs, o_coff, o_soff, o_pbase, o_poff, o_pids, o_wp, o_lines, o_vp};
  return PVB_OK;
}

// AddLidarLineToLineResidual2 (util/Optimization.cpp:329-441) for a whole pose graph with the tails on the device (pvb_lines.cuh): the vote pass above with the
// matrices left in HBM, FindAssociations + track gate per edge (k_line_assoc), and - when the next pvb_frames_point2plane_blocks sizes the block arrays - one
// Point2Line block per member point of every kept neighbour segment (k_line_blocks_emit).  Only the per-edge block COUNTS come back to the host.
int pvb_frames_line2line_blocks_device(pvb_ctx* ctx, int n_frames, const pvb_line_frame* frames, int n_edges, const int* ref, const int* nei, double dist_threshold,
                                       int n_tracks, const int* track_off, const int* feat_frame, const int* feat_line, int angle_residual, int normalize_distance,
                                       double weight, long* n_blocks) {
  if (!ctx || n_frames < 0 || n_edges < 0 || (n_frames > 0 && !frames) || (n_edges > 0 && (!ref || !nei))) return ctx ? ctx->fail(PVB_ERR_ARG, "pvb_frames_line2line_blocks_device: bad arguments") : PVB_ERR_ARG;
  if (n_tracks > 0 && (!track_off || !feat_frame || !feat_line)) return ctx->fail(PVB_ERR_ARG, "pvb_frames_line2line_blocks_device: tracks are NULL");
  if (n_blocks) *n_blocks = 0;
  ctx->line_pending.pending = false;
  if (n_frames == 0 || n_edges == 0) return PVB_OK;
  std::vector<int> soff(n_frames + 1, 0);
  for (int f = 0; f < n_frames; ++f) soff[f + 1] = soff[f] + frames[f].n_segments;
  const int ns = soff[n_frames];
  std::vector<long long> m_off(n_edges + 1, 0); std::vector<int> h_off(n_edges + 1, 0);
  for (int e = 0; e < n_edges; ++e) {
    if (ref[e] < 0 || ref[e] >= n_frames || nei[e] < 0 || nei[e] >= n_frames) return ctx->fail(PVB_ERR_ARG, "pvb_frames_line2line_blocks_device: edge %d out of range", e);
    m_off[e + 1] = m_off[e] + (long long)frames[ref[e]].n_segments * frames[nei[e]].n_segments;
    h_off[e + 1] = h_off[e] + frames[ref[e]].n_segments;
  }
  const int H = h_off[n_edges];
  if (ns == 0 || H == 0) return PVB_OK;
  // world-frame segment lines (TransformLines, LidarFeatureAssociate.cpp:219-236), the local coefficients, segment sizes and the track of every line
  std::vector<double> lines((size_t)ns * 6), coeffs((size_t)ns * 6);
  std::vector<int> seg_size(ns, 0), member(ns, 0), track_of(ns, -1), slot_edge(H);
  for (int f = 0; f < n_frames; ++f) {
    const pvb_line_frame& fr = frames[f];
    for (int s = 0; s < fr.n_segments; ++s) {
      const double* in = fr.segment_coeffs + 6 * s; double* out = &lines[(size_t)(soff[f] + s) * 6];
      for (int r = 0; r < 3; ++r) {
        out[r] = fr.R_wl[r * 3] * in[0] + fr.R_wl[r * 3 + 1] * in[1] + fr.R_wl[r * 3 + 2] * in[2] + fr.t_wl[r];
        out[3 + r] = fr.R_wl[r * 3] * in[3] + fr.R_wl[r * 3 + 1] * in[4] + fr.R_wl[r * 3 + 2] * in[5];
      }
      memcpy(&coeffs[(size_t)(soff[f] + s) * 6], in, 48);
    }
    for (int i = 0; i < fr.n_corner; ++i) {
      for (int q = fr.p2s_off[i]; q < fr.p2s_off[i + 1]; ++q) {
        const int id = fr.p2s_ids[q];
        if (id < 0 || id >= fr.n_segments) return ctx->fail(PVB_ERR_ARG, "pvb_frames_line2line_blocks_device: frame %d point %d names segment %d of %d", f, i, id, fr.n_segments);
        seg_size[soff[f] + id]++;
        bool first = true;                                             // a point is a member once, however often it lists the segment
        for (int q2 = fr.p2s_off[i]; q2 < q; ++q2) first = first && fr.p2s_ids[q2] != id;
        if (first) member[soff[f] + id]++;
      }
    }
  }
  if (n_tracks >= 0)
    for (int t = 0; t < n_tracks; ++t)
      for (int k = track_off[t]; k < track_off[t + 1]; ++k) {
        const int fr = feat_frame[k], ln = feat_line[k];
        if (fr < 0 || fr >= n_frames || ln < 0 || ln >= frames[fr].n_segments) return ctx->fail(PVB_ERR_ARG, "pvb_frames_line2line_blocks_device: track %d names line %d of frame %d", t, ln, fr);
        track_of[soff[fr] + ln] = t;
      }
  for (int e = 0; e < n_edges; ++e) for (int k = h_off[e]; k < h_off[e + 1]; ++k) slot_edge[k] = e;
  int rc = pvb_line_votes_batch(ctx, n_frames, frames, lines.data(), n_edges, ref, nei, m_off.data(), m_off[n_edges], dist_threshold, nullptr, nullptr);
  if (rc) return rc;
  const pvb_ctx::LineLayout& L = ctx->line_layout;
  if (!L.valid) return ctx->fail(PVB_ERR_STATE, "pvb_frames_line2line_blocks_device: the vote pass left no tables");
  // the tail's own tables in one upload: seg_size | member | track_of | h_off | slot_edge | local coefficients
  const size_t o_size = 0, o_member = o_size + (size_t)ns * 4, o_track = o_member + (size_t)ns * 4, o_hoff = o_track + (size_t)ns * 4, o_slot = o_hoff + (size_t)(n_edges + 1) * 4,
               o_coeffs = (o_slot + (size_t)H * 4 + 15) / 16 * 16, total = o_coeffs + (size_t)ns * 48;
  std::vector<unsigned char> hbuf(total);
  memcpy(&hbuf[o_size], seg_size.data(), (size_t)ns * 4); memcpy(&hbuf[o_member], member.data(), (size_t)ns * 4); memcpy(&hbuf[o_track], track_of.data(), (size_t)ns * 4);
  memcpy(&hbuf[o_hoff], h_off.data(), (size_t)(n_edges + 1) * 4); memcpy(&hbuf[o_slot], slot_edge.data(), (size_t)H * 4); memcpy(&hbuf[o_coeffs], coeffs.data(), (size_t)ns * 48);
  CK(ctx->lt_misc.ensure(total)); CK(ctx->lt_hold.ensure((size_t)H * 4)); CK(ctx->lt_kbase.ensure((size_t)H * 4)); CK(ctx->lt_cnt.ensure((size_t)n_edges * 4));
  CK(cudaMemcpyAsync(ctx->lt_misc.p, hbuf.data(), total, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned char* d = ctx->v_misc.as<unsigned char>();
  const unsigned char* t = ctx->lt_misc.as<unsigned char>();
  k_line_assoc<<<(n_edges + 127) / 128, 128, 0, ctx->stream>>>(reinterpret_cast<const VotePair*>(d + L.o_vp), n_edges, ctx->v_M.as<int>(), reinterpret_cast<const int*>(d + L.o_soff),
                                                              reinterpret_cast<const double*>(d + L.o_lines), reinterpret_cast<const int*>(t + o_size), reinterpret_cast<const int*>(t + o_member),
                                                              n_tracks >= 0 ? reinterpret_cast<const int*>(t + o_track) : nullptr, reinterpret_cast<const int*>(t + o_hoff),
                                                              ctx->lt_hold.as<int>(), ctx->lt_kbase.as<int>(), ctx->lt_cnt.as<int>());
  CKL();
  pvb_ctx::LinePending& P = ctx->line_pending;
  P.cnt.assign(n_edges, 0);
  CK(cudaMemcpyAsync(P.cnt.data(), ctx->lt_cnt.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  P.ref.assign(ref, ref + n_edges); P.nei.assign(nei, nei + n_edges);
  P.total = 0;
  for (int e = 0; e < n_edges; ++e) P.total += P.cnt[e];
  P.slots = H; P.type = angle_residual ? PVB_P2LINE_ANGLE : PVB_P2LINE_METER; P.normalize = normalize_distance; P.huber = angle_residual ? 0.0 : 0.2; P.weight = weight;
  P.o_hoff = o_hoff; P.o_slot = o_slot; P.o_coeffs = o_coeffs;
  P.pending = P.total > 0;
  if (n_blocks) *n_blocks = (long)P.total;
  return PVB_OK;
}

// ================================================================ F. camera-LiDAR angle votes
int pvb_angle_votes(pvb_ctx* ctx, int rows, int cols, const float* lines4, int L, const float* cloud_local, int P, const int* p2s_off, const int* p2s_ids, int S, const double* T_cl, int* counts) {
  if (!ctx || !counts || L < 0 || P < 0 || S < 0) return ctx ? ctx->fail(PVB_ERR_ARG, "bad arguments") : PVB_ERR_ARG;
  const size_t csz = (size_t)L * S;
  if (csz == 0) return PVB_OK;
  if (P == 0) { memset(counts, 0, csz * 4); return PVB_OK; }
  if (!lines4 || !cloud_local || !p2s_off || !p2s_ids || !T_cl) return ctx->fail(PVB_ERR_ARG, "null input");
  CK(cudaSetDevice(ctx->device));
  if (int qrc = quiesce_copy(ctx)) return qrc;
  // image-line planes on the host (CameraLidarLineAssociate.cpp:381-387): ImageToCam uses exact sin/cos
  std::vector<ImageLinePlane> lp(L);
  for (int l = 0; l < L; ++l) {
    double p1[3], p2[3];
    image_to_cam_f64(lines4[l * 4], lines4[l * 4 + 1], rows, cols, p1);
    image_to_cam_f64(lines4[l * 4 + 2], lines4[l * 4 + 3], rows, cols, p2);
    // FormPlane(p1, p2, 0): (p2-p1) x (0-p1)
    const double a = (p2[1] - p1[1]) * (0 - p1[2]) - (p2[2] - p1[2]) * (0 - p1[1]);
    const double b = (p2[2] - p1[2]) * (0 - p1[0]) - (p2[0] - p1[0]) * (0 - p1[2]);
    const double c = (p2[0] - p1[0]) * (0 - p1[1]) - (p2[1] - p1[1]) * (0 - p1[0]);
    const double d = -(a * p1[0] + b * p1[1] + c * p1[2]);
    const double nn = std::sqrt(a * a + b * b + c * c + d * d);
    lp[l].n[0] = a / nn; lp[l].n[1] = b / nn; lp[l].n[2] = c / nn; lp[l].n[3] = d / nn;
    for (int k = 0; k < 3; ++k) lp[l].p4[k] = (p1[k] + p2[k]) / 2.0;
    double cs = p1[0] * lp[l].p4[0] + p1[1] * lp[l].p4[1] + p1[2] * lp[l].p4[2];
    const double n1 = std::sqrt(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]);
    const double n2 = std::sqrt(lp[l].p4[0] * lp[l].p4[0] + lp[l].p4[1] * lp[l].p4[1] + lp[l].p4[2] * lp[l].p4[2]);
    cs /= (n1 * n2);
    lp[l].scope = cs >= 1.0 ? 0.0 : (cs <= -1.0 ? M_PI : std::acos(cs));
  }
  const int n_ids = p2s_off[P];
  CK(ctx->m_a.ensure(lp.size() * sizeof(ImageLinePlane))); CK(ctx->m_b.ensure((size_t)P * 16)); CK(ctx->m_c.ensure((size_t)(P + 1) * 4)); CK(ctx->m_d.ensure(std::max<size_t>(16, (size_t)n_ids * 4))); CK(ctx->m_e.ensure(csz * 4));
  CK(cudaMemcpyAsync(ctx->m_a.p, lp.data(), lp.size() * sizeof(ImageLinePlane), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_b.p, cloud_local, (size_t)P * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->m_c.p, p2s_off, (size_t)(P + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (n_ids) CK(cudaMemcpyAsync(ctx->m_d.p, p2s_ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->m_e.p, 0, csz * 4, ctx->stream));
  WorldPose w; T16_to_pose(T_cl, w);
  dim3 grid((P + 127) / 128, L);
  k_angle_votes<<<grid, 128, 0, ctx->stream>>>(ctx->m_a.as<ImageLinePlane>(), L, ctx->m_b.as<F4>(), P, w, ctx->m_c.as<int>(), ctx->m_d.as<int>(), S, ctx->m_e.as<int>());
  CKL();
  CK(cudaMemcpyAsync(counts, ctx->m_e.p, csz * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PVB_OK;
}

}  // extern "C"
