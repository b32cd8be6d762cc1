// TEST INFRASTRUCTURE ONLY — compiles the product's `__host__ __device__` math (panovlm_b200/csrc/*.cuh,
// pvb_host.hpp) with g++ so the not-gpu tests can check the exact device arithmetic against the oracle on the
// CPU.  It is never linked into libpanovlm_b200.so and the product has no CPU path.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../panovlm_b200/csrc/pvb_host.hpp"
#include "../panovlm_b200/csrc/pvb_knn.cuh"

using namespace pvb;

// per-query association on a host-built grid (counting sort), K = 10 or 5
static int g_prune = 1;   // 0: exhaustive block walk (TMA-staged variant), 1 / 2: pruned two-pass walk from a 3x3x3 / 5x5x5 block, 3: buffered single pass (frames mode default), 4: the same over merged super-rows (dense mode default)
static int g_static = 1;                // mode 4: queries without a usable hint take the static bound of the target
static int g_flat = 1;                  // mode 3: hinted walk over the flattened row ranges (device default) or the nested per-row loops
static const float* g_hint = nullptr;   // mode 3: per query {x, y, z, tau} search-radius hints (the device kernel's formula), or null
static float* g_hint_out = nullptr;     // mode 3: the hints the device kernel would store
template <int K>
static void associate_all(const float* tgt, int n, const double* R_ref, const double* t_ref, const float* qry, int m, const double* R_nei, const double* t_nei,
                          double h, float thr, double plane_tol, unsigned char* valid, double* p_local, double* plane, int* nn_idx, float* nn_d2) {
  GridDesc g;
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], tgt[i * 4 + c]); mx[c] = std::max(mx[c], tgt[i * 4 + c]); }
  for (int c = 0; c < 3; ++c) { g.origin[c] = mn[c]; g.dims[c] = std::max(1, (int)std::floor(((double)mx[c] - mn[c]) / h) + 1); }
  g.h = h; g.inv_h = 1.0 / h; g.n_points = n; g.cell_base = 0; g.point_base = 0;
  const long long ncell = (long long)g.dims[0] * g.dims[1] * g.dims[2];
  std::vector<int> start(ncell + 1, 0), cell(n);
  for (int i = 0; i < n; ++i) {
    const int cx = cell_coord(tgt[i * 4], g.origin[0], g.inv_h, g.dims[0]), cy = cell_coord(tgt[i * 4 + 1], g.origin[1], g.inv_h, g.dims[1]),
              cz = cell_coord(tgt[i * 4 + 2], g.origin[2], g.inv_h, g.dims[2]);
    cell[i] = (int)(((long long)cz * g.dims[1] + cy) * g.dims[0] + cx);
    start[cell[i] + 1]++;
  }
  for (long long c = 0; c < ncell; ++c) start[c + 1] += start[c];
  std::vector<int> fill(start.begin(), start.end() - 1);
  std::vector<F4> sorted(n);
  for (int i = 0; i < n; ++i) {
    F4 r; r.x = tgt[i * 4]; r.y = tgt[i * 4 + 1]; r.z = tgt[i * 4 + 2];
    r.w = u2f(((uint32_t)i << 5) | ((uint32_t)tgt[i * 4 + 3] & 31u));
    sorted[fill[cell[i]]++] = r;
  }
  auto cells = [&](long long c) { return (long long)start[c]; };
  auto load = [&](long long i) { return sorted[i]; };
  // merged super-rows (mode 4), built like k_superrow_counts / k_superrow_fill
  struct HostSuperRow {
    std::vector<uint32_t> st, wv; std::vector<float> quads, rkv;
    float rk(uint32_t r) const { return rkv[r]; }
    uint32_t w(uint32_t r) const { return wv[r]; }
    uint32_t start(long long i) const { return st[i]; }
    struct Quad { float v[12]; };
    Quad load3(uint32_t G) const { Quad q; memcpy(q.v, quads.data() + (size_t)G * 12, 48); return q; }
    void sqdist4(const Quad& c, float qx, float qy, float qz, uint32_t (&kb)[4]) const {
      for (int k = 0; k < 4; ++k) kb[k] = f2u(sqdist_f32(qx, qy, qz, c.v[k], c.v[4 + k], c.v[8 + k]));
    }
  } sr;
  if (g_prune == 4) {
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    // static bound of every target record like k_target_rk: squared distance to its 10th nearest target point (itself included) within 1.5 cells, else +inf
    std::vector<float> rk(n);
    { const float rthr = (float)(1.5 * h); const float rthr2 = rthr * rthr; std::vector<float> dd(n);
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) dd[j] = sqdist_f32(sorted[i].x, sorted[i].y, sorted[i].z, sorted[j].x, sorted[j].y, sorted[j].z);
        if (n >= 10) { std::nth_element(dd.begin(), dd.begin() + 9, dd.end()); rk[i] = dd[9] <= rthr2 ? dd[9] : INFINITY; } else rk[i] = INFINITY;
      } }
    // merged super-rows like k_superrow_counts / k_superrow_fill: segments padded to groups of 4 with points at +infinity
    sr.st.assign(ncell + 1, 0u);
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t run = 0;
      auto put = [&](uint32_t r, float px, float py, float pz, uint32_t w, float k) {
        float* q = sr.quads.data() + (size_t)(r >> 2) * 12 + (r & 3u);
        q[0] = px; q[4] = py; q[8] = pz; sr.wv[r] = w; sr.rkv[r] = k;
      };
      for (long long c = 0; c < ncell; ++c) {
        const int x = (int)(c % nx); const long long row = c / nx; const int y = (int)(row % ny), z = (int)(row / ny);
        if (pass == 0) sr.st[c] = run;
        for (int dz = -1; dz <= 1; ++dz) {
          const int zz = z + dz; if (zz < 0 || zz >= nz) continue;
          for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy; if (yy < 0 || yy >= ny) continue;
            const long long cc = ((long long)zz * ny + yy) * nx + x;
            for (int i = start[cc]; i < start[cc + 1]; ++i) {
              if (pass == 1) put(run, sorted[i].x, sorted[i].y, sorted[i].z, ((uint32_t)i << 5) | (f2u(sorted[i].w) & 31u), rk[i]);
              ++run;
            }
          }
        }
        while (run & 3u) { if (pass == 1) put(run, INFINITY, INFINITY, INFINITY, 0u, INFINITY); ++run; }
      }
      if (pass == 0) { sr.st[ncell] = run; sr.quads.assign((size_t)std::max<uint32_t>(run, 4u) * 3, 0.f); sr.wv.assign(run, 0u); sr.rkv.assign(run, 0.f); }
    }
  }
  AssocParams prm; prm.sq_thr = thr * thr; prm.rmax = (int)std::ceil((double)thr / h); prm.plane_tol = plane_tol; prm.collinear_tol = 3.0; prm.r0 = g_prune == 2 ? 2 : 1;
  for (int i = 0; i < m; ++i) {
    uint32_t wpos[K];
    for (int j = 0; j < K; ++j) wpos[j] = 0xFFFFFFFFu;
    const uint32_t qcls = (uint32_t)qry[i * 4 + 3] & 31u;
    auto win = [&](int j) { return wpos[j]; };
    auto set_win = [&](int j, uint32_t pos) { wpos[j] = pos; };
    uint32_t rng[18];
    auto range_set = [&](int k, uint32_t lo, uint32_t hi) { rng[2 * k] = lo; rng[2 * k + 1] = hi; };
    auto range_get = [&](int k, uint32_t& lo, uint32_t& hi) { lo = rng[2 * k]; hi = rng[2 * k + 1]; };
    auto no_map = [](int, int, uint32_t&, uint32_t&) {};
    constexpr int LC = 24;
    U2 lst[LC];
    auto win2 = [&](int j) { return lst[j].y; };
    auto set_win2 = [&](int j, uint32_t pos) { lst[j].y = pos; };
    const float qx = qry[i * 4], qy = qry[i * 4 + 1], qz = qry[i * 4 + 2];
    uint32_t lim_hint = 0u, tau = 0x7F800000u;
    if (g_prune >= 3 && g_hint) {        // same arithmetic as k_associate (pvb_kernels.cuh)
      const float* hq = g_hint + 4 * i;
      if (hq[3] >= 0.f && hq[3] < 3.0e38f) {
        const double dx = (double)qx - (double)hq[0], dy = (double)qy - (double)hq[1], dz = (double)qz - (double)hq[2];
        const double rad = (sqrt((double)hq[3]) + sqrt(dx * dx + dy * dy + dz * dz)) * (1.0 + 1e-5) + 1e-9;
        const double lim2 = rad * rad;
        if (lim2 < (double)prm.sq_thr) lim_hint = f2u((float)lim2) + 2u;
      }
    }
#define PVBH_ASSOC(MODE, W, SW) associate_point2plane<K, false, MODE, LC>(g, cells, load, load, no_map, prm, qx, qy, qz, qcls, R_ref, t_ref, R_nei, t_nei, \
                                                               p_local + 3 * i, plane + 4 * i, W, SW, range_set, range_get, lim_hint, &tau, lst, 1, g_flat != 0)
    if (g_prune >= 3) { for (int j = 0; j < K; ++j) lst[j].y = 0xFFFFFFFFu; }
    if (g_prune == 4)
      valid[i] = associate_point2plane<K, false, 4, LC>(g, cells, load, load, no_map, prm, qx, qy, qz, qcls, R_ref, t_ref, R_nei, t_nei, p_local + 3 * i, plane + 4 * i, win2, set_win2,
                                                        range_set, range_get, lim_hint, &tau, lst, 1, false, sr, g_static != 0) ? 1 : 0;
    else
    valid[i] = (g_prune == 0 ? PVBH_ASSOC(0, win, set_win) : (g_prune == 3 ? PVBH_ASSOC(2, win2, set_win2) : PVBH_ASSOC(1, win, set_win))) ? 1 : 0;
#undef PVBH_ASSOC
    if (g_prune >= 3) {
      for (int j = 0; j < K; ++j) wpos[j] = tau == 0x7F800000u ? 0xFFFFFFFFu : lst[j].y;
      if (g_hint_out) { g_hint_out[4 * i] = qx; g_hint_out[4 * i + 1] = qy; g_hint_out[4 * i + 2] = qz; g_hint_out[4 * i + 3] = u2f(tau); }
    }
    std::vector<std::pair<std::pair<float, uint32_t>, int>> nn;
    for (int j = 0; j < K; ++j) {
      if (wpos[j] == 0xFFFFFFFFu) continue;
      const F4 r = sorted[wpos[j]];
      nn.push_back({{sqdist_f32(qry[i * 4], qry[i * 4 + 1], qry[i * 4 + 2], r.x, r.y, r.z), wpos[j]}, (int)(f2u(r.w) >> 5)});
    }
    std::sort(nn.begin(), nn.end());
    for (int j = 0; j < K; ++j) {
      nn_idx[i * K + j] = j < (int)nn.size() ? nn[j].second : -1;
      nn_d2[i * K + j] = j < (int)nn.size() ? nn[j].first.first : INFINITY;
    }
  }
}

static void associate_lines(const float* tgt, int n, const double* R_ref, const double* t_ref, const float* qry, int m, const double* R_nei, const double* t_nei,
                            double h, float thr, unsigned char* valid, double* p_local, double* pa, double* pb) {
  GridDesc g;
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], tgt[i * 4 + c]); mx[c] = std::max(mx[c], tgt[i * 4 + c]); }
  for (int c = 0; c < 3; ++c) { g.origin[c] = mn[c]; g.dims[c] = std::max(1, (int)std::floor(((double)mx[c] - mn[c]) / h) + 1); }
  g.h = h; g.inv_h = 1.0 / h; g.n_points = n; g.cell_base = 0; g.point_base = 0;
  const long long ncell = (long long)g.dims[0] * g.dims[1] * g.dims[2];
  std::vector<int> start(ncell + 2, 0), cell(n);
  for (int i = 0; i < n; ++i) {
    cell[i] = (int)(((long long)cell_coord(tgt[i * 4 + 2], g.origin[2], g.inv_h, g.dims[2]) * g.dims[1] + cell_coord(tgt[i * 4 + 1], g.origin[1], g.inv_h, g.dims[1])) * g.dims[0] +
                    cell_coord(tgt[i * 4], g.origin[0], g.inv_h, g.dims[0]));
    start[cell[i] + 1]++;
  }
  for (long long c = 0; c <= ncell; ++c) start[c + 1] += start[c];
  std::vector<int> fill(start.begin(), start.end() - 1);
  std::vector<F4> sorted(n);
  for (int i = 0; i < n; ++i) { F4 r; r.x = tgt[i * 4]; r.y = tgt[i * 4 + 1]; r.z = tgt[i * 4 + 2]; r.w = u2f((uint32_t)i << 5); sorted[fill[cell[i]]++] = r; }
  auto cells = [&](long long c) { return (long long)start[c]; };
  auto load = [&](long long i) { return sorted[i]; };
  for (int i = 0; i < m; ++i) {
    uint32_t wpos[5], rng[18];
    auto win = [&](int j) { return wpos[j]; };
    auto set_win = [&](int j, uint32_t pos) { wpos[j] = pos; };
    auto range_set = [&](int k, uint32_t lo, uint32_t hi) { rng[2 * k] = lo; rng[2 * k + 1] = hi; };
    auto range_get = [&](int k, uint32_t& lo, uint32_t& hi) { lo = rng[2 * k]; hi = rng[2 * k + 1]; };
    valid[i] = associate_point2line<5>(g, cells, load, thr * thr, (int)std::ceil((double)thr / h), qry[i * 4], qry[i * 4 + 1], qry[i * 4 + 2], R_ref, t_ref, R_nei, t_nei,
                                       p_local + 3 * i, pa + 3 * i, pb + 3 * i, win, set_win, range_set, range_get) ? 1 : 0;
  }
}

extern "C" {
void pvbh_set_prune(int on) { g_prune = on; }
void pvbh_set_flat(int on) { g_flat = on; }
void pvbh_set_static(int on) { g_static = on; }
void pvbh_set_hints(const float* hint_in, float* hint_out) { g_hint = hint_in; g_hint_out = hint_out; }


void pvbh_associate_lines(const float* tgt, int n, const double* R_ref, const double* t_ref, const float* qry, int m, const double* R_nei, const double* t_nei,
                          double h, float thr, unsigned char* valid, double* p_local, double* pa, double* pb) {
  associate_lines(tgt, n, R_ref, t_ref, qry, m, R_nei, t_nei, h, thr, valid, p_local, pa, pb);
}

void pvbh_pose_prep(const double* pose6, double* out21) {
  PosePrep p; prepare_pose(pose6, p);
  memcpy(out21, p.R, 72); memcpy(out21 + 9, p.Jl, 72); memcpy(out21 + 18, p.t, 24);
}

void pvbh_eval_blocks(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts,
                      const double* poses, int nb, int apply_loss, double* out_r, double* out_J, double* out_cost) {
  std::vector<PosePrep> pp(nb);
  for (int b = 0; b < nb; ++b) prepare_pose(poses + 6 * b, pp[b]);
  for (long i = 0; i < n; ++i) {
    double J[12];
    double r = eval_block(type[i], normalize[i] != 0, consts + 12 * i, pp[ref[i]], pp[nei[i]], J);
    const double c = huber_correct(apply_loss ? huber[i] : 0.0, r, J, 12);
    out_r[i] = r; out_cost[i] = c;
    if (out_J) memcpy(out_J + 12 * i, J, 96);
  }
}

// dense LM through the product's host solver, residuals evaluated with the device math on the CPU
void pvbh_solve_lm(long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts,
                   double* poses, int nb, const unsigned char* is_const, int max_iter, double* summary) {
  EvalFn eval = [&](const double* x, double* H, double* g) {
    std::vector<PosePrep> pp(nb);
    for (int b = 0; b < nb; ++b) prepare_pose(x + 6 * b, pp[b]);
    const int D = 6 * nb;
    if (H) std::fill(H, H + (size_t)D * D, 0.0);
    if (g) std::fill(g, g + D, 0.0);
    double cost = 0;
    for (long i = 0; i < n; ++i) {
      double J[12];
      double r = eval_block(type[i], normalize[i] != 0, consts + 12 * i, pp[ref[i]], pp[nei[i]], J);
      cost += huber_correct(huber[i], r, J, 12);
      if (!H) continue;
      const int o[2] = {6 * ref[i], 6 * nei[i]};
      for (int a = 0; a < 12; ++a) {
        const int ia = o[a / 6] + a % 6;
        g[ia] += J[a] * r;
        for (int c = 0; c < 12; ++c) H[(size_t)ia * D + o[c / 6] + c % 6] += J[a] * J[c];
      }
    }
    return cost;
  };
  LMOptions opt; opt.max_iterations = max_iter;
  LMSummary S = solve_lm(eval, poses, nb, is_const, opt);
  summary[0] = S.initial_cost; summary[1] = S.final_cost; summary[2] = S.iterations; summary[3] = S.successful; summary[4] = S.unsuccessful; summary[5] = S.termination;
}

void pvbh_associate(const float* tgt, int n, const double* R_ref, const double* t_ref, const float* qry, int m, const double* R_nei, const double* t_nei,
                    double h, float thr, double plane_tol, int K, unsigned char* valid, double* p_local, double* plane, int* nn_idx, float* nn_d2) {
  if (K == 10) associate_all<10>(tgt, n, R_ref, t_ref, qry, m, R_nei, t_nei, h, thr, plane_tol, valid, p_local, plane, nn_idx, nn_d2);
  else associate_all<5>(tgt, n, R_ref, t_ref, qry, m, R_nei, t_nei, h, thr, plane_tol, valid, p_local, plane, nn_idx, nn_d2);
}

void pvbh_transform_cloud(const double* R, const double* t, const float* in, int n, float* out) {
  for (int i = 0; i < n; ++i) { transform_point_f32(R, t, in[i * 4], in[i * 4 + 1], in[i * 4 + 2], out[i * 4], out[i * 4 + 1], out[i * 4 + 2]); out[i * 4 + 3] = in[i * 4 + 3]; }
}
// the per-point body of k_undistort on the host: frame prepared exactly like pvb_undistort_clouds does
void pvbh_undistort_cloud(const double* R_wl, const double* t_wl, const double* R_we, const double* t_we, const float* in, long n, float* out) {
  double R_se[9], t_se[3], q[4];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { double acc = 0.0; for (int k = 0; k < 3; ++k) acc += R_wl[k * 3 + r] * R_we[k * 3 + c]; R_se[r * 3 + c] = acc; }
    double acc = 0.0; for (int k = 0; k < 3; ++k) acc += R_wl[k * 3 + r] * (t_we[k] - t_wl[k]); t_se[r] = acc;
  }
  quat_from_matrix_eigen(R_se, q);
  UndistortPrep u; undistort_prepare(q, t_se, u);
  for (long i = 0; i < n; ++i) { undistort_point_f32(u, i, n, in[i * 4], in[i * 4 + 1], in[i * 4 + 2], out[i * 4], out[i * 4 + 1], out[i * 4 + 2]); out[i * 4 + 3] = in[i * 4 + 3]; }
}
void pvbh_fast_atan2_f(long n, const float* y, const float* x, float* out) { for (long i = 0; i < n; ++i) out[i] = fast_atan2_f32(y[i], x[i]); }
void pvbh_fast_atan2_d(long n, const double* y, const double* x, double* out) { for (long i = 0; i < n; ++i) out[i] = fast_atan2_f64(y[i], x[i]); }
void pvbh_cam_to_image_f(int rows, int cols, long n, const float* cam, float* px) { for (long i = 0; i < n; ++i) cam_to_image_f32(cam[3 * i], cam[3 * i + 1], cam[3 * i + 2], rows, cols, px[2 * i], px[2 * i + 1]); }
void pvbh_image_to_cam_d(int rows, int cols, long n, const double* px, double* cam) { for (long i = 0; i < n; ++i) image_to_cam_f64(px[2 * i], px[2 * i + 1], rows, cols, cam + 3 * i); }

}  // extern "C"
