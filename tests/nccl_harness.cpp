// TEST INFRASTRUCTURE.  A C++ multi-GPU caller of the library the way a C++ host (the reference's language) would write it: one process, one thread and one
// pvb_ctx per GPU, ncclCommInitAll, the pose graph's edges sharded by reference frame, the GLOBAL edge list registered on every rank, the NCCL hook of
// libpanovlm_b200_nccl.so as the single exchange step of an evaluation, and pvb_blocks_solve_lm run by every rank.  Returns every rank's poses and summary so the
// test can check that they are identical to each other and agree with a single-GPU solve.  Built by the test with g++ against the two libraries and libnccl.
#include <nccl.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>
#include <set>
#include <thread>
#include <utility>
#include <vector>
#include "panovlm_b200_nccl.h"

extern "C" int nccl_pose_graph_run(int n_gpus, long n, const int* type, const int* ref, const int* nei, const int* normalize, const double* huber, const double* consts, int nb,
                                   const double* poses_in, const unsigned char* is_const, int max_iter, double* poses_out /* n_gpus x nb x 6 */, double* summaries /* n_gpus x 6 */,
                                   int* rank_blocks /* n_gpus */) {
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < n_gpus) return -1;
  std::vector<int> devs(n_gpus);
  for (int r = 0; r < n_gpus; ++r) devs[r] = r;
  std::vector<ncclComm_t> comms(n_gpus);
  if (ncclCommInitAll(comms.data(), n_gpus, devs.data()) != ncclSuccess) return -2;
  // global edge list: sorted, unique (ref, nei)
  std::set<std::pair<int, int>> es;
  for (long i = 0; i < n; ++i) es.insert({ref[i], nei[i]});
  std::vector<int> er, en;
  for (auto& e : es) { er.push_back(e.first); en.push_back(e.second); }
  std::vector<int> rc(n_gpus, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < n_gpus; ++r) {
    th.emplace_back([&, r]() {
      pvb_ctx* ctx = nullptr;
      if (pvb_create(r, &ctx) != PVB_OK) { rc[r] = -3; return; }
      // this rank's shard: blocks whose reference frame lies in [lo, hi)
      const int lo = nb * r / n_gpus, hi = nb * (r + 1) / n_gpus;
      std::vector<int> t, rf, ne, nz; std::vector<double> hb, cs;
      for (long i = 0; i < n; ++i)
        if (ref[i] >= lo && ref[i] < hi) { t.push_back(type[i]); rf.push_back(ref[i]); ne.push_back(nei[i]); nz.push_back(normalize[i]); hb.push_back(huber[i]); cs.insert(cs.end(), consts + 12 * i, consts + 12 * i + 12); }
      rank_blocks[r] = (int)t.size();
      int e = pvb_blocks_set_edge_list(ctx, (int)er.size(), er.data(), en.data());
      if (e == PVB_OK) e = pvb_nccl_attach(ctx, comms[r]);
      if (e == PVB_OK) e = pvb_blocks_set(ctx, (long)t.size(), t.data(), rf.data(), ne.data(), nz.data(), hb.data(), cs.data(), nb);
      std::vector<double> p(poses_in, poses_in + 6 * (size_t)nb);
      if (e == PVB_OK) e = pvb_blocks_solve_lm(ctx, p.data(), is_const, max_iter, summaries + 6 * r);
      std::memcpy(poses_out + 6 * (size_t)nb * r, p.data(), sizeof(double) * 6 * nb);
      rc[r] = e;
      pvb_nccl_detach(ctx);
      pvb_destroy(ctx);
    });
  }
  for (auto& x : th) x.join();
  for (int r = 0; r < n_gpus; ++r) ncclCommDestroy(comms[r]);
  for (int r = 0; r < n_gpus; ++r) if (rc[r]) return rc[r] > 0 ? -100 - rc[r] : rc[r] - 10;
  return 0;
}
