// panovlm_b200 — NCCL exchange hook for C / C++ hosts (include/panovlm_b200_nccl.h).  Separate shared library: libpanovlm_b200.so has no NCCL dependency.
#include <cuda_runtime.h>
#include <nccl.h>
#include "../../include/panovlm_b200_nccl.h"

namespace {
thread_local int g_last = 0;
void allreduce_hook(void* user, double* device_edge_systems, long n_doubles, void* cuda_stream) {
  g_last = (int)ncclAllReduce(device_edge_systems, device_edge_systems, (size_t)n_doubles, ncclDouble, ncclSum, (ncclComm_t)user, (cudaStream_t)cuda_stream);
}
}  // namespace

extern "C" {
int pvb_nccl_attach(pvb_ctx* ctx, void* nccl_comm) {
  if (!ctx || !nccl_comm) return PVB_ERR_ARG;
  return pvb_blocks_set_reduce_hook(ctx, allreduce_hook, nccl_comm);
}
int pvb_nccl_detach(pvb_ctx* ctx) { return ctx ? pvb_blocks_set_reduce_hook(ctx, nullptr, nullptr) : PVB_ERR_ARG; }
int pvb_nccl_allreduce(pvb_ctx* ctx, void* nccl_comm, double* device_buffer, long n_doubles) {
  if (!ctx || !nccl_comm || !device_buffer || n_doubles < 0) return PVB_ERR_ARG;
  g_last = (int)ncclAllReduce(device_buffer, device_buffer, (size_t)n_doubles, ncclDouble, ncclSum, (ncclComm_t)nccl_comm, (cudaStream_t)pvb_stream(ctx));
  return g_last == 0 ? PVB_OK : PVB_ERR_CUDA;
}
int pvb_nccl_last_result(void) { return g_last; }
}
