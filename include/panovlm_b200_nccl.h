/* panovlm_b200 — the exchange step of a sharded pose graph over NCCL, for C / C++ hosts (libpanovlm_b200_nccl.so).
 *
 * SURVEY.md 8e / BASELINE configs[3]: the pose graph's edges are sharded by reference frame, one context per GPU; every rank registers
 * the GLOBAL edge list (pvb_blocks_set_edge_list) and its own residual blocks, and each evaluation ends with ONE sum-allreduce of the
 * per-edge normal equations (n_edges x 92 doubles) so that all ranks continue with the complete system and pvb_blocks_solve_lm takes
 * identical steps everywhere.  libpanovlm_b200.so itself only knows a hook (pvb_blocks_set_reduce_hook); this small library is the hook
 * for NCCL: ncclAllReduce(ncclDouble, ncclSum) on the context's stream.  The reference has no distributed code to mirror; the host stays
 * C++ (one thread per GPU with ncclCommInitAll, or one process per GPU with ncclCommInitRank - both work, the communicator is the caller's).
 */
#ifndef PANOVLM_B200_NCCL_H
#define PANOVLM_B200_NCCL_H
#include "panovlm_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* nccl_comm: an ncclComm_t of the rank that owns ctx's device.  After attach every pvb_blocks_evaluate(want_system = 1) and every evaluation
 * inside pvb_blocks_solve_lm allreduces its edge systems in place.                                                                        */
int pvb_nccl_attach(pvb_ctx* ctx, void* nccl_comm);
int pvb_nccl_detach(pvb_ctx* ctx);
/* the dense sweep's exchange (configs[4]): in-place sum-allreduce of n_doubles doubles on ctx's stream (e.g. the packed 6x6 / 6x1 blocks
 * of all frames, every rank having written only its own frames' rows into a zeroed buffer)                                                */
int pvb_nccl_allreduce(pvb_ctx* ctx, void* nccl_comm, double* device_buffer, long n_doubles);
/* ncclResult_t of the last NCCL call made by this library on the calling thread (0 = ncclSuccess)                                         */
int pvb_nccl_last_result(void);
#ifdef __cplusplus
}
#endif
#endif
