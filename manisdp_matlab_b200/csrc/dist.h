// dist.h -- row sharding across GPUs: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// Rank r owns rows [r*rpr, min(n, (r+1)*rpr)) with rpr = ceil(n / world); every local n x ld array is allocated with
// rpr rows (zero padded) so the all-gather of the thin factor uses equal counts and a gathered row sits at its
// global index.
#pragma once
#include "common.cuh"

int msdp_dist_init(manisdp_handle* h, const void* unique_id);
void msdp_dist_destroy(manisdp_handle* h);
// full[world*rpr x ld] <- all-gather of local[rpr x ld]  (the exchange step before every Hessian product)
int msdp_dist_allgather_rows(manisdp_handle* h, const double* local, double* full);
// st->tmp[0..count) summed over ranks, in place on the device (the scalar packet of one tCG step)
int msdp_dist_allreduce_tmp(manisdp_handle* h, int count);
int msdp_dist_allreduce_buf(manisdp_handle* h, double* buf, int64_t count);
// fx = tmp[0]/2, gradnorm2 = tmp[1] after the all-reduce of the initial cost+grad
int msdp_dist_finish_init(manisdp_handle* h);
inline int64_t msdp_rows_per_rank(int64_t n, int world) { return world > 1 ? (n + world - 1) / world : n; }
// collective, after the work arrays were (re)allocated: export / map the exchange sources (pipeline mode 2)
int msdp_dist_ipc_refresh(manisdp_handle* h);
// collective, BEFORE exported arrays are freed: unmap all peer memory and wait for every rank to have done so
int msdp_dist_ipc_release(manisdp_handle* h);
// device table of G operand pointers (peers' arrays mapped through CUDA IPC, own entry local) for `local` = the
// direction array or SLOT_U; NULL when the mappings are not available
const double* const* msdp_dist_peer_table(manisdp_handle* h, const double* local);
// one-double all-reduce on the main stream: every rank has finished writing its operand
int msdp_dist_barrier(manisdp_handle* h);
// staged all-gather on the exchange stream; ev_stage[s] fires when the chunk of rank (r - s) mod G is in `full`
int msdp_dist_exchange_begin(manisdp_handle* h, const double* local, double* full);
// dst[world*count] <- all-gather of src[count] (eigen step blocks)
int msdp_dist_allgather_block(manisdp_handle* h, const double* src, double* dst, int64_t count);
