// kernels.cuh -- host-side launchers shared between translation units
#pragma once
#include "common.cuh"

struct VecPtrs {
  double *Y0, *Y1, *G0, *G1, *eG0, *eG1, *eta0, *eta1, *r, *d, *Hd;
};
VecPtrs msdp_vecptrs(const manisdp_handle* h);

// tcg.cu
int msdp_launch_tcg_init(manisdp_handle* h);
int msdp_launch_tcg_update(manisdp_handle* h, cudaGraphConditionalHandle cond, int use_cond, int defer);
int msdp_launch_tcg_after_update_scalar(manisdp_handle* h, cudaGraphConditionalHandle cond = 0, int use_cond = 0);
int msdp_launch_tcg_dir(manisdp_handle* h);
int msdp_launch_retract(manisdp_handle* h, const double* Y, const double* eta, double* dst, int from_state);
int msdp_launch_project(manisdp_handle* h, const double* Y, const double* src, double* dst);
int msdp_launch_dot(manisdp_handle* h, const double* a, const double* b, int slot);

// tail modes of the Hessian / cost kernels
enum {
  TAIL_NONE = 0,      // plain closure call (manisdp_hess): no scalar side effects
  TAIL_TCG = 1,       // inside tCG: skip when st->stop != 0, reduce <mdelta, Hmdelta>, run tcg_after_hv
  TAIL_TCG_DEFER = 2  // as TAIL_TCG but leave the local sum in st->tmp[0] (row-sharded: all-reduce follows)
};
enum {
  CG_PLAIN = 0,   // cost + gradient at Ybuf[which]: fill caches, st->tmp[0] = f, st->tmp[1] = |grad|^2
  CG_INIT = 1,    // as PLAIN and also st->fx, st->gradnorm2 (trustregions.m:405)
  CG_TR = 2,      // proposal inside the TR loop: st->fprop, st->gradnorm2_prop, then tr_decide
  CG_TR_DEFER = 3, // row-sharded variant of CG_TR: local sums in st->tmp[0..1]
  CG_COSTONLY = 4  // cost only (line search): st->tmp[0] = f, no cache / gradient side effects for the affine kinds
};

// spmm.cu (ONLYUNITDIAG closures, ManiSDP_onlyunitdiag.m:117-130)
int msdp_maxcut_hess(manisdp_handle* h, const double* Dgather, const double* Down, double* Hout, int from_state,
                     int tail_mode);
int msdp_maxcut_costgrad(manisdp_handle* h, int which_or_neg, int cg_mode);
// row-sharded: exchange of the operand overlapped with one column pass per owner chunk
bool msdp_pipeline_ok(const manisdp_handle* h);
int msdp_costgrad_exchange(manisdp_handle* h, int buf, int which, int cg_mode);
bool msdp_peer_gather_ok(const manisdp_handle* h);  // products gather remote rows in place: a stopped iteration costs ~nothing
int msdp_maxcut_hess_pipelined(manisdp_handle* h, const double* Down, double* Hout, int from_state, int tail_mode);
int msdp_spmm_shift(manisdp_handle* h, const double* Vgather, const double* Vown, double* out, int k_ld,
                    const double* zdiag);
int msdp_launch_tr_decide_scalar(manisdp_handle* h);
int msdp_launch_tcg_after_hv_scalar(manisdp_handle* h);

int msdp_spmm_shift_tcg(manisdp_handle* h, const double* V, double* out, int ld, int in_tcg);
// colshard.cu (column-sharded handle: every rank holds all rows and pl = ceil(p / G) columns)
int msdp_col_init(manisdp_handle* h, const void* unique_id, int world, int rank);
void msdp_col_destroy(manisdp_handle* h);
int msdp_col_split(manisdp_handle* h);
int msdp_col_merge(manisdp_handle* h);
int msdp_col_hess(manisdp_handle* h, const double* D, double* Hout, int from_state, int tail_mode);
int msdp_col_costgrad(manisdp_handle* h, int buf);  // leaves the all-reduced (sum eG, |G|^2) in st->tmp[0..1]
int msdp_col_retract(manisdp_handle* h);
int msdp_col_tcg_dir(manisdp_handle* h);
int msdp_col_tcg_update_dir(manisdp_handle* h, cudaGraphConditionalHandle cond, int use_cond);
int msdp_col_allreduce_tmp(manisdp_handle* h, int count);
int msdp_col_cg_scalar(manisdp_handle* h, int cg_mode);  // scalar tail of a cost+grad call (CG_* modes)
