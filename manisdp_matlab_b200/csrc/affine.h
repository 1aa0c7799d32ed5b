// affine.h -- closures of the three affine drivers (ManiSDP_unitdiag.m:152-171, ManiSDP_unittrace.m:156-177,
// ManiSDP.m:149-165) on f(Y) = <C, YY'> + sigma/2 |A(YY') - b - y/sigma|^2
#pragma once
#include "common.cuh"

int msdp_affine_setup(manisdp_handle* h, const manisdp_problem* pb);
void msdp_affine_free(manisdp_handle* h);
int msdp_affine_costgrad(manisdp_handle* h, int which, int cg_mode);
int msdp_affine_hess(manisdp_handle* h, const double* D, double* Hout, int tail_mode);
int msdp_affine_cost_only(manisdp_handle* h, const double* Z, double* f_host);
// KKT residues of the affine drivers + dual update + dual slack operator set-up (ManiSDP_unitdiag.m:59-71)
int msdp_affine_kkt(manisdp_handle* h, int update_dual, manisdp_kkt_info* out);
// AV = S * V on an n x kld block (S as prepared by the last msdp_affine_kkt)
int msdp_affine_apply_S(manisdp_handle* h, const double* V, double* AV, int kld);
// (i, j) of every entry of At as the kernels index it (CSC order); count = nnz(At)
int msdp_affine_index_split(manisdp_handle* h, int64_t* i_out, int64_t* j_out, int64_t cap, int64_t* count);
