// affine.h -- closures of the three affine drivers (ManiSDP_unitdiag.m:152-171, ManiSDP_unittrace.m:156-177,
// ManiSDP.m:149-165) on f(Y) = <C, YY'> + sigma/2 |A(YY') - b - y/sigma|^2
#pragma once
#include <vector>
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
// helpers shared with dual.cu
int msdp_affine_rowdot(manisdp_handle* h, const double* Y, const double* T, double* zout, int slot);
int msdp_affine_touch(manisdp_handle* h, const double* vec, double coef, const double* base, double* dst);
// dual.cu -- Riemannian ADMM on the SOS form (src/dual/ManiDSDP_unitdiag.m): additions to the closures above
int msdp_dual_setup(manisdp_handle* h, const manisdp_problem* pb, const std::vector<double>& dAAt);
void msdp_dual_free(manisdp_handle* h);
int msdp_dual_refresh(manisdp_handle* h);  // C_eff = bA + x - sigma*c, the cost constant k0, eS <- C_eff
int msdp_dual_cost_extra(manisdp_handle* h, const double* Z, const double* resid, int w);
int msdp_dual_grad_extra(manisdp_handle* h, const double* Z, double* G, int w, const int* pred);
int msdp_dual_hess_extra(manisdp_handle* h, const double* Y, const double* D, double* Hout, int w, int skip);
int msdp_dual_kkt(manisdp_handle* h, int update, manisdp_kkt_info* out);
