// affine.cu -- (stub until the affine closures land)
#include "affine.h"
int msdp_affine_setup(manisdp_handle* h, const manisdp_problem*) { return msdp_fail(h, MANISDP_E_ARG, "affine kinds not built yet"); }
void msdp_affine_free(manisdp_handle*) {}
int msdp_affine_costgrad(manisdp_handle* h, int, int) { return msdp_fail(h, MANISDP_E_ARG, "affine kinds not built yet"); }
int msdp_affine_hess(manisdp_handle* h, const double*, double*, int) { return msdp_fail(h, MANISDP_E_ARG, "affine kinds not built yet"); }
int msdp_affine_cost_only(manisdp_handle* h, const double*, double*) { return msdp_fail(h, MANISDP_E_ARG, "affine kinds not built yet"); }
int msdp_affine_kkt(manisdp_handle* h, int, manisdp_kkt_info*) { return msdp_fail(h, MANISDP_E_ARG, "affine kinds not built yet"); }
int msdp_affine_apply_S(manisdp_handle* h, const double*, double*, int) { return msdp_fail(h, MANISDP_E_ARG, "affine kinds not built yet"); }
