// closures.cu -- dispatch of problem.cost / .grad / .hess (ManiSDP_unitdiag.m:41-43) to the per-driver kernels
#include "affine.h"
#include "kernels.cuh"

int msdp_costgrad(manisdp_handle* h, int which, int cg_mode) {
  if (h->kind == MANISDP_ONLYUNITDIAG) return msdp_maxcut_costgrad(h, which, cg_mode);
  return msdp_affine_costgrad(h, which, cg_mode);
}

int msdp_hess_dir(manisdp_handle* h, const double* D, double* Hout, int tail_mode) {
  if (h->kind == MANISDP_ONLYUNITDIAG) {
    const double* gather = (h->world > 1) ? h->gatherbuf : D;
    return msdp_maxcut_hess(h, gather, D, Hout, tail_mode != TAIL_NONE, tail_mode);
  }
  return msdp_affine_hess(h, D, Hout, tail_mode);
}
