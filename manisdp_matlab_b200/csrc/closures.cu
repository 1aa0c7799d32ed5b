// closures.cu -- dispatch of problem.cost / .grad / .hess (ManiSDP_unitdiag.m:41-43) to the per-driver kernels
#include "affine.h"
#include "kernels.cuh"

int msdp_costgrad(manisdp_handle* h, int which, int cg_mode) {
  if (h->col_split) {
    // column-sharded: the host mirror of pt is exact wherever a cost is requested (between TR iterations / inside one
    // after its retraction), so the device-selected buffers (-1 proposal, -2 current) are known here
    const int buf = which >= 0 ? which : (which == -1 ? (h->pt ^ 1) : h->pt);
    MSDP_TRY(msdp_col_costgrad(h, buf));
    return msdp_col_cg_scalar(h, cg_mode);
  }
  if (h->kind == MANISDP_ONLYUNITDIAG) return msdp_maxcut_costgrad(h, which, cg_mode);
  return msdp_affine_costgrad(h, which, cg_mode);
}

int msdp_hess_dir(manisdp_handle* h, const double* D, double* Hout, int tail_mode) {
  if (h->col_split) return msdp_col_hess(h, D, Hout, tail_mode != TAIL_NONE, tail_mode);
  if (h->kind == MANISDP_ONLYUNITDIAG) {
    const double* gather = (h->world > 1) ? h->gatherbuf : D;
    return msdp_maxcut_hess(h, gather, D, Hout, tail_mode != TAIL_NONE, tail_mode);
  }
  return msdp_affine_hess(h, D, Hout, tail_mode);
}
