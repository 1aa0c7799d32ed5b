// gemm.h -- FP64 DMMA GEMM entry points (gemm_f64.cu)
#pragma once
#include "common.cuh"
// out(n x w, ld = ldo) = alpha * S(n x n) * V(n x w, ld = ldv) + beta * out ; skipped when pred && (*pred == 0) [pred_sense 0] / (*pred != 0) [pred_sense 1]
int msdp_gemm_nn(manisdp_handle* h, const double* S, int n, const double* V, int ldv, int w, double* out, int ldo,
                 double alpha, double beta, const int* pred, int pred_sense = 0);
// M(n x n) = alpha * P(n x w, ld = ldp) * Q(n x w, ld = ldq)'
int msdp_gemm_nt(manisdp_handle* h, const double* P, int ldp, const double* Q, int ldq, int n, int w, double* M,
                 double alpha, const int* pred, int pred_sense = 0);
