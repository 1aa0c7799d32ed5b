// gemm.h -- FP64 DMMA GEMM entry points (gemm_f64.cu)
#pragma once
#include "common.cuh"
// out(n x w, ld = ldo) = alpha * S(n x n) * V(n x w, ld = ldv) + beta * out ; skipped when pred && (*pred == 0) [pred_sense 0] / (*pred != 0) [pred_sense 1]
int msdp_gemm_nn(manisdp_handle* h, const double* S, int n, const double* V, int ldv, int w, double* out, int ldo,
                 double alpha, double beta, const int* pred, int pred_sense = 0);
// M(n x n) = alpha * P(n x w, ld = ldp) * Q(n x w, ld = ldq)'
int msdp_gemm_nt(manisdp_handle* h, const double* P, int ldp, const double* Q, int ldq, int n, int w, double* M,
                 double alpha, const int* pred, int pred_sense = 0);
// G(ka x kb) = P(nrows x ka)' * Q(nrows x kb)  (split-K over the rows, deterministic slice sum)
int msdp_gemm_tn(manisdp_handle* h, const double* P, int ldp, int ka, const double* Q, int ldq, int kb, int64_t nrows,
                 double* G);
// out(nrows x w) = alpha * P(nrows x k) * Cm(k x w, row stride ldc) + beta * out
int msdp_gemm_rows_small(manisdp_handle* h, const double* P, int ldp, int k, const double* Cm, int ldc, int w,
                         int64_t nrows, double* out, int ldo, double alpha, double beta);
