// scalar_logic.cuh -- the scalar control flow of tCG / trustregions, executed on the device by the last block of a
// kernel (or by a one-thread kernel after an NCCL all-reduce on row-sharded handles).
// Follows manopt7.0/manopt/solvers/trustregions/tCG.m:160-287 and trustregions.m:548-726.
#pragma once
#include <float.h>
#include "common.cuh"

__device__ __forceinline__ void set_cond(cudaGraphConditionalHandle cond, int use_cond, unsigned v) {
  if (use_cond) cudaGraphSetConditional(cond, v);
}

// state at tCG entry (tCG.m:103-152): eta = 0, r = z = mdelta = grad
__device__ __forceinline__ void tcg_reset(RtrState* st) {
  const double rr = st->gradnorm2;
  st->r_r = rr;
  st->z_r = rr;
  st->d_Pd = rr;
  st->e_Pd = 0.0;
  st->e_Pe = 0.0;
  st->e_Pe_new = 0.0;
  st->model_value = 0.0;
  st->norm_r0 = sqrt(rr);
  st->eta_g = 0.0;
  st->eta_Heta = 0.0;
  st->norm_eta = 0.0;
  st->y_r = 0.0;
  st->y_d = 0.0;
  st->alpha = st->beta = st->tau = 0.0;
  st->j = 0;
  st->stop = 0;
  st->branch = 0;
  st->eta_cur = 0;
}

// after Hmdelta = Hess[mdelta] and d_Hd = <mdelta, Hmdelta>  (tCG.m:163-211)
__device__ __forceinline__ void tcg_after_hv(RtrState* st, double d_Hd) {
  st->j += 1;
  st->hv_count += 1;
  st->d_Hd = d_Hd;
  const double alpha = st->z_r / d_Hd;  // :170
  const double e_Pe_new = st->e_Pe + 2.0 * alpha * st->e_Pd + alpha * alpha * st->d_Pd;  // :173
  const double D2 = st->Delta * st->Delta;
  st->alpha = alpha;
  st->e_Pe_new = e_Pe_new;
  if (d_Hd <= 0.0 || e_Pe_new >= D2) {  // :183
    st->tau = (-st->e_Pd + sqrt(st->e_Pd * st->e_Pd + st->d_Pd * (D2 - st->e_Pe))) / st->d_Pd;  // :188
    st->branch = (d_Hd <= 0.0) ? 1 : 2;  // :205-209
  } else {
    st->branch = 0;
  }
}

// after the update pass.  q: branch != 0 -> {<eta',g>, <eta',Heta'>, <eta',eta'>}
//                            branch == 0 -> {<eta',g>, <eta',r'-g>, <r',r'>, <eta',eta'>, <Y,r'>, <Y,mdelta>}
__device__ __forceinline__ void tcg_after_update(RtrState* st, const double* q, cudaGraphConditionalHandle cond,
                                                 int use_cond) {
  if (st->branch != 0) {  // tCG.m:188-211 : eta - tau*mdelta leaves through the boundary / negative curvature
    st->eta_g = q[0];
    st->eta_Heta = q[1];
    st->norm_eta = sqrt(q[2]);
    st->eta_cur ^= 1;
    st->stop = st->branch;
    set_cond(cond, use_cond, 0u);
    return;
  }
  const double new_model = q[0] + 0.5 * q[1];  // :227 (model_fun :150)
  if (!(new_model < st->model_value)) {        // :228  (new_model >= model_value; NaN also stops)
    st->stop = 6;
    set_cond(cond, use_cond, 0u);
    return;
  }
  st->eta_cur ^= 1;  // commit new_eta (:233-235)
  st->e_Pe = st->e_Pe_new;
  st->model_value = new_model;
  st->eta_g = q[0];
  st->eta_Heta = q[1];
  st->norm_eta = sqrt(q[3]);
  st->r_r = q[2];  // :241
  const double norm_r = sqrt(q[2]);
  const double nr0 = st->norm_r0;
  const double nr0t = pow(nr0, st->theta);
  if (st->j >= st->mininner && norm_r <= nr0 * fmin(nr0t, st->kappa)) {  // :249
    st->stop = (st->kappa < nr0t) ? 3 : 4;                               // :251-255
    set_cond(cond, use_cond, 0u);
    return;
  }
  if (st->j >= st->maxinner) {  // for-loop exhausted (:160), stop_tCG keeps its initial value 5
    st->stop = 5;
    set_cond(cond, use_cond, 0u);
    return;
  }
  const double zold = st->z_r;  // :267
  st->z_r = q[2];               // :269 (identity preconditioner: z = r)
  const double beta = st->z_r / zold;  // :272
  st->beta = beta;
  st->e_Pd = beta * (st->e_Pd + st->alpha * st->d_Pd);  // :286
  st->d_Pd = st->z_r + beta * beta * st->d_Pd;          // :287
  st->y_r = q[4];
  st->y_d = q[5];
}

// after the cost (+ gradient) at the proposal is known: rho, radius update, accept / reject (trustregions.m:548-726)
__device__ __forceinline__ void tr_decide(RtrState* st) {
  const double fx = st->fx, fp = st->fprop;
  double rhonum = fx - fp;                               // :548
  double rhoden = -(st->eta_g + 0.5 * st->eta_Heta);     // :549-550
  const double reg = fmax(1.0, fabs(fx)) * DBL_EPSILON * st->rho_regularization;  // :579
  rhonum += reg;
  rhoden += reg;
  const bool model_decreased = (rhoden >= 0.0);  // :613
  const double rho = rhonum / rhoden;            // :621
  st->rhonum = rhonum;
  st->rhoden = rhoden;
  st->rho = rho;
  if (rho < 0.25 || !model_decreased || isnan(rho)) {  // :653
    st->Delta = st->Delta / 4.0;
  } else if (rho > 0.75 && (st->stop == 1 || st->stop == 2)) {  // :667
    st->Delta = fmin(2.0 * st->Delta, st->Delta_bar);
  }
  const bool acc = model_decreased && (rho > st->rho_prime);  // :688
  if (acc) {
    st->pt ^= 1;
    st->fx = fp;
    st->gradnorm2 = st->gradnorm2_prop;
  }
  st->accepted = acc ? 1 : 0;
  st->tr_iter += 1;
}
