// rtr.cu -- device-resident Riemannian trust-region driver.
// Replaces trustregions(problem, Y, opts) (manopt7.0/manopt/solvers/trustregions/trustregions.m:395-767) + tCG.m and
// the Manopt dispatch layer (getCost / getGradient / getHessian / StoreDB): the engine keeps exactly
// two point slots (current, proposal) and all scalars in one device struct, so per TR iteration the host launches one
// CUDA graph -- tcg_init -> WHILE{Hv, update, direction} -> retract -> cost+grad+accept -- and reads back one record.
#include <math.h>
#include <string.h>
#include "kernels.cuh"
#include "dist.h"

struct TrSetup {
  double Delta, Delta_bar, rho_prime, rho_regularization, kappa, theta;
  int maxinner, mininner;
};

__global__ void k_tr_setup(RtrState* st, TrSetup s) {
  st->Delta = s.Delta;
  st->Delta_bar = s.Delta_bar;
  st->rho_prime = s.rho_prime;
  st->rho_regularization = s.rho_regularization;
  st->kappa = s.kappa;
  st->theta = s.theta;
  st->maxinner = s.maxinner;
  st->mininner = s.mininner;
  st->tr_iter = 0;
  st->hv_count = 0;
  st->accepted = 1;
  st->stop = 0;
  st->j = 0;
  st->ticket = 0u;
  st->rho = 0.0;
  st->norm_eta = 0.0;
}

int msdp_sync_state(manisdp_handle* h) {
  CUDA_TRY(h, cudaMemcpyAsync(h->st_host, h->st, sizeof(RtrState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->pt = h->st_host->pt;
  return MANISDP_OK;
}

// one Hessian product of the tCG loop: Hd = Hess[d], tail -> tcg_after_hv
static int tcg_hv(manisdp_handle* h) {
  if (h->world > 1) {
    if (msdp_pipeline_ok(h)) {
      MSDP_TRY(msdp_maxcut_hess_pipelined(h, h->d, h->Hd, 1, TAIL_TCG_DEFER));
    } else {
      MSDP_TRY(msdp_dist_allgather_rows(h, h->d, h->gatherbuf));
      MSDP_TRY(msdp_hess_dir(h, h->d, h->Hd, TAIL_TCG_DEFER));
    }
    MSDP_TRY(msdp_dist_allreduce_tmp(h, 1));
    return msdp_launch_tcg_after_hv_scalar(h);
  }
  return msdp_hess_dir(h, h->d, h->Hd, TAIL_TCG);
}

static int tcg_iteration(manisdp_handle* h, cudaGraphConditionalHandle cond, int use_cond) {
  MSDP_TRY(tcg_hv(h));
  if (h->col_split) return msdp_col_tcg_update_dir(h, cond, use_cond);
  if (h->world > 1) {
    MSDP_TRY(msdp_launch_tcg_update(h, cond, 0, 1));
    MSDP_TRY(msdp_dist_allreduce_tmp(h, 6));
    MSDP_TRY(msdp_launch_tcg_after_update_scalar(h));
  } else {
    MSDP_TRY(msdp_launch_tcg_update(h, cond, use_cond, 0));
  }
  return msdp_launch_tcg_dir(h);
}

static int tr_tail(manisdp_handle* h) {
  if (h->col_split) {
    MSDP_TRY(msdp_col_retract(h));
    return msdp_costgrad(h, -1, CG_TR);
  }
  MSDP_TRY(msdp_launch_retract(h, nullptr, nullptr, nullptr, 1));
  if (h->world > 1) {
    // the proposal lives in Ybuf[pt^1]; pt on the host mirrors the device between iterations
    MSDP_TRY(msdp_costgrad_exchange(h, h->pt ^ 1, -1, CG_TR_DEFER));
    MSDP_TRY(msdp_dist_allreduce_tmp(h, 2));
    return msdp_launch_tr_decide_scalar(h);
  }
  return msdp_costgrad(h, -1, CG_TR);
}

static void destroy_graph(manisdp_handle* h) {
  for (int i = 0; i < 2; ++i) {
    if (h->tcg_exec[i]) cudaGraphExecDestroy(h->tcg_exec[i]);
    if (h->tcg_graph[i]) cudaGraphDestroy(h->tcg_graph[i]);
    h->tcg_exec[i] = nullptr;
    h->tcg_graph[i] = nullptr;
  }
  h->tcg_graph_p = -1;
}
void msdp_invalidate_graph(manisdp_handle* h) { destroy_graph(h); }

// Build   tcg_init -> WHILE(cond){ Hv ; update ; direction } -> retract -> cost+grad+accept   as one graph.
// The affine closures resolve their point buffers on the host, so one graph is built per value of pt (which of the two
// point buffers is current); the MaxCut kernels select on the device and would work with either.
static int build_tr_graph(manisdp_handle* h) {
  const int gi = h->pt;
  const int64_t launches0 = h->launches;
  cudaGraph_t g = nullptr;
  CUDA_TRY(h, cudaGraphCreate(&g, 0));
  h->tcg_graph[gi] = g;
  cudaStream_t s = h->stream;
  // segment A
  CUDA_TRY(h, cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  int rc = msdp_launch_tcg_init(h);
  int64_t l_fixed = h->launches - launches0;
  cudaStreamCaptureStatus cst;
  const cudaGraphNode_t* deps = nullptr;
  size_t ndeps = 0;
  cudaGraph_t gtmp = nullptr;
  cudaError_t e = cudaStreamGetCaptureInfo_v2(s, &cst, nullptr, &gtmp, &deps, &ndeps);
  std::vector<cudaGraphNode_t> depv(deps, deps + (e == cudaSuccess ? ndeps : 0));
  cudaGraph_t dummy = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(s, &dummy);
  if (rc != MANISDP_OK) return rc;
  CUDA_TRY(h, e);
  CUDA_TRY(h, e2);
  // WHILE node
  cudaGraphConditionalHandle cond;
  CUDA_TRY(h, cudaGraphConditionalHandleCreate(&cond, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = cond;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t wnode;
  CUDA_TRY(h, cudaGraphAddNode(&wnode, g, depv.data(), depv.size(), &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  CUDA_TRY(h, cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  const int64_t lb0 = h->launches;
  rc = tcg_iteration(h, cond, 1);
  h->graph_l_body = h->launches - lb0;
  e2 = cudaStreamEndCapture(s, &dummy);
  if (rc != MANISDP_OK) return rc;
  CUDA_TRY(h, e2);
  // segment C
  CUDA_TRY(h, cudaStreamBeginCaptureToGraph(s, g, &wnode, nullptr, 1, cudaStreamCaptureModeRelaxed));
  const int64_t lc0 = h->launches;
  rc = tr_tail(h);
  l_fixed += h->launches - lc0;
  h->graph_l_fixed = l_fixed;
  e2 = cudaStreamEndCapture(s, &dummy);
  if (rc != MANISDP_OK) return rc;
  CUDA_TRY(h, e2);
  CUDA_TRY(h, cudaGraphInstantiate(&h->tcg_exec[gi], g, 0));
  h->tcg_graph_p = h->p;
  h->launches = launches0;  // capture-time launches are not executions
  return MANISDP_OK;
}

int msdp_tr_solve(manisdp_handle* h, const manisdp_tr_options* o, manisdp_tr_info* info) {
  NvtxRange nvtx_range("manisdp:tr_solve");
  if (h->p <= 0) return msdp_fail(h, MANISDP_E_STATE, "tr_solve: set_Y / rand_Y first");
  manisdp_tr_options opt;
  memset(&opt, 0, sizeof(opt));
  if (o) opt = *o;
  // defaults: trustregions.m:340-372
  double typicaldist;
  if (h->kind == MANISDP_MULTIBLOCK)
    typicaldist = msdp_mb_typicaldist(h);  // multiblockmanifold.m:11-15
  else if (h->mf == MF_OBLIQUE)
    typicaldist = M_PI * sqrt((double)h->n);  // ManiSDP_unitdiag.m:179
  else if (h->mf == MF_SPHERE)
    typicaldist = M_PI;  // spherefactory.m:111
  else
    typicaldist = sqrt((double)h->n * (double)h->p);  // euclideanfactory.m:59
  if (opt.maxiter <= 0) opt.maxiter = 1000;
  if (opt.maxinner <= 0) {  // M.dim()
    double dim = (h->mf == MF_OBLIQUE) ? (double)(h->p - 1) * h->n
                                       : (h->mf == MF_SPHERE ? (double)h->n * h->p - 1 : (double)h->n * h->p);
    if (h->kind == MANISDP_MULTIBLOCK) dim = msdp_mb_dim(h);  // multiblockmanifold.m:3
    opt.maxinner = (int)fmin(dim, 2.0e9);
    if (opt.maxinner < 1) opt.maxinner = 1;
  }
  if (opt.mininner <= 0) opt.mininner = 1;
  if (opt.tolgradnorm <= 0) opt.tolgradnorm = 1e-6;
  if (opt.kappa <= 0) opt.kappa = 0.1;
  if (opt.theta <= 0) opt.theta = 1.0;
  if (opt.rho_prime <= 0) opt.rho_prime = 0.1;
  if (opt.rho_regularization <= 0) opt.rho_regularization = 1e3;
  if (opt.Delta_bar <= 0) opt.Delta_bar = typicaldist;
  if (opt.Delta0 <= 0) opt.Delta0 = opt.Delta_bar / 8.0;
  // column-split handles: the NCCL all-reduces of the loop are captured into the graph (and into the body of its WHILE
  // node) like the kernels around them -- opt-in for world > 1 (MANISDP_COL_GRAPH=1) until it has been measured
  const bool col_graph_ok = !h->col_split || h->cworld <= 1 || h->col_graph;
  const int use_graph = (opt.use_graph != 0) && (h->world <= 1) && col_graph_ok;
  h->y_version++;

  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  TrSetup s{opt.Delta0, opt.Delta_bar, opt.rho_prime, opt.rho_regularization, opt.kappa, opt.theta,
            opt.maxinner, opt.mininner};
  k_tr_setup<<<1, 1, 0, h->stream>>>(h->st, s);
  KERNEL_CHECK(h);
  // getCostGrad at the initial point (trustregions.m:405).  When the point has not changed since the last closure
  // call -- e.g. consecutive tr_solve calls, or tr_solve right after manisdp_cost -- the device already holds f, the
  // gradient and the per-point caches (the equivalent of Manopt's StoreDB hit, getCost.m:42-55), so nothing is redone.
  if (!(h->cache_valid && h->grad_valid)) {
    if (h->world > 1) {
      MSDP_TRY(msdp_costgrad_exchange(h, h->pt, -2, CG_TR_DEFER));
      MSDP_TRY(msdp_dist_allreduce_tmp(h, 2));
      MSDP_TRY(msdp_dist_finish_init(h));
    } else {
      MSDP_TRY(msdp_costgrad(h, -2, CG_INIT));
    }
  }
  MSDP_TRY(msdp_sync_state(h));
  h->cache_valid = 1;
  h->grad_valid = 1;
  h->log.clear();
  const RtrState* sh = h->st_host;
  manisdp_tr_iter rec;
  memset(&rec, 0, sizeof(rec));
  rec.cost = sh->fx;
  rec.gradnorm = sqrt(sh->gradnorm2);
  rec.Delta = sh->Delta;
  rec.rho = INFINITY;
  rec.stepsize = NAN;
  rec.accepted = 1;
  h->log.push_back(rec);

  if (use_graph && (h->tcg_graph_p != h->p || h->tcg_graph_maxinner != opt.maxinner || h->graph_sigma != h->sigma)) {
    destroy_graph(h);  // kernel arguments (widths, sigma) are baked into the captured launches
    h->tcg_graph_p = h->p;
    h->tcg_graph_maxinner = opt.maxinner;
    h->graph_sigma = h->sigma;
  }
  // kernels per tCG iteration / per TR iteration (for the launch counter in graph mode)
  int k = 0, naccepted = 0, stop_reason = 1;
  while (true) {
    const double gradnorm = sqrt(sh->gradnorm2);
    if (gradnorm < opt.tolgradnorm) {  // stoppingcriterion.m:50-72
      stop_reason = 0;
      break;
    }
    if (k >= opt.maxiter) {
      stop_reason = 1;
      break;
    }
    if (use_graph) {
      if (!h->tcg_exec[h->pt]) MSDP_TRY(build_tr_graph(h));
      CUDA_TRY(h, cudaGraphLaunch(h->tcg_exec[h->pt], h->stream));
    } else {
      MSDP_TRY(msdp_launch_tcg_init(h));
      int done = 0, issued = 0;
      // Iterations queued after tCG has stopped are no-ops for the kernels but NOT for the NCCL exchanges of a
      // row-sharded handle (an all-gather of the factor each), so those are issued one at a time: a 20 us flag read
      // per iteration against a >= 1 ms exchange.  Single-GPU stream mode grows the chunk instead, and so does the
      // direct peer-gather path (no exchange: a stopped iteration is three tiny all-reduces and no-op kernels).
      // (column-sharded: a stopped iteration still runs its three small all-reduces, ~0.1 ms: short chunks)
      const bool col_comm = h->col_split && h->cworld > 1;
      const bool cheap_stop = (h->world <= 1) && !col_comm;
      const int chunk_cap = cheap_stop ? 32 : ((col_comm || msdp_peer_gather_ok(h)) ? 4 : 1);
      int chunk = cheap_stop ? 4 : ((col_comm || msdp_peer_gather_ok(h)) ? 2 : 1);
      // column-sharded: consecutive TR iterations tend to need similar numbers of inner iterations -- issue as many
      // as the previous solve used (minus one) before the first look at the stop flag, then continue one at a time
      if (col_comm && h->last_numinner > 2) chunk = h->last_numinner - 1;
      bool first_chunk = true;
      while (!done && issued < opt.maxinner) {
        int c = chunk;
        if (c > opt.maxinner - issued) c = opt.maxinner - issued;
        for (int i = 0; i < c; ++i) MSDP_TRY(tcg_iteration(h, 0, 0));
        issued += c;
        CUDA_TRY(h, cudaMemcpyAsync(&h->st_host->stop, &h->st->stop, sizeof(int), cudaMemcpyDeviceToHost,
                                    h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        done = (h->st_host->stop != 0);
        if (col_comm && first_chunk) {
          chunk = 1;
          first_chunk = false;
        } else if (chunk < chunk_cap) {
          chunk *= 2;
        }
      }
      MSDP_TRY(tr_tail(h));
    }
    MSDP_TRY(msdp_sync_state(h));
    if (use_graph) h->launches += h->graph_l_fixed + h->graph_l_body * (int64_t)sh->j;
    h->last_numinner = sh->j;
    ++k;
    naccepted += sh->accepted;
    rec.iter = k;
    rec.cost = sh->fx;
    rec.gradnorm = sqrt(sh->gradnorm2);
    rec.Delta = sh->Delta;
    rec.rho = sh->rho;
    rec.stepsize = sh->norm_eta;
    rec.numinner = sh->j;
    rec.stop_inner = sh->stop;
    rec.accepted = sh->accepted;
    h->log.push_back(rec);
    if (!(sh->fx == sh->fx)) return msdp_fail(h, MANISDP_E_NUMERIC, "tr_solve: cost became NaN");
  }
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(h, cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->hv_total += sh->hv_count;
  if (info) {
    info->cost = sh->fx;
    info->gradnorm = sqrt(sh->gradnorm2);
    info->Delta = sh->Delta;
    info->seconds = ms * 1e-3;
    info->hv_count = sh->hv_count;
    info->iters = k;
    info->accepted = naccepted;
    info->stop_reason = stop_reason;
    info->reserved = 0;
  }
  return MANISDP_OK;
}
