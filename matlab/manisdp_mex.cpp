/*
 * manisdp_mex.cpp -- thin MATLAB MEX gateway over the C ABI of libmanisdp_b200.so (include/manisdp_b200.h).
 *
 *   mex -R2017b -I../include manisdp_mex.cpp -L../manisdp_matlab_b200 -lmanisdp_b200
 *
 * (separate-complex API, so mxGetPr / mxGetIr / mxGetJc hand out the CSC arrays without a copy; mwIndex is uint64 on
 * every 64-bit MATLAB, which is exactly the index type the C ABI takes.)  The build container has no mex.h, so this
 * file is source-only there; tests/ drive the same C ABI through ctypes (manisdp_matlab_b200/_lib.py).
 *
 * Usage from MATLAB (see ManiSDP_onlyunitdiag.m etc. in this directory):
 *   h    = manisdp_mex('create', kind, n, C)                      kind 0: C sparse n x n
 *   h    = manisdp_mex('create', 0, n, C, devices)                several GPUs from this one MATLAB session: devices =
 *                                                                 [0 1 ... ] (0-based CUDA ordinals) creates a group
 *                                                                 (manisdp_group_*: one worker thread + one column-sharded
 *                                                                 handle per device inside the library); set_Y, get_Y,
 *                                                                 rand_Y, cost, tr_solve, kkt, rank_cut, escape,
 *                                                                 line_search, stats, destroy work on it unchanged
 *   h    = manisdp_mex('create', kind, n, At, b, c)               kind 1..3: SeDuMi data
 *          manisdp_mex('set_Y', h, Y, layout)                     layout 0: p x n (unit-diag drivers), 1: n x p
 *   Y    = manisdp_mex('get_Y', h, layout)
 *          manisdp_mex('rand_Y', h, p, seed)
 *          manisdp_mex('set_dual', h, y, sigma) / manisdp_mex('set_sigma', h, sigma)
 *   [y, sigma] = manisdp_mex('get_dual', h)
 *   info = manisdp_mex('tr_solve', h, opts)                       opts fields: maxiter, maxinner, tolgradnorm, use_graph
 *   k    = manisdp_mex('kkt', h, delta, eig_tol, update_dual)
 *   [r, p] = manisdp_mex('rank_cut', h, theta, apply)
 *          manisdp_mex('escape', h, nne, alpha, line_search)
 *   a    = manisdp_mex('line_search', h)
 *   [vals, vecs] = manisdp_mex('get_eigs', h, k)
 *   f = manisdp_mex('cost', h); [G, gn] = manisdp_mex('grad', h); H = manisdp_mex('hess', h, U)   (A/B debugging)
 *          manisdp_mex('destroy', h)
 *
 * Ownership: prhs[] stay MATLAB's (read-only, copied to the device by the library); plhs[] are created here with
 * mxCreate* and handed to MATLAB (same convention as the reference's src/C-files/lincombc.cpp:16,30,39).  Errors are
 * raised with mexErrMsgIdAndTxt("ManiSDP:b200:<code>", msg) (reference convention: src/C-files/innerc.cpp:5-10).
 */
#include <string.h>
#include <string>
#include <vector>
#include "manisdp_b200.h"
#include "mex.h"

static std::vector<manisdp_t*> g_handles;
static std::vector<manisdp_group_t*> g_groups;  // same index space as g_handles: exactly one of the two is non-null

static void at_exit() {
  for (manisdp_t* h : g_handles)
    if (h) manisdp_destroy(h);
  for (manisdp_group_t* g : g_groups)
    if (g) manisdp_group_destroy(g);
  g_handles.clear();
  g_groups.clear();
}

static void fail(manisdp_t* h, int rc, const char* what) {
  const char* msg = manisdp_last_error(h);
  char id[64];
  snprintf(id, sizeof(id), "ManiSDP:b200:E%d", -rc);
  mexErrMsgIdAndTxt(id, "%s failed (%d): %s", what, rc, msg ? msg : "");
}
#define CK(h, expr, what)        \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) fail(h, _rc, what); \
  } while (0)

static uint64_t index_of(const mxArray* a) {
  if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "bad handle");
  const uint64_t idx = *(const uint64_t*)mxGetData(a);
  if (idx >= g_handles.size() || (!g_handles[idx] && !g_groups[idx])) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "stale handle");
  return idx;
}

static void gfail(manisdp_group_t* g, int rc, const char* what) {
  const char* msg = manisdp_group_last_error(g);
  char id[64];
  snprintf(id, sizeof(id), "ManiSDP:b200:E%d", -rc);
  mexErrMsgIdAndTxt(id, "%s failed (%d): %s", what, rc, msg ? msg : "");
}
#define GCK(g, expr, what)            \
  do {                                \
    int _rc = (expr);                 \
    if (_rc != 0) gfail(g, _rc, what); \
  } while (0)

// commands on a multi-GPU group (created with a device list); the outputs have the same shapes as on a single handle
static void group_command(manisdp_group_t* g, const std::string& c, const char* cmd, uint64_t idx, int nlhs, mxArray* plhs[],
                          int nrhs, const mxArray* prhs[]);

static double field_or(const mxArray* s, const char* name, double dflt) {
  const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
  return f ? mxGetScalar(f) : dflt;
}

static mxArray* scalar_struct(const char** names, const double* vals, int nf) {
  mxArray* s = mxCreateStructMatrix(1, 1, nf, names);
  for (int i = 0; i < nf; ++i) mxSetFieldByNumber(s, 0, i, mxCreateDoubleScalar(vals[i]));
  return s;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "first argument must be a command string");
  char cmd[32];
  mxGetString(prhs[0], cmd, sizeof(cmd));
  const std::string c(cmd);

  if (c == "create") {
    manisdp_problem pb;
    memset(&pb, 0, sizeof(pb));
    pb.kind = (int32_t)mxGetScalar(prhs[1]);
    pb.n = (int64_t)mxGetScalar(prhs[2]);
    pb.world = 1;
    pb.row_end = pb.n;
    std::vector<double> bdense;
    if (pb.kind == MANISDP_ONLYUNITDIAG) {
      const mxArray* C = prhs[3];
      if (!mxIsSparse(C) || !mxIsDouble(C)) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "C must be sparse double");
      pb.C_jc = (const uint64_t*)mxGetJc(C);
      pb.C_ir = (const uint64_t*)mxGetIr(C);
      pb.C_pr = mxGetPr(C);
    } else {
      const mxArray *At = prhs[3], *b = prhs[4], *cc = prhs[5];
      if (!mxIsSparse(At)) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "At must be sparse (SeDuMi format)");
      pb.m = (int64_t)mxGetN(At);
      pb.At_jc = (const uint64_t*)mxGetJc(At);
      pb.At_ir = (const uint64_t*)mxGetIr(At);
      pb.At_pr = mxGetPr(At);
      if (mxIsSparse(b)) {  // bqpmom.m:37 builds a sparse b
        bdense.assign((size_t)pb.m, 0.0);
        const mwIndex *jc = mxGetJc(b), *ir = mxGetIr(b);
        const double* pr = mxGetPr(b);
        for (mwIndex e = jc[0]; e < jc[mxGetN(b)]; ++e) bdense[ir[e]] = pr[e];
        pb.b = bdense.data();
      } else {
        pb.b = mxGetPr(b);
      }
      if (mxIsSparse(cc)) {
        pb.c_ir = (const uint64_t*)mxGetIr(cc);
        pb.c_pr = mxGetPr(cc);
        pb.c_nnz = (int64_t)mxGetJc(cc)[1];
      } else {
        pb.c_pr = mxGetPr(cc);
        pb.c_nnz = (int64_t)mxGetNumberOfElements(cc);
      }
    }
    std::vector<int64_t> blocks;
    if (pb.kind == MANISDP_MULTIBLOCK) {  // manisdp_mex('create', 4, sum(K.s), At, b, c, K.s, K.nob)
      if (nrhs < 8) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "multi-block create needs K.s and K.nob");
      const size_t t = mxGetNumberOfElements(prhs[6]);
      blocks.resize(t);
      for (size_t i = 0; i < t; ++i) blocks[i] = (int64_t)mxGetPr(prhs[6])[i];
      pb.nblocks = (int32_t)t;
      pb.nob = (int32_t)mxGetScalar(prhs[7]);
      pb.block_sizes = blocks.data();
    }
    if (pb.kind == MANISDP_DUAL_UNITDIAG) {
      // manisdp_mex('create', 5, n, A(:,K.f+1:end)', b, c(K.f+1:end), dAAt, A(:,1:K.f), c(1:K.f))   (ManiDSDP_unitdiag.m:34-40)
      if (nrhs > 6 && mxGetNumberOfElements(prhs[6]) > 0) pb.dAAt = mxGetPr(prhs[6]);
      if (nrhs > 8 && mxGetN(prhs[7]) > 0) {
        if (!mxIsSparse(prhs[7])) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "B must be sparse");
        pb.nfree = (int64_t)mxGetN(prhs[7]);
        pb.B_jc = (const uint64_t*)mxGetJc(prhs[7]);
        pb.B_ir = (const uint64_t*)mxGetIr(prhs[7]);
        pb.B_pr = mxGetPr(prhs[7]);
        pb.cf = mxGetPr(prhs[8]);
      }
    }
    manisdp_t* h = nullptr;
    manisdp_group_t* grp = nullptr;
    if (pb.kind == MANISDP_ONLYUNITDIAG && nrhs > 4 && mxGetNumberOfElements(prhs[4]) > 0) {
      // device list: one MATLAB session drives several GPUs (SURVEY 8b "Threading")
      const size_t nd = mxGetNumberOfElements(prhs[4]);
      std::vector<int32_t> dev(nd);
      for (size_t i = 0; i < nd; ++i) dev[i] = (int32_t)mxGetPr(prhs[4])[i];
      int rc = manisdp_group_create(&grp, &pb, (int32_t)nd, dev.data());
      if (rc != 0) gfail(nullptr, rc, "manisdp_group_create");
    } else {
      int rc = manisdp_create(&h, &pb);
      if (rc != 0) fail(nullptr, rc, "manisdp_create");
    }
    if (g_handles.empty()) {
      mexLock();
      mexAtExit(at_exit);
    }
    g_handles.push_back(h);
    g_groups.push_back(grp);
    plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(g_handles.size() - 1);
    return;
  }

  if (nrhs < 2) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "missing handle");
  const uint64_t hidx = index_of(prhs[1]);
  if (g_groups[hidx]) {
    group_command(g_groups[hidx], c, cmd, hidx, nlhs, plhs, nrhs, prhs);
    return;
  }
  manisdp_t* h = g_handles[hidx];
  manisdp_stats st;

  if (c == "destroy") {
    manisdp_destroy(h);
    g_handles[hidx] = nullptr;
  } else if (c == "set_Y") {
    const int layout = (int)mxGetScalar(prhs[3]);
    const int64_t p = layout == MANISDP_LAYOUT_ROWS ? (int64_t)mxGetM(prhs[2]) : (int64_t)mxGetN(prhs[2]);
    CK(h, manisdp_set_Y(h, mxGetPr(prhs[2]), p, layout), "set_Y");
  } else if (c == "get_Y") {
    const int layout = (int)mxGetScalar(prhs[2]);
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    plhs[0] = layout == MANISDP_LAYOUT_ROWS ? mxCreateDoubleMatrix((mwSize)st.p, (mwSize)st.n, mxREAL)
                                            : mxCreateDoubleMatrix((mwSize)st.n, (mwSize)st.p, mxREAL);
    CK(h, manisdp_get_Y(h, mxGetPr(plhs[0]), layout), "get_Y");
  } else if (c == "rand_Y") {
    CK(h, manisdp_rand_Y(h, (int64_t)mxGetScalar(prhs[2]), (uint64_t)mxGetScalar(prhs[3])), "rand_Y");
  } else if (c == "set_dual") {
    CK(h, manisdp_set_dual(h, mxGetPr(prhs[2]), mxGetScalar(prhs[3])), "set_dual");
  } else if (c == "set_sigma") {
    CK(h, manisdp_set_sigma(h, mxGetScalar(prhs[2])), "set_sigma");
  } else if (c == "get_dual") {
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    plhs[0] = mxCreateDoubleMatrix((mwSize)st.m, 1, mxREAL);
    double sigma = 0;
    CK(h, manisdp_get_dual(h, mxGetPr(plhs[0]), &sigma), "get_dual");
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(sigma);
  } else if (c == "tr_solve") {
    manisdp_tr_options o;
    memset(&o, 0, sizeof(o));
    const mxArray* s = nrhs > 2 ? prhs[2] : nullptr;
    if (s) {
      o.maxiter = (int32_t)field_or(s, "maxiter", 0);
      o.maxinner = (int32_t)field_or(s, "maxinner", 0);
      o.tolgradnorm = field_or(s, "tolgradnorm", 0);
      o.use_graph = (int32_t)field_or(s, "use_graph", 1);
    }
    manisdp_tr_info info;
    CK(h, manisdp_tr_solve(h, &o, &info), "tr_solve");
    const char* names[] = {"cost", "gradnorm", "Delta", "seconds", "hv_count", "iters", "accepted", "stop_reason"};
    const double vals[] = {info.cost, info.gradnorm, info.Delta, info.seconds, (double)info.hv_count,
                           (double)info.iters, (double)info.accepted, (double)info.stop_reason};
    plhs[0] = scalar_struct(names, vals, 8);
  } else if (c == "kkt") {
    manisdp_kkt_info k;
    CK(h, manisdp_kkt(h, (int32_t)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), (int32_t)mxGetScalar(prhs[4]), &k), "kkt");
    const char* names[] = {"obj", "by", "pinf", "dinf", "gap", "lam_min", "lam_max", "z_sum", "nneg", "eig_iters",
                           "eig_resid", "eig_converged"};
    const double vals[] = {k.obj, k.by, k.pinf, k.dinf, k.gap, k.lam_min, k.lam_max, k.z_sum, (double)k.nneg,
                           (double)k.eig_iters, k.eig_resid, (double)k.eig_converged};
    plhs[0] = scalar_struct(names, vals, 12);
  } else if (c == "rank_cut") {
    int64_t r = 0, p = 0;
    CK(h, manisdp_rank_cut(h, mxGetScalar(prhs[2]), (int32_t)mxGetScalar(prhs[3]), &r, &p), "rank_cut");
    plhs[0] = mxCreateDoubleScalar((double)r);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar((double)p);
  } else if (c == "escape") {
    CK(h, manisdp_escape(h, (int32_t)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), (int32_t)mxGetScalar(prhs[4])),
       "escape");
  } else if (c == "line_search") {
    double a = 0;
    CK(h, manisdp_line_search(h, &a), "line_search");
    plhs[0] = mxCreateDoubleScalar(a);
  } else if (c == "get_eigs") {
    const int k = (int)mxGetScalar(prhs[2]);
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    plhs[0] = mxCreateDoubleMatrix((mwSize)k, 1, mxREAL);
    mxArray* V = mxCreateDoubleMatrix((mwSize)k, (mwSize)st.n, mxREAL);  // k x n column-major == n x k rows
    CK(h, manisdp_get_eigs(h, mxGetPr(plhs[0]), mxGetPr(V), k), "get_eigs");
    if (nlhs > 1)
      plhs[1] = V;
    else
      mxDestroyArray(V);
  } else if (c == "cost") {
    double f = 0;
    CK(h, manisdp_cost(h, &f), "cost");
    plhs[0] = mxCreateDoubleScalar(f);
  } else if (c == "grad") {
    double gn = 0;
    CK(h, manisdp_grad(h, &gn), "grad");
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    plhs[0] = mxCreateDoubleMatrix((mwSize)st.p, (mwSize)st.n, mxREAL);
    CK(h, manisdp_slot_get(h, MANISDP_SLOT_G, mxGetPr(plhs[0]), MANISDP_LAYOUT_ROWS), "slot_get");
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(gn);
  } else if (c == "hess") {
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    CK(h, manisdp_slot_set(h, MANISDP_SLOT_U, mxGetPr(prhs[2]), MANISDP_LAYOUT_ROWS), "slot_set");
    CK(h, manisdp_hess(h), "hess");
    plhs[0] = mxCreateDoubleMatrix((mwSize)st.p, (mwSize)st.n, mxREAL);
    CK(h, manisdp_slot_get(h, MANISDP_SLOT_H, mxGetPr(plhs[0]), MANISDP_LAYOUT_ROWS), "slot_get");
  } else if (c == "dual_state") {
    // [x, w] = manisdp_mex('dual_state', h, nfree): the ADMM multipliers of a dual handle (x: n*n vector)
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    const size_t nf = (size_t)mxGetScalar(prhs[2]);
    plhs[0] = mxCreateDoubleMatrix((mwSize)(st.n * st.n), 1, mxREAL);
    mxArray* w = mxCreateDoubleMatrix((mwSize)nf, 1, mxREAL);
    CK(h, manisdp_dual_get_state(h, mxGetPr(plhs[0]), nf ? mxGetPr(w) : nullptr), "dual_get_state");
    if (nlhs > 1)
      plhs[1] = w;
    else
      mxDestroyArray(w);
  } else if (c == "mb_set_Y") {
    // Y: cell array, Y{i} is p_i x n_i (the layout of ManiSDP_multiblock.m; its memory image is n_i rows of p_i doubles)
    const mxArray* Y = prhs[2];
    if (!mxIsCell(Y)) mexErrMsgIdAndTxt("ManiSDP:b200:arg", "mb_set_Y: Y must be a cell array");
    const size_t t = mxGetNumberOfElements(Y);
    std::vector<int64_t> p(t);
    std::vector<double> cat;
    for (size_t i = 0; i < t; ++i) {
      const mxArray* Yi = mxGetCell(Y, i);
      p[i] = (int64_t)mxGetM(Yi);
      cat.insert(cat.end(), mxGetPr(Yi), mxGetPr(Yi) + mxGetNumberOfElements(Yi));
    }
    CK(h, manisdp_mb_set_Y(h, cat.data(), p.data()), "mb_set_Y");
  } else if (c == "mb_get_Y" || c == "mb_widths") {
    // sizes: prhs[2] = K.s
    const size_t t = mxGetNumberOfElements(prhs[2]);
    std::vector<int64_t> p(t);
    CK(h, manisdp_mb_get_widths(h, p.data()), "mb_get_widths");
    if (c == "mb_widths") {
      plhs[0] = mxCreateDoubleMatrix((mwSize)t, 1, mxREAL);
      for (size_t i = 0; i < t; ++i) mxGetPr(plhs[0])[i] = (double)p[i];
    } else {
      size_t total = 0;
      for (size_t i = 0; i < t; ++i) total += (size_t)p[i] * (size_t)mxGetPr(prhs[2])[i];
      std::vector<double> cat(total);
      CK(h, manisdp_mb_get_Y(h, cat.data()), "mb_get_Y");
      plhs[0] = mxCreateCellMatrix((mwSize)t, 1);
      size_t o = 0;
      for (size_t i = 0; i < t; ++i) {
        const size_t ni = (size_t)mxGetPr(prhs[2])[i];
        mxArray* Yi = mxCreateDoubleMatrix((mwSize)p[i], (mwSize)ni, mxREAL);
        memcpy(mxGetPr(Yi), cat.data() + o, (size_t)p[i] * ni * sizeof(double));
        o += (size_t)p[i] * ni;
        mxSetCell(plhs[0], (mwIndex)i, Yi);
      }
    }
  } else if (c == "mb_rand_Y") {
    const size_t t = mxGetNumberOfElements(prhs[2]);
    std::vector<int64_t> p(t);
    for (size_t i = 0; i < t; ++i) p[i] = (int64_t)mxGetPr(prhs[2])[i];
    CK(h, manisdp_mb_rand_Y(h, p.data(), (uint64_t)mxGetScalar(prhs[3])), "mb_rand_Y");
  } else if (c == "mb_kkt") {
    // [k, dinfs] = manisdp_mex('mb_kkt', h, update_dual, K.s)
    const size_t t = mxGetNumberOfElements(prhs[3]);
    manisdp_kkt_info k;
    mxArray* d = mxCreateDoubleMatrix((mwSize)t, 1, mxREAL);
    CK(h, manisdp_mb_kkt(h, (int32_t)mxGetScalar(prhs[2]), &k, mxGetPr(d), nullptr), "mb_kkt");
    const char* names[] = {"obj", "by", "pinf", "dinf", "gap", "lam_min", "lam_max", "z_sum", "nneg"};
    const double vals[] = {k.obj, k.by, k.pinf, k.dinf, k.gap, k.lam_min, k.lam_max, k.z_sum, (double)k.nneg};
    plhs[0] = scalar_struct(names, vals, 9);
    if (nlhs > 1)
      plhs[1] = d;
    else
      mxDestroyArray(d);
  } else if (c == "mb_update") {
    // p = manisdp_mex('mb_update', h, theta, delta, alpha, line_search, min_facsize, K.s)
    const size_t t = mxGetNumberOfElements(prhs[7]);
    std::vector<int64_t> p(t);
    CK(h, manisdp_mb_update(h, mxGetScalar(prhs[2]), (int32_t)mxGetScalar(prhs[3]), mxGetScalar(prhs[4]),
                            (int32_t)mxGetScalar(prhs[5]), (int32_t)mxGetScalar(prhs[6]), p.data()), "mb_update");
    plhs[0] = mxCreateDoubleMatrix((mwSize)t, 1, mxREAL);
    for (size_t i = 0; i < t; ++i) mxGetPr(plhs[0])[i] = (double)p[i];
  } else if (c == "stats") {
    CK(h, manisdp_get_stats(h, &st), "get_stats");
    const char* names[] = {"n", "m", "p", "nnzC", "nnzA", "s_mode", "a_mode", "hv_total", "launches_total",
                           "bytes_per_hv", "flops_per_hv"};
    const double vals[] = {(double)st.n, (double)st.m, (double)st.p, (double)st.nnzC, (double)st.nnzA,
                           (double)st.s_mode, (double)st.a_mode, (double)st.hv_total, (double)st.launches_total,
                           st.bytes_per_hv, st.flops_per_hv};
    plhs[0] = scalar_struct(names, vals, 11);
  } else {
    mexErrMsgIdAndTxt("ManiSDP:b200:arg", "unknown command '%s'", cmd);
  }
}

static void group_command(manisdp_group_t* g, const std::string& c, const char* cmd, uint64_t idx, int nlhs, mxArray* plhs[],
                          int nrhs, const mxArray* prhs[]) {
  manisdp_stats st;
  if (c == "destroy") {
    manisdp_group_destroy(g);
    g_groups[idx] = nullptr;
  } else if (c == "set_Y") {
    const int layout = (int)mxGetScalar(prhs[3]);
    const int64_t p = layout == MANISDP_LAYOUT_ROWS ? (int64_t)mxGetM(prhs[2]) : (int64_t)mxGetN(prhs[2]);
    GCK(g, manisdp_group_set_Y(g, mxGetPr(prhs[2]), p, layout), "set_Y");
  } else if (c == "get_Y") {
    const int layout = (int)mxGetScalar(prhs[2]);
    GCK(g, manisdp_group_get_stats(g, &st), "get_stats");
    plhs[0] = layout == MANISDP_LAYOUT_ROWS ? mxCreateDoubleMatrix((mwSize)st.p, (mwSize)st.n, mxREAL)
                                            : mxCreateDoubleMatrix((mwSize)st.n, (mwSize)st.p, mxREAL);
    GCK(g, manisdp_group_get_Y(g, mxGetPr(plhs[0]), layout), "get_Y");
  } else if (c == "rand_Y") {
    GCK(g, manisdp_group_rand_Y(g, (int64_t)mxGetScalar(prhs[2]), (uint64_t)mxGetScalar(prhs[3])), "rand_Y");
  } else if (c == "cost") {
    double f = 0;
    GCK(g, manisdp_group_cost(g, &f), "cost");
    plhs[0] = mxCreateDoubleScalar(f);
  } else if (c == "tr_solve") {
    manisdp_tr_options o;
    memset(&o, 0, sizeof(o));
    const mxArray* s = nrhs > 2 ? prhs[2] : nullptr;
    if (s) {
      o.maxiter = (int32_t)field_or(s, "maxiter", 0);
      o.maxinner = (int32_t)field_or(s, "maxinner", 0);
      o.tolgradnorm = field_or(s, "tolgradnorm", 0);
      o.use_graph = (int32_t)field_or(s, "use_graph", 1);
    }
    manisdp_tr_info info;
    GCK(g, manisdp_group_tr_solve(g, &o, &info), "tr_solve");
    const char* names[] = {"cost", "gradnorm", "Delta", "seconds", "hv_count", "iters", "accepted", "stop_reason"};
    const double vals[] = {info.cost, info.gradnorm, info.Delta, info.seconds, (double)info.hv_count,
                           (double)info.iters, (double)info.accepted, (double)info.stop_reason};
    plhs[0] = scalar_struct(names, vals, 8);
  } else if (c == "kkt") {
    manisdp_kkt_info k;
    GCK(g, manisdp_group_kkt(g, (int32_t)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), (int32_t)mxGetScalar(prhs[4]), &k),
        "kkt");
    const char* names[] = {"obj", "by", "pinf", "dinf", "gap", "lam_min", "lam_max", "z_sum", "nneg", "eig_iters",
                           "eig_resid", "eig_converged"};
    const double vals[] = {k.obj, k.by, k.pinf, k.dinf, k.gap, k.lam_min, k.lam_max, k.z_sum, (double)k.nneg,
                           (double)k.eig_iters, k.eig_resid, (double)k.eig_converged};
    plhs[0] = scalar_struct(names, vals, 12);
  } else if (c == "rank_cut") {
    int64_t r = 0, p = 0;
    GCK(g, manisdp_group_rank_cut(g, mxGetScalar(prhs[2]), (int32_t)mxGetScalar(prhs[3]), &r, &p), "rank_cut");
    plhs[0] = mxCreateDoubleScalar((double)r);
    if (nlhs > 1) plhs[1] = mxCreateDoubleScalar((double)p);
  } else if (c == "escape") {
    GCK(g, manisdp_group_escape(g, (int32_t)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), (int32_t)mxGetScalar(prhs[4])),
        "escape");
  } else if (c == "line_search") {
    double a = 0;
    GCK(g, manisdp_group_line_search(g, &a), "line_search");
    plhs[0] = mxCreateDoubleScalar(a);
  } else if (c == "stats") {
    GCK(g, manisdp_group_get_stats(g, &st), "get_stats");
    const char* names[] = {"n", "m", "p", "nnzC", "nnzA", "s_mode", "a_mode", "hv_total", "launches_total",
                           "bytes_per_hv", "flops_per_hv", "n_devices"};
    const double vals[] = {(double)st.n, (double)st.m, (double)st.p, (double)st.nnzC, (double)st.nnzA,
                           (double)st.s_mode, (double)st.a_mode, (double)st.hv_total, (double)st.launches_total,
                           st.bytes_per_hv, st.flops_per_hv, (double)manisdp_group_size(g)};
    plhs[0] = scalar_struct(names, vals, 12);
  } else {
    mexErrMsgIdAndTxt("ManiSDP:b200:arg", "command '%s' is not available on a multi-GPU group", cmd);
  }
}
