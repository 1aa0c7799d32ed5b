/*
 * manisdp_b200.h -- C ABI of libmanisdp_b200.so, the B200 (sm_100a) engine for ManiSDP's inner hot path.
 *
 * What it replaces in the reference (wangjie212/ManiSDP-matlab; paths relative to the reference root):
 *   - Manopt's Riemannian trust-region solver and its truncated CG
 *       manopt7.0/manopt/solvers/trustregions/trustregions.m:395-767, tCG.m:95-292
 *   - the cost / grad / hess closures and manifold structs the four primal drivers hand to it
 *       src/primal/ManiSDP_onlyunitdiag.m:117-156, ManiSDP_unitdiag.m:152-198,
 *       ManiSDP_unittrace.m:156-177 (+ spherefactory.m), ManiSDP.m:149-165 (+ euclideanfactory.m)
 *   - the eig(S) saddle-escape step and the KKT / rank / escape pieces of the outer loops
 *       ManiSDP_onlyunitdiag.m:45-84, ManiSDP_unitdiag.m:59-112, ManiSDP_unittrace.m:59-117, ManiSDP.m:59-113
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no exceptions cross the boundary.  Every function returns an
 *     int status: 0 = ok, < 0 = error (see MANISDP_E_*); manisdp_last_error() gives the message.
 *   - Sparse inputs are MATLAB-style CSC with 64-bit unsigned indices (mwIndex): jc[ncols+1], ir[nnz], pr[nnz],
 *     0-based, exactly what mxGetJc / mxGetIr / mxGetPr return.  The library copies everything at create time
 *     and never retains host pointers.
 *   - `At` is the SeDuMi constraint matrix, n*n rows by m columns; row index r = j*n + i addresses X(i,j)
 *     (column-major vec).  The index split r -> (i, j) is done in 64-bit integer arithmetic.
 *   - The factor Y lives on the device as n rows ("vertex-major"), each row p doubles padded to ld = 4*ceil(p/4).
 *     Host buffers passed to set/get are either MANISDP_LAYOUT_ROWS (n x p row-major == MATLAB p x n column-major,
 *     the layout of the unit-diagonal drivers, ManiSDP_unitdiag.m:53) or MANISDP_LAYOUT_COLS (n x p column-major,
 *     MATLAB's layout in ManiSDP.m / ManiSDP_unittrace.m).
 *   - Limits of this implementation (the reference has none): factor width p <= 1024 on the affine and dual kinds,
 *     p <= 512 on ONLYUNITDIAG and MULTIBLOCK handles (row-group geometry of the fused kernels; manisdp_set_Y / escape /
 *     mb_update fail with MANISDP_E_ARG beyond it), n and m < 2^31, dense S for
 *     n <= 40000, options.delta <= 60, block orders <= 1024 for the device block eigensolver.
 *   - A handle is single-caller (not re-entrant); different handles may be used from different threads.
 *   - There is no CPU fallback: every compute entry point fails with MANISDP_E_CUDA when no sm_100 device works.
 */
#ifndef MANISDP_B200_H
#define MANISDP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct manisdp_handle manisdp_t;

/* status codes */
#define MANISDP_OK 0
#define MANISDP_E_ARG (-1)     /* bad argument / inconsistent sizes */
#define MANISDP_E_CUDA (-2)    /* CUDA runtime / launch / allocation failure */
#define MANISDP_E_NCCL (-3)    /* NCCL failure (row-sharded handles) */
#define MANISDP_E_STATE (-4)   /* call out of order (e.g. solve before set_Y) */
#define MANISDP_E_NUMERIC (-5) /* NaN / breakdown inside an eigen step */

/* which reference driver's closures + manifold the handle implements */
enum {
  MANISDP_ONLYUNITDIAG = 0, /* src/primal/ManiSDP_onlyunitdiag.m : oblique rows, f = 1/2 <C, YY'> */
  MANISDP_UNITDIAG = 1,     /* src/primal/ManiSDP_unitdiag.m     : oblique rows + AL on A(X) = b   */
  MANISDP_UNITTRACE = 2,    /* src/primal/ManiSDP_unittrace.m    : unit Frobenius sphere + AL      */
  MANISDP_GENERAL = 3,      /* src/primal/ManiSDP.m              : Euclidean + AL                  */
  MANISDP_MULTIBLOCK = 4,   /* src/primal/ManiSDP_multiblock.m   : product of oblique / Euclidean blocks + AL
                               (manifold ops: src/basicfunction/multiblockmanifold.m:1-42, src/C-files/ sources) */
  MANISDP_DUAL_UNITDIAG = 5 /* src/dual/ManiDSDP_unitdiag.m      : Riemannian ADMM on the SOS (dual) form, oblique rows */
};

enum { MANISDP_LAYOUT_ROWS = 0, MANISDP_LAYOUT_COLS = 1 };

/* device-resident n x p work arrays ("slots") addressable through the fine-grained closure calls */
enum {
  MANISDP_SLOT_Y = 0,     /* current point */
  MANISDP_SLOT_YPROP = 1, /* proposal x_prop = retr(Y, eta) */
  MANISDP_SLOT_G = 2,     /* Riemannian gradient at Y */
  MANISDP_SLOT_ETA = 3,   /* tCG iterate */
  MANISDP_SLOT_R = 4,     /* tCG residual */
  MANISDP_SLOT_D = 5,     /* tCG direction (mdelta) */
  MANISDP_SLOT_HD = 6,    /* Hess[mdelta] */
  MANISDP_SLOT_U = 7,     /* user scratch (hess_apply input, escape direction) */
  MANISDP_SLOT_H = 8,     /* user scratch (hess_apply output) */
  MANISDP_NUM_SLOTS = 9
};

/* problem description (host pointers, copied at create) */
typedef struct {
  int32_t kind;   /* MANISDP_ONLYUNITDIAG ... MANISDP_GENERAL */
  int32_t device; /* CUDA device ordinal */
  int64_t n;      /* order of X (K.s) */
  int64_t m;      /* number of affine constraints (0 for ONLYUNITDIAG) */
  /* ONLYUNITDIAG: C, n x n sparse symmetric CSC (ManiSDP_onlyunitdiag.m:6) */
  const uint64_t *C_jc, *C_ir;
  const double *C_pr;
  /* affine kinds: At (n*n x m, CSC), b (m, dense), c (n*n): dense if c_ir == NULL, else sparse vector with
   * c_nnz entries (c_ir = row indices into vec(X), c_pr = values)  (ManiSDP_unitdiag.m:7) */
  const uint64_t *At_jc, *At_ir;
  const double *At_pr;
  const double *b;
  const uint64_t *c_ir;
  const double *c_pr;
  int64_t c_nnz;
  /* row sharding: rank r of `world` owns rows [row_begin, row_end) of Y and of C.  world <= 1 means
   * unsharded.  For a sharded handle C_jc/C_ir/C_pr describe only the OWNED columns (= rows, C symmetric) of C:
   * C_jc has (row_end - row_begin + 1) entries and ir holds GLOBAL row indices. */
  int32_t rank, world;
  int64_t row_begin, row_end;
  const void *nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks (world > 1), else NULL */
  int32_t force_mode;         /* 0 auto; bit0 force dense S; bit1 force sparse S; bit2 force dense A; bit3 force sparse A */
  /* world > 1 only.  MANISDP_SHARD_ROWS (0): the row sharding described above.  MANISDP_SHARD_COLS (1): every rank
   * holds ALL rows of C (C_jc/C_ir/C_pr describe the whole matrix, row_begin = 0, row_end = n) and, between
   * manisdp_col_split and manisdp_col_merge, ceil(p/world) COLUMNS of the factor: the Hessian product needs no exchange
   * of the factor, only all-reduces of per-row scalars (n doubles) -- SURVEY 8e "p-sharding". */
  int32_t shard_layout;
  /* MANISDP_MULTIBLOCK only (ManiSDP_multiblock.m:7-9): K.s = block_sizes[0..nblocks), the first `nob` blocks have a
   * unit diagonal (K.nob).  n must equal sum(block_sizes); At has sum(block_sizes[i]^2) rows -- the stacked column-major
   * vecs of the blocks, exactly the reference's x (:44, :67-72) -- and c that many entries.  The library maps row r to
   * (block, i, j) in 64-bit integer arithmetic at create time. */
  int32_t nblocks, nob;
  const int64_t *block_sizes;
  /* MANISDP_DUAL_UNITDIAG only (ManiDSDP_unitdiag.m:8,35-44).  The caller's A (m x (K.f + n*n)) is passed split:
   * At = A(:, K.f+1:end)' (n*n x m, CSC -- MATLAB builds it with one transpose), c = c(K.f+1:end), b = b;
   * B = A(:, 1:K.f) (m x nfree, CSC) with its cost cf = c(1:K.f) when nfree = K.f > 0; dAAt = options.dAAt (m doubles)
   * or NULL for diag(A*A') computed by the library (:40). */
  const double *dAAt;
  int64_t nfree;
  const uint64_t *B_jc, *B_ir;
  const double *B_pr;
  const double *cf;
} manisdp_problem;
enum { MANISDP_SHARD_ROWS = 0, MANISDP_SHARD_COLS = 1 };

/* trust-region options: trustregions.m:340-372 defaults are applied for fields left at 0 */
typedef struct {
  int32_t maxiter;        /* opts.maxiter     (ManiSDP: options.TR_maxiter) */
  int32_t maxinner;       /* opts.maxinner    (ManiSDP: options.TR_maxinner) */
  int32_t mininner;       /* default 1 */
  int32_t use_graph;      /* 1: tCG loop runs as one CUDA graph with a device-side WHILE node; 0: stream launches */
  double tolgradnorm;     /* opts.tolgradnorm */
  double kappa;           /* default 0.1 */
  double theta;           /* default 1.0 */
  double rho_prime;       /* default 0.1 */
  double rho_regularization; /* default 1e3 */
  double Delta_bar;       /* default M.typicaldist() */
  double Delta0;          /* default Delta_bar / 8 */
} manisdp_tr_options;

typedef struct {
  double cost;       /* f at the returned point */
  double gradnorm;   /* info(end).gradnorm, the only field the reference drivers consume */
  double Delta;      /* final trust-region radius */
  double seconds;    /* device time of the call (CUDA events) */
  int64_t hv_count;  /* Hessian-vector products (tCG.m:163 calls) */
  int32_t iters;     /* outer TR iterations performed */
  int32_t accepted;  /* accepted steps */
  int32_t stop_reason; /* 0 gradnorm < tol, 1 maxiter */
  int32_t reserved;
} manisdp_tr_info;

/* one record per TR iteration, the analogue of Manopt's info struct array (trustregions.m:790-826) */
typedef struct {
  double cost, gradnorm, Delta, rho, stepsize;
  int32_t iter, numinner, stop_inner, accepted;
} manisdp_tr_iter;

/* KKT residues of the outer loop (ManiSDP_unitdiag.m:59-71 and siblings) */
typedef struct {
  double obj;     /* <C, X> */
  double by;      /* dual objective */
  double pinf, dinf, gap;
  double lam_min, lam_max; /* extreme eigenvalues of the dual slack S */
  double z_sum;   /* sum(z) (unit-diag kinds) or z (unit-trace) */
  int32_t nneg;   /* min(#negative eigenvalues found, delta requested) */
  int32_t eig_iters;
  double eig_resid; /* largest residual norm among the returned eigenpairs */
  int32_t eig_converged; /* 1: the residual test |S v - lambda v| <= eig_tol*(1+lambda_max) was met for all of them */
  int32_t reserved;
} manisdp_kkt_info;

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
int manisdp_create(manisdp_t **out, const manisdp_problem *prob);
int manisdp_destroy(manisdp_t *h);
const char *manisdp_last_error(const manisdp_t *h); /* h may be NULL: message of the last failed create */
int manisdp_version(void);

/* ---- state ------------------------------------------------------------------------------------------------- */
/* set the factor width p and upload Y (n_local x p); replaces M = factory(p, n) + x0 (ManiSDP_unitdiag.m:53,57) */
int manisdp_set_Y(manisdp_t *h, const double *Y, int64_t p, int32_t layout);
int manisdp_get_Y(manisdp_t *h, double *Y, int32_t layout);
int manisdp_get_p(manisdp_t *h, int64_t *p);
/* M.rand() of the manifold (ManiSDP_unitdiag.m:194-197 / spherefactory.m:249 / euclideanfactory.m:81) from a
 * counter-based generator (Philox4x32-10 + Box-Muller); reproducible for a given (seed, n, p) on any shard layout */
int manisdp_rand_Y(manisdp_t *h, int64_t p, uint64_t seed);
/* AL state: y (m) and sigma (ManiSDP_unitdiag.m:34-36) */
int manisdp_set_dual(manisdp_t *h, const double *y, double sigma);
int manisdp_get_dual(manisdp_t *h, double *y, double *sigma);
/* copy a slot to / from the host (tests, A/B debugging) */
int manisdp_slot_set(manisdp_t *h, int32_t slot, const double *src, int32_t layout);
int manisdp_slot_get(manisdp_t *h, int32_t slot, double *dst, int32_t layout);

/* ---- device-resident closures (problem.cost / .grad / .hess, ManiSDP_unitdiag.m:41-43) ---------------------- */
/* f(Y) at SLOT_Y; also refreshes the per-point caches the reference keeps in shared closure variables
 * (YC, eG / Axb) */
int manisdp_cost(manisdp_t *h, double *f);
/* Riemannian gradient at SLOT_Y into SLOT_G; gradnorm returned.  Requires manisdp_cost at the same point first,
 * exactly like the reference closures -- the library performs it itself if the cache is stale. */
int manisdp_grad(manisdp_t *h, double *gradnorm);
/* SLOT_H = Hess f(Y)[SLOT_U] (Riemannian, projection included) */
int manisdp_hess(manisdp_t *h);
/* benchmark hook: `reps` back-to-back Hessian products SLOT_U -> SLOT_H; returns mean device ms per product */
int manisdp_hess_bench(manisdp_t *h, int32_t reps, double *ms_per_hv);
/* benchmark hook for the fused vector kernels: which = 0 retraction, 1 tangent projection, 2 tCG update pass,
 * 3 tCG direction pass; returns mean device ms per launch and the algorithmic bytes of one launch */
int manisdp_vec_bench(manisdp_t *h, int32_t which, int32_t reps, double *ms_per_launch, double *bytes);
/* manifold ops on slots: dst = retr(Y, eta) ; dst = proj_Y(src) */
int manisdp_retract(manisdp_t *h, int32_t eta_slot, int32_t dst_slot);
int manisdp_project(manisdp_t *h, int32_t src_slot, int32_t dst_slot);

/* ---- the solver (trustregions(problem, Y, opts), ManiSDP_unitdiag.m:57) -------------------------------------- */
int manisdp_tr_solve(manisdp_t *h, const manisdp_tr_options *opts, manisdp_tr_info *info);
/* per-iteration log of the last tr_solve; returns the number of records written (<= cap) in *count */
int manisdp_tr_log(manisdp_t *h, manisdp_tr_iter *buf, int32_t cap, int32_t *count);

/* ---- outer-loop pieces ------------------------------------------------------------------------------------- */
/* KKT residues + dual update y <- y - sigma*(A(X) - b) + the `delta` smallest eigenpairs of S (device LOBPCG) and
 * lambda_max (Lanczos).  The eigenvectors stay on the device for manisdp_escape.
 * update_dual = 0 evaluates the residues without changing y.
 * eig_tol: residual tolerance of the eigen step relative to 1 + lambda_max.  0: 1e-9.  Negative: ADAPTIVE, |eig_tol| is
 * the caller's KKT tolerance -- the tolerance follows the previous dinf (1e-2 * dinf_prev in [1e-9, 1e-5]) and the block is
 * continued to 1e-9 whenever the new dinf comes within 100x of the KKT tolerance, so an accepted dinf < tol rests on
 * the tight tolerance (what the drivers pass by default). */
int manisdp_kkt(manisdp_t *h, int32_t delta, double eig_tol, int32_t update_dual, manisdp_kkt_info *out);
/* eigenvalues (ascending, count = min(cap, delta of the last kkt call)) and optionally vectors (n x count, ROWS) */
int manisdp_get_eigs(manisdp_t *h, double *vals, double *vecs, int32_t cap);
/* rank estimate + cut through the p x p Gram matrix (replaces svd(Y), ManiSDP_unitdiag.m:72-74,93-96):
 * r = #{sigma_i >= theta*sigma_1}; if r <= p-1 the factor is replaced by its rank-r truncation.
 * apply = 0 only reports r. */
int manisdp_rank_cut(manisdp_t *h, double theta, int32_t apply, int64_t *r, int64_t *p_new);
/* append nne escape directions (ManiSDP_unitdiag.m:97-107): line_search = 0: Y <- normalise([Y, alpha*V]);
 * line_search = 1: backtracking search along U = [0, V] (ManiSDP_unitdiag.m:138-150) */
int manisdp_escape(manisdp_t *h, int32_t nne, double alpha, int32_t line_search);
/* backtracking line search along the direction staged by manisdp_escape(..., line_search = 1); the reference runs
 * it at the top of the next outer iteration, i.e. after sigma was updated (ManiSDP_unitdiag.m:54-56,138-150).
 * *alpha receives the accepted step. */
int manisdp_line_search(manisdp_t *h, double *alpha);
/* AL penalty update hook: set sigma only (y untouched) */
int manisdp_set_sigma(manisdp_t *h, double sigma);
/* row sharding: rank 0 creates the 128-byte NCCL id, the host language broadcasts it (torch.distributed, MPI, ...) */
int manisdp_nccl_unique_id(void *out128);

/* ---- introspection ----------------------------------------------------------------------------------------- */
typedef struct {
  int64_t n, n_local, m, p, ld;
  int64_t nnzC, nnzA;
  int32_t kind, s_mode, a_mode; /* modes: 0 none, 1 sparse, 2 dense */
  int32_t rank, world;
  int64_t hv_total;      /* Hessian products since create */
  int64_t launches_total;/* kernels launched by this handle since create */
  double bytes_per_hv;   /* algorithmic bytes of one Hessian product at the current p (SURVEY 8d formulas) */
  double flops_per_hv;
} manisdp_stats;
int manisdp_get_stats(manisdp_t *h, manisdp_stats *out);
/* Column-sharded handles (shard_layout = MANISDP_SHARD_COLS), collective over the ranks.  A handle is created MERGED:
 * every rank holds the same full-width factor and set_Y / rand_Y / kkt / rank_cut / escape run identically
 * (deterministically) on every rank.  manisdp_col_split keeps columns [rank*pl, (rank+1)*pl), pl = ceil(p/world), of the
 * current point for the trust-region solve (the factor width becomes world*pl: zero columns are appended when world
 * does not divide p); manisdp_col_merge all-gathers the slices back.  While split, only tr_solve, cost, grad, hess,
 * hess_bench, get_Y / set_Y / slot access (on the local n x pl slice) and get_stats are available. */
int manisdp_col_split(manisdp_t *h);
int manisdp_col_merge(manisdp_t *h);

/* ---- single-process multi-GPU group (SURVEY 8b "Threading": one caller thread drives G devices) -----------------
 * For host languages that call synchronously from ONE thread (MATLAB's MEX interface): the group owns one worker thread
 * and one column-sharded ONLYUNITDIAG handle per device, creates the NCCL id in-process, and every group call runs the
 * corresponding handle call on all devices and returns when all are done.  `prob` describes the WHOLE problem (kind
 * ONLYUNITDIAG, full C; rank / world / device / row_* / nccl_unique_id / shard_layout are filled in by the group).
 * Between calls the factor is merged (full width on every device); manisdp_group_tr_solve = split -> tr_solve -> merge.
 * Results (Y, info, KKT record, rank) are those of device 0 -- identical on all devices by construction. */
typedef struct manisdp_group manisdp_group_t;
int manisdp_group_create(manisdp_group_t **out, const manisdp_problem *prob, int32_t ndev, const int32_t *devices);
int manisdp_group_destroy(manisdp_group_t *g);
int manisdp_group_size(const manisdp_group_t *g);
const char *manisdp_group_last_error(const manisdp_group_t *g);
int manisdp_group_set_Y(manisdp_group_t *g, const double *Y, int64_t p, int32_t layout);
int manisdp_group_rand_Y(manisdp_group_t *g, int64_t p, uint64_t seed);
int manisdp_group_get_Y(manisdp_group_t *g, double *Y, int32_t layout);
int manisdp_group_get_stats(manisdp_group_t *g, manisdp_stats *out);
int manisdp_group_cost(manisdp_group_t *g, double *f);
int manisdp_group_tr_solve(manisdp_group_t *g, const manisdp_tr_options *opts, manisdp_tr_info *info);
int manisdp_group_kkt(manisdp_group_t *g, int32_t delta, double eig_tol, int32_t update_dual, manisdp_kkt_info *out);
int manisdp_group_rank_cut(manisdp_group_t *g, double theta, int32_t apply, int64_t *r, int64_t *p_new);
int manisdp_group_escape(manisdp_group_t *g, int32_t nne, double alpha, int32_t line_search);
int manisdp_group_line_search(manisdp_group_t *g, double *alpha);
/* read-back of the integer index split r = j*n + i -> (i, j) of every stored entry of At, in CSC order, exactly as
 * the device kernels index with it (SURVEY 7 "index exactness"; BASELINE north_star: "A(YY') index handling must be
 * bit-exact").  *count = nnz(At); at most `cap` pairs are written; i / j may be NULL to query the count. */
int manisdp_get_index_split(manisdp_t *h, int64_t *i, int64_t *j, int64_t cap, int64_t *count);
/* host-only diagnostic (no GPU): eigen-decomposition of a small dense symmetric matrix A (n x n, row-major) with the
 * solver the eigen / rank steps use for their projected problems; w ascending, eigenvectors in the columns of V */
int manisdp_test_sym_eig(const double *A, int32_t n, double *w, double *V);


/* ---- multi-block handles (kind MANISDP_MULTIBLOCK; src/primal/ManiSDP_multiblock.m) ------------------------------
 * The blocks live on the device as ONE factor of N = sum(n_i) rows: block i owns rows [off_i, off_i + n_i) and the
 * leading p_i of the ld columns (the rest are zero and stay zero under every closure and manifold operation), so the
 * product-manifold operations of src/C-files/ projc.cpp, retrc.cpp, innerc.cpp, lincombc.cpp are the fused row kernels of the
 * single-block drivers with a per-row manifold switch, batched over all blocks in one launch.  cost / grad / hess /
 * tr_solve / line_search / set_dual / set_sigma / slot access work as for the other affine kinds (slots are N x pmax).
 * Host layout of a multi-block point ("cat"): the blocks one after the other, block i as n_i x p_i row-major -- the
 * memory image of the reference's p_i x n_i column-major cell Y{i}. */
int manisdp_mb_set_Y(manisdp_t *h, const double *Ycat, const int64_t *p);
int manisdp_mb_get_Y(manisdp_t *h, double *Ycat);     /* sum(n_i * p_i) doubles */
int manisdp_mb_get_widths(manisdp_t *h, int64_t *p);  /* nblocks entries */
/* M.rand() of multiblockmanifold.m:32-35 (randc.cpp:52-80): N(0,1) entries, unit rows on the first nob blocks */
int manisdp_mb_rand_Y(manisdp_t *h, const int64_t *p, uint64_t seed);
/* KKT step of ManiSDP_multiblock.m:66-97: obj, pinf, y <- y - sigma*(A x - b), by (with the z of the unit-diagonal blocks),
 * and eig(S{i}) of every block (batched, one eigen-decomposition per block); dinf = max_i dinfs[i].  dinfs / nneg
 * (number of negative eigenvalues of each block, uncapped) may be NULL.  The eigenvectors stay on the device for
 * manisdp_mb_update. */
int manisdp_mb_kkt(manisdp_t *h, int32_t update_dual, manisdp_kkt_info *out, double *dinfs, int32_t *nneg);
/* all eigenvalues of the dual slack of block `blk` from the last manisdp_mb_kkt (ascending, n_blk doubles) and,
 * if vecs != NULL, the eigenvectors (n_blk x n_blk, column k = k-th eigenvector, row-major) */
int manisdp_mb_get_block_eigs(manisdp_t *h, int32_t blk, double *vals, double *vecs);
/* rank cut + escape of every block, ManiSDP_multiblock.m:114-153: blocks with n_i >= min_facsize get
 * r_i = #{s >= theta*s_1} from the p_i x p_i Gram matrix (replaces svd(Y{i})), the rank-r_i truncation when r_i < p_i,
 * and nne_i escape directions (the lowest eigenvectors of S{i} kept by manisdp_mb_kkt): appended scaled by alpha and
 * (unit-diagonal blocks) renormalised when line_search = 0, staged in SLOT_U for manisdp_line_search when 1.
 * p_new (nblocks, may be NULL) receives the new widths. */
int manisdp_mb_update(manisdp_t *h, double theta, int32_t delta, double alpha, int32_t line_search,
                      int32_t min_facsize, int64_t *p_new);

/* ---- dual handles (kind MANISDP_DUAL_UNITDIAG; src/dual/ManiDSDP_unitdiag.m) ---------------------------------------
 * The factor Y (n x p, unit rows) is the factor of the DUAL slack S = Y Y'; cost / grad / hess / tr_solve / line_search /
 * rank_cut / escape work as on a UNITDIAG handle with the closures of ManiDSDP_unitdiag.m:171-191.  manisdp_kkt performs
 * the ADMM step of :71-89 (y = D^-1 A (vec(S) - c), x <- x - sigma*As, w <- w - sigma*Af) and the eigen step on the primal
 * matrix X = mat(x + bA) - diag(z); manisdp_get_dual returns that y, manisdp_set_dual / manisdp_set_sigma only change sigma.
 * The ADMM multipliers: x (n*n, column-major vec of the n x n matrix) and w (nfree). */
int manisdp_dual_get_state(manisdp_t *h, double *x, double *w);       /* either may be NULL */
int manisdp_dual_set_state(manisdp_t *h, const double *x, const double *w);

#ifdef __cplusplus
}
#endif
#endif /* MANISDP_B200_H */
