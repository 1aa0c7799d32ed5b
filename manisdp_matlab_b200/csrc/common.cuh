// common.cuh -- shared declarations of the B200 ManiSDP engine (internal; the public surface is include/manisdp_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/manisdp_b200.h"

#define MSDP_MAX_BLOCKS 2048   // upper bound on the grid of any reducing kernel (partials buffer rows)
#define MSDP_NQ 8              // reduction lanes per kernel (distinct scalars reduced in one pass)
#define MSDP_THREADS 256

enum { MF_OBLIQUE = 0, MF_SPHERE = 1, MF_EUCLID = 2 };
enum { MODE_NONE = 0, MODE_SPARSE = 1, MODE_DENSE = 2 };

// ---- device-resident scalar state of one trust-region solve ------------------------------------------------------
// Only the LAST block of a kernel (ticket pattern) writes it, after every block of that kernel has read what it
// needs, so kernels need no double-buffering of scalars and no host round trip (SURVEY 7 "hard parts": latency).
struct RtrState {
  // trust-region level (trustregions.m:441-767)
  double fx, gradnorm2, Delta, Delta_bar, rho_prime, rho_regularization;
  double fprop, gradnorm2_prop, rho, rhonum, rhoden, norm_eta;
  // tCG level (tCG.m:95-292)
  double z_r, d_Pd, e_Pd, e_Pe, e_Pe_new, model_value, norm_r0, r_r, alpha, beta, tau, d_Hd;
  double eta_g, eta_Heta;   // <eta, grad>, <eta, Hess eta> of the returned tCG iterate
  double y_r, y_d;          // sphere only: <Y, r>, <Y, mdelta>
  double kappa, theta;
  // per-point scalars of the affine closures: [pt] = value at point buffer pt
  double zsph[2];           // unittrace: z = <Y, eS Y>   (ManiSDP_unittrace.m:162)
  double cx[2];             // <C, YY'>
  double rr[2];             // |A(YY') - b - y/sigma|^2
  double sigma;
  double tmp[8];            // scratch scalars between kernels of one closure
  int pt;                   // which of the two point buffers holds the current (accepted) point
  int accepted, tr_iter;
  int stop, j, maxinner, mininner, branch, eta_cur;
  int hv_count;
  unsigned int ticket;
  // multi-block handles (multiblockmanifold.m:1-42): rows [0, nob_rows) are unit vectors (oblique blocks), the rows behind
  // them are Euclidean.  Every other oblique handle keeps LLONG_MAX here (all rows oblique).
  long long nob_rows;
  double dual_k0, dual_sigma;  // dual handles: constant of the cost and the ADMM penalty (read by the dual.cu kernels)
};

struct Csr {  // row lists on the device (int32 indices; n, nnz < 2^31)
  int *rowptr = nullptr, *col = nullptr;
  double* val = nullptr;
  int64_t nrows = 0, nnz = 0;
};

// constraint pattern of At for the sparse ("SDDMM / scatter-SpMM") path
struct ASparse {
  // by constraint k (CSC order of At): entry e in [kptr[k], kptr[k+1]) -> (ei[e], ej[e], ea[e])
  int *kptr = nullptr, *ei = nullptr, *ej = nullptr;
  double* ea = nullptr;
  // by row i: entry e in [rptr[i], rptr[i+1]) -> (rj[e], rk[e], ra[e])
  int *rptr = nullptr, *rj = nullptr, *rk = nullptr;
  double* ra = nullptr;
  int64_t nnz = 0;
  // long constraints (e.g. the trace row of a theta SDP: n entries in ONE constraint) are cut into segments of at most
  // SDDMM_SEG entries so that no row group walks thousands of dependent gathers alone: segment s covers the entries
  // [sptr[s], sptr[s+1]) and constraint k owns the segments [ksegs[k], ksegs[k+1]).  nseg == 0: no constraint is long.
  int *sptr = nullptr, *ksegs = nullptr;
  double* segval = nullptr;
  int64_t nseg = 0;
};
#define SDDMM_SEG 32

// constraint pattern of At for the dense path: A as CSR over k with linear indices into the n x n matrix, and its
// transpose as CSR over the linear index
struct ADense {
  int* kptr = nullptr;      // m+1
  int* klin = nullptr;      // nnz: lin = j*n + i  (int32 is enough: n <= 46340 on this path)
  double* ka = nullptr;
  // transpose restricted to the TOUCHED positions of the n x n matrix (dense S mode, either A mode):
  // position u -> linear index upos[u], entries [lptr[u], lptr[u+1]) -> (constraint lk[e], value la[e])
  int* upos = nullptr;      // nu
  int* lptr = nullptr;      // nu+1
  int* lk = nullptr;        // nnz
  double* la = nullptr;
  int64_t nnz = 0, nu = 0;
};

// dual (ManiDSDP) handles, dual.cu
struct DualData {
  int on = 0, dirty = 0;
  int64_t nfree = 0;
  double *bA = nullptr, *cpsd = nullptr, *x = nullptr;  // n x n: (A' D^-1 b), PSD part of c, ADMM multiplier
  double *isd = nullptr, *borig = nullptr;              // m: 1/sqrt(dAAt), the caller's b
  double* gram[2] = {nullptr, nullptr};                 // ld x ld: Y'Y of point buffer w
  double *gramS = nullptr, *gramW = nullptr, *gram_part = nullptr;  // scratch Gram (line search), Y'U, chunk partials
  int gram_ld = 0;
  size_t gram_part_cap = 0;
  int *B_jc = nullptr, *B_ir = nullptr;                 // free part of A (m x nfree, CSC)
  double *B_pr = nullptr, *cf = nullptr, *w = nullptr;  // its values, cost of the free variables, their multiplier
  double normc = 1.0;
};

struct manisdp_handle {
  int kind = 0, mf = 0, device = 0;
  int64_t n = 0, nloc = 0, m = 0, p = 0, ld = 0;
  int64_t row_begin = 0, row_end = 0;
  int rank = 0, world = 1;
  int s_mode = MODE_NONE, a_mode = MODE_NONE;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // problem data
  Csr C;                    // sparse C (row lists of the owned rows; columns are global)
  double* Cdense = nullptr; // n x n (dense S mode): C itself
  double* eS = nullptr;     // n x n (dense S mode): C + sigma*At*r at the current point / C - At*y in kkt
  double* Mbuf = nullptr;   // n x n scratch (dense A mode): Y U'
  double* Tbuf = nullptr;   // n x n (dense A mode): mat(At * w); untouched positions stay zero
  ASparse As;
  ADense Ad;
  double *b = nullptr, *y = nullptr;   // m
  double *resid[2] = {nullptr, nullptr}; // m: r = A(YY') - b - y/sigma at point buffer pt
  double *wU = nullptr, *wtmp = nullptr; // m
  double* y_kkt = nullptr;               // y used by the last kkt call (h->y, or h->wtmp when the dual was not updated)
  double sigma = 1.0;
  double normb = 0.0;
  // n x ld arrays
  double* Ybuf[2] = {nullptr, nullptr};
  double* Gbuf[2] = {nullptr, nullptr};
  double* eG[2] = {nullptr, nullptr};    // n: oblique row multipliers (eG / YeG) at point buffer pt
  double* eta[2] = {nullptr, nullptr};
  double *r = nullptr, *d = nullptr, *Hd = nullptr, *Uslot = nullptr, *Hslot = nullptr;
  double* gatherbuf = nullptr;           // sharded: n x ld all-gathered direction
  size_t cap_elems = 0;                  // capacity (elements) of each n x ld array
  // scalars
  RtrState* st = nullptr;                // device
  RtrState* st_host = nullptr;           // pinned
  double* partials = nullptr;            // MSDP_NQ x MSDP_MAX_BLOCKS
  int cache_valid = 0;                   // cost caches valid for Ybuf[pt]
  int grad_valid = 0;
  int pt = 0;                            // host mirror of st->pt
  // eig step
  double* eigvecs = nullptr;             // n x kcap (ROWS layout, ld = kld)
  double* eigvals_host = nullptr;
  int eig_k = 0, eig_kld = 0;
  double* zdiag = nullptr;               // n: z of the last kkt (diag shift of S)
  double zshift = 0.0;                   // unittrace: S = eS - z*I
  // log + stats
  std::vector<manisdp_tr_iter> log;
  int64_t hv_total = 0, launches = 0;
  int64_t y_version = 0;                 // bumped whenever the current point may have changed (rank-step cache key)
  int num_sms = 148;
  // CUDA graph cache for the tCG loop
  cudaGraphExec_t tcg_exec[2] = {nullptr, nullptr};  // one per value of pt (which buffer is the current point)
  cudaGraph_t tcg_graph[2] = {nullptr, nullptr};
  int64_t tcg_graph_p = -1;
  int tcg_graph_maxinner = -1;
  double graph_sigma = -1.0;
  int64_t graph_l_fixed = 0, graph_l_body = 0;  // kernels per graph launch: outside / inside the WHILE body
  // column-blocked SpMM (spmm.cu): pass pointers for the current operand width
  int spmm_ld = -1, spmm_B = 1;
  int* spmm_bptr = nullptr;
  size_t spmm_bptr_cap = 0;
  int spmm_use_bulk = 1;               // 1: cp.async.bulk gather kernel for ld >= 32 ; 0: register gathers only
  // block-major entry stream of C (spmm.cu: launch_bm) for the product on large graphs without locality: entries
  // stored column block by column block as (col, val, row), row-aligned chunks per warp, per-block partial rows
  int bm_mode = 1;                     // MANISDP_SPMM_BM: 0 off, 1 auto (no locality, operand >= 2x L2; default), 2 always
  int bm_B = 0;                        // number of column blocks (0: format not built)
  int *bm_col = nullptr, *bm_row = nullptr, *bm_chunk = nullptr;
  double* bm_val = nullptr;
  std::vector<int> bm_chunk_off;       // first chunk of each block (B + 1)
  double* bm_part = nullptr;           // B x nloc x ld partial rows
  size_t bm_part_cap = 0;
  int64_t bm_part_ld = -1;             // row length the (zero-initialised) partial buffers are laid out for
  int spmm_narrow = 1;                // 1: k_spmm_narrow for ld <= 32 (MANISDP_SPMM_NARROW=0: generic kernel everywhere)
  int spmm_block_mode = 0;             // 0 never (default), 1 auto (only for matrices without locality), 2 always
  int64_t spmm_l2_target = 64ll << 20; // bytes of operand rows per column block
  int C_sorted = 0;                    // rows of C are column-sorted
  int C_lowdeg = 0;                    // every 32-row batch of C has <= 320 entries and the mean degree is <= 8
  int C_maxdeg = 0;                    // largest number of stored entries in a row of C
  int spmm_lowdeg = 1;                 // 1: batched low-degree kernel when C_lowdeg (MANISDP_SPMM_LOWDEG=0: off)
  double C_far_fraction = 0.0;         // share of entries whose column is farther than an L2 window from the row
  // split-K workspace of the DMMA GEMM (gemm_f64.cu)
  double* gemm_ws = nullptr;
  size_t gemm_ws_cap = 0;
  // NCCL
  void* nccl_comm = nullptr;
  // pipelined exchange (dist.cu / spmm.cu): second communicator + stream, one event per stage, owner-chunk pass pointers
  void* nccl_comm2 = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr;
  std::vector<cudaEvent_t> ev_stage;
  int pipeline = 0;                     // 0 off, 1 NCCL point-to-point stages, 2 peer-memory (CUDA IPC) copy stages
  int* owner_bptr = nullptr;
  int ipc_ready = 0;                    // all ranks mapped all peers (decided collectively in msdp_dist_ipc_refresh)
  void* ipc_dev = nullptr;              // device scratch: world x 128 bytes of IPC handles + one barrier double
  std::vector<double*> peer_d, peer_u;  // peers' direction array / SLOT_U mapped into this process
  std::vector<double*> peer_y[2];       // peers' point buffers
  double** peer_tab_dev = nullptr;      // device copy of the four tables (d, SLOT_U, Ybuf[0], Ybuf[1]), G pointers each
  const double* const* cg_peer_tab = nullptr;  // set around a cost+grad product that gathers from peers in place
  double C_remote_fraction = 1.0;       // share of the shard's entries whose column is owned by another rank
  double peer_gather_max_remote = 0.05; // direct peer gathers only below this share (MANISDP_PEER_GATHER_MAX)
  // column-sharded layout (colshard.cu): every rank holds all rows; split = only pl = ceil(p/G) columns of the factor
  int col_mode = 0, cworld = 1, crank = 0;
  int col_split = 0;                    // 1 between manisdp_col_split and manisdp_col_merge
  double last_dinf = 1e-3;              // dinf of the previous kkt call (adaptive eigen-step tolerance)
  int last_numinner = 0;                // inner iterations of the previous TR iteration (chunking of the stream loop)
  int col_graph = 0;                    // world > 1: capture the split tCG loop (with its all-reduces) in a CUDA graph
  int col_pfull = 0;                    // width of the factor at the last split
  void* col_comm = nullptr;             // ncclComm_t
  double* col_rowvec = nullptr;         // n: per-row partial sums on their way through the all-reduce
  double* col_pack = nullptr;           // 8 + 2n: update packet [6 scalars, pad | rowsum(Y.*r') | rowsum(Y.*mdelta)]
  // grow-only device scratch buffers of the outer-loop steps (Gram partials, combination staging): the rank step and the
  // escape step run once per outer iteration and used to pay two to four cudaMalloc / cudaFree pairs each
  void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t scratch_cap[4] = {0, 0, 0, 0};
  // multi-block handles (multiblock.cu): block orders, row offsets (prefix sums of n_i), offsets into the stacked vec
  // (prefix sums of n_i^2), current widths; n = mb_roff.back(); rows [0, mb_nob_rows) are unit vectors
  std::vector<int64_t> mb_n, mb_roff, mb_off2, mb_p;
  int mb_nob = 0;
  int64_t mb_nob_rows = 0;
  int* mb_rowblk = nullptr;              // N: block of each row (device)
  int* mb_roff_dev = nullptr;            // t + 1: row offsets (device)
  int* mb_pw = nullptr;                  // 4t ints: current widths, then work vectors of mb_update (device)
  int mb_eig_device = 0;                 // 1: batched Jacobi on the device for eig(S{i}) (MANISDP_MB_EIG=device at create)
  int mb_have_eigs = 0;                  // mb_evals / mb_evecs belong to the current point
  // device: S{i} of the last mb_kkt, overwritten by their eigenvectors (block i at mb_off2[i], row-major n_i x n_i, column k =
  // k-th vector); Jacobi scratch; eigenvalues
  double *mb_S = nullptr, *mb_V = nullptr, *mb_w = nullptr;
  int64_t* mb_off2_dev = nullptr;        // t + 1: offsets of the blocks in mb_S
  int *mb_n_dev = nullptr, *mb_sweeps = nullptr;  // t: block orders, Jacobi sweeps used
  std::vector<double> mb_evals;          // eigenvalues of every S{i} of the last mb_kkt, stacked by block (ascending)
  DualData dual;
  int rank_strict = 0;                   // rank estimate counts e > theta*e1 (ManiDSDP_unitdiag.m:91) instead of >=
  std::string err;
};
int msdp_scratch(manisdp_handle* h, int slot, size_t bytes, void** out);  // api.cu
// multiblock.cu
int msdp_mb_setup(manisdp_handle* h, const manisdp_problem* pb);
void msdp_mb_free(manisdp_handle* h);
double msdp_mb_typicaldist(const manisdp_handle* h);
double msdp_mb_dim(const manisdp_handle* h);
// jacobi.cu: batched symmetric eigen-decomposition, one CTA per matrix (eigenvalues ascending in w, eigenvectors in A)
int msdp_jacobi_batched(manisdp_handle* h, double* A, double* V, double* w, const int64_t* off_dev, const int* woff_dev,
                        const int* n_dev, int* sweeps_dev, int count, int max_n);

// ---- tracing: one NVTX range per phase of the hot path (visible in Nsight Systems / ncu --nvtx; no cost when no
// tool is attached: NVTX v3 is header-only and resolves its injection library lazily) -----------------------------
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// ---- error plumbing ---------------------------------------------------------------------------------------------
int msdp_fail(manisdp_handle* h, int code, const std::string& msg);
#define CUDA_TRY(h, expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return msdp_fail(h, MANISDP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)
#define MSDP_TRY(expr)          \
  do {                          \
    int _s = (expr);            \
    if (_s != MANISDP_OK) return _s; \
  } while (0)
#define KERNEL_CHECK(h)                                                                                \
  do {                                                                                                 \
    (h)->launches++;                                                                                   \
    cudaError_t _e = cudaPeekAtLastError();                                                            \
    if (_e != cudaSuccess)                                                                             \
      return msdp_fail(h, MANISDP_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e));    \
  } while (0)

// ---- row-group geometry -----------------------------------------------------------------------------------------
// A row of ld doubles (ld % 4 == 0) is handled by GS lanes, each holding VPL double2 vectors: vector index
// v = lane + GS*t.  GS in {2,4,8,16,32}, VPL in {1,2,4,8,16}  =>  ld <= 1024 for the affine / dual kinds (the VPL = 16
// instances spill registers and exist for completeness: BQP d >= 120 needs p > 512); the MaxCut kernels of spmm.cu and
// the multi-block tables stop at ld = 512.
struct RowGeom {
  int gs, vpl;
};
inline RowGeom row_geom(int64_t ld) {
  int nvec = (int)(ld / 2);
  RowGeom g{2, 1};
  while (g.gs < 32 && g.gs < nvec) g.gs *= 2;
  while (g.gs * g.vpl < nvec) g.vpl *= 2;
  return g;
}
#define MSDP_MAX_LD 512          // ONLYUNITDIAG (spmm.cu kernels), multi-block
#define MSDP_MAX_LD_AFFINE 1024  // UNITDIAG / UNITTRACE / GENERAL / DUAL_UNITDIAG

#define DISPATCH_GEOM(geom, ...)                                            \
  do {                                                                      \
    const RowGeom _g = (geom);                                              \
    if (_g.vpl == 1) {                                                      \
      switch (_g.gs) {                                                      \
        case 2: { constexpr int GS = 2, VPL = 1; __VA_ARGS__; } break;      \
        case 4: { constexpr int GS = 4, VPL = 1; __VA_ARGS__; } break;      \
        case 8: { constexpr int GS = 8, VPL = 1; __VA_ARGS__; } break;      \
        case 16: { constexpr int GS = 16, VPL = 1; __VA_ARGS__; } break;    \
        default: { constexpr int GS = 32, VPL = 1; __VA_ARGS__; } break;    \
      }                                                                     \
    } else if (_g.vpl == 2) { constexpr int GS = 32, VPL = 2; __VA_ARGS__; } \
    else if (_g.vpl == 4) { constexpr int GS = 32, VPL = 4; __VA_ARGS__; }  \
    else if (_g.vpl == 8) { constexpr int GS = 32, VPL = 8; __VA_ARGS__; }  \
    else { constexpr int GS = 32, VPL = 16; __VA_ARGS__; }                  \
  } while (0)

// grid for a row-group kernel over nrows rows
inline int rows_grid(const manisdp_handle* h, int64_t nrows, int gs, int blocks_per_sm = 8) {
  int64_t rows_per_block = (MSDP_THREADS / gs);
  int64_t nb = (nrows + rows_per_block - 1) / rows_per_block;
  int64_t cap = (int64_t)h->num_sms * blocks_per_sm;
  if (cap > MSDP_MAX_BLOCKS) cap = MSDP_MAX_BLOCKS;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

// ---- cross-file entry points ------------------------------------------------------------------------------------
// closures.cu
// which >= 0: explicit point buffer; -1: proposal (st->pt^1, chosen on the device); -2: current point (st->pt)
int msdp_costgrad(manisdp_handle* h, int which, int cg_mode);
// Hout = Hess f(Y)[D]; tail_mode != TAIL_NONE: inside tCG (point chosen on the device, scalars updated)
int msdp_hess_dir(manisdp_handle* h, const double* D, double* Hout, int tail_mode);
// cost of an arbitrary n x ld array Z without touching the closure caches (the reference's `co`, line search)
int msdp_eval_cost_only(manisdp_handle* h, const double* Z, double* f_host);
void msdp_invalidate_graph(manisdp_handle* h);
// rtr.cu
int msdp_tr_solve(manisdp_handle* h, const manisdp_tr_options* o, manisdp_tr_info* info);
// eig.cu
int msdp_kkt(manisdp_handle* h, int delta, double eig_tol, int update_dual, manisdp_kkt_info* out);
int msdp_rank_cut(manisdp_handle* h, double theta, int apply, int64_t* r, int64_t* pnew);
int msdp_escape(manisdp_handle* h, int nne, double alpha, int line_search);
int msdp_line_search(manisdp_handle* h, double* alpha_out);
void msdp_eig_release(manisdp_handle* h);
extern "C" int msdp_ensure_costgrad(manisdp_handle* h);  // cost + gradient caches at the current point (one sync)
// api.cu helpers
int msdp_resize(manisdp_handle* h, int64_t p);
int msdp_sync_state(manisdp_handle* h);  // device RtrState -> st_host (synchronises the stream)
