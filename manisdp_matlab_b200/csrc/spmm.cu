// spmm.cu -- K1: CSR SpMM with fused row epilogues, the closures of ManiSDP_onlyunitdiag.m:117-130.
//
//   hess      (:128-129)  eH = U*C ; H = eH - Y.*sum(Y.*eH) - U.*eG          -> + <U,H> for tCG.m:166
//   cost+grad (:118-124)  YC = Y*C ; eG = sum(YC.*Y) ; f = .5*sum(eG) ; G = YC - Y.*eG   -> + |G|^2
//   S*V       (:49)       (C - diag(z)) * V for the eigen step
//
// Layout: the factor is vertex-major (row i = the p numbers of vertex i, ld = 4*ceil(p/4) doubles = the reference's
// p x n column-major array), so one gathered operand row is one contiguous ld*8-byte segment.  The row epilogue
// (projection onto the tangent space of the oblique manifold, ManiSDP_onlyunitdiag.m:138-139) needs only the finished
// row, so it is fused and H is written exactly once.
//
// Two inner loops:
//   k_spmm_bulk (ld >= 32, one warp per row): operand rows are fetched with cp.async.bulk (the TMA engine's 1-D bulk
//     copy, SASS UBLKCP) into a per-warp shared-memory ring of 2 x SLOTS rows completed through mbarriers, so the
//     gathers in flight are bounded by shared memory (192 KB/SM = 384 rows of 512 B) instead of registers; the next
//     batch is issued before the current one is consumed and the row pointers are read one row ahead.
//   k_spmm (any ld, GS lanes per row): register gathers, four in flight per group; used for thin blocks (eigen step,
//     small p) where a full warp per row would idle.
//
// Round-1 measurements that shaped this (profiles/): on the n = 1e6 Erdos-Renyi instance the product needs
// nnz*p*8 = 25 GB of gathers against 2.1 GB of compulsory bytes and the operand (512 MB) is 4x the L2, so ncu shows
// 22.6 GB of DRAM reads per launch; random 512 B gathers reach 7.7 TB/s from a 512 MB table and 19 TB/s from an
// L2-resident one (r1_gather_microbench.jsonl).  COLUMN PASSES (B > 1: pass b adds the entries with col in block b to
// the partial rows kept in `out`, the last pass applies the epilogue) cut the DRAM traffic to ~10 GB but are latency
// bound per pass in this form (tools/spmm_lab.cu: best 3.18 ms at B = 4 vs 3.33 ms at B = 1), so they stay opt-in
// (MANISDP_SPMM_BLOCK) until the per-pass pipeline is deeper.
//
// Algorithmic bytes per product: 12*nnz + 4*(n+1) + 24*n*p + 8*n.
#include <algorithm>
#include <vector>
#include "dist.h"
#include "rowops.cuh"
#include "scalar_logic.cuh"
#include "kernels.cuh"

enum { EPI_HESS = 0, EPI_COSTGRAD = 1, EPI_SHIFT = 2 };

struct SpmmArgs {
  const int* col;
  const double* val;
  const int* bptr0;    // per row: first entry of this pass
  const int* bptr1;    // per row: one past the last entry of this pass
  int first, last;     // first pass: accumulators start at 0 ; last pass: epilogue + reductions
  int64_t nrows;
  int ld;
  int slots;           // bulk kernel: rows per ring buffer
  const double* Ug;    // gather source, indexed by GLOBAL row
  const double* Uown;  // the same array restricted to the owned rows (local row index)
  const double* Y;     // point, owned rows
  const double* eG;    // row multipliers (eG of the point, or z for EPI_SHIFT)
  double* out;
  double* eGout;
  VecPtrs v;
  int sel;             // 0: pointers above are final; 1: hess inside tCG (select by st->pt); 2: cost at proposal
                       // (pt^1); 3: cost at current (pt)
  int sharded;         // 1: Ug is the all-gathered array, fixed
  RtrState* st;
  double* partials;
  int mode;            // tail_mode (EPI_HESS) or cg_mode (EPI_COSTGRAD)
  // direct peer gathers (row-sharded, graphs with locality): operand row c lives at peer_tab[c / rpr] + (c % rpr)*ld,
  // where peer_tab[q] is rank q's operand array mapped through CUDA IPC (own entry = the local array)
  const double* const* peer_tab;
  int rpr;
  int row0;            // global index of local row 0 (set by launch_pass)
};

struct SpmmPtrs {
  const double *Y, *eG, *Uown, *Ug;
  double *out, *eGout;
};

// device-side buffer selection (no host round trip between TR iterations)
__device__ __forceinline__ SpmmPtrs select_ptrs(const SpmmArgs& a) {
  SpmmPtrs p{a.Y, a.eG, a.Uown, a.Ug, a.out, a.eGout};
  if (a.sel == 1) {
    const int pt = a.st->pt;
    p.Y = pt ? a.v.Y1 : a.v.Y0;
    p.eG = pt ? a.v.eG1 : a.v.eG0;
  } else if (a.sel >= 2) {
    const int w = (a.sel == 2) ? (a.st->pt ^ 1) : a.st->pt;
    p.Uown = w ? a.v.Y1 : a.v.Y0;
    if (!a.sharded) p.Ug = p.Uown;
    p.out = w ? a.v.G1 : a.v.G0;
    p.eGout = w ? a.v.eG1 : a.v.eG0;
  }
  return p;
}

__device__ __forceinline__ double2 ldcs2(const double* p) { return __ldcs(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stcs2(double* p, double2 v) { __stcs(reinterpret_cast<double2*>(p), v); }

// fused row epilogue on the finished accumulators of one row (lane holds vectors gl + GS*t)
template <int GS, int VPL, int EPI>
__device__ __forceinline__ void row_epilogue(double2 (&acc)[VPL], const SpmmPtrs& p, int64_t row, int ld, int gl,
                                             unsigned mask, double (&q)[2]) {
  const int nvec = ld / 2;
  const size_t rb = (size_t)row * ld;
  if (EPI == EPI_HESS) {
    double2 y[VPL], u[VPL];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = gl + GS * t;
      if (cv < nvec) {
        y[t] = ld2(p.Y + rb + 2 * cv);
        u[t] = ld2(p.Uown + rb + 2 * cv);
        dot += y[t].x * acc[t].x + y[t].y * acc[t].y;
      }
    }
    dot = group_sum<GS>(dot, mask);  // sum(Y.*eH), :129
    const double eg = p.eG[row];
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = gl + GS * t;
      if (cv < nvec) {
        double2 hv;
        hv.x = acc[t].x - y[t].x * dot - u[t].x * eg;
        hv.y = acc[t].y - y[t].y * dot - u[t].y * eg;
        st2(p.out + rb + 2 * cv, hv);
        q[0] += u[t].x * hv.x + u[t].y * hv.y;  // <mdelta, Hmdelta>, tCG.m:166
      }
    }
  } else if (EPI == EPI_COSTGRAD) {
    double2 y[VPL];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = gl + GS * t;
      if (cv < nvec) {
        y[t] = ld2(p.Uown + rb + 2 * cv);
        dot += y[t].x * acc[t].x + y[t].y * acc[t].y;
      }
    }
    dot = group_sum<GS>(dot, mask);  // eG(row) = sum(YC.*Y), :119
    if (gl == 0) {
      p.eGout[row] = dot;
      q[0] += dot;
    }
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = gl + GS * t;
      if (cv < nvec) {
        double2 g;
        g.x = acc[t].x - y[t].x * dot;  // G = YC - Y.*eG, :124
        g.y = acc[t].y - y[t].y * dot;
        st2(p.out + rb + 2 * cv, g);
        q[1] += g.x * g.x + g.y * g.y;
      }
    }
  } else {  // EPI_SHIFT: out = C*V - z.*V
    const double z = p.eG ? p.eG[row] : 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = gl + GS * t;
      if (cv < nvec) {
        const double2 u = ld2(p.Uown + rb + 2 * cv);
        double2 o;
        o.x = acc[t].x - z * u.x;
        o.y = acc[t].y - z * u.y;
        st2(p.out + rb + 2 * cv, o);
      }
    }
  }
}

// grid reductions + scalar logic of the last pass
template <int EPI>
__device__ __forceinline__ void spmm_tail(const SpmmArgs& a, double (&q)[2], double* sm) {
  RtrState* st = a.st;
  if (EPI == EPI_HESS) {
    if (a.mode == TAIL_NONE) return;
    double tot[1], q1[1] = {q[0]};
    __syncwarp();
    if (grid_sum_last<1>(q1, a.partials, &st->ticket, sm, tot)) {
      if (threadIdx.x == 0) {
        if (a.mode == TAIL_TCG)
          tcg_after_hv(st, tot[0]);
        else
          st->tmp[0] = tot[0];
      }
    }
  } else if (EPI == EPI_COSTGRAD) {
    double tot[2];
    __syncwarp();
    if (grid_sum_last<2>(q, a.partials, &st->ticket, sm, tot)) {
      if (threadIdx.x == 0) {
        const double f = 0.5 * tot[0];  // :120
        st->tmp[0] = f;
        st->tmp[1] = tot[1];
        if (a.mode == CG_INIT) {
          st->fx = f;
          st->gradnorm2 = tot[1];
        } else if (a.mode == CG_TR) {
          st->fprop = f;
          st->gradnorm2_prop = tot[1];
          tr_decide(st);
        } else if (a.mode == CG_TR_DEFER) {
          st->tmp[0] = tot[0];  // un-halved local sum; the scalar kernel halves after the all-reduce
        }
      }
    }
  }
}

// ---- register-gather kernel (any ld) --------------------------------------------------------------------------------
template <int GS, int VPL, int EPI, bool PEER>
__global__ void __launch_bounds__(MSDP_THREADS) k_spmm(const SpmmArgs a) {
  __shared__ double sm[2 * 32];
  if ((EPI == EPI_HESS || EPI == EPI_SHIFT) && a.mode != TAIL_NONE && a.st->stop != 0) return;
  const SpmmPtrs p = select_ptrs(a);
  const int* __restrict__ col = a.col;
  const double* __restrict__ val = a.val;
  const double* __restrict__ Ug = p.Ug;
  const int ld = a.ld, nvec = ld / 2;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const bool first = a.first != 0, last = a.last != 0;
  double q[2] = {0.0, 0.0};
  // where operand row c lives: the local / gathered array, or (PEER) straight in the owner's HBM over NVLink
  auto operand_row = [&](int c) -> const double* {
    if (PEER) {
      const int owner = c / a.rpr;
      return a.peer_tab[owner] + (size_t)(c - owner * a.rpr) * ld;
    }
    return Ug + (size_t)c * ld;
  };

  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < a.nrows; row += ngroups) {
    const int e0 = __ldg(a.bptr0 + row), e1 = __ldg(a.bptr1 + row);
    const size_t rb = (size_t)row * ld;
    double2 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = gl + GS * t;
      acc[t] = (!first && cv < nvec) ? ldcs2(p.out + rb + 2 * cv) : make_double2(0.0, 0.0);
    }
    for (int base = e0; base < e1; base += GS) {
      const int e = base + gl;
      int c = 0;
      double w = 0.0;
      if (e < e1) {
        c = __ldg(col + e);
        w = __ldg(val + e);
      }
      const int cnt = min(GS, e1 - base);
      int k = 0;
      for (; k + 4 <= cnt; k += 4) {
        const int c0 = __shfl_sync(mask, c, k, GS), c1 = __shfl_sync(mask, c, k + 1, GS),
                  c2 = __shfl_sync(mask, c, k + 2, GS), c3 = __shfl_sync(mask, c, k + 3, GS);
        const double w0 = __shfl_sync(mask, w, k, GS), w1 = __shfl_sync(mask, w, k + 1, GS),
                     w2 = __shfl_sync(mask, w, k + 2, GS), w3 = __shfl_sync(mask, w, k + 3, GS);
        const double* p0 = operand_row(c0);
        const double* p1 = operand_row(c1);
        const double* p2 = operand_row(c2);
        const double* p3 = operand_row(c3);
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int cv = gl + GS * t;
          if (cv < nvec) {
            const double2 u0 = ldg2(p0 + 2 * cv), u1 = ldg2(p1 + 2 * cv), u2 = ldg2(p2 + 2 * cv),
                          u3 = ldg2(p3 + 2 * cv);
            acc[t].x = fma(w0, u0.x, acc[t].x);
            acc[t].y = fma(w0, u0.y, acc[t].y);
            acc[t].x = fma(w1, u1.x, acc[t].x);
            acc[t].y = fma(w1, u1.y, acc[t].y);
            acc[t].x = fma(w2, u2.x, acc[t].x);
            acc[t].y = fma(w2, u2.y, acc[t].y);
            acc[t].x = fma(w3, u3.x, acc[t].x);
            acc[t].y = fma(w3, u3.y, acc[t].y);
          }
        }
      }
      for (; k < cnt; ++k) {
        const int c0 = __shfl_sync(mask, c, k, GS);
        const double w0 = __shfl_sync(mask, w, k, GS);
        const double* p0 = operand_row(c0);
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int cv = gl + GS * t;
          if (cv < nvec) {
            const double2 u0 = ldg2(p0 + 2 * cv);
            acc[t].x = fma(w0, u0.x, acc[t].x);
            acc[t].y = fma(w0, u0.y, acc[t].y);
          }
        }
      }
    }
    if (!last) {  // partial rows of an inner pass: streamed back, re-read by the next pass
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int cv = gl + GS * t;
        if (cv < nvec) stcs2(p.out + rb + 2 * cv, acc[t]);
      }
      continue;
    }
    row_epilogue<GS, VPL, EPI>(acc, p, row, ld, gl, mask, q);
  }
  if (last) spmm_tail<EPI>(a, q, sm);
}

// ---- register-gather kernel for narrow rows (ld <= 32: one double2 per lane, 2..16 lanes per row) ---------------------
// Same arithmetic, entry order and results as k_spmm.  With few lanes per row a row is a chain of short dependent
// rounds (indices -> gathers -> epilogue operands -> store), so the chain is shortened instead of widened
// (tools/spmm_narrow_lab.cu, profiles/r1_spmm_narrow_lab.txt; 1e6 rows of 1 + Poisson(48) entries):
//   * the indices of the next round are fetched before the gathers of the current one;
//   * the tail of a round is padded to the unroll width with (own row, weight 0) entries -- the one-by-one remainder
//     loop paid a full memory latency per entry (own row: the line the epilogue reads anyway, and never remote);
//   * the epilogue operands (Y row, U row, multiplier) are requested at the start of the row.
//   ld = 8: 0.573 -> 0.437 ms, ld = 16: 0.979 -> 0.808 ms, ld = 32: 1.761 -> 1.680 ms.  At ld = 64 (32 lanes per row,
//   the bench configuration) the gain is inside the noise (3.60 -> 3.54 ms), so k_spmm stays as it is.
template <int GS, int EPI, bool PEER>
__global__ void __launch_bounds__(MSDP_THREADS, 4) k_spmm_narrow(const SpmmArgs a) {
  __shared__ double sm[2 * 32];
  if ((EPI == EPI_HESS || EPI == EPI_SHIFT) && a.mode != TAIL_NONE && a.st->stop != 0) return;
  const SpmmPtrs p = select_ptrs(a);
  const int* __restrict__ col = a.col;
  const double* __restrict__ val = a.val;
  const double* __restrict__ Ug = p.Ug;
  const int ld = a.ld;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const bool act = gl < ld / 2;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const bool first = a.first != 0, last = a.last != 0;
  double q[2] = {0.0, 0.0};
  auto operand_row = [&](int c) -> const double* {
    if (PEER) {
      const int owner = c / a.rpr;
      return a.peer_tab[owner] + (size_t)(c - owner * a.rpr) * ld;
    }
    return Ug + (size_t)c * ld;
  };

  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < a.nrows; row += ngroups) {
    const int e0 = __ldg(a.bptr0 + row), e1 = __ldg(a.bptr1 + row);
    const size_t off = (size_t)row * ld + 2 * gl;
    double2 y = make_double2(0.0, 0.0), u = make_double2(0.0, 0.0);
    double eg = 0.0;
    if (last) {
      if (EPI == EPI_HESS) {
        if (act) {
          y = ld2(p.Y + off);
          u = ld2(p.Uown + off);
        }
        eg = p.eG[row];
      } else if (EPI == EPI_COSTGRAD) {
        if (act) y = ld2(p.Uown + off);
      } else {
        if (act) u = ld2(p.Uown + off);
        eg = p.eG ? p.eG[row] : 0.0;
      }
    }
    double2 acc = (!first && act) ? ldcs2(p.out + off) : make_double2(0.0, 0.0);
    const int padc = a.row0 + (int)row;
    int c = padc;
    double w = 0.0;
    if (e0 + gl < e1) {
      c = __ldg(col + e0 + gl);
      w = __ldg(val + e0 + gl);
    }
    for (int base = e0; base < e1; base += GS) {
      int cn = padc;
      double wn = 0.0;
      const int en = base + GS + gl;
      if (en < e1) {
        cn = __ldg(col + en);
        wn = __ldg(val + en);
      }
      int cnt = min(GS, e1 - base);
      if (GS >= 4) cnt = (cnt + 3) & ~3;  // <= GS; the extra lanes hold (own row, 0)
      int k = 0;
      for (; k + 4 <= cnt; k += 4) {
        const int c0 = __shfl_sync(mask, c, k, GS), c1 = __shfl_sync(mask, c, k + 1, GS),
                  c2 = __shfl_sync(mask, c, k + 2, GS), c3 = __shfl_sync(mask, c, k + 3, GS);
        const double w0 = __shfl_sync(mask, w, k, GS), w1 = __shfl_sync(mask, w, k + 1, GS),
                     w2 = __shfl_sync(mask, w, k + 2, GS), w3 = __shfl_sync(mask, w, k + 3, GS);
        const double* p0 = operand_row(c0);
        const double* p1 = operand_row(c1);
        const double* p2 = operand_row(c2);
        const double* p3 = operand_row(c3);
        if (act) {
          const double2 u0 = ldg2(p0 + 2 * gl), u1 = ldg2(p1 + 2 * gl), u2 = ldg2(p2 + 2 * gl), u3 = ldg2(p3 + 2 * gl);
          acc.x = fma(w0, u0.x, acc.x);
          acc.y = fma(w0, u0.y, acc.y);
          acc.x = fma(w1, u1.x, acc.x);
          acc.y = fma(w1, u1.y, acc.y);
          acc.x = fma(w2, u2.x, acc.x);
          acc.y = fma(w2, u2.y, acc.y);
          acc.x = fma(w3, u3.x, acc.x);
          acc.y = fma(w3, u3.y, acc.y);
        }
      }
      for (; k < cnt; ++k) {  // GS == 2 only
        const int c0 = __shfl_sync(mask, c, k, GS);
        const double w0 = __shfl_sync(mask, w, k, GS);
        if (act) {
          const double2 u0 = ldg2(operand_row(c0) + 2 * gl);
          acc.x = fma(w0, u0.x, acc.x);
          acc.y = fma(w0, u0.y, acc.y);
        }
      }
      c = cn;
      w = wn;
    }
    if (!last) {
      if (act) stcs2(p.out + off, acc);
      continue;
    }
    if (EPI == EPI_HESS) {
      const double dot = group_sum<GS>(act ? y.x * acc.x + y.y * acc.y : 0.0, mask);  // sum(Y.*eH), :129
      if (act) {
        double2 hv;
        hv.x = acc.x - y.x * dot - u.x * eg;
        hv.y = acc.y - y.y * dot - u.y * eg;
        st2(p.out + off, hv);
        q[0] += u.x * hv.x + u.y * hv.y;  // <mdelta, Hmdelta>, tCG.m:166
      }
    } else if (EPI == EPI_COSTGRAD) {
      const double dot = group_sum<GS>(act ? y.x * acc.x + y.y * acc.y : 0.0, mask);  // eG(row) = sum(YC.*Y), :119
      if (gl == 0) {
        p.eGout[row] = dot;
        q[0] += dot;
      }
      if (act) {
        double2 g;
        g.x = acc.x - y.x * dot;  // G = YC - Y.*eG, :124
        g.y = acc.y - y.y * dot;
        st2(p.out + off, g);
        q[1] += g.x * g.x + g.y * g.y;
      }
    } else if (act) {  // EPI_SHIFT: out = C*V - z.*V
      double2 o;
      o.x = acc.x - eg * u.x;
      o.y = acc.y - eg * u.y;
      st2(p.out + off, o);
    }
  }
  if (last) spmm_tail<EPI>(a, q, sm);
}

// ---- batched kernel for LOW-DEGREE rows (32 < ld <= 64: one warp per row, one double2 per lane) ------------------------
// On the toroidal-grid profile (G11/G32/G81 shape: 5 entries per row) 96 % of the algorithmic bytes are plain streams,
// yet k_spmm reached only 0.40 of the HBM roofline: per row a warp walked a chain of dependent round trips -- row
// pointers -> (col, val) -> four gathers -> the remaining gather -> epilogue operands -> store -- and nothing of the
// next row was in flight meanwhile (round-1 VERDICT "what's weak" 4).  Here a warp takes a BATCH of 32 consecutive rows:
//   * the 33 row pointers and then all (col, val) pairs of the batch (contiguous in the CSR arrays) are fetched with
//     coalesced loads and staged in shared memory -- two round trips per 32 rows instead of two per row;
//   * per row, the epilogue operands (Y row, U row, multiplier) and up to LB_GW gathers are issued together, the tail
//     of a row padded with (own row, weight 0) entries so that the loads stay unconditional (the own row is the line
//     the epilogue reads anyway);
//   * Y is read and H written with streaming (evict-first) hints: the L2 is kept for the operand rows, which the
//     neighbouring rows of the grid re-use;
//   * the gather width GW is a template parameter chosen from the largest row degree of C (5 on the torus: four
//     neighbours + the diagonal): no padded gathers, 20 instead of 32 gather registers, four resident blocks per SM;
//   * (tried and removed: `prefetch.global.L2` of the lines row r + pf will read -- measured counter-productive on
//     B200, torus n = 1e6, p = 64: 0.44 ms without, 0.52-0.55 ms with pf = 2..16, profiles/r2_sweep_torus_pf.txt; the
//     ncu capture profiles/r2_ncu_lowdeg_torus_p64.md shows why: the kernel issues 219 warp instructions per row, half
//     of them address arithmetic, so extra instructions cost more than the latency they hide.)
// Same entry order as k_spmm (which keeps several partial accumulators per row, so rows agree to rounding).  Chosen by launch_pass when every
// 32-row batch of C has at most LB_CAP entries and the mean degree is <= 8 (api.cu: C_lowdeg).
#define LB_CAP 320
#define LB_GW 8
template <int EPI, bool PEER, int GW>
__global__ void __launch_bounds__(MSDP_THREADS, GW <= 6 ? 4 : 3) k_spmm_lowdeg(const SpmmArgs a) {
  constexpr int CAP = GW < 8 ? 32 * GW : LB_CAP;  // GW < 8: every row has <= GW entries (api.cu: C_maxdeg)
  __shared__ double sm[2 * 32];
  __shared__ int s_col[MSDP_THREADS / 32][CAP];
  __shared__ double s_val[MSDP_THREADS / 32][CAP];
  if ((EPI == EPI_HESS || EPI == EPI_SHIFT) && a.mode != TAIL_NONE && a.st->stop != 0) return;
  const SpmmPtrs p = select_ptrs(a);
  const int* __restrict__ col = a.col;
  const double* __restrict__ val = a.val;
  const double* __restrict__ Ug = p.Ug;
  const int ld = a.ld;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool act = lane < ld / 2;
  const int lo = act ? 2 * lane : 0;  // inactive lanes (ld < 64) shadow vector 0 and never store
  int* sc = s_col[wid];
  double* sv = s_val[wid];
  // gather address = per-lane base + column * row bytes: one 32 x 32 -> 64-bit multiply-add per gather
  const char* const Ugl = reinterpret_cast<const char*>(Ug + lo);
  const unsigned rowbytes = (unsigned)ld * 8u;
  auto operand_row = [&](int c) -> const double* {
    if (PEER) {
      const int owner = c / a.rpr;
      return a.peer_tab[owner] + (size_t)(c - owner * a.rpr) * ld + lo;
    }
    return reinterpret_cast<const double*>(Ugl + (size_t)(unsigned)c * rowbytes);
  };
  const int64_t nbatches = (a.nrows + 31) / 32;
  const int64_t nw = (int64_t)gridDim.x * (MSDP_THREADS / 32);
  double q[2] = {0.0, 0.0};
  for (int64_t batch = (int64_t)blockIdx.x * (MSDP_THREADS / 32) + wid; batch < nbatches; batch += nw) {
    const int64_t r0 = batch * 32;
    const int nr = (int)min((int64_t)32, a.nrows - r0);
    int e0 = 0, e1 = 0;
    if (lane < nr) {
      e0 = __ldg(a.bptr0 + r0 + lane);
      e1 = __ldg(a.bptr1 + r0 + lane);
    }
    const int E0 = __shfl_sync(0xffffffffu, e0, 0);
    const int E1 = __shfl_sync(0xffffffffu, e1, nr - 1);
    __syncwarp();  // every lane is done with the previous batch's staged entries
    for (int t = lane; t < E1 - E0; t += 32) {
      sc[t] = __ldg(col + E0 + t);
      sv[t] = __ldg(val + E0 + t);
    }
    __syncwarp();
    for (int r = 0; r < nr; ++r) {
      const int re0 = __shfl_sync(0xffffffffu, e0, r) - E0, re1 = __shfl_sync(0xffffffffu, e1, r) - E0;
      const int64_t row = r0 + r;
      const size_t off = (size_t)row * ld + lo;
      const int ownc = a.row0 + (int)row;
      double2 y = make_double2(0.0, 0.0), u = make_double2(0.0, 0.0);
      double eg = 0.0;
      if (EPI == EPI_HESS) {
        y = ldcs2(p.Y + off);
        u = ld2(p.Uown + off);
        eg = p.eG[row];
      } else if (EPI == EPI_COSTGRAD) {
        y = ld2(p.Uown + off);
      } else {
        u = ld2(p.Uown + off);
        eg = p.eG ? p.eG[row] : 0.0;
      }
      double2 acc = make_double2(0.0, 0.0);
      for (int base = re0; base < re1; base += GW) {
        double2 g[GW];
        double w[GW];
#pragma unroll
        for (int s = 0; s < GW; ++s) {
          const bool ok = base + s < re1;
          const int c = ok ? sc[base + s] : ownc;
          w[s] = ok ? sv[base + s] : 0.0;
          g[s] = ldg2(operand_row(c));
        }
#pragma unroll
        for (int s = 0; s < GW; ++s) {
          acc.x = fma(w[s], g[s].x, acc.x);
          acc.y = fma(w[s], g[s].y, acc.y);
        }
      }
      if (EPI == EPI_HESS) {
        const double dot = warp_sum(act ? y.x * acc.x + y.y * acc.y : 0.0);  // sum(Y.*eH), :129
        double2 hv;
        hv.x = acc.x - y.x * dot - u.x * eg;
        hv.y = acc.y - y.y * dot - u.y * eg;
        if (act) {
          stcs2(p.out + off, hv);
          q[0] += u.x * hv.x + u.y * hv.y;  // <mdelta, Hmdelta>, tCG.m:166
        }
      } else if (EPI == EPI_COSTGRAD) {
        const double dot = warp_sum(act ? y.x * acc.x + y.y * acc.y : 0.0);  // eG(row) = sum(YC.*Y), :119
        if (lane == 0) {
          p.eGout[row] = dot;
          q[0] += dot;
        }
        if (act) {
          double2 gv;
          gv.x = acc.x - y.x * dot;  // G = YC - Y.*eG, :124
          gv.y = acc.y - y.y * dot;
          st2(p.out + off, gv);
          q[1] += gv.x * gv.x + gv.y * gv.y;
        }
      } else if (act) {  // EPI_SHIFT: out = C*V - z.*V
        double2 o;
        o.x = acc.x - eg * u.x;
        o.y = acc.y - eg * u.y;
        st2(p.out + off, o);
      }
    }
  }
  spmm_tail<EPI>(a, q, sm);
}

// ---- bulk-async gather kernel (ld >= 32, one warp per row) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#define BULK_WARPS 8
#define BULK_MAXSLOTS 8

template <int VPL, int EPI>
__global__ void __launch_bounds__(BULK_WARPS * 32) k_spmm_bulk(const SpmmArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ double sm[2 * 32];
  __shared__ uint64_t bars[BULK_WARPS * 2];
  if ((EPI == EPI_HESS || EPI == EPI_SHIFT) && a.mode != TAIL_NONE && a.st->stop != 0) return;
  const SpmmPtrs p = select_ptrs(a);
  const int* __restrict__ col = a.col;
  const double* __restrict__ val = a.val;
  const double* __restrict__ Ug = p.Ug;
  const int ld = a.ld, nvec = ld / 2, slots = a.slots;
  const uint32_t rowbytes = (uint32_t)ld * 8u;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* myring = reinterpret_cast<double*>(smraw) + (size_t)wid * 2 * slots * ld;
  uint64_t* mybar = bars + wid * 2;
  if (lane == 0) {
    mbar_init(&mybar[0], 1);
    mbar_init(&mybar[1], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const bool first = a.first != 0, last = a.last != 0;
  const int64_t nw = (int64_t)gridDim.x * BULK_WARPS;
  int64_t row = (int64_t)blockIdx.x * BULK_WARPS + wid;
  uint32_t it = 0;  // batch counter of this warp: buffer = it & 1, parity = (it >> 1) & 1
  double q[2] = {0.0, 0.0};
  int e0 = 0, e1 = 0;
  if (row < a.nrows) {
    e0 = __ldg(a.bptr0 + row);
    e1 = __ldg(a.bptr1 + row);
  }
  for (; row < a.nrows; row += nw) {
    const int64_t rown = row + nw;
    int e0n = 0, e1n = 0;
    if (rown < a.nrows) {  // row pointers one row ahead
      e0n = __ldg(a.bptr0 + rown);
      e1n = __ldg(a.bptr1 + rown);
    }
    const size_t rb = (size_t)row * ld;
    double2 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int cv = lane + 32 * t;
      acc[t] = (!first && cv < nvec) ? ldcs2(p.out + rb + 2 * cv) : make_double2(0.0, 0.0);
    }
    int pos = e0;
    int c = 0;
    double w = 0.0;
    int nb = min(slots, e1 - pos);
    if (nb > 0) {
      if (lane < nb) {
        c = __ldcs(col + pos + lane);
        w = __ldcs(val + pos + lane);
      }
      if (lane == 0) mbar_expect_tx(&mybar[it & 1], (uint32_t)nb * rowbytes);
      __syncwarp();
      if (lane < nb)
        bulk_g2s(myring + ((size_t)(it & 1) * slots + lane) * ld, Ug + (size_t)c * ld, rowbytes, &mybar[it & 1]);
    }
    while (nb > 0) {
      const int posn = pos + nb;
      const int nbn = min(slots, e1 - posn);
      int cn = 0;
      double wn = 0.0;
      if (nbn > 0) {  // issue the next batch before consuming the current one
        if (lane < nbn) {
          cn = __ldcs(col + posn + lane);
          wn = __ldcs(val + posn + lane);
        }
        if (lane == 0) mbar_expect_tx(&mybar[(it + 1) & 1], (uint32_t)nbn * rowbytes);
        __syncwarp();
        if (lane < nbn)
          bulk_g2s(myring + ((size_t)((it + 1) & 1) * slots + lane) * ld, Ug + (size_t)cn * ld, rowbytes,
                   &mybar[(it + 1) & 1]);
      }
      mbar_wait(&mybar[it & 1], (it >> 1) & 1);
      const double* slot = myring + (size_t)(it & 1) * slots * ld;
#pragma unroll
      for (int s = 0; s < BULK_MAXSLOTS; ++s) {
        if (s < nb) {
          const double ws = __shfl_sync(0xffffffffu, w, s);
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int cv = lane + 32 * t;
            if (cv < nvec) {
              const double2 u = ld2(slot + (size_t)s * ld + 2 * cv);
              acc[t].x = fma(ws, u.x, acc[t].x);
              acc[t].y = fma(ws, u.y, acc[t].y);
            }
          }
        }
      }
      __syncwarp();  // every lane has read the slots before they are refilled
      ++it;
      pos = posn;
      nb = nbn;
      c = cn;
      w = wn;
    }
    if (!last) {
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int cv = lane + 32 * t;
        if (cv < nvec) stcs2(p.out + rb + 2 * cv, acc[t]);
      }
    } else {
      row_epilogue<32, VPL, EPI>(acc, p, row, ld, lane, 0xffffffffu, q);
    }
    e0 = e0n;
    e1 = e1n;
  }
  if (last) spmm_tail<EPI>(a, q, sm);
}

__global__ void k_tr_decide_scalar(RtrState* st) {
  st->fprop = 0.5 * st->tmp[0];
  st->gradnorm2_prop = st->tmp[1];
  tr_decide(st);
}
__global__ void k_tcg_after_hv_scalar(RtrState* st) {
  if (st->stop != 0) return;
  tcg_after_hv(st, st->tmp[0]);
}
int msdp_launch_tr_decide_scalar(manisdp_handle* h) {
  k_tr_decide_scalar<<<1, 1, 0, h->stream>>>(h->st);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}
int msdp_launch_tcg_after_hv_scalar(manisdp_handle* h) {
  k_tcg_after_hv_scalar<<<1, 1, 0, h->stream>>>(h->st);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// ---- column passes ------------------------------------------------------------------------------------------------------
// bptr[b*nrows + row] = first entry of `row` whose column is >= b*jrows  (b = 0..B); rows must be column-sorted
__global__ void k_block_ptrs(const int* __restrict__ rowptr, const int* __restrict__ col, int64_t nrows, int64_t jrows,
                             int B, int* __restrict__ bptr) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * blockDim.x) {
    const int e0 = rowptr[row], e1 = rowptr[row + 1];
    int e = e0;
    for (int b = 0; b <= B; ++b) {
      const int64_t lim = (int64_t)b * jrows;
      while (e < e1 && col[e] < lim) ++e;
      bptr[(size_t)b * nrows + row] = (b == B) ? e1 : e;
    }
  }
}

// decide the number of column passes for operand rows of `ld` doubles and (re)build the pass pointers
int msdp_spmm_prepare(manisdp_handle* h, int ld) {
  if (h->spmm_ld == ld) return MANISDP_OK;
  h->spmm_ld = ld;
  h->spmm_B = 1;
  const int64_t ncols = (h->world > 1) ? msdp_rows_per_rank(h->n, h->world) * h->world : h->n;
  const double operand_bytes = (double)ncols * ld * 8.0;
  int B = 1;
  if (h->spmm_block_mode != 0 && h->C_sorted && operand_bytes > (double)h->spmm_l2_target * 1.5) {
    if (h->spmm_block_mode == 2 || h->C_far_fraction > 0.3) {
      B = (int)((operand_bytes + h->spmm_l2_target - 1) / h->spmm_l2_target);
      if (B > 64) B = 64;
    }
  }
  if (B <= 1) return MANISDP_OK;
  const int64_t jrows = (ncols + B - 1) / B;
  if (h->spmm_bptr_cap < (size_t)(B + 1) * (size_t)h->nloc) {
    if (h->spmm_bptr) cudaFree(h->spmm_bptr);
    h->spmm_bptr = nullptr;
    h->spmm_bptr_cap = (size_t)(B + 1) * (size_t)h->nloc;
    CUDA_TRY(h, cudaMalloc((void**)&h->spmm_bptr, h->spmm_bptr_cap * sizeof(int)));
  }
  k_block_ptrs<<<std::max(1, (int)std::min<int64_t>(h->num_sms * 8, (h->nloc + 255) / 256)), 256, 0, h->stream>>>(
      h->C.rowptr, h->C.col, h->nloc, jrows, B, h->spmm_bptr);
  KERNEL_CHECK(h);
  h->spmm_B = B;
  return MANISDP_OK;
}

template <int VPL, int EPI>
static int launch_bulk(manisdp_handle* h, const SpmmArgs& a) {
  const size_t smem = (size_t)BULK_WARPS * 2 * a.slots * a.ld * sizeof(double);
  // per-device attribute, cheap to set: no process-wide "done" flag (handles may live on several GPUs)
  CUDA_TRY(h, cudaFuncSetAttribute(k_spmm_bulk<VPL, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const int64_t rows_per_block = BULK_WARPS;
  int64_t nb = (a.nrows + rows_per_block - 1) / rows_per_block;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 1024)));
  nb = std::min<int64_t>(nb, (int64_t)h->num_sms * per_sm);
  nb = std::min<int64_t>(std::max<int64_t>(nb, 1), MSDP_MAX_BLOCKS);
  k_spmm_bulk<VPL, EPI><<<(int)nb, BULK_WARPS * 32, smem, h->stream>>>(a);
  return MANISDP_OK;
}

// ---- block-major product (32 < ld <= 64, single GPU, graphs without locality) ----------------------------------------
// The operand of the bench instance (n = 1e6, p = 64: 512 MB) is 4x the L2 and the graph has no locality, so the row
// kernels above re-read it ~10x from HBM.  Here the entries are stored column block by column block (api.cu:
// build_block_major) and one launch per block streams them: a warp takes a row-aligned chunk of ~224 consecutive
// entries = hundreds of independent gathers from a block that stays in L2, accumulates while the row stays the same and
// stores the partial row of (block, row) into the block's own buffer when the row changes (no read inside a pass, so
// no dependent load).  k_bm_finish then adds the partial rows a row really has (bit mask per row) and applies the
// usual fused epilogue and reductions.  Measured in tools/spmm_blockmajor_lab.cu (profiles/r1_spmm_blockmajor_lab.txt):
// 2.86 ms against 3.44-3.88 ms for the single pass; per-pass times sit on the L2 gather rate (13-19 TB/s).
// STATUS (end of round 1): correct (tests/test_gpu_maxcut.py forces it with MANISDP_SPMM_BM=2) but OPT-IN -- inside the
// product the pass kernel runs at 0.84 ms against 0.62 ms in the lab and the epilogue pass costs 0.66 ms, 4.09 ms per
// product against 3.69 ms for k_spmm (profiles/r1_blockmajor_product_ab.txt); the default stays the row kernel until
// the pass reaches the lab's rate.
// Entry order within a row is unchanged; only the association of the sum changes (per-block partial sums).
#define BM_U 8
template <int EPI>
__global__ void __launch_bounds__(MSDP_THREADS, 3)
    k_bm_pass(const SpmmArgs a, const int* __restrict__ ecol, const double* __restrict__ eval_,
              const int* __restrict__ erow, const int* __restrict__ chunk_ptr, int nchunks, double* __restrict__ part) {
  if ((EPI == EPI_HESS || EPI == EPI_SHIFT) && a.mode != TAIL_NONE && a.st->stop != 0) return;
  const SpmmPtrs p = select_ptrs(a);
  const double* __restrict__ Ug = p.Ug;
  const int ld = a.ld;
  const int lane = threadIdx.x & 31;
  const bool act = lane < ld / 2;
  // The gathers are UNCONDITIONAL: a lane past ld/2 (only when ld < 64) reads vector 0 of the same operand row and
  // never stores.  A predicated `act ? ldg2(..) : 0` made ptxas route every load through a temporary and issue the
  // eight gathers of a group in ~3 dependent batches (round-1 SASS, profiles/r1_blockmajor_product_ab.txt).
  // The entry stream is read and the partial rows are written with streaming (evict-first) hints: both are touched once
  // per pass, and write-allocating 8np bytes of partial rows per pass would push the operand block out of the L2.
  const int lo = act ? 2 * lane : 0;
  const int nw = gridDim.x * (MSDP_THREADS / 32);
  for (int ch = blockIdx.x * (MSDP_THREADS / 32) + (threadIdx.x >> 5); ch < nchunks; ch += nw) {
    const int e0 = __ldg(chunk_ptr + ch), e1 = __ldg(chunk_ptr + ch + 1);
    int cur = -1, prev_last = -1;
    double2 acc = make_double2(0.0, 0.0);
    int c = 0, r = -1;
    double w = 0.0;
    if (e0 + lane < e1) {
      c = __ldcs(ecol + e0 + lane);
      w = __ldcs(eval_ + e0 + lane);
      r = __ldcs(erow + e0 + lane);
    }
    for (int base = e0; base < e1; base += 32) {
      int cn = 0, rn = -1;
      double wn = 0.0;
      if (base + 32 + lane < e1) {  // next group's entries before this group's gathers
        cn = __ldcs(ecol + base + 32 + lane);
        wn = __ldcs(eval_ + base + 32 + lane);
        rn = __ldcs(erow + base + 32 + lane);
      }
      const int cnt = min(32, e1 - base);
      int rprev = __shfl_up_sync(0xffffffffu, r, 1);
      if (lane == 0) rprev = prev_last;
      const unsigned chg = __ballot_sync(0xffffffffu, lane < cnt && r != rprev);  // bit k: entry k starts a new row
      prev_last = __shfl_sync(0xffffffffu, r, cnt - 1);
      // lanes past cnt hold (column 0, weight 0): the count is rounded up to the unroll width
      const int cnt_pad = min(32, (cnt + BM_U - 1) / BM_U * BM_U);
      for (int k = 0; k < cnt_pad; k += BM_U) {
        double2 u[BM_U];
#pragma unroll
        for (int s = 0; s < BM_U; ++s) {
          const int cj = __shfl_sync(0xffffffffu, c, k + s);
          u[s] = ldg2(Ug + (size_t)cj * ld + lo);
        }
#pragma unroll
        for (int s = 0; s < BM_U; ++s) {
          if ((chg >> (k + s)) & 1u) {  // warp-uniform
            if (cur >= 0 && act) stcs2(part + (size_t)cur * ld + 2 * lane, acc);
            cur = __shfl_sync(0xffffffffu, r, k + s);
            acc = make_double2(0.0, 0.0);
          }
          const double ws = __shfl_sync(0xffffffffu, w, k + s);
          acc.x = fma(ws, u[s].x, acc.x);
          acc.y = fma(ws, u[s].y, acc.y);
        }
      }
      c = cn;
      w = wn;
      r = rn;
    }
    if (cur >= 0 && act) stcs2(part + (size_t)cur * ld + 2 * lane, acc);
  }
}

// out(row) = epilogue( sum over ALL blocks b of part[b][row] ), + the reductions / scalar tail.  The partial buffers
// are zeroed whenever the row length changes (msdp_resize) and a pass only ever writes the (block, row) pairs that have
// entries, so a pair without entries reads as an exact 0 -- no per-row mask, no predicated loads: the NB loads of a
// row (plus the epilogue operands) are in flight together.
template <int EPI, int NB>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_bm_finish(const SpmmArgs a, const double* __restrict__ part) {
  __shared__ double sm[2 * 32];
  if ((EPI == EPI_HESS || EPI == EPI_SHIFT) && a.mode != TAIL_NONE && a.st->stop != 0) return;
  const SpmmPtrs p = select_ptrs(a);
  const int ld = a.ld;
  const int gl = threadIdx.x & 31;
  const bool act = gl < ld / 2;
  const int lo = act ? 2 * gl : 0;
  const size_t pstride = (size_t)a.nrows * ld;
  const int64_t ngroups = (int64_t)gridDim.x * (MSDP_THREADS / 32);
  double q[2] = {0.0, 0.0};
  for (int64_t row = (int64_t)blockIdx.x * (MSDP_THREADS / 32) + threadIdx.x / 32; row < a.nrows; row += ngroups) {
    double2 v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = ldcs2(part + (size_t)b * pstride + (size_t)row * ld + lo);
    double2 acc[1];
    acc[0] = v[0];
#pragma unroll
    for (int b = 1; b < NB; ++b) {
      acc[0].x += v[b].x;
      acc[0].y += v[b].y;
    }
    row_epilogue<32, 1, EPI>(acc, p, row, ld, gl, 0xffffffffu, q);
  }
  spmm_tail<EPI>(a, q, sm);
}

static bool bm_applies(const manisdp_handle* h, const SpmmArgs& a) {
  return h->bm_B > 0 && h->bm_part && a.ld > 32 && a.ld <= 64 && !a.sharded && !a.peer_tab &&
         h->bm_part_cap >= (size_t)h->bm_B * (size_t)a.nrows * (size_t)a.ld;
}

template <int EPI>
static int launch_bm(manisdp_handle* h, const SpmmArgs& a) {
  const int B = h->bm_B;
  const size_t pstride = (size_t)a.nrows * (size_t)a.ld;
  const int grid = h->num_sms * 3;
  for (int b = 0; b < B; ++b) {
    const int nch = h->bm_chunk_off[(size_t)b + 1] - h->bm_chunk_off[(size_t)b] - 1;
    if (nch <= 0) continue;
    k_bm_pass<EPI><<<std::min(grid, (nch + 7) / 8), MSDP_THREADS, 0, h->stream>>>(
        a, h->bm_col, h->bm_val, h->bm_row, h->bm_chunk + h->bm_chunk_off[(size_t)b], nch, h->bm_part + b * pstride);
    KERNEL_CHECK(h);
  }
  const int gf = rows_grid(h, a.nrows, 32);
  switch (B) {
    case 2: k_bm_finish<EPI, 2><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
    case 3: k_bm_finish<EPI, 3><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
    case 4: k_bm_finish<EPI, 4><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
    case 5: k_bm_finish<EPI, 5><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
    case 6: k_bm_finish<EPI, 6><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
    case 7: k_bm_finish<EPI, 7><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
    default: k_bm_finish<EPI, 8><<<gf, MSDP_THREADS, 0, h->stream>>>(a, h->bm_part); break;
  }
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

template <int GS, int VPL, int EPI>
static void launch_rows(manisdp_handle* h, const SpmmArgs& a) {
  const int nb = rows_grid(h, a.nrows, GS);
  if constexpr (VPL == 1 && GS <= 16) {
    if (h->spmm_narrow && !a.peer_tab) {  // (the peer-table variant would spill at the 64-register cap)
      k_spmm_narrow<GS, EPI, false><<<nb, MSDP_THREADS, 0, h->stream>>>(a);
      return;
    }
  }
  if (a.peer_tab)
    k_spmm<GS, VPL, EPI, true><<<nb, MSDP_THREADS, 0, h->stream>>>(a);
  else
    k_spmm<GS, VPL, EPI, false><<<nb, MSDP_THREADS, 0, h->stream>>>(a);
}

// one pass (a.bptr0 / a.bptr1 / a.first / a.last set by the caller)
template <int EPI>
static int launch_pass(manisdp_handle* h, SpmmArgs a) {
  // bulk path: measured on B200 (profiles/r1_sweep_bulk_vs_regs.txt) it wins from ld = 128 on (7.7 vs 10.2 ms at
  // p = 128), ties at p = 64 and loses below, where four register gathers per group already cover the latency
  const bool bulk = !a.peer_tab && (h->spmm_use_bulk == 2 ? (a.ld >= 32) : (h->spmm_use_bulk == 1 && a.ld >= 96));
  a.slots = std::max(1, std::min(BULK_MAXSLOTS, 4096 / (a.ld * 8)));
  if (h->spmm_lowdeg && h->C_lowdeg && a.first && a.last && a.bptr0 == h->C.rowptr && a.ld > 32 && a.ld <= 64 &&
      h->spmm_use_bulk != 2) {
    a.row0 = (int)h->row_begin;
    const int nb = (int)std::max<int64_t>(
        1, std::min<int64_t>((int64_t)h->num_sms * 4, (a.nrows + 32 * (MSDP_THREADS / 32) - 1) / (32 * (MSDP_THREADS / 32))));
    const int gw = h->C_maxdeg <= 4 ? 4 : h->C_maxdeg <= 5 ? 5 : h->C_maxdeg <= 6 ? 6 : 8;
#define LOWDEG_LAUNCH(GW_)                                                      \
  do {                                                                          \
    const int nbw = std::min(nb, h->num_sms * ((GW_) <= 6 ? 4 : 3));            \
    if (a.peer_tab)                                                             \
      k_spmm_lowdeg<EPI, true, GW_><<<nbw, MSDP_THREADS, 0, h->stream>>>(a);    \
    else                                                                        \
      k_spmm_lowdeg<EPI, false, GW_><<<nbw, MSDP_THREADS, 0, h->stream>>>(a);   \
  } while (0)
    switch (gw) {
      case 4: LOWDEG_LAUNCH(4); break;
      case 5: LOWDEG_LAUNCH(5); break;
      case 6: LOWDEG_LAUNCH(6); break;
      default: LOWDEG_LAUNCH(8); break;
    }
#undef LOWDEG_LAUNCH
    KERNEL_CHECK(h);
    return MANISDP_OK;
  }
  if (bulk) {
    const int vpl = row_geom(a.ld).vpl;
    if (vpl == 1)
      launch_bulk<1, EPI>(h, a);
    else if (vpl == 2)
      launch_bulk<2, EPI>(h, a);
    else if (vpl == 4)
      launch_bulk<4, EPI>(h, a);
    else
      launch_bulk<8, EPI>(h, a);
  } else {
    a.row0 = (int)h->row_begin;
    DISPATCH_GEOM(row_geom(a.ld), { launch_rows<GS, VPL, EPI>(h, a); });
  }
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

template <int EPI>
static int launch_spmm(manisdp_handle* h, SpmmArgs a) {
  if (EPI != EPI_SHIFT && bm_applies(h, a)) return launch_bm<EPI>(h, a);
  MSDP_TRY(msdp_spmm_prepare(h, a.ld));
  const int B = h->spmm_B;
  for (int b = 0; b < B; ++b) {
    if (B == 1) {
      a.bptr0 = h->C.rowptr;
      a.bptr1 = h->C.rowptr + 1;
    } else {
      a.bptr0 = h->spmm_bptr + (size_t)b * h->nloc;
      a.bptr1 = h->spmm_bptr + (size_t)(b + 1) * h->nloc;
    }
    a.first = (b == 0);
    a.last = (b == B - 1);
    MSDP_TRY(launch_pass<EPI>(h, a));
  }
  return MANISDP_OK;
}

// ---- exchange-overlapped product of a row-sharded handle ------------------------------------------------------------
// The operand arrives chunk by chunk (dist.cu: msdp_dist_exchange_begin, one chunk per owner rank, the own one first);
// the product runs as one column pass per owner chunk in arrival order, each pass waiting only for its chunk, the
// partial rows staying in `out` (n/G rows: L2-sized at the scales where this matters), the last pass applying the
// epilogue.  Owner chunks are column blocks of width ceil(n/G), so the pass pointers are built once per handle.
static int owner_ptrs(manisdp_handle* h) {
  if (h->owner_bptr) return MANISDP_OK;
  const int G = h->world;
  CUDA_TRY(h, cudaMalloc((void**)&h->owner_bptr, (size_t)(G + 1) * (size_t)h->nloc * sizeof(int)));
  k_block_ptrs<<<std::max(1, (int)std::min<int64_t>(h->num_sms * 8, (h->nloc + 255) / 256)), 256, 0, h->stream>>>(
      h->C.rowptr, h->C.col, h->nloc, msdp_rows_per_rank(h->n, G), G, h->owner_bptr);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

bool msdp_pipeline_ok(const manisdp_handle* h) {
  return h->world > 1 && h->C_sorted && (h->pipeline == 1 || (h->pipeline == 2 && h->ipc_ready));
}

bool msdp_peer_gather_ok(const manisdp_handle* h) {
  return msdp_pipeline_ok(h) && h->pipeline == 2 && h->peer_tab_dev &&
         h->C_remote_fraction < h->peer_gather_max_remote;
}

int msdp_maxcut_hess_pipelined(manisdp_handle* h, const double* Down, double* Hout, int from_state, int tail_mode) {
  // Graphs with locality (few entries point outside the owned rows, e.g. the torus profile): no exchange at all --
  // after the cross-rank "operand written" barrier the SpMM gathers the few remote rows directly from the owner's
  // memory (peer loads over NVLink).  Without locality every shard needs ~all rows and each would cross NVLink ~nnz/n
  // times (peer loads bypass the local L2), so the staged chunk copies + column passes below are used instead.
  const double* const* tab = msdp_dist_peer_table(h, Down);
  if (tab && msdp_peer_gather_ok(h)) {
    MSDP_TRY(msdp_dist_barrier(h));
    SpmmArgs a{};
    a.col = h->C.col;
    a.val = h->C.val;
    a.nrows = h->nloc;
    a.ld = (int)h->ld;
    a.v = msdp_vecptrs(h);
    a.st = h->st;
    a.partials = h->partials;
    a.sharded = 1;
    a.Ug = Down;
    a.Uown = Down;
    a.out = Hout;
    a.sel = from_state ? 1 : 0;
    a.Y = h->Ybuf[h->pt];
    a.eG = h->eG[h->pt];
    a.mode = tail_mode;
    a.peer_tab = tab;
    a.rpr = (int)msdp_rows_per_rank(h->n, h->world);
    a.bptr0 = h->C.rowptr;
    a.bptr1 = h->C.rowptr + 1;
    a.first = 1;
    a.last = 1;
    return launch_pass<EPI_HESS>(h, a);
  }
  MSDP_TRY(owner_ptrs(h));
  MSDP_TRY(msdp_dist_exchange_begin(h, Down, h->gatherbuf));
  SpmmArgs a{};
  a.col = h->C.col;
  a.val = h->C.val;
  a.nrows = h->nloc;
  a.ld = (int)h->ld;
  a.v = msdp_vecptrs(h);
  a.st = h->st;
  a.partials = h->partials;
  a.sharded = 1;
  a.Ug = h->gatherbuf;
  a.Uown = Down;
  a.out = Hout;
  a.sel = from_state ? 1 : 0;
  a.Y = h->Ybuf[h->pt];
  a.eG = h->eG[h->pt];
  a.mode = tail_mode;
  const int G = h->world, r = h->rank;
  for (int s = 0; s < G; ++s) {
    const int q = (r - s + G) % G;
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_stage[(size_t)s], 0));
    a.bptr0 = h->owner_bptr + (size_t)q * h->nloc;
    a.bptr1 = h->owner_bptr + (size_t)(q + 1) * h->nloc;
    a.first = (s == 0);
    a.last = (s == G - 1);
    MSDP_TRY(launch_pass<EPI_HESS>(h, a));
  }
  return MANISDP_OK;
}

static SpmmArgs base_args(manisdp_handle* h) {
  SpmmArgs a{};
  a.col = h->C.col;
  a.val = h->C.val;
  a.nrows = h->nloc;
  a.ld = (int)h->ld;
  a.v = msdp_vecptrs(h);
  a.st = h->st;
  a.partials = h->partials;
  a.sharded = (h->world > 1);
  return a;
}

// Hout = Hess f(Y)[D].  from_state = 1: Y / eG chosen on the device from st->pt (inside tr_solve);
// otherwise the host mirror h->pt is used.
int msdp_maxcut_hess(manisdp_handle* h, const double* Dgather, const double* Down, double* Hout, int from_state,
                     int tail_mode) {
  SpmmArgs a = base_args(h);
  a.Ug = Dgather;
  a.Uown = Down;
  a.out = Hout;
  a.sel = from_state ? 1 : 0;
  a.Y = h->Ybuf[h->pt];
  a.eG = h->eG[h->pt];
  a.mode = tail_mode;
  return launch_spmm<EPI_HESS>(h, a);
}

// which >= 0: explicit point buffer; -1: proposal (st->pt ^ 1, device-selected); -2: current (st->pt).
// Row-sharded handles must have all-gathered the point into h->gatherbuf first.
int msdp_maxcut_costgrad(manisdp_handle* h, int which, int cg_mode) {
  SpmmArgs a = base_args(h);
  a.mode = cg_mode;
  if (which >= 0) {
    a.sel = 0;
    a.Uown = h->Ybuf[which];
    a.Ug = a.sharded ? h->gatherbuf : a.Uown;
    a.out = h->Gbuf[which];
    a.eGout = h->eG[which];
  } else {
    a.sel = (which == -1) ? 2 : 3;
    a.Ug = h->gatherbuf;
  }
  if (h->cg_peer_tab) {
    a.peer_tab = h->cg_peer_tab;
    a.rpr = (int)msdp_rows_per_rank(h->n, h->world);
  }
  return launch_spmm<EPI_COSTGRAD>(h, a);
}

// exchange + cost/grad product of a row-sharded handle at host-known point buffer `buf` (which: as msdp_costgrad)
int msdp_costgrad_exchange(manisdp_handle* h, int buf, int which, int cg_mode) {
  const double* const* tab = msdp_peer_gather_ok(h) ? msdp_dist_peer_table(h, h->Ybuf[buf]) : nullptr;
  if (!tab) {
    MSDP_TRY(msdp_dist_allgather_rows(h, h->Ybuf[buf], h->gatherbuf));
    return msdp_costgrad(h, which, cg_mode);
  }
  MSDP_TRY(msdp_dist_barrier(h));  // every rank has written its rows of the point
  h->cg_peer_tab = tab;
  const int rc = msdp_costgrad(h, which, cg_mode);
  h->cg_peer_tab = nullptr;
  return rc;
}

// out = C*V - zdiag.*V on an n x k_ld block (eigen step); zdiag may be NULL
int msdp_spmm_shift(manisdp_handle* h, const double* Vgather, const double* Vown, double* out, int k_ld,
                    const double* zdiag) {
  SpmmArgs a = base_args(h);
  a.ld = k_ld;
  a.Ug = Vgather;
  a.Uown = Vown;
  a.out = out;
  a.eG = zdiag;
  a.sel = 0;
  return launch_spmm<EPI_SHIFT>(h, a);
}

// out = C*V, the raw product of the column-sharded closures (colshard.cu).  in_tcg: a no-op once tCG has stopped.
int msdp_spmm_shift_tcg(manisdp_handle* h, const double* V, double* out, int ld, int in_tcg) {
  SpmmArgs a = base_args(h);
  a.ld = ld;
  a.Ug = V;
  a.Uown = V;
  a.out = out;
  a.eG = nullptr;
  a.sel = 0;
  a.mode = in_tcg ? TAIL_TCG : TAIL_NONE;
  return launch_spmm<EPI_SHIFT>(h, a);
}
