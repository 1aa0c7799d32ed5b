// spmm.cu -- K1: CSR SpMM with fused row epilogues, the closures of ManiSDP_onlyunitdiag.m:117-130.
//
//   hess      (:128-129)  eH = U*C ; H = eH - Y.*sum(Y.*eH) - U.*eG          -> one kernel, + <U,H> for tCG.m:166
//   cost+grad (:118-124)  YC = Y*C ; eG = sum(YC.*Y) ; f = .5*sum(eG) ; G = YC - Y.*eG   -> one kernel, + |G|^2
//   S*V       (:49)       (C - diag(z)) * V for the eigen step                -> one kernel
//
// Layout: the factor is vertex-major (row i = the p numbers of vertex i, ld = 4*ceil(p/4) doubles = the reference's
// p x n column-major array), so one gathered operand row is one contiguous ld*8-byte segment.  A row of the output is
// produced by a group of GS lanes, each holding VPL double2 accumulators; the row's (col, val) entries are loaded
// coalesced by the group and broadcast with shuffles; gathers are issued four at a time for memory-level parallelism.
// The row epilogue (projection onto the tangent space of the oblique manifold, ManiSDP_onlyunitdiag.m:138-139) needs
// only the finished row, so it is fused and H is written exactly once.
//
// Algorithmic bytes per product: 12*nnz + 4*(n+1) + 24*n*p + 8*n.
#include "rowops.cuh"
#include "scalar_logic.cuh"
#include "kernels.cuh"

enum { EPI_HESS = 0, EPI_COSTGRAD = 1, EPI_SHIFT = 2 };

struct SpmmArgs {
  const int* rowptr;
  const int* col;
  const double* val;
  int64_t nrows;
  int ld;
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
};

template <int GS, int VPL, int EPI>
__global__ void __launch_bounds__(MSDP_THREADS) k_spmm(SpmmArgs a) {
  __shared__ double sm[2 * 32];
  RtrState* st = a.st;
  if (EPI == EPI_HESS && a.mode != TAIL_NONE && st->stop != 0) return;
  // ---- device-side buffer selection (no host round trip between TR iterations)
  if (a.sel == 1) {
    const int pt = st->pt;
    a.Y = pt ? a.v.Y1 : a.v.Y0;
    a.eG = pt ? a.v.eG1 : a.v.eG0;
  } else if (a.sel >= 2) {
    const int w = (a.sel == 2) ? (st->pt ^ 1) : st->pt;
    a.Uown = w ? a.v.Y1 : a.v.Y0;
    if (!a.sharded) a.Ug = a.Uown;
    a.out = w ? a.v.G1 : a.v.G0;
    a.eGout = w ? a.v.eG1 : a.v.eG0;
  }
  const int* __restrict__ rowptr = a.rowptr;
  const int* __restrict__ col = a.col;
  const double* __restrict__ val = a.val;
  const double* __restrict__ Ug = a.Ug;
  const int ld = a.ld, nvec = ld / 2;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  double q[2] = {0.0, 0.0};

  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < a.nrows; row += ngroups) {
    const int e0 = rowptr[row], e1 = rowptr[row + 1];
    double2 acc[VPL];
#pragma unroll
    for (int t = 0; t < VPL; ++t) acc[t] = make_double2(0.0, 0.0);

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
        const double* p0 = Ug + (size_t)c0 * ld;
        const double* p1 = Ug + (size_t)c1 * ld;
        const double* p2 = Ug + (size_t)c2 * ld;
        const double* p3 = Ug + (size_t)c3 * ld;
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
        const double* p0 = Ug + (size_t)c0 * ld;
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

    // ---- fused row epilogue
    const size_t rb = (size_t)row * ld;
    if (EPI == EPI_HESS) {
      double2 y[VPL], u[VPL];
      double dot = 0.0;
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int cv = gl + GS * t;
        if (cv < nvec) {
          y[t] = ld2(a.Y + rb + 2 * cv);
          u[t] = ld2(a.Uown + rb + 2 * cv);
          dot += y[t].x * acc[t].x + y[t].y * acc[t].y;
        }
      }
      dot = group_sum<GS>(dot, mask);  // sum(Y.*eH), :129
      const double eg = a.eG[row];
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int cv = gl + GS * t;
        if (cv < nvec) {
          double2 hv;
          hv.x = acc[t].x - y[t].x * dot - u[t].x * eg;
          hv.y = acc[t].y - y[t].y * dot - u[t].y * eg;
          st2(a.out + rb + 2 * cv, hv);
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
          y[t] = ld2(a.Uown + rb + 2 * cv);
          dot += y[t].x * acc[t].x + y[t].y * acc[t].y;
        }
      }
      dot = group_sum<GS>(dot, mask);  // eG(row) = sum(YC.*Y), :119
      if (gl == 0) {
        a.eGout[row] = dot;
        q[0] += dot;
      }
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int cv = gl + GS * t;
        if (cv < nvec) {
          double2 g;
          g.x = acc[t].x - y[t].x * dot;  // G = YC - Y.*eG, :124
          g.y = acc[t].y - y[t].y * dot;
          st2(a.out + rb + 2 * cv, g);
          q[1] += g.x * g.x + g.y * g.y;
        }
      }
    } else {  // EPI_SHIFT: out = C*V - z.*V
      const double z = a.eG ? a.eG[row] : 0.0;
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        const int cv = gl + GS * t;
        if (cv < nvec) {
          const double2 u = ld2(a.Uown + rb + 2 * cv);
          double2 o;
          o.x = acc[t].x - z * u.x;
          o.y = acc[t].y - z * u.y;
          st2(a.out + rb + 2 * cv, o);
        }
      }
    }
  }

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

template <int EPI>
static int launch_spmm(manisdp_handle* h, const SpmmArgs& a) {
  DISPATCH_GEOM(row_geom(a.ld), {
    const int nb = rows_grid(h, a.nrows, GS);
    k_spmm<GS, VPL, EPI><<<nb, MSDP_THREADS, 0, h->stream>>>(a);
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

static SpmmArgs base_args(manisdp_handle* h) {
  SpmmArgs a{};
  a.rowptr = h->C.rowptr;
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
  return launch_spmm<EPI_COSTGRAD>(h, a);
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
