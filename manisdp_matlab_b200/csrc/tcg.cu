// tcg.cu -- fused vector kernels of the truncated-CG / trust-region loop.
// Replaces the ~15 MATLAB expressions (≈30 array passes) per inner iteration of
// manopt7.0/manopt/solvers/trustregions/tCG.m:166-287 with two passes:
//   update pass  (tCG.m:192-241): eta' = eta - a*mdelta, r' = r - a*Hmdelta and all inner products   56 np bytes
//   direction    (tCG.m:273-283): mdelta' = P_Y(r' + beta*mdelta)                                       32 np bytes
// Heta is never stored: with eta0 = 0 the recurrences give Heta = r - grad exactly (tCG.m:220,238 apply the same
// update -alpha*Hmdelta to both), so <eta,Heta> = <eta, r - grad>.
#include "rowops.cuh"
#include "scalar_logic.cuh"
#include "kernels.cuh"

// ---- tCG init: r = mdelta = grad, eta = 0 (tCG.m:103-134) --------------------------------------------------------
__global__ void __launch_bounds__(MSDP_THREADS) k_tcg_init(VecPtrs v, RtrState* st, int64_t nvec) {
  const double* g = st->pt ? v.G1 : v.G0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const double2 gv = ld2(g + 2 * i);
    st2(v.r + 2 * i, gv);
    st2(v.d + 2 * i, gv);
    st2(v.eta0 + 2 * i, make_double2(0.0, 0.0));
  }
  if (grid_last(&st->ticket)) {
    if (threadIdx.x == 0) tcg_reset(st);
  }
}

// ---- update pass ----------------------------------------------------------------------------------------------------
template <int MF>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_tcg_update(VecPtrs v, RtrState* st, double* partials, int64_t nvec, cudaGraphConditionalHandle cond,
                 int use_cond, int defer) {
  __shared__ double sm[MSDP_NQ * 32];
  if (st->stop != 0) return;
  const int branch = st->branch, cur = st->eta_cur;
  const double a = (branch != 0) ? st->tau : st->alpha;
  const double* __restrict__ g = st->pt ? v.G1 : v.G0;
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const double* __restrict__ eo = cur ? v.eta1 : v.eta0;
  double* __restrict__ en = cur ? v.eta0 : v.eta1;
  double q[6] = {0, 0, 0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (branch != 0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      const double2 e = ld2(eo + 2 * i), d = ld2(v.d + 2 * i), r = ld2(v.r + 2 * i), hd = ld2(v.Hd + 2 * i),
                    gv = ld2(g + 2 * i);
      double2 n;
      n.x = e.x - a * d.x;  // tCG.m:192
      n.y = e.y - a * d.y;
      st2(en + 2 * i, n);
      const double hx = (r.x - gv.x) - a * hd.x, hy = (r.y - gv.y) - a * hd.y;  // Heta - tau*Hmdelta (:196)
      q[0] += n.x * gv.x + n.y * gv.y;
      q[1] += n.x * hx + n.y * hy;
      q[2] += n.x * n.x + n.y * n.y;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      const double2 e = ld2(eo + 2 * i), d = ld2(v.d + 2 * i), r = ld2(v.r + 2 * i), hd = ld2(v.Hd + 2 * i),
                    gv = ld2(g + 2 * i);
      double2 n, rn;
      n.x = e.x - a * d.x;  // :215
      n.y = e.y - a * d.y;
      rn.x = r.x - a * hd.x;  // :238
      rn.y = r.y - a * hd.y;
      st2(en + 2 * i, n);
      st2(v.r + 2 * i, rn);
      q[0] += n.x * gv.x + n.y * gv.y;
      q[1] += n.x * (rn.x - gv.x) + n.y * (rn.y - gv.y);
      q[2] += rn.x * rn.x + rn.y * rn.y;
      q[3] += n.x * n.x + n.y * n.y;
      if (MF == MF_SPHERE) {
        const double2 y = ld2(Y + 2 * i);
        q[4] += y.x * rn.x + y.y * rn.y;
        q[5] += y.x * d.x + y.y * d.y;
      }
    }
  }
  double tot[6];
  if (grid_sum_last<6>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) {
      if (defer) {
        for (int k = 0; k < 6; ++k) st->tmp[k] = tot[k];
      } else {
        tcg_after_update(st, tot, cond, use_cond);
      }
    }
  }
}

// ---- direction pass: mdelta = tangent(r + beta*mdelta) (tCG.m:273,283) ----------------------------------------------
template <int GS, int VPL, int MF>
__global__ void __launch_bounds__(MSDP_THREADS) k_tcg_dir(VecPtrs v, RtrState* st, int64_t nrows, int ld) {
  if (st->stop != 0) return;
  const double beta = st->beta;
  const double* __restrict__ Y = st->pt ? v.Y1 : v.Y0;
  const double sph = (MF == MF_SPHERE) ? (st->y_r + beta * st->y_d) : 0.0;
  const long long nob = (MF == MF_OBLIQUE) ? st->nob_rows : 0;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const int nvec = ld / 2;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double2 dn[VPL], y[VPL];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        const double2 r = ld2(v.r + base + 2 * c), d = ld2(v.d + base + 2 * c);
        dn[t].x = r.x + beta * d.x;
        dn[t].y = r.y + beta * d.y;
        if (MF != MF_EUCLID) {
          y[t] = ld2(Y + base + 2 * c);
          dot += y[t].x * dn[t].x + y[t].y * dn[t].y;
        }
      }
    }
    if (MF == MF_OBLIQUE) dot = rowsel(group_sum<GS>(dot, mask), row < nob);  // ManiSDP_unitdiag.m:181 / projc.cpp:34-48
    if (MF == MF_SPHERE) dot = sph;                        // spherefactory.m:113
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        if (MF != MF_EUCLID) {
          dn[t].x -= y[t].x * dot;
          dn[t].y -= y[t].y * dot;
        }
        st2(v.d + base + 2 * c, dn[t]);
      }
    }
  }
}

// ---- retraction (ManiSDP_unitdiag.m:184-187 ; spherefactory.m:220-232 ; euclideanfactory.m:65) ----------------------
// dst = retr(Y, eta).  Oblique: row-wise normalisation.  Sphere: pass 1 writes Y+eta and reduces |Y+eta|_F^2 into
// st->tmp[0]; k_scale_by_tmp0 finishes.  Euclid: Y+eta.
template <int GS, int VPL, int MF>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_retract(const double* Yin, const double* eta, double* dst, VecPtrs v, RtrState* st, double* partials,
              int64_t nrows, int ld, int from_state) {
  __shared__ double sm[32];
  const double* __restrict__ Y = from_state ? (st->pt ? v.Y1 : v.Y0) : Yin;
  const double* __restrict__ E = from_state ? (st->eta_cur ? v.eta1 : v.eta0) : eta;
  double* __restrict__ out = from_state ? (st->pt ? v.Y0 : v.Y1) : dst;
  const long long nob = (MF == MF_OBLIQUE) ? st->nob_rows : 0;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const int nvec = ld / 2;
  double q[1] = {0.0};
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double2 x[VPL];
    double ss = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        const double2 y = ld2(Y + base + 2 * c), e = ld2(E + base + 2 * c);
        x[t].x = y.x + e.x;
        x[t].y = y.y + e.y;
        ss += x[t].x * x[t].x + x[t].y * x[t].y;
      }
    }
    double scale = 1.0;
    if (MF == MF_OBLIQUE) {
      ss = group_sum<GS>(ss, mask);
      scale = (row < nob) ? 1.0 / sqrt(ss) : 1.0;  // Euclidean blocks of a multi-block point: Y + eta (retrc.cpp)
    }
    if (MF == MF_SPHERE) q[0] += ss;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        if (MF == MF_OBLIQUE) {
          x[t].x *= scale;
          x[t].y *= scale;
        }
        st2(out + base + 2 * c, x[t]);
      }
    }
  }
  if (MF == MF_SPHERE) {
    double tot[1];
    __syncwarp();
    if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
      if (threadIdx.x == 0) st->tmp[0] = tot[0];
    }
  }
}

// x *= 1/sqrt(st->tmp[0])   (second half of the sphere retraction / normalisation)
__global__ void __launch_bounds__(MSDP_THREADS)
    k_scale_by_tmp0(double* x, VecPtrs v, RtrState* st, int64_t nvec, int from_state) {
  double* __restrict__ out = from_state ? (st->pt ? v.Y0 : v.Y1) : x;
  const double s = 1.0 / sqrt(st->tmp[0]);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    double2 a = ld2(out + 2 * i);
    a.x *= s;
    a.y *= s;
    st2(out + 2 * i, a);
  }
}

// dst = P_Y(src)  (M.proj / M.tangent)
template <int GS, int VPL, int MF>
__global__ void __launch_bounds__(MSDP_THREADS)
    k_project(const double* Y, const double* src, double* dst, RtrState* st, int64_t nrows, int ld, int use_tmp0) {
  const long long nob = (MF == MF_OBLIQUE) ? st->nob_rows : 0;
  const unsigned mask = group_mask<GS>();
  const int gl = threadIdx.x % GS;
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / GS);
  const int nvec = ld / 2;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x / GS) + threadIdx.x / GS; row < nrows; row += ngroups) {
    const size_t base = (size_t)row * ld;
    double2 x[VPL], y[VPL];
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        x[t] = ld2(src + base + 2 * c);
        y[t] = ld2(Y + base + 2 * c);
        dot += x[t].x * y[t].x + x[t].y * y[t].y;
      }
    }
    if (MF == MF_OBLIQUE) dot = rowsel(group_sum<GS>(dot, mask), row < nob);
    if (MF == MF_SPHERE) dot = use_tmp0 ? st->tmp[0] : 0.0;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const int c = gl + GS * t;
      if (c < nvec) {
        if (MF != MF_EUCLID) {
          x[t].x -= y[t].x * dot;
          x[t].y -= y[t].y * dot;
        }
        st2(dst + base + 2 * c, x[t]);
      }
    }
  }
}

// generic reductions: q0 = <a,b> over nvec double2 -> st->tmp[slot]
__global__ void __launch_bounds__(MSDP_THREADS)
    k_dot(const double* a, const double* b, RtrState* st, double* partials, int64_t nvec, int slot) {
  __shared__ double sm[32];
  double q[1] = {0.0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const double2 x = ld2(a + 2 * i), y = ld2(b + 2 * i);
    q[0] += x.x * y.x + x.y * y.y;
  }
  double tot[1];
  if (grid_sum_last<1>(q, partials, &st->ticket, sm, tot)) {
    if (threadIdx.x == 0) st->tmp[slot] = tot[0];
  }
}

__global__ void k_tcg_after_update_scalar(RtrState* st, cudaGraphConditionalHandle cond, int use_cond) {
  if (st->stop != 0) return;
  tcg_after_update(st, st->tmp, cond, use_cond);
}

// ---- host launchers -------------------------------------------------------------------------------------------------
static inline int flat_grid(const manisdp_handle* h, int64_t nvec) {
  int64_t nb = (nvec + MSDP_THREADS - 1) / MSDP_THREADS;
  int64_t cap = (int64_t)h->num_sms * 8;
  if (cap > MSDP_MAX_BLOCKS) cap = MSDP_MAX_BLOCKS;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

VecPtrs msdp_vecptrs(const manisdp_handle* h) {
  VecPtrs v;
  v.Y0 = h->Ybuf[0];
  v.Y1 = h->Ybuf[1];
  v.G0 = h->Gbuf[0];
  v.G1 = h->Gbuf[1];
  v.eG0 = h->eG[0];
  v.eG1 = h->eG[1];
  v.eta0 = h->eta[0];
  v.eta1 = h->eta[1];
  v.r = h->r;
  v.d = h->d;
  v.Hd = h->Hd;
  return v;
}

int msdp_launch_tcg_init(manisdp_handle* h) {
  const int64_t nvec = h->nloc * h->ld / 2;
  k_tcg_init<<<flat_grid(h, nvec), MSDP_THREADS, 0, h->stream>>>(msdp_vecptrs(h), h->st, nvec);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_launch_tcg_update(manisdp_handle* h, cudaGraphConditionalHandle cond, int use_cond, int defer) {
  const int64_t nvec = h->nloc * h->ld / 2;
  const int nb = flat_grid(h, nvec);
  const VecPtrs v = msdp_vecptrs(h);
  switch (h->mf) {
    case MF_OBLIQUE:
      k_tcg_update<MF_OBLIQUE><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->partials, nvec, cond, use_cond, defer);
      break;
    case MF_SPHERE:
      k_tcg_update<MF_SPHERE><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->partials, nvec, cond, use_cond, defer);
      break;
    default:
      k_tcg_update<MF_EUCLID><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->partials, nvec, cond, use_cond, defer);
  }
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_launch_tcg_after_update_scalar(manisdp_handle* h, cudaGraphConditionalHandle cond, int use_cond) {
  k_tcg_after_update_scalar<<<1, 1, 0, h->stream>>>(h->st, cond, use_cond);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

int msdp_launch_tcg_dir(manisdp_handle* h) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    switch (h->mf) {
      case MF_OBLIQUE:
        k_tcg_dir<GS, VPL, MF_OBLIQUE><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->nloc, ld);
        break;
      case MF_SPHERE:
        k_tcg_dir<GS, VPL, MF_SPHERE><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->nloc, ld);
        break;
      default:
        k_tcg_dir<GS, VPL, MF_EUCLID><<<nb, MSDP_THREADS, 0, h->stream>>>(v, h->st, h->nloc, ld);
    }
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// dst = retr(Y, eta); from_state = 1 uses (Ybuf[pt], eta[eta_cur]) -> Ybuf[pt^1] chosen on the device
int msdp_launch_retract(manisdp_handle* h, const double* Y, const double* eta, double* dst, int from_state) {
  const VecPtrs v = msdp_vecptrs(h);
  const int ld = (int)h->ld;
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    switch (h->mf) {
      case MF_OBLIQUE:
        k_retract<GS, VPL, MF_OBLIQUE>
            <<<nb, MSDP_THREADS, 0, h->stream>>>(Y, eta, dst, v, h->st, h->partials, h->nloc, ld, from_state);
        break;
      case MF_SPHERE:
        k_retract<GS, VPL, MF_SPHERE>
            <<<nb, MSDP_THREADS, 0, h->stream>>>(Y, eta, dst, v, h->st, h->partials, h->nloc, ld, from_state);
        break;
      default:
        k_retract<GS, VPL, MF_EUCLID>
            <<<nb, MSDP_THREADS, 0, h->stream>>>(Y, eta, dst, v, h->st, h->partials, h->nloc, ld, from_state);
    }
  });
  KERNEL_CHECK(h);
  if (h->mf == MF_SPHERE) {
    const int64_t nvec = h->nloc * h->ld / 2;
    k_scale_by_tmp0<<<flat_grid(h, nvec), MSDP_THREADS, 0, h->stream>>>(dst, v, h->st, nvec, from_state);
    KERNEL_CHECK(h);
  }
  return MANISDP_OK;
}

// st->tmp[slot] = <a, b>
int msdp_launch_dot(manisdp_handle* h, const double* a, const double* b, int slot) {
  const int64_t nvec = h->nloc * h->ld / 2;
  k_dot<<<flat_grid(h, nvec), MSDP_THREADS, 0, h->stream>>>(a, b, h->st, h->partials, nvec, slot);
  KERNEL_CHECK(h);
  return MANISDP_OK;
}

// dst = P_Y(src)
int msdp_launch_project(manisdp_handle* h, const double* Y, const double* src, double* dst) {
  const int ld = (int)h->ld;
  if (h->mf == MF_SPHERE) MSDP_TRY(msdp_launch_dot(h, Y, src, 0));
  DISPATCH_GEOM(row_geom(h->ld), {
    const int nb = rows_grid(h, h->nloc, GS);
    switch (h->mf) {
      case MF_OBLIQUE:
        k_project<GS, VPL, MF_OBLIQUE><<<nb, MSDP_THREADS, 0, h->stream>>>(Y, src, dst, h->st, h->nloc, ld, 0);
        break;
      case MF_SPHERE:
        k_project<GS, VPL, MF_SPHERE><<<nb, MSDP_THREADS, 0, h->stream>>>(Y, src, dst, h->st, h->nloc, ld, 1);
        break;
      default:
        k_project<GS, VPL, MF_EUCLID><<<nb, MSDP_THREADS, 0, h->stream>>>(Y, src, dst, h->st, h->nloc, ld, 0);
    }
  });
  KERNEL_CHECK(h);
  return MANISDP_OK;
}
