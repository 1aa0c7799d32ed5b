// small_eig.h -- host-side dense symmetric eigensolver and Cholesky for the SMALL projected problems of the eigen
// step (Rayleigh-Ritz blocks of <= ~64 columns) and of the rank step (p x p Gram, p <= 512).  These replace what the
// reference gets from MATLAB's LAPACK-backed eig / svd on tiny matrices; everything that scales with n stays on the
// device.  Algorithm: Householder tridiagonalisation followed by implicit-shift QL (the classical EISPACK scheme).
#pragma once
#include <math.h>
#include <algorithm>
#include <thread>
#include <vector>

// The O(n^3) loops below are unit-stride over columns; MSDP_SIMD lets the host compiler vectorise their reductions
// (build.py passes -fopenmp-simd) and MSDP_EIG_CLONES builds an AVX2/FMA clone next to the baseline x86-64 one, chosen at
// load time by the CPU (the library is built in one container and run on another).
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDA_ARCH__)
#define MSDP_EIG_CLONES __attribute__((target_clones("arch=haswell", "default")))
#else
#define MSDP_EIG_CLONES
#endif
#define MSDP_DO_PRAGMA(x) _Pragma(#x)
#define MSDP_SIMD_SUM(var) MSDP_DO_PRAGMA(omp simd reduction(+ : var))
#define MSDP_SIMD _Pragma("omp simd")

struct EigRotation {
  int i;
  double c, s;
};

// Apply the recorded QL rotations to rows [k0, k1) of the column-major n x n matrix V (column i and i + 1 per rotation)
MSDP_EIG_CLONES inline void sym_eig_apply_rotations(double* V, int n, const std::vector<EigRotation>& rot, int k0, int k1) {
  // Row blocks of RB rows go through the WHOLE rotation sequence (their RB x n slice stays in L1).  Inside a QL sweep the
  // rotations walk down the columns, (i, i+1) then (i-1, i): the updated column i is carried in registers to the next
  // rotation instead of going through a store -> load round trip.
  constexpr int RB = 32;
  const size_t nrot = rot.size();
  for (int kb = k0; kb < k1; kb += RB) {
    const int kr = std::min(RB, k1 - kb);
    double carry[RB];
    size_t t = 0;
    while (t < nrot) {
      int i = rot[t].i;
      {
        const double* __restrict__ b0 = V + (size_t)(i + 1) * n + kb;
        for (int k = 0; k < kr; ++k) carry[k] = b0[k];
      }
      while (true) {
        double* __restrict__ a = V + (size_t)i * n + kb;
        double* __restrict__ b = V + (size_t)(i + 1) * n + kb;
        const double c = rot[t].c, s = rot[t].s;
        if (kr == RB) {
          MSDP_SIMD
          for (int k = 0; k < RB; ++k) {
            const double av = a[k], h = carry[k];
            b[k] = s * av + c * h;
            carry[k] = c * av - s * h;
          }
        } else {
          for (int k = 0; k < kr; ++k) {
            const double av = a[k], h = carry[k];
            b[k] = s * av + c * h;
            carry[k] = c * av - s * h;
          }
        }
        ++t;
        if (t < nrot && rot[t].i == i - 1) {
          --i;
          continue;
        }
        for (int k = 0; k < kr; ++k) a[k] = carry[k];
        break;
      }
    }
  }
}

// A: n x n symmetric, row-major.  On return V (n x n, row-major) holds eigenvectors in COLUMNS and w the eigenvalues in
// ascending order.  Returns false if QL fails to converge (never observed; 60 sweeps per eigenvalue allowed).
// want_vectors = false: eigenvalues only (skips the accumulation of the Householder reflectors and the rotation
// updates, ~5x cheaper); V is then scratch.
// max_threads: upper bound on the host threads of the rotation phase (0: automatic; 1 when the caller already runs one
// decomposition per thread, e.g. the blocks of a multi-block SDP).
MSDP_EIG_CLONES inline bool sym_eig(const std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& V,
                    bool want_vectors = true, int max_threads = 0) {
  V = A;
  w.assign(n, 0.0);
  std::vector<double> e(n, 0.0);
  if (n == 0) return true;
  // internal storage is COLUMN-major (A is symmetric, so V = A is valid either way): the hot loops of both phases
  // run down columns (v(k, j) over k), which is then unit stride; the result is transposed once at the end
  auto v = [&](int i, int j) -> double& { return V[(size_t)j * n + i]; };
  // --- Householder reduction to tridiagonal form, accumulating the transformation in V
  for (int j = 0; j < n; ++j) w[j] = v(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, hh = 0.0;
    for (int k = 0; k < i; ++k) scale += fabs(w[k]);
    if (scale == 0.0) {
      e[i] = w[i - 1];
      for (int j = 0; j < i; ++j) {
        w[j] = v(i - 1, j);
        v(i, j) = 0.0;
        v(j, i) = 0.0;
      }
    } else {
      for (int k = 0; k < i; ++k) {
        w[k] /= scale;
        hh += w[k] * w[k];
      }
      double f = w[i - 1];
      double g = sqrt(hh);
      if (f > 0) g = -g;
      e[i] = scale * g;
      hh -= f * g;
      w[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = w[j];
        v(j, i) = f;
        g = e[j] + v(j, j) * f;
        {
          const double* __restrict__ col = &v(0, j);
          double* __restrict__ ee = e.data();
          const double* __restrict__ ww = w.data();
          double gs = 0.0;
          MSDP_SIMD_SUM(gs)
          for (int k = j + 1; k <= i - 1; ++k) {
            gs += col[k] * ww[k];
            ee[k] += col[k] * f;
          }
          g += gs;
        }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) {
        e[j] /= hh;
        f += e[j] * w[j];
      }
      const double hk = f / (hh + hh);
      for (int j = 0; j < i; ++j) e[j] -= hk * w[j];
      for (int j = 0; j < i; ++j) {
        f = w[j];
        g = e[j];
        {
          double* __restrict__ col = &v(0, j);
          const double* __restrict__ ee = e.data();
          const double* __restrict__ ww = w.data();
          MSDP_SIMD
          for (int k = j; k <= i - 1; ++k) col[k] -= (f * ee[k] + g * ww[k]);
        }
        w[j] = v(i - 1, j);
        v(i, j) = 0.0;
      }
    }
    w[i] = hh;
  }
  for (int i = 0; i < n - 1; ++i) {
    v(n - 1, i) = v(i, i);  // the diagonal of the tridiagonal matrix is parked in the last row
    if (!want_vectors) continue;
    v(i, i) = 1.0;
    const double hh = w[i + 1];
    if (hh != 0.0) {
      for (int k = 0; k <= i; ++k) w[k] = v(k, i + 1) / hh;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        const double* __restrict__ ci = &v(0, i + 1);
        double* __restrict__ cj = &v(0, j);
        const double* __restrict__ ww = w.data();
        MSDP_SIMD_SUM(g)
        for (int k = 0; k <= i; ++k) g += ci[k] * cj[k];
        MSDP_SIMD
        for (int k = 0; k <= i; ++k) cj[k] -= g * ww[k];
      }
    }
    for (int k = 0; k <= i; ++k) v(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) {
    w[j] = v(n - 1, j);
    v(n - 1, j) = 0.0;
  }
  v(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
  // --- implicit QL on the tridiagonal (w = diagonal, e = sub-diagonal)
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  // the rotations of the QL sweeps are recorded and applied to the accumulated transformation afterwards, rows split over
  // a few host threads (every row of V sees the same sequence of rotations, independently of the other rows)
  // small matrices (the <= 48 x 48 Rayleigh-Ritz problems of LOBPCG, thousands per solve): the whole V sits in L1, so each
  // rotation is applied on the spot (two unit-stride columns) instead of being recorded and replayed in row blocks
  const bool direct = n <= 64;
  std::vector<EigRotation> rot;
  if (want_vectors && !direct) rot.reserve((size_t)n * (size_t)n);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, fabs(w[l]) + fabs(e[l]));
    int m = l;
    while (m < n) {
      if (fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m == n) m = n - 1;
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 60) return false;
        double g = w[l];
        double p = (w[l + 1] - g) / (2.0 * e[l]);
        double r = hypot(p, 1.0);
        if (p < 0) r = -r;
        w[l] = e[l] / (p + r);
        w[l + 1] = e[l] * (p + r);
        const double dl1 = w[l + 1];
        double h = g - w[l];
        for (int i = l + 2; i < n; ++i) w[i] -= h;
        f += h;
        p = w[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * w[i] - s * g;
          w[i + 1] = h + s * (c * g + s * w[i]);
          if (want_vectors) {
            if (direct) {
              double* __restrict__ ci = &v(0, i);
              double* __restrict__ cj = &v(0, i + 1);
              MSDP_SIMD
              for (int k = 0; k < n; ++k) {
                const double hk = cj[k];
                cj[k] = s * ci[k] + c * hk;
                ci[k] = c * ci[k] - s * hk;
              }
            } else {
              rot.push_back(EigRotation{i, c, s});
            }
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        w[l] = c * p;
      } while (fabs(e[l]) > eps * tst1);
    }
    w[l] = w[l] + f;
    e[l] = 0.0;
  }
  if (want_vectors && !rot.empty()) {
    int T = 1;
    if ((size_t)n * rot.size() > (size_t)4000000) {
      const unsigned hw = std::thread::hardware_concurrency();
      T = (int)std::max(1u, std::min(8u, hw / 2));
      T = std::min(T, std::max(1, n / 32));
      if (max_threads > 0) T = std::min(T, max_threads);
    }
    if (T <= 1) {
      sym_eig_apply_rotations(V.data(), n, rot, 0, n);
    } else {
      std::vector<std::thread> th;
      const int chunk = (n + T - 1) / T;
      for (int t = 1; t < T; ++t) {
        const int k0 = std::min(n, t * chunk), k1 = std::min(n, k0 + chunk);
        th.emplace_back([&, k0, k1]() { sym_eig_apply_rotations(V.data(), n, rot, k0, k1); });
      }
      sym_eig_apply_rotations(V.data(), n, rot, 0, std::min(n, chunk));
      for (auto& t : th) t.join();
    }
  }
  // --- sort ascending
  for (int i = 0; i < n - 1; ++i) {
    int k = i;
    double p = w[i];
    for (int j = i + 1; j < n; ++j)
      if (w[j] < p) {
        k = j;
        p = w[j];
      }
    if (k != i) {
      w[k] = w[i];
      w[i] = p;
      for (int j = 0; j < n; ++j) std::swap(v(j, i), v(j, k));
    }
  }
  for (int i = 0; i < n; ++i)  // column-major -> row-major
    for (int j = i + 1; j < n; ++j) std::swap(V[(size_t)i * n + j], V[(size_t)j * n + i]);
  return true;
}

// In-place Cholesky G = L L' of an n x n row-major SPD matrix (lower triangle returned, upper zeroed).
// Returns false when a pivot falls below tol * (largest diagonal entry): G is numerically rank deficient.
inline bool cholesky(std::vector<double>& G, int n, double tol) {
  double dmax = 0.0;
  for (int i = 0; i < n; ++i) dmax = std::max(dmax, G[(size_t)i * n + i]);
  for (int j = 0; j < n; ++j) {
    double d = G[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= G[(size_t)j * n + k] * G[(size_t)j * n + k];
    if (!(d > tol * dmax)) return false;
    const double ljj = sqrt(d);
    G[(size_t)j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double s = G[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= G[(size_t)i * n + k] * G[(size_t)j * n + k];
      G[(size_t)i * n + j] = s / ljj;
    }
    for (int i = 0; i < j; ++i) G[(size_t)i * n + j] = 0.0;
  }
  return true;
}
