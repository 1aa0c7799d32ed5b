// small_eig.h -- host-side dense symmetric eigensolver and Cholesky for the SMALL projected problems of the eigen
// step (Rayleigh-Ritz blocks of <= ~64 columns) and of the rank step (p x p Gram, p <= 512).  These replace what the
// reference gets from MATLAB's LAPACK-backed eig / svd on tiny matrices; everything that scales with n stays on the
// device.  Algorithm: Householder tridiagonalisation followed by implicit-shift QL (the classical EISPACK scheme).
#pragma once
#include <math.h>
#include <algorithm>
#include <vector>

// A: n x n symmetric, row-major.  On return V (n x n, row-major) holds eigenvectors in COLUMNS and w the eigenvalues in
// ascending order.  Returns false if QL fails to converge (never observed; 60 sweeps per eigenvalue allowed).
// want_vectors = false: eigenvalues only (skips the accumulation of the Householder reflectors and the rotation
// updates, ~5x cheaper); V is then scratch.
inline bool sym_eig(const std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& V,
                    bool want_vectors = true) {
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
        for (int k = j + 1; k <= i - 1; ++k) {
          g += v(k, j) * w[k];
          e[k] += v(k, j) * f;
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
        for (int k = j; k <= i - 1; ++k) v(k, j) -= (f * e[k] + g * w[k]);
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
        for (int k = 0; k <= i; ++k) g += v(k, i + 1) * v(k, j);
        for (int k = 0; k <= i; ++k) v(k, j) -= g * w[k];
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
          for (int k = 0; want_vectors && k < n; ++k) {
            h = v(k, i + 1);
            v(k, i + 1) = s * v(k, i) + c * h;
            v(k, i) = c * v(k, i) - s * h;
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
