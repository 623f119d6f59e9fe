// Small dense complex Hermitian algebra for the Rayleigh-Ritz step of the block eigensolver
// (the reference delegates this to hypre's LOBPCG / LAPACK dsygv, meta_material_solver.cpp:3285).
// Sizes are <= 3 * block (a few dozen), so clarity beats blocking.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace bloch_b200 {
namespace dense {

using cplx = std::complex<double>;
using Mat = std::vector<cplx>;   // row-major n x n (or n x m)

// In-place lower Cholesky A = L L^H (upper part ignored / left untouched).  Returns false if a
// pivot is <= tol * max diagonal.
inline bool cholesky(int n, Mat &A, double tol = 1e-14) {
  double dmax = 0;
  for (int i = 0; i < n; i++) dmax = std::max(dmax, A[i * n + i].real());
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j].real();
    for (int k = 0; k < j; k++) d -= std::norm(A[j * n + k]);
    if (!(d > tol * dmax)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      cplx s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * std::conj(A[j * n + k]);
      A[i * n + j] = s / d;
    }
  }
  return true;
}

// B <- L^-1 B L^-H for Hermitian B (full storage), L lower from cholesky()
inline void reduce_to_standard(int n, const Mat &L, Mat &B) {
  // X = L^-1 B  (forward substitution, row oriented: X[i,:] = (B[i,:] - sum_{k<i} L[i,k] X[k,:]) / L[i,i])
  for (int i = 0; i < n; i++) {
    cplx *bi = &B[(size_t)i * n];
    for (int k = 0; k < i; k++) {
      const cplx l = L[(size_t)i * n + k];
      const cplx *bk = &B[(size_t)k * n];
      for (int c = 0; c < n; c++) bi[c] -= l * bk[c];
    }
    const double inv = 1.0 / L[(size_t)i * n + i].real();
    for (int c = 0; c < n; c++) bi[c] *= inv;
  }
  // Y = X L^-H : solve Y L^H = X row by row  (Y[r][i] = (X[r][i] - sum_{k<i} Y[r][k] conj(L[i][k])) / L[i][i])
  for (int r = 0; r < n; r++)
    for (int i = 0; i < n; i++) {
      cplx s = B[r * n + i];
      for (int k = 0; k < i; k++) s -= B[r * n + k] * std::conj(L[i * n + k]);
      B[r * n + i] = s / L[i * n + i].real();
    }
  for (int i = 0; i < n; i++) {   // symmetrise
    B[i * n + i] = B[i * n + i].real();
    for (int j = i + 1; j < n; j++) {
      cplx a = 0.5 * (B[i * n + j] + std::conj(B[j * n + i]));
      B[i * n + j] = a;
      B[j * n + i] = std::conj(a);
    }
  }
}

// C (n x m) <- L^-H C   (back substitution with the conjugate transpose of L)
inline void back_transform(int n, int m, const Mat &L, Mat &C) {
  for (int c = 0; c < m; c++)
    for (int i = n - 1; i >= 0; i--) {
      cplx s = C[i * m + c];
      for (int k = i + 1; k < n; k++) s -= std::conj(L[k * n + i]) * C[k * m + c];
      C[i * m + c] = s / L[i * n + i].real();
    }
}

// Real symmetric tridiagonal QL with implicit shifts; d (diag, n), e (sub-diag, e[i] couples i and
// i+1, e[n-1] unused); Z (n x n real, row-major) is post-multiplied by the rotations (an empty Z
// skips the eigenvector accumulation).
inline bool tridiag_ql(int n, std::vector<double> &d, std::vector<double> &e, std::vector<double> &Z) {
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (int k = 0; k < n && !Z.empty(); k++) {
            double f2 = Z[k * n + i + 1];
            Z[k * n + i + 1] = s * Z[k * n + i] + c * f2;
            Z[k * n + i] = c * Z[k * n + i] - s * f2;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

// Hermitian eigen-decomposition A = V diag(w) V^H, w ascending, V (n x n) columns = eigenvectors.
// Householder tridiagonalisation, phase normalisation to a real tridiagonal, QL.
// want_vectors = false skips the accumulation of Q and of the QL rotations (V is left empty).
inline bool heev(int n, Mat A, std::vector<double> &w, Mat &V, bool want_vectors = true) {
  Mat Q;
  if (want_vectors) {
    Q.assign((size_t)n * n, cplx(0));
    for (int i = 0; i < n; i++) Q[i * n + i] = 1.0;
  }
  std::vector<cplx> v(n), pv(n), qv(n);
  for (int k = 0; k + 2 < n; k++) {
    double xn = 0;
    for (int i = k + 1; i < n; i++) xn += std::norm(A[i * n + k]);
    xn = std::sqrt(xn);
    double tail = 0;
    for (int i = k + 2; i < n; i++) tail += std::norm(A[i * n + k]);
    if (tail < 1e-300) continue;
    cplx x0 = A[(k + 1) * n + k];
    cplx ph = std::abs(x0) > 0 ? x0 / std::abs(x0) : cplx(1.0);
    cplx alpha = -ph * xn;
    double vn = 0;
    for (int i = k + 1; i < n; i++) {
      v[i] = A[i * n + k];
      if (i == k + 1) v[i] -= alpha;
      vn += std::norm(v[i]);
    }
    vn = std::sqrt(vn);
    if (vn < 1e-300) continue;
    for (int i = k + 1; i < n; i++) v[i] /= vn;
    // A <- H A H with H = I - 2 v v^H on the trailing block (and the k-th row/column)
    // column k / row k
    {
      cplx s = 0;
      for (int i = k + 1; i < n; i++) s += std::conj(v[i]) * A[i * n + k];
      for (int i = k + 1; i < n; i++) {
        A[i * n + k] -= 2.0 * v[i] * s;
        A[k * n + i] = std::conj(A[i * n + k]);
      }
    }
    // trailing block: p = A v ; K = v^H p ; q = p - K v ; A -= 2 (v q^H + q v^H)
    for (int i = k + 1; i < n; i++) {
      cplx s = 0;
      for (int j = k + 1; j < n; j++) s += A[i * n + j] * v[j];
      pv[i] = s;
    }
    cplx K = 0;
    for (int i = k + 1; i < n; i++) K += std::conj(v[i]) * pv[i];
    for (int i = k + 1; i < n; i++) qv[i] = pv[i] - K * v[i];
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j < n; j++)
        A[i * n + j] -= 2.0 * (v[i] * std::conj(qv[j]) + qv[i] * std::conj(v[j]));
    // Q <- Q H
    for (int r = 0; r < n && want_vectors; r++) {
      cplx s = 0;
      for (int j = k + 1; j < n; j++) s += Q[r * n + j] * v[j];
      for (int j = k + 1; j < n; j++) Q[r * n + j] -= 2.0 * s * std::conj(v[j]);
    }
  }
  std::vector<double> d(n), e(n, 0.0);
  std::vector<cplx> ph(n);
  ph[0] = 1.0;
  for (int i = 0; i < n; i++) d[i] = A[i * n + i].real();
  for (int i = 0; i + 1 < n; i++) {
    cplx ek = A[(i + 1) * n + i];
    double a = std::abs(ek);
    e[i] = a;
    ph[i + 1] = a > 0 ? ph[i] * ek / a : ph[i];
  }
  std::vector<double> Z;
  if (want_vectors) {
    Z.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) Z[i * n + i] = 1.0;
  }
  if (!tridiag_ql(n, d, e, Z)) return false;
  if (!want_vectors) {
    std::sort(d.begin(), d.end());
    w = d;
    V.clear();
    return true;
  }
  std::vector<int> ord(n);
  for (int i = 0; i < n; i++) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return d[a] < d[b]; });
  w.resize(n);
  V.assign(n * n, cplx(0));
  for (int c = 0; c < n; c++) {
    const int src = ord[c];
    w[c] = d[src];
    for (int r = 0; r < n; r++) {
      cplx s = 0;
      for (int j = 0; j < n; j++) s += Q[r * n + j] * ph[j] * Z[j * n + src];
      V[r * n + c] = s;
    }
  }
  return true;
}

// Same QL iteration with the eigenvector matrix stored TRANSPOSED (Zt[i * n + k] = Z[k][i]): a rotation of
// columns (i, i+1) of Z touches two contiguous rows of Zt, which the compiler vectorises.
inline bool tridiag_ql_t(int n, std::vector<double> &d, std::vector<double> &e, std::vector<double> &Zt) {
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i], b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          double *z0 = &Zt[(size_t)i * n], *z1 = &Zt[(size_t)(i + 1) * n];
          for (int k = 0; k < n; k++) {
            const double f2 = z1[k];
            z1[k] = s * z0[k] + c * f2;
            z0[k] = c * z0[k] - s * f2;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

// The m LOWEST eigenpairs of a Hermitian matrix (w ascending, V is n x m row-major).  Householder
// tridiagonalisation keeping the reflectors instead of accumulating Q, QL on the transposed eigenvector
// matrix, and back-transformation of the m wanted vectors only: ~3x cheaper than heev() for m = n/3, which is
// the Rayleigh-Ritz case (lowest block of a 3-block basis) executed once per outer iteration on the host.
inline bool heev_lowest(int n, int m, Mat A, std::vector<double> &w, Mat &V) {
  Mat Rf((size_t)n * n, cplx(0));            // row k: reflector v_k (entries k+1 .. n-1)
  std::vector<char> has(n, 0);
  std::vector<cplx> v(n), pv(n), qv(n);
  for (int k = 0; k + 2 < n; k++) {
    double xn = 0, tail = 0;
    for (int i = k + 1; i < n; i++) xn += std::norm(A[(size_t)i * n + k]);
    for (int i = k + 2; i < n; i++) tail += std::norm(A[(size_t)i * n + k]);
    xn = std::sqrt(xn);
    if (tail < 1e-300) continue;
    const cplx x0 = A[(size_t)(k + 1) * n + k];
    const cplx ph = std::abs(x0) > 0 ? x0 / std::abs(x0) : cplx(1.0);
    const cplx alpha = -ph * xn;
    double vn = 0;
    for (int i = k + 1; i < n; i++) {
      v[i] = A[(size_t)i * n + k];
      if (i == k + 1) v[i] -= alpha;
      vn += std::norm(v[i]);
    }
    vn = std::sqrt(vn);
    if (vn < 1e-300) continue;
    for (int i = k + 1; i < n; i++) { v[i] /= vn; Rf[(size_t)k * n + i] = v[i]; }
    has[k] = 1;
    {   // column k / row k
      cplx sum = 0;
      for (int i = k + 1; i < n; i++) sum += std::conj(v[i]) * A[(size_t)i * n + k];
      for (int i = k + 1; i < n; i++) {
        A[(size_t)i * n + k] -= 2.0 * v[i] * sum;
        A[(size_t)k * n + i] = std::conj(A[(size_t)i * n + k]);
      }
    }
    // trailing block: p = A v ; K = v^H p ; q = p - K v ; A -= 2 (v q^H + q v^H)
    for (int i = k + 1; i < n; i++) {
      const cplx *row = &A[(size_t)i * n];
      cplx sum = 0;
      for (int j = k + 1; j < n; j++) sum += row[j] * v[j];
      pv[i] = sum;
    }
    cplx K = 0;
    for (int i = k + 1; i < n; i++) K += std::conj(v[i]) * pv[i];
    for (int i = k + 1; i < n; i++) qv[i] = pv[i] - K * v[i];
    for (int i = k + 1; i < n; i++) {
      cplx *row = &A[(size_t)i * n];
      const cplx vi2 = 2.0 * v[i], qi2 = 2.0 * qv[i];
      for (int j = k + 1; j < n; j++) row[j] -= vi2 * std::conj(qv[j]) + qi2 * std::conj(v[j]);
    }
  }
  std::vector<double> d(n), e(n, 0.0);
  std::vector<cplx> ph(n);
  ph[0] = 1.0;
  for (int i = 0; i < n; i++) d[i] = A[(size_t)i * n + i].real();
  for (int i = 0; i + 1 < n; i++) {
    const cplx ek = A[(size_t)(i + 1) * n + i];
    const double a = std::abs(ek);
    e[i] = a;
    ph[i + 1] = a > 0 ? ph[i] * ek / a : ph[i];
  }
  std::vector<double> Zt((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) Zt[(size_t)i * n + i] = 1.0;
  if (!tridiag_ql_t(n, d, e, Zt)) return false;
  std::vector<int> ord(n);
  for (int i = 0; i < n; i++) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return d[a] < d[b]; });
  w.resize(m);
  V.assign((size_t)n * m, cplx(0));
  std::vector<cplx> y(n);
  for (int c = 0; c < m; c++) {
    const int src = ord[c];
    w[c] = d[src];
    for (int r = 0; r < n; r++) y[r] = ph[r] * Zt[(size_t)src * n + r];
    for (int k = n - 3; k >= 0; k--) {
      if (!has[k]) continue;
      const cplx *vk = &Rf[(size_t)k * n];
      cplx sum = 0;
      for (int i = k + 1; i < n; i++) sum += std::conj(vk[i]) * y[i];
      sum *= 2.0;
      for (int i = k + 1; i < n; i++) y[i] -= vk[i] * sum;
    }
    for (int r = 0; r < n; r++) V[(size_t)r * m + c] = y[r];
  }
  return true;
}

// Generalised Hermitian problem GA c = lambda GM c (GM positive semi-definite Gram matrix),
// lowest m pairs, C is n x m (row-major).
//   fast path : diagonal scaling + Cholesky, accepted only if every pivot is > chol_tol (well
//               conditioned basis, error amplification <= 1/sqrt(chol_tol));
//   fallback  : (only if drop_tol > 0) spectral whitening GM = U S U^H, directions with
//               S_i <= drop_tol * S_max are discarded.  The eigensolver does NOT use it: dropping
//               the whole P block and restarting proved more robust on tiny meshes whose
//               constrained space is almost exhausted by [X W P].
// Returns false only if fewer than m independent directions remain or QL fails.
inline bool hegv_lowest(int n, int m, const Mat &GA, const Mat &GM, std::vector<double> &lam, Mat &C,
                        double chol_tol = 1e-9, double drop_tol = 0.0) {
  std::vector<double> sc(n);
  for (int i = 0; i < n; i++) {
    double dii = GM[i * n + i].real();
    if (!(dii > 0)) return false;
    sc[i] = 1.0 / std::sqrt(dii);
  }
  Mat L = GM, B = GA;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { L[i * n + j] *= sc[i] * sc[j]; B[i * n + j] *= sc[i] * sc[j]; }
  const Mat GMs = L, GAs = B;
  std::vector<double> w;
  Mat V;
  if (cholesky(n, L, chol_tol)) {
    reduce_to_standard(n, L, B);
    if (!heev_lowest(n, m, B, w, C)) return false;
    lam.assign(w.begin(), w.begin() + m);
    back_transform(n, m, L, C);
  } else if (drop_tol <= 0.0) {
    return false;   // caller shrinks the basis (drops the P block) and retries
  } else {
    std::vector<double> sig;
    Mat U;
    if (!heev(n, GMs, sig, U)) return false;
    const double smax = sig[n - 1];
    int first = 0;
    while (first < n && !(sig[first] > drop_tol * smax)) first++;
    const int r = n - first;
    if (r < m) return false;
    Mat W((size_t)n * r);                       // whitening basis: columns U_i / sqrt(sig_i)
    for (int i = 0; i < n; i++)
      for (int k = 0; k < r; k++) W[(size_t)i * r + k] = U[(size_t)i * n + first + k] / std::sqrt(sig[first + k]);
    Mat T((size_t)n * r, cplx(0)), R((size_t)r * r, cplx(0));
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const cplx a = GAs[(size_t)i * n + j];
        for (int k = 0; k < r; k++) T[(size_t)i * r + k] += a * W[(size_t)j * r + k];
      }
    for (int i = 0; i < n; i++)
      for (int a = 0; a < r; a++) {
        const cplx wc = std::conj(W[(size_t)i * r + a]);
        for (int k = 0; k < r; k++) R[(size_t)a * r + k] += wc * T[(size_t)i * r + k];
      }
    for (int a = 0; a < r; a++) {
      R[(size_t)a * r + a] = R[(size_t)a * r + a].real();
      for (int k = a + 1; k < r; k++) {
        const cplx x = 0.5 * (R[(size_t)a * r + k] + std::conj(R[(size_t)k * r + a]));
        R[(size_t)a * r + k] = x;
        R[(size_t)k * r + a] = std::conj(x);
      }
    }
    if (!heev(r, R, w, V)) return false;
    lam.assign(w.begin(), w.begin() + m);
    C.assign((size_t)n * m, cplx(0));
    for (int i = 0; i < n; i++)
      for (int j = 0; j < m; j++) {
        cplx sum = 0;
        for (int k = 0; k < r; k++) sum += W[(size_t)i * r + k] * V[(size_t)k * r + j];
        C[(size_t)i * m + j] = sum;
      }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) C[i * m + j] *= sc[i];
  return true;
}


// Lowest m eigenvalues only of GA c = lambda GM c for a LARGE, possibly nearly dependent basis (the
// reduced-basis sweep: a few hundred vectors collected at neighbouring k-points).  Diagonal scaling,
// then Cholesky with diagonal pivoting picks a well-conditioned subset of the basis vectors and stops
// when the largest remaining Schur-complement diagonal falls to drop_tol (a vector whose component
// outside the span of the chosen ones is below sqrt(drop_tol) of its norm adds nothing but noise);
// the pencil restricted to the subset is reduced to standard form and only tridiagonalised.
inline bool hegv_lowest_values(int n, int m, const Mat &GA, const Mat &GM, std::vector<double> &lam,
                               double drop_tol = 1e-10, int *rank_out = nullptr) {
  std::vector<double> sc(n);
  for (int i = 0; i < n; i++) {
    const double dii = GM[(size_t)i * n + i].real();
    if (!(dii > 0)) return false;
    sc[i] = 1.0 / std::sqrt(dii);
  }
  Mat S((size_t)n * n);                        // scaled GM, overwritten column by column with L (pivoted order)
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) S[(size_t)i * n + j] = GM[(size_t)i * n + j] * (sc[i] * sc[j]);
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  std::vector<double> dg(n, 1.0);
  Mat L((size_t)n * n, cplx(0));               // L[row in pivoted order][k]
  int r = 0;
  for (; r < n; r++) {
    int piv = r;
    for (int i = r + 1; i < n; i++)
      if (dg[i] > dg[piv]) piv = i;
    if (!(dg[piv] > drop_tol)) break;
    if (piv != r) {
      std::swap(perm[piv], perm[r]);
      std::swap(dg[piv], dg[r]);
      for (int k = 0; k < r; k++) std::swap(L[(size_t)piv * n + k], L[(size_t)r * n + k]);
    }
    const double lrr = std::sqrt(dg[r]);
    L[(size_t)r * n + r] = lrr;
    for (int i = r + 1; i < n; i++) {
      cplx v = S[(size_t)perm[i] * n + perm[r]];
      for (int k = 0; k < r; k++) v -= L[(size_t)i * n + k] * std::conj(L[(size_t)r * n + k]);
      v /= lrr;
      L[(size_t)i * n + r] = v;
      dg[i] -= std::norm(v);
    }
  }
  if (rank_out) *rank_out = r;
  if (r < m) return false;
  Mat Lr((size_t)r * r, cplx(0)), B((size_t)r * r);
  for (int i = 0; i < r; i++) {
    for (int k = 0; k <= i; k++) Lr[(size_t)i * r + k] = L[(size_t)i * n + k];
    for (int j = 0; j < r; j++)
      B[(size_t)i * r + j] = GA[(size_t)perm[i] * n + perm[j]] * (sc[perm[i]] * sc[perm[j]]);
  }
  reduce_to_standard(r, Lr, B);
  std::vector<double> w;
  Mat V;
  if (!heev(r, B, w, V, false)) return false;
  lam.assign(w.begin(), w.begin() + m);
  return true;
}

}  // namespace dense
}  // namespace bloch_b200
