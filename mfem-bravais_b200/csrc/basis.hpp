// 1-D tables behind the sum-factorised element kernels.
//
// closed 1-D basis c_j : Lagrange on the p+1 Gauss-Lobatto points l_j of [0,1]
// open   1-D basis o_a : Lagrange on the p   Gauss-Legendre points g_a of [0,1]
// (MFEM's default ND/RT/H1 hexahedral bases, see oracle/bloch_oracle.py header).
//
// "Mode" representation of a closed-direction polynomial f (degree p):
//     f~_a = f(g_a), a < p        (values at the Gauss points)
//     f~_p = alpha(f) = (2p+1) int_0^1 f L~_p   (coefficient of the shifted Legendre L~_p)
// In it every 1-D mass matrix is diagonal:  int f g = sum_a w_a f~_a g~_a + f~_p g~_p/(2p+1),
// and the open/closed and open/open products only see the first p entries.  TI is the map
// nodal(closed) -> mode, Dt = (d/dx at Gauss points) o TI^-1 and om the diagonal weights.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

namespace bloch_b200 {

struct Basis1D {
  int p = 0;
  std::vector<double> g, w;      // Gauss-Legendre points / weights on [0,1] (p)
  std::vector<double> l;         // Gauss-Lobatto points on [0,1] (p+1)
  std::vector<double> I;         // [p][p+1]   c_j(g_a)
  std::vector<double> D;         // [p][p+1]   c_j'(g_a)
  std::vector<double> TI;        // [p+1][p+1] nodal -> mode
  std::vector<double> TIinv;     // [p+1][p+1]
  std::vector<double> Dt;        // [p][p+1]   D * TI^-1
  std::vector<double> om;        // [p+1]      w_0..w_{p-1}, 1/(2p+1)
};

namespace detail {
inline void legendre(int n, double x, double &P, double &dP) {   // on [-1,1]
  double p0 = 1.0, p1 = x;
  if (n == 0) { P = 1; dP = 0; return; }
  for (int k = 2; k <= n; k++) {
    double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1; p1 = pk;
  }
  P = p1;
  dP = n * (x * p1 - p0) / (x * x - 1.0);
}
inline void gauss_legendre01(int n, std::vector<double> &x, std::vector<double> &w) {
  x.resize(n); w.resize(n);
  for (int i = 0; i < n; i++) {
    double z = -std::cos(M_PI * (i + 0.75) / (n + 0.5)), P, dP;
    for (int it = 0; it < 100; it++) {
      legendre(n, z, P, dP);
      double dz = P / dP;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    legendre(n, z, P, dP);
    x[i] = 0.5 * (z + 1.0);
    w[i] = 1.0 / ((1.0 - z * z) * dP * dP);     // (2/((1-z^2) P'^2)) / 2
  }
  for (int i = 0; i < n / 2; i++) {              // symmetrise
    double a = 0.5 * (x[i] + 1.0 - x[n - 1 - i]);
    x[i] = a; x[n - 1 - i] = 1.0 - a;
    double b = 0.5 * (w[i] + w[n - 1 - i]);
    w[i] = w[n - 1 - i] = b;
  }
  if (n % 2) x[n / 2] = 0.5;
}
inline void gauss_lobatto01(int n, std::vector<double> &x) {      // n >= 2 points
  x.resize(n);
  x[0] = 0.0; x[n - 1] = 1.0;
  int m = n - 1;                                                   // interior: roots of P'_m
  for (int i = 1; i < n - 1; i++) {
    double z = -std::cos(M_PI * i / m);
    for (int it = 0; it < 100; it++) {
      double P, dP;
      legendre(m, z, P, dP);
      // P'' from the Legendre ODE: (1-z^2) P'' = 2 z P' - m(m+1) P
      double ddP = (2.0 * z * dP - m * (m + 1) * P) / (1.0 - z * z);
      double dz = dP / ddP;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[i] = 0.5 * (z + 1.0);
  }
  for (int i = 0; i < n / 2; i++) {
    double a = 0.5 * (x[i] + 1.0 - x[n - 1 - i]);
    x[i] = a; x[n - 1 - i] = 1.0 - a;
  }
  if (n % 2) x[n / 2] = 0.5;
}
inline void lagrange(const std::vector<double> &nodes, double x, std::vector<double> &v,
                     std::vector<double> &dv) {
  int n = (int)nodes.size();
  v.assign(n, 0.0); dv.assign(n, 0.0);
  for (int i = 0; i < n; i++) {
    double den = 1.0, num = 1.0;
    for (int j = 0; j < n; j++)
      if (j != i) { den *= nodes[i] - nodes[j]; num *= x - nodes[j]; }
    double d = 0.0;
    for (int k = 0; k < n; k++) {
      if (k == i) continue;
      double t = 1.0;
      for (int j = 0; j < n; j++)
        if (j != i && j != k) t *= x - nodes[j];
      d += t;
    }
    v[i] = num / den; dv[i] = d / den;
  }
}
inline void invert(int n, const std::vector<double> &A, std::vector<double> &Ainv) {
  std::vector<long double> a(n * 2 * n, 0.0L);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) a[i * 2 * n + j] = A[i * n + j];
    a[i * 2 * n + n + i] = 1.0L;
  }
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (fabsl(a[r * 2 * n + c]) > fabsl(a[piv * 2 * n + c])) piv = r;
    if (fabsl(a[piv * 2 * n + c]) < 1e-300L) throw std::runtime_error("singular 1-D table");
    if (piv != c)
      for (int j = 0; j < 2 * n; j++) std::swap(a[c * 2 * n + j], a[piv * 2 * n + j]);
    long double d = a[c * 2 * n + c];
    for (int j = 0; j < 2 * n; j++) a[c * 2 * n + j] /= d;
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      long double f = a[r * 2 * n + c];
      if (f != 0.0L)
        for (int j = 0; j < 2 * n; j++) a[r * 2 * n + j] -= f * a[c * 2 * n + j];
    }
  }
  Ainv.resize(n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Ainv[i * n + j] = (double)a[i * 2 * n + n + j];
}
}  // namespace detail

inline Basis1D make_basis(int p) {
  using namespace detail;
  Basis1D B;
  B.p = p;
  const int q = p + 1;
  gauss_legendre01(p, B.g, B.w);
  gauss_lobatto01(q, B.l);
  B.I.resize(p * q); B.D.resize(p * q);
  std::vector<double> v, dv;
  for (int a = 0; a < p; a++) {
    lagrange(B.l, B.g[a], v, dv);
    for (int j = 0; j < q; j++) { B.I[a * q + j] = v[j]; B.D[a * q + j] = dv[j]; }
  }
  B.TI.assign(q * q, 0.0);
  for (int a = 0; a < p; a++)
    for (int j = 0; j < q; j++) B.TI[a * q + j] = B.I[a * q + j];
  // alpha_j = (2p+1) int c_j L~_p with a (p+1)-point Gauss rule (degree 2p: exact)
  std::vector<double> xq, wq;
  gauss_legendre01(q, xq, wq);
  for (int k = 0; k < q; k++) {
    lagrange(B.l, xq[k], v, dv);
    double P, dP;
    legendre(p, 2.0 * xq[k] - 1.0, P, dP);
    for (int j = 0; j < q; j++) B.TI[p * q + j] += (2 * p + 1) * wq[k] * v[j] * P;
  }
  invert(q, B.TI, B.TIinv);
  B.Dt.assign(p * q, 0.0);
  for (int a = 0; a < p; a++)
    for (int r = 0; r < q; r++) {
      double s = 0;
      for (int j = 0; j < q; j++) s += B.D[a * q + j] * B.TIinv[j * q + r];
      B.Dt[a * q + r] = s;
    }
  B.om.resize(q);
  for (int a = 0; a < p; a++) B.om[a] = B.w[a];
  B.om[p] = 1.0 / (2 * p + 1);
  return B;
}

}  // namespace bloch_b200
