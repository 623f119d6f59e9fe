// Geometric h-multigrid for the projector's inner system  S0 phi = G^H M x,  S0 = G^H M1(eps) G
// (H1_p Bloch Laplacian).  The reference solves it with MINRES to 1e-13 per vector per LOBPCG
// iteration (maxwell/maxwell_bloch.cpp:2146-2152, 2280-2290) and that solve dominates its run
// time; Jacobi-PCG needs O(n p) iterations.  The Wigner-Seitz meshes are n^3 subdivisions of a
// few coarse hexes, so n -> n/2 -> ... gives nested H1_p spaces for free:
//   * levels share the element classes (J_l = J_0 * n_0 / n_l); eps is averaged over children;
//   * prolongation = element-wise tensor interpolation parent -> 8 children (1-D matrices
//     c_j((a + l_i)/2)), restriction = its transpose with 1/multiplicity weights;
//   * smoother = Chebyshev(degree 3) in D^-1 S0 on [lmax/4, lmax], symmetric V(1,1) cycle;
//   * coarsest level: dense inverse computed on the host from one block apply of the level
//     operator to the identity (a few dozen unknowns), or a long Chebyshev sweep if it is big;
//   * outer iteration: block PCG (columns independent, scalars on the device) preconditioned by
//     one V-cycle.  Every level operator is the same matrix-free kernel (k_h1_op mode 3).
#include <algorithm>
#include <deque>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <functional>

#include "core.hpp"
#include "dense.hpp"
#include "mg.hpp"

using namespace bloch_b200;
using D2 = double2;

namespace {

constexpr int TPB = 256;
inline unsigned grid_for(long total) {
  long g = (total + TPB - 1) / TPB;
  const long cap = 148L * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

struct Transfer1D { double P[2][kMaxP + 1][kMaxP + 1]; };   // P[a][i][j] = c_j((a + l_i)/2)

__device__ __forceinline__ void parent_of(int e, int n, int &parent, int &a0, int &a1, int &a2) {
  const int n3 = n * n * n, nc = n / 2;
  const int blk = e / n3, rem = e - blk * n3;
  const int ix = rem % n, jy = (rem / n) % n, kz = rem / (n * n);
  parent = blk * nc * nc * nc + ((kz / 2) * nc + jy / 2) * nc + ix / 2;
  a0 = ix & 1; a1 = jy & 1; a2 = kz & 1;
}

// xf[fine dof][v] = interpolation of the parent's nodal polynomial (plain stores, conforming)
template <int P>
__global__ void k_h1_prolong(const __grid_constant__ Transfer1D T, const int32_t *__restrict__ map_f,
                             const int32_t *__restrict__ map_c, int n_elem_f, int n_f,
                             const D2 *__restrict__ xc, D2 *__restrict__ xf, int m) {
  constexpr int Q = P + 1, L = Q * Q * Q;
  const long total = (long)n_elem_f * L * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t % m);
    const long r = t / m;
    const int i = (int)(r % L), e = (int)(r / L);
    const int i0 = i / (Q * Q), i1 = (i / Q) % Q, i2 = i % Q;
    int par, a0, a1, a2;
    parent_of(e, n_f, par, a0, a1, a2);
    const int32_t *mc = map_c + (long)par * L;
    D2 acc = make_double2(0.0, 0.0);
    for (int j0 = 0; j0 < Q; j0++) {
      const double w0 = T.P[a0][i0][j0];
      if (w0 == 0.0) continue;
      for (int j1 = 0; j1 < Q; j1++) {
        const double w1 = w0 * T.P[a1][i1][j1];
        if (w1 == 0.0) continue;
        for (int j2 = 0; j2 < Q; j2++) {
          const double w = w1 * T.P[a2][i2][j2];
          const D2 c = xc[(long)(__ldg(mc + (j0 * Q + j1) * Q + j2) - 1) * m + v];
          acc.x = fma(w, c.x, acc.x); acc.y = fma(w, c.y, acc.y);
        }
      }
    }
    xf[(long)(__ldg(map_f + (long)e * L + i) - 1) * m + v] = acc;
  }
}

// rc[coarse dof][v] += sum_i P[i][j] * rf[fine dof i][v] / multiplicity(fine dof i)
template <int P>
__global__ void k_h1_restrict(const __grid_constant__ Transfer1D T, const int32_t *__restrict__ map_f,
                              const int32_t *__restrict__ map_c, int n_elem_f, int n_f,
                              const double *__restrict__ invmult_f, const D2 *__restrict__ rf,
                              D2 *__restrict__ rc, int m) {
  constexpr int Q = P + 1, L = Q * Q * Q;
  const long total = (long)n_elem_f * L * m;
  double *rcd = reinterpret_cast<double *>(rc);
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t % m);
    const long r = t / m;
    const int j = (int)(r % L), e = (int)(r / L);
    const int j0 = j / (Q * Q), j1 = (j / Q) % Q, j2 = j % Q;
    int par, a0, a1, a2;
    parent_of(e, n_f, par, a0, a1, a2);
    const int32_t *mf = map_f + (long)e * L;
    D2 acc = make_double2(0.0, 0.0);
    for (int i0 = 0; i0 < Q; i0++) {
      const double w0 = T.P[a0][i0][j0];
      if (w0 == 0.0) continue;
      for (int i1 = 0; i1 < Q; i1++) {
        const double w1 = w0 * T.P[a1][i1][j1];
        if (w1 == 0.0) continue;
        for (int i2 = 0; i2 < Q; i2++) {
          const long g = __ldg(mf + (i0 * Q + i1) * Q + i2) - 1;
          const double w = w1 * T.P[a2][i2][j2] * invmult_f[g];
          const D2 c = rf[g * m + v];
          acc.x = fma(w, c.x, acc.x); acc.y = fma(w, c.y, acc.y);
        }
      }
    }
    const long gc = __ldg(map_c + (long)par * L + j) - 1;
    atomicAdd(rcd + 2 * (gc * m + v), acc.x);
    atomicAdd(rcd + 2 * (gc * m + v) + 1, acc.y);
  }
}

// ---- sum-factorised nested-mesh transfers ------------------------------------------------------------------
// One thread per (parent element, vector, re/im part).  A parent edge carries the 2P+1 distinct fine nodes of its
// two children; W[i][j] = c_j(x_i) is the 1-D interpolation from the parent's P+1 nodes to them, and the 3-D
// transfer is W (x) W (x) W: streamed plane by plane it costs (2P+1)(Q^2 Q + (2P+1) Q Q + (2P+1)^2 Q) multiply-adds
// per parent (2604 at order 3) against Q^3 (2P+1)^3 = 21952 of the entry-by-entry forms (the element-wise kernels
// above, the CSR rows), with the Q^3 parent values (prolongation) / accumulators (restriction) in registers.
// Prolongation stores (fine nodes shared by several parents receive the same value from each); restriction
// weights every fine node by 1 / (number of parent positions that hold it) and adds into the cleared coarse vector.
struct TransferW { double W[2 * kMaxP + 1][kMaxP + 1]; };

template <int P>
__device__ __forceinline__ long sf_fine_dof(const int32_t *__restrict__ map_f, int n_f, int blk, int ix, int jy, int kz,
                                            int i0, int i1, int i2) {
  constexpr int Q = P + 1, L = Q * Q * Q;
  const int a0 = i0 > P ? 1 : 0, a1 = i1 > P ? 1 : 0, a2 = i2 > P ? 1 : 0;
  const long e = (long)blk * n_f * n_f * n_f + ((long)(2 * kz + a2) * n_f + (2 * jy + a1)) * n_f + 2 * ix + a0;
  return (long)__ldg(map_f + e * L + ((i0 - P * a0) * Q + (i1 - P * a1)) * Q + (i2 - P * a2)) - 1;
}

template <int P>
__global__ void __launch_bounds__(128) k_h1_prolong_sf(const __grid_constant__ TransferW T, const int32_t *__restrict__ map_f,
                                                       const int32_t *__restrict__ map_c, int n_par, int n_f,
                                                       const double *__restrict__ xc, double *__restrict__ xf, int m) {
  constexpr int Q = P + 1, F = 2 * P + 1, L = Q * Q * Q;
  const int nc = n_f / 2, nc3 = nc * nc * nc, m2 = 2 * m;
  const long total = (long)n_par * m2;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int lane2 = (int)(t % m2);
    const int par = (int)(t / m2);
    const int blk = par / nc3, rem = par - blk * nc3;
    const int ix = rem % nc, jy = (rem / nc) % nc, kz = rem / (nc * nc);
    double c[Q][Q][Q];
#pragma unroll
    for (int j0 = 0; j0 < Q; j0++)
#pragma unroll
      for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++)
          c[j0][j1][j2] = xc[((long)__ldg(map_c + (long)par * L + (j0 * Q + j1) * Q + j2) - 1) * m2 + lane2];
#pragma unroll 1
    for (int i0 = 0; i0 < F; i0++) {
      double pl[Q][Q];
#pragma unroll
      for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) {
          double a = 0.0;
#pragma unroll
          for (int j0 = 0; j0 < Q; j0++) a = fma(T.W[i0][j0], c[j0][j1][j2], a);
          pl[j1][j2] = a;
        }
#pragma unroll
      for (int i1 = 0; i1 < F; i1++) {
        double row[Q];
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) {
          double a = 0.0;
#pragma unroll
          for (int j1 = 0; j1 < Q; j1++) a = fma(T.W[i1][j1], pl[j1][j2], a);
          row[j2] = a;
        }
#pragma unroll
        for (int i2 = 0; i2 < F; i2++) {
          double a = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < Q; j2++) a = fma(T.W[i2][j2], row[j2], a);
          xf[sf_fine_dof<P>(map_f, n_f, blk, ix, jy, kz, i0, i1, i2) * m2 + lane2] = a;
        }
      }
    }
  }
}

template <int P>
__global__ void __launch_bounds__(128) k_h1_restrict_sf(const __grid_constant__ TransferW T, const int32_t *__restrict__ map_f,
                                                        const int32_t *__restrict__ map_c, int n_par, int n_f,
                                                        const double *__restrict__ invmult_par,
                                                        const double *__restrict__ rf, double *__restrict__ rc, int m) {
  constexpr int Q = P + 1, F = 2 * P + 1, L = Q * Q * Q;
  const int nc = n_f / 2, nc3 = nc * nc * nc, m2 = 2 * m;
  const long total = (long)n_par * m2;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int lane2 = (int)(t % m2);
    const int par = (int)(t / m2);
    const int blk = par / nc3, rem = par - blk * nc3;
    const int ix = rem % nc, jy = (rem / nc) % nc, kz = rem / (nc * nc);
    double acc[Q][Q][Q];
#pragma unroll
    for (int j0 = 0; j0 < Q; j0++)
#pragma unroll
      for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) acc[j0][j1][j2] = 0.0;
#pragma unroll 1
    for (int i0 = 0; i0 < F; i0++) {
      double s2[Q][Q];
#pragma unroll
      for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) s2[j1][j2] = 0.0;
#pragma unroll
      for (int i1 = 0; i1 < F; i1++) {
        double s1[Q];
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) s1[j2] = 0.0;
#pragma unroll
        for (int i2 = 0; i2 < F; i2++) {
          const long g = sf_fine_dof<P>(map_f, n_f, blk, ix, jy, kz, i0, i1, i2);
          const double r = rf[g * m2 + lane2] * __ldg(invmult_par + g);
#pragma unroll
          for (int j2 = 0; j2 < Q; j2++) s1[j2] = fma(T.W[i2][j2], r, s1[j2]);
        }
#pragma unroll
        for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
          for (int j2 = 0; j2 < Q; j2++) s2[j1][j2] = fma(T.W[i1][j1], s1[j2], s2[j1][j2]);
      }
#pragma unroll
      for (int j0 = 0; j0 < Q; j0++)
#pragma unroll
        for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
          for (int j2 = 0; j2 < Q; j2++) acc[j0][j1][j2] = fma(T.W[i0][j0], s2[j1][j2], acc[j0][j1][j2]);
    }
#pragma unroll
    for (int j0 = 0; j0 < Q; j0++)
#pragma unroll
      for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++)
          atomicAdd(rc + ((long)__ldg(map_c + (long)par * L + (j0 * Q + j1) * Q + j2) - 1) * m2 + lane2, acc[j0][j1][j2]);
  }
}

inline unsigned grid_sf(long total) {
  long g = (total + 127) / 128;
  const long cap = 148L * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
template <int P>
void prolong_sf_t(const TransferW &T, const int32_t *mf, const int32_t *mc, int nef, int nf, const D2 *xc, D2 *xf, int m, cudaStream_t s) {
  const int n_par = nef / 8;
  k_h1_prolong_sf<P><<<grid_sf((long)n_par * 2 * m), 128, 0, s>>>(T, mf, mc, n_par, nf, reinterpret_cast<const double *>(xc),
                                                                   reinterpret_cast<double *>(xf), m);
}
template <int P>
void restrict_sf_t(const TransferW &T, const int32_t *mf, const int32_t *mc, int nef, int nf, const double *w, const D2 *rf, D2 *rc, int m, cudaStream_t s) {
  const int n_par = nef / 8;
  k_h1_restrict_sf<P><<<grid_sf((long)n_par * 2 * m), 128, 0, s>>>(T, mf, mc, n_par, nf, w, reinterpret_cast<const double *>(rf),
                                                                    reinterpret_cast<double *>(rc), m);
}

// Y[row][v] (+)= sum_k val[k] X[col[k]][v]: the nested-mesh transfers as explicit sparse matrices (real weights,
// complex block vectors).  Prolongation: rows = fine dofs (1 - 27 entries each at p = 2), restriction = its
// transpose: rows = coarse dofs gathering their fine neighbours - no atomics, every output written once, the
// gathered rows are contiguous in v (coalesced) and L2 resident.
__global__ void k_csr_apply(const int *__restrict__ ptr, const int32_t *__restrict__ col, const double *__restrict__ val,
                            const D2 *__restrict__ X, D2 *__restrict__ Y, long nrows, int m, int accumulate) {
  const long total = nrows * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long r = t / m;
    const int v = (int)(t - r * m);
    D2 acc0 = accumulate ? Y[t] : make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
    const int k1 = __ldg(ptr + r + 1);
    int k = __ldg(ptr + r);
    for (; k + 4 <= k1; k += 4) {      // four independent index -> value chains in flight
      int c[4];
      double w[4];
      D2 x[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { c[u] = __ldg(col + k + u); w[u] = __ldg(val + k + u); }
#pragma unroll
      for (int u = 0; u < 4; u++) x[u] = X[(long)c[u] * m + v];
      acc0.x = fma(w[0], x[0].x, acc0.x); acc0.y = fma(w[0], x[0].y, acc0.y);
      acc1.x = fma(w[1], x[1].x, acc1.x); acc1.y = fma(w[1], x[1].y, acc1.y);
      acc0.x = fma(w[2], x[2].x, acc0.x); acc0.y = fma(w[2], x[2].y, acc0.y);
      acc1.x = fma(w[3], x[3].x, acc1.x); acc1.y = fma(w[3], x[3].y, acc1.y);
    }
    for (; k < k1; k++) {
      const double w = __ldg(val + k);
      const D2 x = X[(long)__ldg(col + k) * m + v];
      acc0.x = fma(w, x.x, acc0.x); acc0.y = fma(w, x.y, acc0.y);
    }
    Y[t] = make_double2(acc0.x + acc1.x, acc0.y + acc1.y);
  }
}

__global__ void k_count(const int32_t *__restrict__ map, long total, double *__restrict__ cnt) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x)
    atomicAdd(cnt + (map[t] - 1), 1.0);
}
__global__ void k_invert(double *x, long n) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x)
    x[t] = x[t] > 0 ? 1.0 / x[t] : 0.0;
}
// x[r][v] = sum_c inv_k[r][c] b[c][v], k = v / cpk   (small dense coarse solve; one inverse per k-point)
__global__ void k_dense_apply(const D2 *__restrict__ inv_all, const D2 *__restrict__ b, D2 *__restrict__ x, int n, int m,
                              int cpk) {
  const int total = n * m;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int r = t / m, v = t - r * m;
    const D2 *inv = inv_all + (size_t)(v / cpk) * n * n;
    D2 acc0 = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
    int c = 0;
    for (; c + 8 <= n; c += 8) {               // 8 independent load pairs in flight per step
      D2 a[8], y[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { a[u] = __ldg(inv + (long)r * n + c + u); y[u] = __ldg(b + (long)(c + u) * m + v); }
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        acc0.x = fma(a[u].x, y[u].x, acc0.x); acc0.x = fma(-a[u].y, y[u].y, acc0.x);
        acc0.y = fma(a[u].x, y[u].y, acc0.y); acc0.y = fma(a[u].y, y[u].x, acc0.y);
        acc1.x = fma(a[u + 1].x, y[u + 1].x, acc1.x); acc1.x = fma(-a[u + 1].y, y[u + 1].y, acc1.x);
        acc1.y = fma(a[u + 1].x, y[u + 1].y, acc1.y); acc1.y = fma(a[u + 1].y, y[u + 1].x, acc1.y);
      }
    }
    for (; c < n; c++) {
      const D2 a = inv[(long)r * n + c], y = b[(long)c * m + v];
      acc0.x = fma(a.x, y.x, acc0.x); acc0.x = fma(-a.y, y.y, acc0.x);
      acc0.y = fma(a.x, y.y, acc0.y); acc0.y = fma(a.y, y.x, acc0.y);
    }
    x[t] = make_double2(acc0.x + acc1.x, acc0.y + acc1.y);
  }
}
// r = b - q
__global__ void k_resid(const D2 *__restrict__ b, const D2 *__restrict__ q, D2 *__restrict__ r, long total) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x)
    r[t] = make_double2(b[t].x - q[t].x, b[t].y - q[t].y);
}
// x += y
__global__ void k_add(D2 *__restrict__ x, const D2 *__restrict__ y, long total) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    D2 a = x[t]; const D2 b = y[t];
    a.x += b.x; a.y += b.y;
    x[t] = a;
  }
}
// Chebyshev: d = c0 jac r ; x (+)= d
// (all Chebyshev coefficients are read from device memory - coef[0] = 1/theta, coef[2k-1], coef[2k] for step
// k - so that the captured CUDA graphs stay valid when Setup() changes them with kappa)
__global__ void k_cheb_first(const double *__restrict__ jac, const D2 *__restrict__ r, D2 *__restrict__ d,
                             D2 *__restrict__ x, const double *__restrict__ coef, long n, int m, int accumulate, int nk,
                             int cpk) {
  const double c0 = coef[0];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = c0 * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 v = r[t];
    const D2 o = make_double2(s * v.x, s * v.y);
    d[t] = o;
    if (accumulate) { D2 a = x[t]; a.x += o.x; a.y += o.y; x[t] = a; } else x[t] = o;
  }
}
// r -= q ; d = a d + b jac r ; x += d
__global__ void k_cheb_step(const double *__restrict__ jac, const D2 *__restrict__ q, D2 *__restrict__ r,
                            D2 *__restrict__ d, D2 *__restrict__ x, const double *__restrict__ coef, int step, long n,
                            int m, int nk, int cpk) {
  const double a = coef[2 * step - 1], b = coef[2 * step];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = b * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 qq = q[t];
    D2 rr = r[t];
    rr.x -= qq.x; rr.y -= qq.y;
    r[t] = rr;
    D2 dd = d[t];
    dd.x = a * dd.x + s * rr.x; dd.y = a * dd.y + s * rr.y;
    d[t] = dd;
    D2 xx = x[t];
    xx.x += dd.x; xx.y += dd.y;
    x[t] = xx;
  }
}
// ---- degree-2 Chebyshev smoothing of the plain schedule written out (coef: c0 = coef[0], a = coef[1], b = coef[2]):
// the (r, d, x) recurrence collapses to  x2 = (1 + a) d0 + b D^-1 (rhs - S0 d0),  d0 = c0 D^-1 rhs, which needs
// 2 + 4 vector passes from a zero guess (instead of copy 2 + first 3 + step 7) and 4 + 5 as a correction of x
// (instead of residual 3 + first 4 + step 7). ----
// x = c0 jac b
__global__ void k_sm2_pre1(const double *__restrict__ jac, const D2 *__restrict__ b, D2 *__restrict__ x,
                           const double *__restrict__ coef, long n, int m, int nk, int cpk) {
  const double c0 = coef[0];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = c0 * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 v = b[t];
    x[t] = make_double2(s * v.x, s * v.y);
  }
}
// x = (1 + a) x + b jac (rhs - q),  q = S0 x
__global__ void k_sm2_pre2(const double *__restrict__ jac, const D2 *__restrict__ rhs, const D2 *__restrict__ q,
                           D2 *__restrict__ x, const double *__restrict__ coef, long n, int m, int nk, int cpk) {
  const double a1 = 1.0 + coef[1], bb = coef[2];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = bb * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 r = rhs[t], qq = q[t];
    D2 xx = x[t];
    xx.x = a1 * xx.x + s * (r.x - qq.x);
    xx.y = a1 * xx.y + s * (r.y - qq.y);
    x[t] = xx;
  }
}
// r = b - q ; d = c0 jac r        (q = S0 x)
__global__ void k_sm2_post1(const double *__restrict__ jac, const D2 *__restrict__ b, const D2 *__restrict__ q,
                            D2 *__restrict__ r, D2 *__restrict__ d, const double *__restrict__ coef, long n, int m, int nk,
                            int cpk) {
  const double c0 = coef[0];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = c0 * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 bb = b[t], qq = q[t];
    const D2 rr = make_double2(bb.x - qq.x, bb.y - qq.y);
    r[t] = rr;
    d[t] = make_double2(s * rr.x, s * rr.y);
  }
}
// x += (1 + a) d + b jac (r - q)   (q = S0 d)
__global__ void k_sm2_post2(const double *__restrict__ jac, const D2 *__restrict__ r, const D2 *__restrict__ q,
                            const D2 *__restrict__ d, D2 *__restrict__ x, const double *__restrict__ coef, long n, int m,
                            int nk, int cpk) {
  const double a1 = 1.0 + coef[1], bb = coef[2];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = bb * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 rr = r[t], qq = q[t], dd = d[t];
    D2 xx = x[t];
    xx.x += a1 * dd.x + s * (rr.x - qq.x);
    xx.y += a1 * dd.y + s * (rr.y - qq.y);
    x[t] = xx;
  }
}
__global__ void k_jacobi(const double *d, double *jac, long n) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x)
    jac[t] = d[t] > 0 ? 1.0 / d[t] : 0.0;
}
// column means removed (Gamma point: S0 singular on constants)
__global__ void k_col_sum(const D2 *__restrict__ X, long n, int m, double *__restrict__ sums) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const D2 v = X[t];
    atomicAdd(sums + 2 * (t % m), v.x);
    atomicAdd(sums + 2 * (t % m) + 1, v.y);
  }
}
__global__ void k_col_shift(D2 *__restrict__ X, long n, int m, const double *__restrict__ sums,
                            const int *__restrict__ gflag, int cpk) {
  const long total = n * m;
  const double inv = 1.0 / (double)n;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    if (!gflag[(int)(t % m) / cpk]) continue;
    D2 v = X[t];
    v.x -= sums[2 * (t % m)] * inv; v.y -= sums[2 * (t % m) + 1] * inv;
    X[t] = v;
  }
}


// ---- fused variants (BLOCH_MG_FUSED, default on): fewer, fatter graph nodes --------------------
// The V-cycle at these sizes is a chain of ~70 dependent few-microsecond nodes per PCG iteration,
// i.e. bound by node-to-node latency.  Invariant used below: every operator-output buffer (level q,
// the fine qvec, every coarse b) is ZERO between uses - the kernel that consumes it clears it -
// so no memset node precedes the scatter-add kernels.

// pre-smoothing from a zero guess: r = b ; d = c0 jac b ; x = d
__global__ void k_cheb_first_b(const double *__restrict__ jac, const D2 *__restrict__ b, D2 *__restrict__ r,
                               D2 *__restrict__ d, D2 *__restrict__ x, const double *__restrict__ coef, long n, int m,
                               int nk, int cpk) {
  const double c0 = coef[0];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = c0 * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 v = b[t];
    const D2 o = make_double2(s * v.x, s * v.y);
    r[t] = v; d[t] = o; x[t] = o;
  }
}
// r -= q ; q = 0 ; d = a d + b jac r ; x += d
__global__ void k_cheb_step_z(const double *__restrict__ jac, D2 *__restrict__ q, D2 *__restrict__ r,
                              D2 *__restrict__ d, D2 *__restrict__ x, const double *__restrict__ coef, int step, long n,
                              int m, int nk, int cpk) {
  const double a = coef[2 * step - 1], b = coef[2 * step];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = b * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 qq = q[t];
    q[t] = make_double2(0.0, 0.0);
    D2 rr = r[t];
    rr.x -= qq.x; rr.y -= qq.y;
    r[t] = rr;
    D2 dd = d[t];
    dd.x = a * dd.x + s * rr.x; dd.y = a * dd.y + s * rr.y;
    d[t] = dd;
    D2 xx = x[t];
    xx.x += dd.x; xx.y += dd.y;
    x[t] = xx;
  }
}
// r = b - q ; q = 0
__global__ void k_resid_z(const D2 *__restrict__ b, D2 *__restrict__ q, D2 *__restrict__ r, long total) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const D2 bb = b[t], qq = q[t];
    q[t] = make_double2(0.0, 0.0);
    r[t] = make_double2(bb.x - qq.x, bb.y - qq.y);
  }
}
// post-smoothing start: r = b - q ; q = 0 ; (b = 0) ; d = c0 jac r ; x += d
__global__ void k_resid_cheb_first(const double *__restrict__ jac, D2 *__restrict__ b, D2 *__restrict__ q,
                                   D2 *__restrict__ r, D2 *__restrict__ d, D2 *__restrict__ x,
                                   const double *__restrict__ coef, long n, int m, int clear_b, int nk, int cpk) {
  const double c0 = coef[0];
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = c0 * jac[(t / m) * nk + (int)(t % m) / cpk];
    const D2 bb = b[t], qq = q[t];
    q[t] = make_double2(0.0, 0.0);
    if (clear_b) b[t] = make_double2(0.0, 0.0);
    const D2 rr = make_double2(bb.x - qq.x, bb.y - qq.y);
    r[t] = rr;
    const D2 o = make_double2(s * rr.x, s * rr.y);
    d[t] = o;
    D2 xx = x[t];
    xx.x += o.x; xx.y += o.y;
    x[t] = xx;
  }
}
// xf[g][v] += interpolation of the parent's polynomial at fine dof g, evaluated once per dof through its
// representative local copy rep[g] = e*L + i (the copies agree: conforming spaces).  No atomics, no
// duplicate work.  `zero` (optional): buffer cleared as a side job.
template <int P>
__global__ void k_h1_prolong_rep(const __grid_constant__ Transfer1D T, const int32_t *__restrict__ rep,
                                 const int32_t *__restrict__ map_c, long n0_f, int n_f,
                                 const D2 *__restrict__ xc, D2 *__restrict__ xf, int m, D2 *__restrict__ zero,
                                 long zero_count) {
  constexpr int Q = P + 1, L = Q * Q * Q;
  const long total = n0_f * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t % m);
    const long g = t / m;
    const int ei = __ldg(rep + g);
    const int i = ei % L, e = ei / L;
    const int i0 = i / (Q * Q), i1 = (i / Q) % Q, i2 = i % Q;
    int par, a0, a1, a2;
    parent_of(e, n_f, par, a0, a1, a2);
    const int32_t *mc = map_c + (long)par * L;
    D2 acc = make_double2(0.0, 0.0);
    for (int j0 = 0; j0 < Q; j0++) {
      const double w0 = T.P[a0][i0][j0];
      if (w0 == 0.0) continue;
      for (int j1 = 0; j1 < Q; j1++) {
        const double w1 = w0 * T.P[a1][i1][j1];
        if (w1 == 0.0) continue;
        for (int j2 = 0; j2 < Q; j2++) {
          const double w = w1 * T.P[a2][i2][j2];
          if (w == 0.0) continue;
          const D2 c = xc[(long)(__ldg(mc + (j0 * Q + j1) * Q + j2) - 1) * m + v];
          acc.x = fma(w, c.x, acc.x); acc.y = fma(w, c.y, acc.y);
        }
      }
    }
    D2 x = xf[t];
    x.x += acc.x; x.y += acc.y;
    xf[t] = x;
  }
  if (zero)
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < zero_count; t += (long)gridDim.x * blockDim.x)
      zero[t] = make_double2(0.0, 0.0);
}

// ---- PCG scalars without memset / atomics on global memory: every dot product is written as
// per-block partial sums part[block][m]; consumers add the PCG_BLOCKS partials themselves. ----
constexpr int PCG_BLOCKS = 148 * 4;   // enough threads in flight to stream the large levels from HBM
// all threads take part: thread t adds the partials of column t % m from blocks t / m, t / m + T / m, ...
// (callers __syncthreads() afterwards; tot must not alias other live shared data)
__device__ __forceinline__ void col_totals(const double *__restrict__ part, int m, double *tot /* smem [m] */) {
  for (int j = threadIdx.x; j < m; j += blockDim.x) tot[j] = 0.0;
  __syncthreads();
  const int per = blockDim.x / m;              // threads per column
  if (per >= 1 && (int)threadIdx.x < per * m) {
    const int j = threadIdx.x % m, b0 = threadIdx.x / m;
    double sum = 0.0;
    for (int b = b0; b < PCG_BLOCKS; b += per) sum += part[b * m + j];
    atomicAdd(&tot[j], sum);
  } else if (per < 1) {
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
      double sum = 0.0;
      for (int b = 0; b < PCG_BLOCKS; b++) sum += part[b * m + j];
      tot[j] = sum;
    }
  }
}
// part[block][j] = partial Re <A_j, B_j>
__global__ void __launch_bounds__(TPB) k_dot_part(const D2 *__restrict__ A, const D2 *__restrict__ B, long n, int m,
                                                  double *__restrict__ part) {
  extern __shared__ double sm[];
  const long total = n * m, usable = ((long)PCG_BLOCKS * TPB / m) * m;
  const long start = blockIdx.x * (long)TPB + threadIdx.x;
  for (int j = threadIdx.x; j < m; j += TPB) sm[j] = 0.0;
  __syncthreads();
  if (start < usable) {
    double acc = 0.0;
    for (long t = start; t < total; t += usable) {
      const D2 a = A[t], b = B[t];
      acc = fma(a.x, b.x, acc);
      acc = fma(a.y, b.y, acc);
    }
    atomicAdd(&sm[start % m], acc);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < m; j += TPB) part[blockIdx.x * m + j] = sm[j];
}
// alpha = <r,z> / <p,q> ; phi += alpha p ; r -= alpha q ; q = 0 ; part_rr = partial |r|^2 ; block 0 saves <r,z>
__global__ void __launch_bounds__(TPB) k_pcg_a(const D2 *__restrict__ p, D2 *__restrict__ q, D2 *__restrict__ phi,
                                               D2 *__restrict__ r, const double *__restrict__ part_rz,
                                               const double *__restrict__ part_pq, double *__restrict__ part_rr,
                                               double *__restrict__ rz_saved, long n, int m) {
  extern __shared__ double sm[];            // rz[m] | pq[m] | rr[m]
  double *trz = sm, *tpq = sm + m, *trr = sm + 2 * m;
  col_totals(part_rz, m, trz);
  col_totals(part_pq, m, tpq);
  for (int j = threadIdx.x; j < m; j += TPB) trr[j] = 0.0;
  __syncthreads();
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < m; j += TPB) rz_saved[j] = trz[j];
  const long total = n * m, usable = ((long)PCG_BLOCKS * TPB / m) * m;
  const long start = blockIdx.x * (long)TPB + threadIdx.x;
  if (start < usable) {
    const int j = (int)(start % m);
    const double alpha = tpq[j] != 0.0 ? trz[j] / tpq[j] : 0.0;
    double acc = 0.0;
    for (long t = start; t < total; t += usable) {
      const D2 pp = p[t], qq = q[t];
      q[t] = make_double2(0.0, 0.0);
      D2 x = phi[t];
      x.x = fma(alpha, pp.x, x.x); x.y = fma(alpha, pp.y, x.y);
      phi[t] = x;
      D2 rr = r[t];
      rr.x = fma(-alpha, qq.x, rr.x); rr.y = fma(-alpha, qq.y, rr.y);
      r[t] = rr;
      acc = fma(rr.x, rr.x, acc);
      acc = fma(rr.y, rr.y, acc);
    }
    atomicAdd(&trr[j], acc);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < m; j += TPB) part_rr[blockIdx.x * m + j] = trr[j];
}
// beta = <r,z>_new / <r,z>_old ; p = z + beta p
__global__ void __launch_bounds__(TPB) k_pcg_b(const D2 *__restrict__ z, D2 *__restrict__ p,
                                               const double *__restrict__ part_rzn, const double *__restrict__ rz_saved,
                                               long n, int m) {
  extern __shared__ double sm[];
  col_totals(part_rzn, m, sm);
  __syncthreads();
  const long total = n * m, usable = ((long)PCG_BLOCKS * TPB / m) * m;
  const long start = blockIdx.x * (long)TPB + threadIdx.x;
  if (start < usable) {
    const int j = (int)(start % m);
    const double beta = rz_saved[j] != 0.0 ? sm[j] / rz_saved[j] : 0.0;
    for (long t = start; t < total; t += usable) {
      const D2 zz = z[t];
      D2 pp = p[t];
      pp.x = fma(beta, pp.x, zz.x); pp.y = fma(beta, pp.y, zz.y);
      p[t] = pp;
    }
  }
}

template <int P>
void prolong_rep_t(const Transfer1D &T, const int32_t *rep, const int32_t *mc, long n0f, int nf, const D2 *xc, D2 *xf,
                   int m, D2 *zero, long zero_count, cudaStream_t s) {
  k_h1_prolong_rep<P><<<grid_for(n0f * m), TPB, 0, s>>>(T, rep, mc, n0f, nf, xc, xf, m, zero, zero_count);
}

template <int P>
void prolong_t(const Transfer1D &T, const int32_t *mf, const int32_t *mc, int nef, int nf, const D2 *xc, D2 *xf, int m, cudaStream_t s) {
  constexpr int L = (P + 1) * (P + 1) * (P + 1);
  k_h1_prolong<P><<<grid_for((long)nef * L * m), TPB, 0, s>>>(T, mf, mc, nef, nf, xc, xf, m);
}
template <int P>
void restrict_t(const Transfer1D &T, const int32_t *mf, const int32_t *mc, int nef, int nf, const double *w, const D2 *rf, D2 *rc, int m, cudaStream_t s) {
  constexpr int L = (P + 1) * (P + 1) * (P + 1);
  k_h1_restrict<P><<<grid_for((long)nef * L * m), TPB, 0, s>>>(T, mf, mc, nef, nf, w, rf, rc, m);
}

double env_double(const char *name, double dflt) {
  const char *s = std::getenv(name);
  return s ? std::atof(s) : dflt;
}

}  // namespace

namespace bloch_b200 {

struct H1Level {
  int n = 0;
  long N0 = 0;
  HexMesh mesh;                       // levels >= 1 (level 0 lives in the handle)
  DofMaps maps;
  DevBuf<int32_t> map_h1;
  DevBuf<int> cls;
  DevBuf<double> eps, cpar, diag, jac, invmult;
  DevBuf<double> invmult_par;         // 1 / (number of parent-element positions holding the dof): sum-factorised restriction
  std::vector<double> eps_h;
  ElemData E{};
  double lmax = 0;
  DevBuf<D2> x, b, r, d, q;           // V-cycle work vectors, N0 x m
  DevBuf<D2> inv, dI, dA;             // dense inverse on the coarsest level (+ identity / operator scratch)
  DevBuf<double> dloc;                // element-local diagonals (scratch)
  DevBuf<double> cheb, cheb40;        // Chebyshev coefficient tables (smoother / coarsest-level fallback)
  DevBuf<int> tp_ptr;                 // transpose of map_h1: dof -> its local copies (atomic-free apply)
  DevBuf<int32_t> tp_loc;
  DevBuf<D2> evec;                    // E-vector [n_elem * L][m]
  bool use_evec = false;
  DevBuf<D2> Sloc;                    // dense element matrices per class [n_class][L][L] (small levels, p <= 2)
  bool have_S = false;
  DevBuf<int32_t> rep;                // one (element, local index) copy e*L + i per dof (owner of the prolongation)
  DevBuf<int> P_ptr, R_ptr;           // CSR of the prolongation from the next coarser level (rows = this level's dofs)
  DevBuf<int32_t> P_col, R_col;       // and of its transpose, the restriction (rows = coarser level's dofs)
  DevBuf<double> P_val, R_val;
  bool have_csr = false;
  bool dense = false;
};

struct H1Multigrid {
  int kind = 0;                       // level operators: 0 = eps-Laplacian (S0 of the projector), 1 = mu^-1-Laplacian (auxiliary space of the
                                      // ND preconditioner), 2 = eps-slot Laplacian as preconditioner of the scalar H1 problem; kinds 1, 2 carry the
                                      // sigma * mass shift on the constant mode of the coarsest level
  int smooth_degree = 2;              // Chebyshev-Jacobi smoother degree (BLOCH_MG_SMOOTH_DEGREE / BLOCH_AUX_MG_DEGREE)
  std::deque<H1Level> lev;
  Transfer1D T;
  TransferW TW;
  int m_alloc = 0;
  DevBuf<D2> z, pvec, qvec;           // PCG vectors on the fine level
  DevBuf<double> scal;                // rz | pq | rz_new | rr | alpha | beta  (m each) + 2m sums
  DevBuf<double> part;                // fused path: per-block partial dots rz | pq | rr (PCG_BLOCKS x m each) + saved <r,z>
  int zero_m = 0;
  bool zero_valid = false;            // fused path: operator-output buffers are known to be zero
  // CUDA graphs of the two halves of one PCG iteration (re-captured after every mg_setup: the
  // Chebyshev coefficients are kernel arguments captured by value)
  cudaGraphExec_t gA = nullptr, gB = nullptr;
  const void *g_rhs = nullptr, *g_phi = nullptr;
  int g_m = 0, g_nk = 0;
  int nodes_A = 0, nodes_B = 0;       // kernels inside each graph (added to the launch counter per graph launch)
  int nk_built = 0;
  void drop_graphs() {
    if (gA) cudaGraphExecDestroy(gA);
    if (gB) cudaGraphExecDestroy(gB);
    gA = gB = nullptr;
  }
  ~H1Multigrid() { drop_graphs(); }
};

static void alloc_work(H1Multigrid *mg, int m) {
  if (m <= mg->m_alloc) return;
  for (auto &L : mg->lev) {
    const size_t sz = (size_t)L.N0 * m;
    L.x.alloc(sz); L.b.alloc(sz); L.r.alloc(sz); L.d.alloc(sz); L.q.alloc(sz);
    if (L.use_evec) L.evec.alloc(L.tp_loc.n * (size_t)m);
  }
  const size_t sz0 = (size_t)mg->lev[0].N0 * m;
  mg->z.alloc(sz0); mg->pvec.alloc(sz0); mg->qvec.alloc(sz0);
  mg->scal.alloc(8 * (size_t)m);
  mg->part.alloc((size_t)4 * PCG_BLOCKS * m + m);
  mg->m_alloc = m;
  mg->zero_valid = false;
}

void launch_csr_apply(const int *ptr, const int32_t *col, const double *val, const D2 *X, D2 *Y, long nrows, int m,
                      int accumulate, cudaStream_t s) {
  k_csr_apply<<<grid_for(nrows * m), TPB, 0, s>>>(ptr, col, val, X, Y, nrows, m, accumulate);
  BLOCH_CUDA(cudaGetLastError());
}

H1Multigrid *mg_create(bloch_handle_s *h, int kind) {
  const int p = h->p, Q = p + 1;
  if (h->mesh.n_sub % 2 != 0 || h->mesh.n_sub < 2) return nullptr;   // no coarser nested mesh
  H1Multigrid *mg = new H1Multigrid();
  mg->kind = kind;
  mg->smooth_degree = kind == 1 ? (int)env_double("BLOCH_AUX_MG_DEGREE", 2.0) : (int)env_double("BLOCH_MG_SMOOTH_DEGREE", 2);
  cudaStream_t s = h->stream;
  // 1-D transfer tables
  for (int a = 0; a < 2; a++)
    for (int i = 0; i < Q; i++) {
      std::vector<double> v, dv;
      detail::lagrange(h->basis.l, 0.5 * (a + h->basis.l[i]), v, dv);
      for (int j = 0; j < Q; j++) mg->T.P[a][i][j] = std::fabs(v[j]) < 1e-15 ? 0.0 : v[j];
    }
  for (int i = 0; i <= 2 * p; i++)
    for (int j = 0; j < Q; j++) mg->TW.W[i][j] = i <= p ? mg->T.P[0][i][j] : mg->T.P[1][i - p][j];
  // level 0 = the handle's own mesh
  int n = h->mesh.n_sub;
  mg->lev.emplace_back();
  {
    H1Level &L = mg->lev.back();
    L.n = n; L.N0 = h->N0;
  }
  while (n % 2 == 0) {
    n /= 2;
    mg->lev.emplace_back();
    H1Level &L = mg->lev.back();
    L.n = n;
    build_mesh(h->coarse_vert, h->coarse_hex, h->mesh.rec.data(), n, L.mesh);
    build_dofmaps(L.mesh, p, L.maps);
    L.N0 = L.maps.n_h1;
    const int ne = L.mesh.n_elem, LH = L.maps.l_h1;
    std::vector<int32_t> kh1((size_t)ne * LH);
    for (int e = 0; e < ne; e++)
      for (int i0 = 0; i0 < Q; i0++)
        for (int i1 = 0; i1 < Q; i1++)
          for (int i2 = 0; i2 < Q; i2++)
            kh1[(size_t)e * LH + (i0 * Q + i1) * Q + i2] = L.maps.h1[(size_t)e * LH + i0 + Q * (i1 + Q * i2)];
    L.map_h1.upload(kh1, s);
    L.cls.upload(L.mesh.cls, s);
    h_sync(s);
    if (L.N0 <= 8) break;
  }
  // transpose maps for the atomic-free operator apply (element-local results + owner reduction)
  {
    static const double evec_min = env_double("BLOCH_H1_EVEC_MIN_ELEMS", 1024.0);
    static const double evec_max = env_double("BLOCH_H1_EVEC_MAX_ELEMS", 4096.0);   // beyond: E-vector leaves L2
    static const bool evec_on = env_double("BLOCH_H1_EVEC", 1.0) != 0.0;
    for (size_t l = 0; l < mg->lev.size(); l++) {
      H1Level &L = mg->lev[l];
      const std::vector<int32_t> &nat = l == 0 ? h->maps.h1 : L.maps.h1;
      const long ne = l == 0 ? h->mesh.n_elem : L.mesh.n_elem;
      const int LH = h->L_h1;
      L.use_evec = evec_on && p <= 2 && ne >= (long)evec_min && ne <= (long)evec_max;
      // kernel local order k = (i0*Q + i1)*Q + i2  <->  natural i0 + Q*(i1 + Q*i2)
      auto gid = [&](long e, int k) {
        const int i0 = k / (Q * Q), i1 = (k / Q) % Q, i2 = k % Q;
        return (long)nat[(size_t)e * LH + i0 + Q * (i1 + Q * i2)] - 1;
      };
      {   // representative copy of every dof (prolongation owner)
        std::vector<int32_t> rep(L.N0, 0);
        for (long e = 0; e < ne; e++)
          for (int k = 0; k < LH; k++) rep[gid(e, k)] = (int32_t)(e * LH + k);
        L.rep.upload(rep, s);
        h_sync(s);
      }
      if (!L.use_evec) continue;
      std::vector<int> ptr(L.N0 + 1, 0);
      for (long e = 0; e < ne; e++)
        for (int k = 0; k < LH; k++) ptr[gid(e, k) + 1]++;
      for (long g = 0; g < L.N0; g++) ptr[g + 1] += ptr[g];
      std::vector<int32_t> loc((size_t)ne * LH);
      std::vector<int> fill(ptr.begin(), ptr.end() - 1);
      for (long e = 0; e < ne; e++)
        for (int k = 0; k < LH; k++) loc[fill[gid(e, k)]++] = (int32_t)(e * LH + k + 1);
      L.tp_ptr.upload(ptr, s);
      L.tp_loc.upload(loc, s);
      h_sync(s);
    }
  }
  // explicit transfer matrices between consecutive levels (kappa independent, built once)
  for (size_t l = 0; l + 1 < mg->lev.size(); l++) {
    H1Level &F = mg->lev[l], &Cl = mg->lev[l + 1];
    const std::vector<int32_t> &nat_f = l == 0 ? h->maps.h1 : F.maps.h1;
    const std::vector<int32_t> &nat_c = Cl.maps.h1;
    const long ne = l == 0 ? h->mesh.n_elem : F.mesh.n_elem;
    const int LH = h->L_h1, nf = F.n, nc = nf / 2, n3 = nf * nf * nf;
    std::vector<int32_t> rep(F.N0, -1);
    for (long e = 0; e < ne; e++)
      for (int k = 0; k < LH; k++) {
        const int i0 = k / (Q * Q), i1 = (k / Q) % Q, i2 = k % Q;
        rep[nat_f[(size_t)e * LH + i0 + Q * (i1 + Q * i2)] - 1] = (int32_t)(e * LH + k);
      }
    std::vector<int> pptr(F.N0 + 1, 0);
    std::vector<int32_t> pcol;
    std::vector<double> pval;
    for (long g = 0; g < F.N0; g++) {
      const int ei = rep[g], k = ei % LH, e = ei / LH;
      const int i0 = k / (Q * Q), i1 = (k / Q) % Q, i2 = k % Q;
      const int blk = e / n3, rem = e - blk * n3;
      const int ix = rem % nf, jy = (rem / nf) % nf, kz = rem / (nf * nf);
      const int par = blk * nc * nc * nc + ((kz / 2) * nc + jy / 2) * nc + ix / 2;
      const int a0 = ix & 1, a1 = jy & 1, a2 = kz & 1;
      for (int j0 = 0; j0 < Q; j0++)
        for (int j1 = 0; j1 < Q; j1++)
          for (int j2 = 0; j2 < Q; j2++) {
            const double w = mg->T.P[a0][i0][j0] * mg->T.P[a1][i1][j1] * mg->T.P[a2][i2][j2];
            if (w == 0.0) continue;
            pcol.push_back(nat_c[(size_t)par * LH + j0 + Q * (j1 + Q * j2)] - 1);
            pval.push_back(w);
          }
      pptr[g + 1] = (int)pcol.size();
    }
    // transpose (counting sort by coarse dof)
    std::vector<int> rptr(Cl.N0 + 1, 0);
    for (int32_t c : pcol) rptr[c + 1]++;
    for (long c = 0; c < Cl.N0; c++) rptr[c + 1] += rptr[c];
    std::vector<int32_t> rcol(pcol.size());
    std::vector<double> rval(pcol.size());
    std::vector<int> fill(rptr.begin(), rptr.end() - 1);
    for (long g = 0; g < F.N0; g++)
      for (int k = pptr[g]; k < pptr[g + 1]; k++) {
        const int pos = fill[pcol[k]]++;
        rcol[pos] = (int32_t)g;
        rval[pos] = pval[k];
      }
    F.P_ptr.upload(pptr, s); F.P_col.upload(pcol, s); F.P_val.upload(pval, s);
    F.R_ptr.upload(rptr, s); F.R_col.upload(rcol, s); F.R_val.upload(rval, s);
    h_sync(s);
    F.have_csr = true;
  }
  // parent-position multiplicities for the sum-factorised restriction
  for (size_t l = 0; l + 1 < mg->lev.size(); l++) {
    H1Level &F = mg->lev[l];
    const std::vector<int32_t> &nat_f = l == 0 ? h->maps.h1 : F.maps.h1;
    const long ne = l == 0 ? h->mesh.n_elem : F.mesh.n_elem;
    const int LH = h->L_h1, nf = F.n, nc = nf / 2, nc3 = nc * nc * nc, FF = 2 * p + 1;
    std::vector<double> cnt(F.N0, 0.0);
    for (long par = 0; par < ne / 8; par++) {
      const int blk = (int)(par / nc3), rem = (int)(par - (long)blk * nc3);
      const int ix = rem % nc, jy = (rem / nc) % nc, kz = rem / (nc * nc);
      for (int i0 = 0; i0 < FF; i0++)
        for (int i1 = 0; i1 < FF; i1++)
          for (int i2 = 0; i2 < FF; i2++) {
            const int a0 = i0 > p, a1 = i1 > p, a2 = i2 > p;
            const long e = (long)blk * nf * nf * nf + ((long)(2 * kz + a2) * nf + (2 * jy + a1)) * nf + 2 * ix + a0;
            cnt[nat_f[(size_t)e * LH + (i0 - p * a0) + Q * ((i1 - p * a1) + Q * (i2 - p * a2))] - 1] += 1.0;
          }
    }
    for (double &c : cnt) c = c > 0.0 ? 1.0 / c : 0.0;
    F.invmult_par.upload(cnt, s);
    h_sync(s);
  }
  // multiplicity weights of every level that restricts (all but the coarsest)
  for (size_t l = 0; l + 1 < mg->lev.size(); l++) {
    H1Level &L = mg->lev[l];
    const int32_t *map = l == 0 ? h->d_map_h1.p : L.map_h1.p;
    const long ne = l == 0 ? h->mesh.n_elem : L.mesh.n_elem;
    L.invmult.alloc(L.N0);
    BLOCH_CUDA(cudaMemsetAsync(L.invmult.p, 0, sizeof(double) * L.N0, s));
    k_count<<<grid_for(ne * h->L_h1), TPB, 0, s>>>(map, ne * h->L_h1, L.invmult.p);
    k_invert<<<grid_for(L.N0), TPB, 0, s>>>(L.invmult.p, L.N0);
  }
  h_sync(s);
  return mg;
}

void mg_destroy(H1Multigrid *mg) { delete mg; }

long mg_level_size(H1Multigrid *mg, int level) { return level >= 0 && level < (int)mg->lev.size() ? mg->lev[level].N0 : -1; }

// test hook: the transfer between level 0 and level 1 in each of its three implementations
// (variant 0: explicit CSR matrices, 1: sum-factorised parent kernels, 2: element-wise kernels);
// dir 0: y(fine) = P x(coarse), dir 1: y(coarse) = P^T x(fine)
void mg_debug_transfer(H1Multigrid *mg, bloch_handle_s *h, int variant, int dir, const D2 *x, D2 *y, int m) {
  if (mg->lev.size() < 2) throw std::invalid_argument("no coarse level");
  H1Level &L = mg->lev[0], &C = mg->lev[1];
  cudaStream_t s = h->stream;
  const int32_t *map_f = h->d_map_h1.p, *map_c = C.map_h1.p;
  const int nef = h->mesh.n_elem;
  if (dir == 0) {
    if (variant == 0) {
      k_csr_apply<<<grid_for(L.N0 * m), TPB, 0, s>>>(L.P_ptr.p, L.P_col.p, L.P_val.p, x, y, L.N0, m, 0);
    } else if (variant == 1) {
      switch (h->p) {
        case 1: prolong_sf_t<1>(mg->TW, map_f, map_c, nef, L.n, x, y, m, s); break;
        case 2: prolong_sf_t<2>(mg->TW, map_f, map_c, nef, L.n, x, y, m, s); break;
        case 3: prolong_sf_t<3>(mg->TW, map_f, map_c, nef, L.n, x, y, m, s); break;
        default: prolong_sf_t<4>(mg->TW, map_f, map_c, nef, L.n, x, y, m, s); break;
      }
    } else {
      switch (h->p) {
        case 1: prolong_t<1>(mg->T, map_f, map_c, nef, L.n, x, y, m, s); break;
        case 2: prolong_t<2>(mg->T, map_f, map_c, nef, L.n, x, y, m, s); break;
        case 3: prolong_t<3>(mg->T, map_f, map_c, nef, L.n, x, y, m, s); break;
        default: prolong_t<4>(mg->T, map_f, map_c, nef, L.n, x, y, m, s); break;
      }
    }
  } else {
    BLOCH_CUDA(cudaMemsetAsync(y, 0, sizeof(D2) * C.N0 * m, s));
    if (variant == 0) {
      k_csr_apply<<<grid_for(C.N0 * m), TPB, 0, s>>>(L.R_ptr.p, L.R_col.p, L.R_val.p, x, y, C.N0, m, 0);
    } else if (variant == 1) {
      switch (h->p) {
        case 1: restrict_sf_t<1>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, x, y, m, s); break;
        case 2: restrict_sf_t<2>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, x, y, m, s); break;
        case 3: restrict_sf_t<3>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, x, y, m, s); break;
        default: restrict_sf_t<4>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, x, y, m, s); break;
      }
    } else {
      switch (h->p) {
        case 1: restrict_t<1>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, x, y, m, s); break;
        case 2: restrict_t<2>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, x, y, m, s); break;
        case 3: restrict_t<3>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, x, y, m, s); break;
        default: restrict_t<4>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, x, y, m, s); break;
      }
    }
  }
  BLOCH_CUDA(cudaGetLastError());
}

static std::vector<double> cheb_coefs(double lmax, double ratio, int degree);

// element-local diagonal of S0 per class on one level (probe launch of the production kernel)
static void probe_level(bloch_handle_s *h, const ElemData &Elev, std::vector<double> &dloc, double *bound,
                        std::vector<D2> *Sloc = nullptr);

void mg_setup(H1Multigrid *mg, bloch_handle_s *h) {
  cudaStream_t s = h->stream;
  // the captured graphs survive: everything kappa-dependent they use lives in device buffers that are
  // rewritten in place (class tables, Jacobi diagonals, Chebyshev tables, coarse inverse)
  const bool was_dense = mg->lev.back().dense;
  const int p = h->p;
  const int nl = (int)mg->lev.size();
  for (int l = 0; l < nl; l++) {
    H1Level &L = mg->lev[l];
    const HexMesh &mesh = l == 0 ? h->mesh : L.mesh;
    // coefficient: level 0 = the user's eps, coarser = mean over the 8 children
    if (l == 0) {
      L.eps_h = mg->kind == 1 ? h->muinv : h->eps;
    } else {
      const H1Level &F = mg->lev[l - 1];
      const int nf = F.n, nc = L.n, n3 = nf * nf * nf;
      L.eps_h.assign(mesh.n_elem, 0.0);
      for (int e = 0; e < (int)F.eps_h.size(); e++) {
        const int blk = e / n3, rem = e - blk * n3;
        const int ix = rem % nf, jy = (rem / nf) % nf, kz = rem / (nf * nf);
        L.eps_h[blk * nc * nc * nc + ((kz / 2) * nc + jy / 2) * nc + ix / 2] += 0.125 * F.eps_h[e];
      }
      L.eps.upload(L.eps_h, s);
    }
    // kappa-dependent class tables with this level's Jacobians
    const int nk = h->nk;
    std::vector<double> cp((size_t)nk * mesh.n_class * kClassParDoubles);
    for (int k = 0; k < nk; k++)
      for (int c = 0; c < mesh.n_class; c++)
        class_params(&mesh.J[9 * c], &h->kappas[3 * k], &cp[((size_t)k * mesh.n_class + c) * kClassParDoubles]);
    L.cpar.upload(cp, s);
    ElemData &E = L.E;
    E = h->E;
    E.n_elem = mesh.n_elem;
    E.n_class = mesh.n_class;
    E.nk = nk;
    E.cpar = L.cpar.p;
    if (l == 0 && mg->kind == 1) E.eps = h->d_muinv.p;   // stiffness coefficient slot of the scalar kernels
    if (l > 0) {
      E.cls = L.cls.p;
      E.eps = L.eps.p;
      E.muinv = L.eps.p;
      E.map_h1 = L.map_h1.p;
      E.map_nd = nullptr;
      E.map_rt = nullptr;
    }
    // Jacobi diagonal and the rigorous local bound of lambda_max(D^-1 S0)
    std::vector<double> dloc;
    double bound = 0.0;
    static const double dense_max = env_double("BLOCH_H1_DENSE_MAX_ELEMS", 1023.0);
    std::vector<D2> Sl;
    probe_level(h, E, dloc, &bound, (p <= 2 && E.n_elem <= (int)dense_max) ? &Sl : nullptr);
    L.have_S = !Sl.empty();
    if (L.have_S) L.Sloc.upload(Sl, s);
    L.lmax = 1.05 * bound;
    L.cheb.upload(cheb_coefs(L.lmax, env_double("BLOCH_MG_SMOOTH_RATIO", 5.0), mg->smooth_degree), s);
    if (l == nl - 1) L.cheb40.upload(cheb_coefs(L.lmax, 2000.0, 40), s);
    DevBuf<double> &dl = L.dloc;
    dl.upload(dloc, s);
    L.diag.alloc((size_t)L.N0 * nk); L.jac.alloc((size_t)L.N0 * nk);     // [N0][nk]
    BLOCH_CUDA(cudaMemsetAsync(L.diag.p, 0, sizeof(double) * L.N0 * nk, s));
    BLOCH_CUDA(launch_scatter_diag(E.map_h1, h->L_h1, E.cls, E.eps, dl.p, E.n_elem, L.diag.p, s, nk, E.n_class));
    k_jacobi<<<grid_for(L.N0 * nk), TPB, 0, s>>>(L.diag.p, L.jac.p, L.N0 * nk);
    h_sync(s);
  }
  // coarsest level: dense inverse per k-point from one block apply to the identity (columns k * n + j = e_j)
  H1Level &C = mg->lev[nl - 1];
  C.dense = C.N0 <= 512;
  if (C.dense) {
    const int n = (int)C.N0, nk = h->nk, w = nk * n;
    DevBuf<D2> &dI = C.dI, &dA = C.dA;
    if (dI.n < (size_t)n * w || mg->nk_built != nk) {
      std::vector<D2> I((size_t)n * w, make_double2(0.0, 0.0));
      for (int k = 0; k < nk; k++)
        for (int i = 0; i < n; i++) I[(size_t)i * w + k * n + i].x = 1.0;
      dI.upload(I, s);
      h_sync(s);
    }
    dA.alloc((size_t)n * w);
    BLOCH_CUDA(cudaMemsetAsync(dA.p, 0, sizeof(D2) * n * w, s));
    BLOCH_CUDA(launch_h1_op(p, 3, h->tabs, C.E, dI.p, w, dA.p, w, w, s, 1.0, 0.0));
    std::vector<D2> A((size_t)n * w);
    BLOCH_CUDA(cudaMemcpyAsync(A.data(), dA.p, sizeof(D2) * n * w, cudaMemcpyDeviceToHost, s));
    h_sync(s);
    std::vector<D2> inv((size_t)nk * n * n);
    for (int k = 0; k < nk && C.dense; k++) {
      dense::Mat Am((size_t)n * n);
      double tr = 0;
      for (int i = 0; i < n; i++) tr += A[(size_t)i * w + k * n + i].x;
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          const D2 a = A[(size_t)i * w + k * n + j], at = A[(size_t)j * w + k * n + i];
          Am[(size_t)i * n + j] = 0.5 * dense::cplx(a.x + at.x, a.y - at.y);
        }
      if (mg->kind >= 1) {
        // auxiliary-space hierarchy: the ND problem it preconditions is A + sigma M, whose auxiliary operator is
        // L + sigma * mass.  The mass term only matters on the constant mode (eigenvalue |kappa|^2 of L against
        // >= (2 pi / a)^2 for every other mode), so it is added there alone, as the rank-one term
        // sigma * mass(1, 1) / n^2 * 1 1^T on the coarsest level: without it the correction Pi B Pi^H amplifies the
        // near-constant fields by 1 / |kappa|^2 next to Gamma (CUB order 1 n_sub 16 at kappa = 0.01 (1,1,1):
        // 20 outer iterations against 9) and is singular at Gamma itself.
        const double vol23 = std::cbrt(h->mesh.volume) * std::cbrt(h->mesh.volume);
        // mass coefficient of the problem the hierarchy preconditions: eps for the ND problem (kind 1), the scalar
        // problem's m (muinv slot) for kind 2
        const std::vector<double> &mass_coef = mg->kind == 1 ? h->eps : h->muinv;
        double eps_mean = 0.0;
        for (double e : mass_coef) eps_mean += e;
        eps_mean /= (double)mass_coef.size();
        const double lift = env_double("BLOCH_SIGMA_SCALE", 1.0) / vol23 * h->mesh.volume * eps_mean / ((double)n * n);
        for (int i = 0; i < n; i++)
          for (int j = 0; j < n; j++) Am[(size_t)i * n + j] += lift;
      } else if (h->betas[k] == 0.0)   // singular on constants: lift that single mode
        for (int i = 0; i < n; i++)
          for (int j = 0; j < n; j++) Am[(size_t)i * n + j] += tr / ((double)n * n);
      dense::Mat Lc = Am;
      if (!dense::cholesky(n, Lc, 1e-14)) {
        C.dense = false;
        break;
      }
      // inverse = L^-H L^-1: solve for the identity
      dense::Mat X((size_t)n * n, dense::cplx(0));
      for (int c = 0; c < n; c++) {
        std::vector<dense::cplx> y(n, dense::cplx(0));
        for (int i = 0; i < n; i++) {           // L y = e_c
          dense::cplx sacc = (i == c) ? 1.0 : 0.0;
          for (int kk = 0; kk < i; kk++) sacc -= Lc[(size_t)i * n + kk] * y[kk];
          y[i] = sacc / Lc[(size_t)i * n + i].real();
        }
        for (int i = n - 1; i >= 0; i--) {      // L^H x = y
          dense::cplx sacc = y[i];
          for (int kk = i + 1; kk < n; kk++) sacc -= std::conj(Lc[(size_t)kk * n + i]) * X[(size_t)kk * n + c];
          X[(size_t)i * n + c] = sacc / Lc[(size_t)i * n + i].real();
        }
      }
      for (size_t i = 0; i < (size_t)n * n; i++) inv[(size_t)k * n * n + i] = make_double2(X[i].real(), X[i].imag());
    }
    if (C.dense) {
      C.inv.upload(inv, s);
      h_sync(s);
    }
  }
  if (mg->nk_built != h->nk) { mg->drop_graphs(); mg->m_alloc = 0; mg->zero_valid = false; }
  mg->nk_built = h->nk;
  if (C.dense != was_dense) mg->drop_graphs();   // the coarse solve is a different node list
  h_sync(s);                                      // host-side tables uploaded above are stack temporaries
}

static void probe_level(bloch_handle_s *h, const ElemData &Elev, std::vector<double> &dloc, double *bound,
                        std::vector<D2> *Sloc) {
  // same probe problem as Setup() (core.cu): one private element per local unit vector; the probe
  // buffers were built there and are reused (no allocation per k-point)
  const int nc = Elev.n_class, L = h->L_h1, nk = h->nk;
  cudaStream_t s = h->stream;
  const int ne = nc * L;
  bloch_handle_s::ProbeWork &pw = h->probe_h1;
  if (!pw.built || pw.built_nk != nk) throw std::runtime_error("multigrid setup before Setup()");
  ElemData E = Elev;
  E.n_elem = ne; E.cls = pw.cls.p; E.eps = pw.one.p; E.muinv = pw.one.p; E.map_h1 = pw.map.p;
  const size_t xs = (size_t)ne * L * nk;
  BLOCH_CUDA(cudaMemsetAsync(pw.y.p, 0, sizeof(D2) * xs, s));
  BLOCH_CUDA(launch_h1_op(h->p, 3, h->tabs, E, pw.x.p, nk, pw.y.p, nk, nk, s, 1.0, 0.0));
  std::vector<D2> &y = pw.hy;
  BLOCH_CUDA(cudaMemcpyAsync(y.data(), pw.y.p, sizeof(D2) * xs, cudaMemcpyDeviceToHost, s));
  h_sync(s);
  // y[((c*L + k)*L + l)*nk + kk] = S_{kk,c}[l][k]
  dloc.resize((size_t)nk * nc * L);
  std::vector<D2> loc((size_t)L * L);
  for (int kk = 0; kk < nk; kk++)
    for (int c = 0; c < nc; c++) {
      for (int k = 0; k < L; k++) dloc[((size_t)kk * nc + c) * L + k] = y[(((size_t)(c * L + k)) * L + k) * nk + kk].x;
      if (bound) {
        for (size_t t = 0; t < (size_t)L * L; t++) loc[t] = y[((size_t)c * L * L + t) * nk + kk];
        *bound = std::max(*bound, local_scaled_lmax(L, loc.data()));
      }
    }
  if (Sloc) {   // row-major S_{kk,c}[l][k], matrices ordered [kk][c]
    Sloc->resize((size_t)nk * nc * L * L);
    for (int kk = 0; kk < nk; kk++)
      for (int c = 0; c < nc; c++)
        for (int k = 0; k < L; k++)
          for (int l = 0; l < L; l++)
            (*Sloc)[(((size_t)kk * nc + c) * L + l) * L + k] = y[(((size_t)(c * L + k)) * L + l) * nk + kk];
  }
}

// ---- V-cycle pieces --------------------------------------------------------------------------
static void level_apply(bloch_handle_s *h, H1Level &L, const D2 *x, D2 *y, int m) {
  BLOCH_CUDA(cudaMemsetAsync(y, 0, sizeof(D2) * L.N0 * m, h->stream));
  BLOCH_CUDA(launch_h1_op(h->p, 3, h->tabs, L.E, x, m, y, m, m, h->stream, 1.0, 0.0));
  h->count_launch();
}
// Chebyshev coefficients of degree `degree` on [lmax / ratio, lmax]: out[0] = 1/theta, out[2k-1], out[2k] = the
// (d, r) weights of step k
static std::vector<double> cheb_coefs(double lmax, double ratio, int degree) {
  const double lmin = lmax / ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;
  std::vector<double> c(2 * std::max(degree, 1), 0.0);
  c[0] = 1.0 / theta;
  double rho = 1.0 / sigma1;
  for (int k = 1; k < degree; k++) {
    const double rho_n = 1.0 / (2.0 * sigma1 - rho);
    c[2 * k - 1] = rho_n * rho;
    c[2 * k] = 2.0 * rho_n / delta;
    rho = rho_n;
  }
  return c;
}
// x (+)= Cheb_deg(D^-1 S0) applied to the residual held in L.r (destroyed); cf = device coefficient table
static void chebyshev(bloch_handle_s *h, H1Level &L, D2 *x, int m, int degree, const double *cf, bool accumulate) {
  cudaStream_t s = h->stream;
  const unsigned g = grid_for(L.N0 * m);
  k_cheb_first<<<g, TPB, 0, s>>>(L.jac.p, L.r.p, L.d.p, x, cf, L.N0, m, accumulate ? 1 : 0, h->nk, m / h->nk);
  h->count_launch();
  for (int k = 1; k < degree; k++) {
    level_apply(h, L, L.d.p, L.q.p, m);
    k_cheb_step<<<g, TPB, 0, s>>>(L.jac.p, L.q.p, L.r.p, L.d.p, x, cf, k, L.N0, m, h->nk, m / h->nk);
    h->count_launch();
  }
}

static void vcycle(H1Multigrid *mg, bloch_handle_s *h, int l, int m, int deg, double ratio) {
  // input: lev[l].b ; output: lev[l].x
  cudaStream_t s = h->stream;
  H1Level &L = mg->lev[l];
  const int nl = (int)mg->lev.size();
  const long tot = L.N0 * m;
  if (l == nl - 1) {
    if (L.dense) {
      k_dense_apply<<<grid_for(tot), TPB, 0, s>>>(L.inv.p, L.b.p, L.x.p, (int)L.N0, m, m / h->nk);
      h->count_launch();
    } else {
      BLOCH_CUDA(cudaMemcpyAsync(L.r.p, L.b.p, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
      chebyshev(h, L, L.x.p, m, 40, L.cheb40.p, false);
    }
    return;
  }
  H1Level &C = mg->lev[l + 1];
  const int32_t *map_f = L.E.map_h1, *map_c = C.E.map_h1;
  const int nef = L.E.n_elem;
  static const bool sm2 = env_double("BLOCH_MG_SM2", 1.0) != 0.0;
  const int nk = h->nk, cpk = m / h->nk;
  const unsigned gv = grid_for(tot);
  // pre-smoothing from a zero initial guess
  if (deg == 2 && sm2) {
    k_sm2_pre1<<<gv, TPB, 0, s>>>(L.jac.p, L.b.p, L.x.p, L.cheb.p, L.N0, m, nk, cpk);
    level_apply(h, L, L.x.p, L.q.p, m);
    k_sm2_pre2<<<gv, TPB, 0, s>>>(L.jac.p, L.b.p, L.q.p, L.x.p, L.cheb.p, L.N0, m, nk, cpk);
    h->count_launch(2);
  } else {
    BLOCH_CUDA(cudaMemcpyAsync(L.r.p, L.b.p, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
    chebyshev(h, L, L.x.p, m, deg, L.cheb.p, false);
  }
  // residual, restriction
  level_apply(h, L, L.x.p, L.q.p, m);
  k_resid<<<grid_for(tot), TPB, 0, s>>>(L.b.p, L.q.p, L.r.p, tot);
  // explicit-matrix transfers win at orders 1-2 (measured: restrict 83 -> 45 us, prolong + add 79 -> 41 us on FCC
  // order 2 n_sub 8 x 160 columns); at order 3 the coarse rows of the restriction hold up to 343 entries and the
  // element-wise kernels are faster (BCC order 3 n_sub 8: projector 596 vs 696 ms per solve)
  // default: the sum-factorised parent kernels (k_h1_*_sf: 2604 instead of 21952 multiply-adds per parent at order 3).
  // BCC order 3 n_sub 8 warm solve 537 -> 399 ms (the element-wise restrict / prolong took 326 / 230 us per call on
  // 16 columns, 32 % of the solve); FCC order 2 bench workload 50.1 (CSR) -> 52.0 k-points/s.  Order 4 spills, it keeps
  // the element-wise kernels.
  static const double csr_env = env_double("BLOCH_MG_CSR_TRANSFER", -1.0);
  static const double sf_env = env_double("BLOCH_MG_SF_TRANSFER", -1.0);
  const bool use_sf = sf_env < 0.0 ? h->p <= 3 : sf_env != 0.0;
  const bool use_csr = !use_sf && (csr_env < 0.0 ? h->p <= 2 : csr_env != 0.0);
  if (use_sf) {
    BLOCH_CUDA(cudaMemsetAsync(C.b.p, 0, sizeof(D2) * C.N0 * m, s));
    switch (h->p) {
      case 1: restrict_sf_t<1>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, L.r.p, C.b.p, m, s); break;
      case 2: restrict_sf_t<2>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, L.r.p, C.b.p, m, s); break;
      case 3: restrict_sf_t<3>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, L.r.p, C.b.p, m, s); break;
      default: restrict_sf_t<4>(mg->TW, map_f, map_c, nef, L.n, L.invmult_par.p, L.r.p, C.b.p, m, s); break;
    }
  } else if (use_csr && L.have_csr) {
    k_csr_apply<<<grid_for(C.N0 * m), TPB, 0, s>>>(L.R_ptr.p, L.R_col.p, L.R_val.p, L.r.p, C.b.p, C.N0, m, 0);
  } else {
    BLOCH_CUDA(cudaMemsetAsync(C.b.p, 0, sizeof(D2) * C.N0 * m, s));
    switch (h->p) {
      case 1: restrict_t<1>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
      case 2: restrict_t<2>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
      case 3: restrict_t<3>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
      default: restrict_t<4>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
    }
  }
  h->count_launch(2);
  vcycle(mg, h, l + 1, m, deg, ratio);
  // prolongation + correction
  if (use_sf) {
    switch (h->p) {
      case 1: prolong_sf_t<1>(mg->TW, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
      case 2: prolong_sf_t<2>(mg->TW, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
      case 3: prolong_sf_t<3>(mg->TW, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
      default: prolong_sf_t<4>(mg->TW, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
    }
    k_add<<<grid_for(tot), TPB, 0, s>>>(L.x.p, L.d.p, tot);
  } else if (use_csr && L.have_csr) {
    k_csr_apply<<<grid_for(tot), TPB, 0, s>>>(L.P_ptr.p, L.P_col.p, L.P_val.p, C.x.p, L.x.p, L.N0, m, 1);
  } else {
    switch (h->p) {
      case 1: prolong_t<1>(mg->T, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
      case 2: prolong_t<2>(mg->T, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
      case 3: prolong_t<3>(mg->T, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
      default: prolong_t<4>(mg->T, map_f, map_c, nef, L.n, C.x.p, L.d.p, m, s); break;
    }
    k_add<<<grid_for(tot), TPB, 0, s>>>(L.x.p, L.d.p, tot);
  }
  // post-smoothing
  level_apply(h, L, L.x.p, L.q.p, m);
  if (deg == 2 && sm2) {
    k_sm2_post1<<<gv, TPB, 0, s>>>(L.jac.p, L.b.p, L.q.p, L.r.p, L.d.p, L.cheb.p, L.N0, m, nk, cpk);
    level_apply(h, L, L.d.p, L.q.p, m);
    k_sm2_post2<<<gv, TPB, 0, s>>>(L.jac.p, L.r.p, L.q.p, L.d.p, L.x.p, L.cheb.p, L.N0, m, nk, cpk);
    h->count_launch(4);
    return;
  }
  k_resid<<<grid_for(tot), TPB, 0, s>>>(L.b.p, L.q.p, L.r.p, tot);
  h->count_launch(3);
  chebyshev(h, L, L.x.p, m, deg, L.cheb.p, true);
}


// ---- fused V-cycle (see the kernel block above for the zero-buffer invariant) ----
static void apply_nz(bloch_handle_s *h, H1Level &L, const D2 *x, D2 *y, int m) {   // y must be zero on entry
  if (L.use_evec) {   // atomic-free: element-local results, then every dof sums its copies (y is overwritten)
    BLOCH_CUDA(launch_h1_op(h->p, 4, h->tabs, L.E, x, m, L.evec.p, m, m, h->stream, 1.0, 0.0));
    BLOCH_CUDA(launch_h1_reduce(L.tp_ptr.p, L.tp_loc.p, L.evec.p, y, L.N0, m, h->stream));
    h->count_launch(2);
    return;
  }
  if (L.have_S) BLOCH_CUDA(launch_h1_dense(h->p, L.E, L.Sloc.p, x, m, y, m, m, 1.0, h->stream));
  else BLOCH_CUDA(launch_h1_op(h->p, 3, h->tabs, L.E, x, m, y, m, m, h->stream, 1.0, 0.0));
  h->count_launch();
}
static void vcycle_fused(H1Multigrid *mg, bloch_handle_s *h, int l, int m, int deg, double ratio, D2 *b, D2 *x) {
  cudaStream_t s = h->stream;
  H1Level &L = mg->lev[l];
  const int nl = (int)mg->lev.size();
  const long tot = L.N0 * m;
  const unsigned g = grid_for(tot);
  if (l == nl - 1) {
    if (L.dense) {
      k_dense_apply<<<g, TPB, 0, s>>>(L.inv.p, b, x, (int)L.N0, m, m / h->nk);
      h->count_launch();
    } else {   // rare: coarsest level too large for a dense inverse
      BLOCH_CUDA(cudaMemcpyAsync(L.r.p, b, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
      BLOCH_CUDA(cudaMemsetAsync(b, 0, sizeof(D2) * tot, s));
      chebyshev(h, L, x, m, 40, L.cheb40.p, false);
    }
    return;
  }
  H1Level &C = mg->lev[l + 1];
  const int32_t *map_f = L.E.map_h1, *map_c = C.E.map_h1;
  const int nef = L.E.n_elem;
  auto cheb_tail = [&]() {   // steps 1 .. deg-1 of the Chebyshev recurrence on (r, d, x)
    for (int k = 1; k < deg; k++) {
      apply_nz(h, L, L.d.p, L.q.p, m);
      k_cheb_step_z<<<g, TPB, 0, s>>>(L.jac.p, L.q.p, L.r.p, L.d.p, x, L.cheb.p, k, L.N0, m, h->nk, m / h->nk);
      h->count_launch();
    }
  };
  // pre-smoothing from a zero guess
  k_cheb_first_b<<<g, TPB, 0, s>>>(L.jac.p, b, L.r.p, L.d.p, x, L.cheb.p, L.N0, m, h->nk, m / h->nk);
  cheb_tail();
  // residual, restriction (coarse b is zero on entry)
  apply_nz(h, L, x, L.q.p, m);
  k_resid_z<<<g, TPB, 0, s>>>(b, L.q.p, L.r.p, tot);
  switch (h->p) {
    case 1: restrict_t<1>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
    case 2: restrict_t<2>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
    case 3: restrict_t<3>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
    default: restrict_t<4>(mg->T, map_f, map_c, nef, L.n, L.invmult.p, L.r.p, C.b.p, m, s); break;
  }
  h->count_launch(3);
  vcycle_fused(mg, h, l + 1, m, deg, ratio, C.b.p, C.x.p);
  // prolongation added straight into x; a dense coarse level has its right-hand side cleared here
  const bool coarse_dense = (l + 1 == nl - 1) && C.dense;
  D2 *zero = coarse_dense ? C.b.p : nullptr;
  const long zc = coarse_dense ? C.N0 * m : 0;
  switch (h->p) {
    case 1: prolong_rep_t<1>(mg->T, L.rep.p, map_c, L.N0, L.n, C.x.p, x, m, zero, zc, s); break;
    case 2: prolong_rep_t<2>(mg->T, L.rep.p, map_c, L.N0, L.n, C.x.p, x, m, zero, zc, s); break;
    case 3: prolong_rep_t<3>(mg->T, L.rep.p, map_c, L.N0, L.n, C.x.p, x, m, zero, zc, s); break;
    default: prolong_rep_t<4>(mg->T, L.rep.p, map_c, L.N0, L.n, C.x.p, x, m, zero, zc, s); break;
  }
  // post-smoothing; levels >= 1 clear their right-hand side for the next cycle's restriction
  apply_nz(h, L, x, L.q.p, m);
  k_resid_cheb_first<<<g, TPB, 0, s>>>(L.jac.p, b, L.q.p, L.r.p, L.d.p, x, L.cheb.p, L.N0, m, l > 0 ? 1 : 0, h->nk, m / h->nk);
  h->count_launch(2);
  cheb_tail();
}

static void ensure_zero_buffers(H1Multigrid *mg, bloch_handle_s *h, int m) {
  if (mg->zero_valid && mg->zero_m == m) return;   // establish the zero-buffer invariant for this block width
  cudaStream_t s = h->stream;
  for (auto &L : mg->lev) {
    BLOCH_CUDA(cudaMemsetAsync(L.q.p, 0, sizeof(D2) * L.N0 * m, s));
    BLOCH_CUDA(cudaMemsetAsync(L.b.p, 0, sizeof(D2) * L.N0 * m, s));
  }
  BLOCH_CUDA(cudaMemsetAsync(mg->qvec.p, 0, sizeof(D2) * mg->lev[0].N0 * m, s));
  mg->zero_valid = true;
  mg->zero_m = m;
}

static int mg_solve_fused(H1Multigrid *mg, bloch_handle_s *h, D2 *rhs, D2 *phi, int m, const double *rel_tol_k, int max_it) {
  cudaStream_t s = h->stream;
  alloc_work(mg, m);
  H1Level &F = mg->lev[0];
  const long N0 = F.N0, tot = N0 * m;
  const int deg = mg->smooth_degree;
  const double ratio = env_double("BLOCH_MG_SMOOTH_RATIO", 5.0);
  ensure_zero_buffers(mg, h, m);
  double *part_rz = mg->part.p, *part_pq = part_rz + PCG_BLOCKS * m, *part_rr = part_pq + PCG_BLOCKS * m;
  double *rz_saved = part_rr + PCG_BLOCKS * m;
  double *sums = mg->scal.p + 6 * m;
  if (h->any_gamma) {
    BLOCH_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * m, s));
    k_col_sum<<<grid_for(tot), TPB, 0, s>>>(rhs, N0, m, sums);
    k_col_shift<<<grid_for(tot), TPB, 0, s>>>(rhs, N0, m, sums, h->d_gflag.p, m / h->nk);
    h->count_launch(2);
  }
  BLOCH_CUDA(cudaMemsetAsync(phi, 0, sizeof(D2) * tot, s));
  std::vector<double> hp((size_t)PCG_BLOCKS * m), rr0(m, 0.0), rrh(m);
  auto fetch_rr = [&](std::vector<double> &out) {
    BLOCH_CUDA(cudaMemcpyAsync(hp.data(), part_rr, sizeof(double) * PCG_BLOCKS * m, cudaMemcpyDeviceToHost, s));
    h_sync(s);
    for (int j = 0; j < m; j++) {
      double sum = 0.0;
      for (int b = 0; b < PCG_BLOCKS; b++) sum += hp[(size_t)b * m + j];
      out[j] = sum;
    }
  };
  const size_t sm1 = sizeof(double) * m, sm3 = 3 * sm1;
  k_dot_part<<<PCG_BLOCKS, TPB, sm1, s>>>(rhs, rhs, N0, m, part_rr);
  h->count_launch();
  fetch_rr(rr0);
  double mx = 0;
  for (double v : rr0) mx = std::max(mx, v);
  if (mx == 0.0) return 0;
  vcycle_fused(mg, h, 0, m, deg, ratio, rhs, mg->z.p);
  BLOCH_CUDA(cudaMemcpyAsync(mg->pvec.p, mg->z.p, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
  k_dot_part<<<PCG_BLOCKS, TPB, sm1, s>>>(rhs, mg->z.p, N0, m, part_rz);
  h->count_launch();
  auto half_a = [&]() {      // q = S0 p ; alpha ; phi += alpha p ; r -= alpha q ; rr = <r,r>
    apply_nz(h, F, mg->pvec.p, mg->qvec.p, m);
    k_dot_part<<<PCG_BLOCKS, TPB, sm1, s>>>(mg->pvec.p, mg->qvec.p, N0, m, part_pq);
    k_pcg_a<<<PCG_BLOCKS, TPB, sm3, s>>>(mg->pvec.p, mg->qvec.p, phi, rhs, part_rz, part_pq, part_rr, rz_saved, N0, m);
    h->count_launch(2);
  };
  auto half_b = [&]() {      // z = V(r) ; beta ; p = z + beta p
    vcycle_fused(mg, h, 0, m, deg, ratio, rhs, mg->z.p);
    k_dot_part<<<PCG_BLOCKS, TPB, sm1, s>>>(rhs, mg->z.p, N0, m, part_rz);
    k_pcg_b<<<PCG_BLOCKS, TPB, sm1, s>>>(mg->z.p, mg->pvec.p, part_rz, rz_saved, N0, m);
    h->count_launch(2);
  };
  static const bool use_graph = env_double("BLOCH_MG_GRAPH", 1.0) != 0.0;
  auto capture = [&](cudaGraphExec_t *exec, const std::function<void()> &body) -> bool {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    bool ok = true;
    try { body(); } catch (...) { ok = false; }
    if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return false; }
    if (ok) ok = cudaGraphInstantiate(exec, graph, 0) == cudaSuccess;
    cudaGraphDestroy(graph);
    if (!ok) { cudaGetLastError(); *exec = nullptr; }
    return ok;
  };
  static const bool timing = std::getenv("BLOCH_MG_TIMING") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); }
  double tA = 0, tB = 0, tW = 0;
  int nA = 0, nB = 0;
  auto w0 = std::chrono::steady_clock::now();
  int it = 0;
  for (it = 1; it <= max_it; it++) {
    bool graphs_ok = use_graph && mg->gA && mg->gB && mg->g_rhs == rhs && mg->g_phi == phi && mg->g_m == m && mg->g_nk == h->nk;
    if (use_graph && !graphs_ok && it == 2) {
      mg->drop_graphs();
      const int64_t l0 = h->stats.launches;
      const bool okA = capture(&mg->gA, half_a);
      const int64_t l1 = h->stats.launches;
      const bool okB = okA && capture(&mg->gB, half_b);
      const int64_t l2 = h->stats.launches;
      h->stats.launches = l0;                      // captured, not launched: counted per graph launch below
      if (okA && okB) {
        mg->nodes_A = (int)(l1 - l0); mg->nodes_B = (int)(l2 - l1);
        mg->g_rhs = rhs; mg->g_phi = phi; mg->g_m = m; mg->g_nk = h->nk;
        graphs_ok = true;
      } else {
        mg->drop_graphs();
      }
    }
    if (timing) cudaEventRecord(e0, s);
    if (graphs_ok) { BLOCH_CUDA(cudaGraphLaunch(mg->gA, s)); h->count_launch(mg->nodes_A); } else half_a();
    if (timing) cudaEventRecord(e1, s);
    fetch_rr(rrh);
    if (timing && it > 2) { float ms; cudaEventElapsedTime(&ms, e0, e1); tA += ms; nA++; }
    bool done = true;
    for (int j = 0; j < m; j++) {
      const double rel_tol = rel_tol_k[j / (m / h->nk)];
      if (rrh[j] > rel_tol * rel_tol * rr0[j]) done = false;
    }
    if (done) break;
    if (timing) cudaEventRecord(e1, s);
    if (graphs_ok) { BLOCH_CUDA(cudaGraphLaunch(mg->gB, s)); h->count_launch(mg->nodes_B); } else half_b();
    if (timing) {
      cudaEventRecord(e2, s);
      cudaEventSynchronize(e2);
      if (it > 2) { float ms; cudaEventElapsedTime(&ms, e1, e2); tB += ms; nB++; }
    }
  }
  if (timing) {
    tW = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    std::printf("[mg] %d its, wall %.3f ms; graph A avg %.1f us, graph B avg %.1f us (m = %d)\n", it, tW,
                nA ? 1e3 * tA / nA : 0.0, nB ? 1e3 * tB / nB : 0.0, m);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  }
  return it > max_it ? max_it : it;
}

void mg_vcycle(H1Multigrid *mg, bloch_handle_s *h, const D2 *b, D2 *x, int m) {
  cudaStream_t s = h->stream;
  alloc_work(mg, m);
  const int deg = mg->smooth_degree;
  const double ratio = env_double("BLOCH_MG_SMOOTH_RATIO", 5.0);
  static const double fused_max = env_double("BLOCH_MG_FUSED_MAX_ENTRIES", 8.0e5);
  static const bool fused = env_double("BLOCH_MG_FUSED", 1.0) != 0.0;
  H1Level &F = mg->lev[0];
  if (fused && (double)F.N0 * m <= fused_max) {
    ensure_zero_buffers(mg, h, m);
    vcycle_fused(mg, h, 0, m, deg, ratio, const_cast<D2 *>(b), x);   // the level-0 right-hand side is only read
  } else {
    mg->zero_valid = false;
    BLOCH_CUDA(cudaMemcpyAsync(F.b.p, b, sizeof(D2) * F.N0 * m, cudaMemcpyDeviceToDevice, s));
    vcycle(mg, h, 0, m, deg, ratio);
    BLOCH_CUDA(cudaMemcpyAsync(x, F.x.p, sizeof(D2) * F.N0 * m, cudaMemcpyDeviceToDevice, s));
  }
}

// block PCG on the fine level, one V-cycle as preconditioner; rhs is overwritten by the residual
int mg_solve(H1Multigrid *mg, bloch_handle_s *h, D2 *rhs, D2 *phi, int m, const double *rel_tol_k, int max_it) {
  // Two schedules of the same iteration.  The fused one (fewer graph nodes, consume-and-clear buffers,
  // atomic-free fine-level apply) wins while the fine level is a few hundred thousand entries and every
  // node sits at its latency floor; once the level vectors outgrow L2 the plain schedule is faster
  // (memset nodes are nearly free there, the extra stores of consume-and-clear are not): measured
  // crossover at ~0.8 M entries (FCC p2: n_sub 12).
  static const double fused_max = env_double("BLOCH_MG_FUSED_MAX_ENTRIES", 8.0e5);
  static const bool fused = env_double("BLOCH_MG_FUSED", 1.0) != 0.0;
  if (m % h->nk != 0) throw std::invalid_argument("block width must be a multiple of the k-point batch size");
  if (fused && (double)mg->lev[0].N0 * m <= fused_max) return mg_solve_fused(mg, h, rhs, phi, m, rel_tol_k, max_it);
  cudaStream_t s = h->stream;
  alloc_work(mg, m);
  mg->zero_valid = false;   // this schedule leaves the operator-output buffers dirty
  H1Level &F = mg->lev[0];
  const long N0 = F.N0, tot = N0 * m;
  const int deg = mg->smooth_degree;
  const double ratio = env_double("BLOCH_MG_SMOOTH_RATIO", 5.0);
  double *rz = mg->scal.p, *pq = rz + m, *rzn = pq + m, *rr = rzn + m, *alpha = rr + m, *beta = alpha + m;
  double *sums = beta + m;
  auto remove_mean = [&](D2 *v) {
    if (!h->any_gamma) return;
    BLOCH_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * m, s));
    k_col_sum<<<grid_for(tot), TPB, 0, s>>>(v, N0, m, sums);
    k_col_shift<<<grid_for(tot), TPB, 0, s>>>(v, N0, m, sums, h->d_gflag.p, m / h->nk);
    h->count_launch(2);
  };
  auto precond = [&](const D2 *r, D2 *z) {
    BLOCH_CUDA(cudaMemcpyAsync(F.b.p, r, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
    vcycle(mg, h, 0, m, deg, ratio);
    BLOCH_CUDA(cudaMemcpyAsync(z, F.x.p, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
  };
  remove_mean(rhs);
  BLOCH_CUDA(cudaMemsetAsync(phi, 0, sizeof(D2) * tot, s));
  std::vector<double> rr0(m), rrh(m);
  BLOCH_CUDA(launch_col_dot(rhs, rhs, N0, m, rr, s));
  BLOCH_CUDA(cudaMemcpyAsync(rr0.data(), rr, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
  h_sync(s);
  double mx = 0;
  for (double v : rr0) mx = std::max(mx, v);
  if (mx == 0.0) return 0;
  precond(rhs, mg->z.p);
  BLOCH_CUDA(cudaMemcpyAsync(mg->pvec.p, mg->z.p, sizeof(D2) * tot, cudaMemcpyDeviceToDevice, s));
  BLOCH_CUDA(launch_col_dot(rhs, mg->z.p, N0, m, rz, s));
  auto half_a = [&]() {      // q = S0 p ; alpha ; phi += alpha p ; r -= alpha q ; rr = <r,r>
    level_apply(h, F, mg->pvec.p, mg->qvec.p, m);
    BLOCH_CUDA(launch_col_dot(mg->pvec.p, mg->qvec.p, N0, m, pq, s));
    BLOCH_CUDA(launch_scalar_div(rz, pq, alpha, m, s));
    BLOCH_CUDA(launch_col_axpy(alpha, 1.0, mg->pvec.p, phi, N0, m, s));
    BLOCH_CUDA(launch_col_axpy(alpha, -1.0, mg->qvec.p, rhs, N0, m, s));
    BLOCH_CUDA(launch_col_dot(rhs, rhs, N0, m, rr, s));
    h->count_launch(5);
  };
  auto half_b = [&]() {      // z = V(r) ; beta ; p = z + beta p
    precond(rhs, mg->z.p);
    BLOCH_CUDA(launch_col_dot(rhs, mg->z.p, N0, m, rzn, s));
    BLOCH_CUDA(launch_scalar_div(rzn, rz, beta, m, s));
    BLOCH_CUDA(launch_col_xpby(mg->z.p, beta, mg->pvec.p, N0, m, s));
    BLOCH_CUDA(cudaMemcpyAsync(rz, rzn, sizeof(double) * m, cudaMemcpyDeviceToDevice, s));
    h->count_launch(3);
  };
  static const bool use_graph = env_double("BLOCH_MG_GRAPH", 1.0) != 0.0;
  auto capture = [&](cudaGraphExec_t *exec, const std::function<void()> &body) -> bool {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    bool ok = true;
    try { body(); } catch (...) { ok = false; }
    if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !graph) { cudaGetLastError(); return false; }
    if (ok) ok = cudaGraphInstantiate(exec, graph, 0) == cudaSuccess;
    cudaGraphDestroy(graph);
    if (!ok) { cudaGetLastError(); *exec = nullptr; }
    return ok;
  };
  int it = 0;
  for (it = 1; it <= max_it; it++) {
    bool graphs_ok = use_graph && mg->gA && mg->gB && mg->g_rhs == rhs && mg->g_phi == phi && mg->g_m == m && mg->g_nk == h->nk;
    if (use_graph && !graphs_ok && it == 2) {
      // iteration 1 ran eagerly (kernel attributes set, buffers allocated): capture now
      mg->drop_graphs();
      const int64_t l0 = h->stats.launches;
      const bool okA = capture(&mg->gA, half_a);
      const int64_t l1 = h->stats.launches;
      const bool okB = okA && capture(&mg->gB, half_b);
      const int64_t l2 = h->stats.launches;
      h->stats.launches = l0;                      // captured, not launched: counted per graph launch below
      if (okA && okB) {
        mg->nodes_A = (int)(l1 - l0); mg->nodes_B = (int)(l2 - l1);
        mg->g_rhs = rhs; mg->g_phi = phi; mg->g_m = m; mg->g_nk = h->nk;
        graphs_ok = true;
      } else {
        mg->drop_graphs();
      }
    }
    if (graphs_ok) { BLOCH_CUDA(cudaGraphLaunch(mg->gA, s)); h->count_launch(mg->nodes_A); } else half_a();
    BLOCH_CUDA(cudaMemcpyAsync(rrh.data(), rr, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
    h_sync(s);
    bool done = true;
    for (int j = 0; j < m; j++) {
      const double rel_tol = rel_tol_k[j / (m / h->nk)];
      if (rrh[j] > rel_tol * rel_tol * rr0[j]) done = false;
    }
    if (done) break;
    if (graphs_ok) { BLOCH_CUDA(cudaGraphLaunch(mg->gB, s)); h->count_launch(mg->nodes_B); } else half_b();
  }
  return it > max_it ? max_it : it;
}

}  // namespace bloch_b200
