// Launch wrappers of the sm_100a kernels (definitions in kernels.cu).
//
// Device block-vector layout: interleaved complex, dof-major:  X[(dof*ld + v)] is a double2
// (re, im) for dof < n, v < nvec <= ld (ld = row pitch in double2 units, so a kernel can work
// on a column sub-block of a wider array).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace bloch_b200 {

constexpr int kMaxP = 4;
constexpr int kMaxDevices = 64;
// index of the current device for per-device one-time launcher state (cudaFuncSetAttribute and occupancy
// results are per device; a process may hold handles on several devices)
inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= kMaxDevices) ? 0 : dev;
}
constexpr int kMaxClasses = 64;
constexpr int kClassParDoubles = 22;   // kh[3], G[3][3], H[3][3], det J

// 1-D tables, sized for the largest supported order; indexed with compile-time indices in the
// kernels so they are read straight from the constant bank holding the kernel parameters.
struct Tabs {
  double TI[kMaxP + 1][kMaxP + 1];      // nodal(closed) -> mode
  double TIinv[kMaxP + 1][kMaxP + 1];
  double Dt[kMaxP][kMaxP + 1];          // derivative at Gauss points, from mode space
  double om[kMaxP + 1];                 // diagonal 1-D mass weights in mode space
};

// 1-D tables of the field-average quadrature (field_avg.cu): q = p + 1 Gauss points
struct AvgTabs {
  double bo[kMaxP + 1][kMaxP];          // open (Gauss-Lagrange) basis at the quadrature points
  double bc[kMaxP + 1][kMaxP + 1];      // closed (GLL-Lagrange) basis at the quadrature points
  double xq[kMaxP + 1], wq[kMaxP + 1];
};

// 1-D tables of the nested ND prolongation (nd_transfer.cu)
struct NdTransfer1D {
  double Pc[2][kMaxP + 1][kMaxP + 1];   // closed: c_j((a + l_i)/2)
  double Po[2][kMaxP][kMaxP];           // open:   o_k((a + g_o)/2) / 2
};

struct ElemData {            // device pointers, element order = mesh order
  int n_elem = 0, n_class = 0;
  // k-point batching: a handle may carry nk Bloch vectors at once.  Every block vector then has nvec = nk * cpk
  // columns, column v belongs to k-point v / cpk and uses the class table cpar[v / cpk][class]; the launchers fill
  // in cpk = nvec / nk (with_cpk below), nk == 1 gives the plain single-kappa behaviour.
  int nk = 1, cpk = 1 << 30;
  // 1: y was cleared by the caller and receives nothing but this launch - dofs INTERIOR to an element (touched by no
  // other element) are then written with plain stores instead of fp64 reductions: no read-for-ownership of those
  // lines, 3p(p-1)^2 fewer reductions per element (25 % of them at p = 3).  0: y += (accumulating launches).
  int fresh_y = 0;
  // 1 (only with fresh_y): the caller cleared just the rows of dofs SHARED between elements (launch_clear_rows); the
  // element-interior rows are left to the plain stores of the lane-pair / six-lane kernels.  The first-generation
  // kernel reduces into every row, so launch_nd_apply clears all of y itself before falling back to it.
  int partial_clear = 0;
  long n_rows_y = 0;                   // rows of y (ND dofs), for that fallback
  const int *cls = nullptr;            // [n_elem]
  const double *eps = nullptr;         // [n_elem]
  const double *muinv = nullptr;       // [n_elem]
  const double *cpar = nullptr;        // [nk][n_class][22]   (kappa dependent)
  const int32_t *map_nd = nullptr;     // [n_elem][L_nd]  kernel ("cyclic") local order, signed 1-based
  const int32_t *map_h1 = nullptr;     // [n_elem][L_h1]  kernel local order, 1-based
  const int32_t *map_rt = nullptr;     // [n_elem][L_rt]  kernel local order, signed 1-based
};
// copy of E with the columns-per-k-point of an nvec-column block vector filled in
inline ElemData with_cpk(const ElemData &E, int nvec) {
  ElemData K = E;
  if (K.nk < 1) K.nk = 1;
  K.cpk = nvec / K.nk;
  if (K.cpk < 1) K.cpk = 1;
  return K;
}

// y = ca * A x + cm * M x   (ND -> ND), A = (C - i Z_kappa)^H M2(muinv) (C - i Z_kappa), M = M1(eps)
// z == nullptr: signed scatter-ADD into y with fp64 atomics (y must be zero-initialised).
// z != nullptr: atomic-free first pass - the element-local results go to the E-vector
//               z[(e*L_nd + j)*nvec + v] with plain coalesced stores and y is not touched;
//               launch_nd_reduce then sums the copies of every dof (deterministic).
cudaError_t launch_nd_apply(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx,
                            double2 *y, int ldy, int nvec, double ca, double cm, cudaStream_t s,
                            double2 *z = nullptr);
// Orders 1 and 2, z == nullptr: the barrier-free lane-pair-per-item kernel (nd_item.cu); launch_nd_apply
// tries it first.  *launched = false (and nothing done) when it does not apply or BLOCH_ND_ITEM=0.
cudaError_t launch_nd_item(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy,
                           int nvec, double ca, double cm, cudaStream_t s, bool *launched);
// Orders 1..3, z == nullptr: the barrier-free six-lanes-per-item kernel (nd_comp.cu), tried before launch_nd_item
// for the orders selected by BLOCH_ND_COMP (bit mask, default: order 3).
cudaError_t launch_nd_comp(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy,
                           int nvec, double ca, double cm, cudaStream_t s, bool *launched);
// y[g][v] = sum_{k in [ptr[g], ptr[g+1])} sign(loc[k]) * z[(|loc[k]|-1)*m + v]
cudaError_t launch_nd_reduce(const int *ptr, const int32_t *loc, const double2 *z, double2 *y, long n,
                             int m, int ldy, cudaStream_t s);
// H1 <-> ND operators of the projector.  mode 0: y_h1 += S0 x_h1 (S0 = G^H M1 G);
// mode 1: y_nd = G x_h1 (interpolation, plain stores); mode 2: y_h1 += G^H M1 x_nd;
// mode 3 (scalar H1 Bloch Helmholtz, misc/scalar3d.cpp:662-818): y_h1 += ca G^H M1(k) G x + cm M0(m) x
// with k = E.eps, m = E.muinv; orders 1..4 (modes 0-2: orders 1..3)
cudaError_t launch_h1_op(int p, int mode, const Tabs &T, const ElemData &E, const double2 *x, int ldx,
                         double2 *y, int ldy, int nvec, cudaStream_t s, double ca = 1.0, double cm = 0.0);
// y_rt = (C - i Z_kappa) x_nd  (interpolation into nodal RT dofs, plain stores)
// Per-vector cell integrals of e^{i kappa.x} {E, C E, eps E, mu^-1 C E} (12 complex numbers per vector,
// out[nvec][12]); X = ND dofs, Y = RT dofs of (C - i Z_kappa) X; part = scratch [n_elem*nvec][12].
cudaError_t launch_field_avg(int p, const AvgTabs &A, const ElemData &E, const double *x0, const double *geom,
                             const double kappa[3], const double2 *X, int ldx, const double2 *Y, int ldy, int nvec,
                             double2 *part, double2 *out, cudaStream_t s);

// y += ca * eps_e * S_class(e) x_e with dense element matrices S [n_class][L_h1][L_h1] (row-major), p <= 2
cudaError_t launch_h1_dense(int p, const ElemData &E, const double2 *S, const double2 *x, int ldx, double2 *y,
                            int ldy, int nvec, double ca, cudaStream_t s);

// launch_h1_op mode 4 (p <= 2, cm == 0): like mode 3 but the element-local results are stored to the E-vector
// y[(e*L_h1 + k)*nvec + v] (no atomics, y need not be zeroed); launch_h1_reduce sums the copies of each dof.
cudaError_t launch_h1_reduce(const int *ptr, const int32_t *loc, const double2 *Z, double2 *Y, long n, int m,
                             cudaStream_t s);

// xf = nodal interpolation of the coarse ND block xc on the once-refined mesh (n_f = 2 n_c subdivisions)
cudaError_t launch_nd_prolong(int p, const NdTransfer1D &T, const int32_t *map_f, const int32_t *map_c, int n_elem_f,
                              int n_f, const double2 *xc, int ldc, double2 *xf, int ldf, int m, cudaStream_t s);

cudaError_t launch_curl(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y,
                        int ldy, int nvec, cudaStream_t s);

// ---- layout conversion: boundary [re(N); im(N)] per vector  <->  block [N][nvec] complex ----
cudaError_t launch_pack(const double *reim, double2 *blk, long n, int nvec, cudaStream_t s);
// y[rows[i]][0 .. nvec) = 0 for the n_rows listed rows (0-based) of an [n][ldy] block vector
cudaError_t launch_clear_rows(double2 *y, int ldy, int nvec, const int32_t *rows, long n_rows, cudaStream_t s);
cudaError_t launch_unpack(const double2 *blk, double *reim, long n, int nvec, cudaStream_t s);

// ---- dense block-vector algebra (tall-skinny), all complex ----
// C[i][j] = sum_r conj(A[r][i]) B[r][j]  (ma x mb, row-major, double2) ; C is overwritten
cudaError_t launch_gram(const double2 *A, int ma, int lda, const double2 *B, int mb, int ldb, long n,
                        double2 *C, cudaStream_t s);
// Y[r][j] = beta*Y[r][j] + sum_i X[r][i] * C[i][j]   (X: n x k, C: k x m row-major, Y: n x m)
cudaError_t launch_block_mult(const double2 *X, int k, const double2 *C, int m, double2 *Y,
                              double beta, long n, cudaStream_t s);
// R[r][j] = AX[r][j] - lambda[j] * MX[r][j]
cudaError_t launch_residual(const double2 *AX, const double2 *MX, const double *lambda, double2 *R,
                            long n, int m, cudaStream_t s);
// Y[r][j] = a*X[r][j] + b*Y[r][j]   (real scalars)
cudaError_t launch_axpby(double a, const double2 *X, double b, double2 *Y, long total, cudaStream_t s);
// Y[r][j] = X[r][j] * d[r]     (real diagonal scaling, e.g. Jacobi)
cudaError_t launch_diag_scale(const double *d, const double2 *X, double2 *Y, long n, int m, cudaStream_t s);
// per-column CG style updates with per-column device scalars:
//   Y[r][j] += sign * alpha[j] * X[r][j]      (alpha real)
cudaError_t launch_col_axpy(const double *alpha, double sign, const double2 *X, double2 *Y, long n,
                            int m, cudaStream_t s);
//   P[r][j] = Z[r][j] + beta[j] * P[r][j]
cudaError_t launch_col_xpby(const double2 *Z, const double *beta, double2 *P, long n, int m,
                            cudaStream_t s);
// d[j] = sum_r Re(conj(A[r][j]) B[r][j])   (column-wise real dot), d overwritten
cudaError_t launch_col_dot(const double2 *A, const double2 *B, long n, int m, double *d,
                           cudaStream_t s);
// tiny helpers on per-column scalars (device side, avoids host round trips in CG)
//   out[j] = (den[j] != 0) ? num[j]/den[j] : 0
cudaError_t launch_scalar_div(const double *num, const double *den, double *out, int m, cudaStream_t s);
// diagonal accumulation: d[gid][k] += coef_e * dloc[k][cls][l] over elements (real); d is [n][nk], k = k-point
cudaError_t launch_scatter_diag(const int32_t *map, int L, const int *cls, const double *coef,
                                const double *dloc, int n_elem, double *d, cudaStream_t s, int nk = 1,
                                int n_class = 0);
// Whole block Jacobi-PCG solve of S0 phi = r0 in ONE cooperative launch (proj_cg.cu).  On entry
// phi = 0, r = right-hand side, scal (8*m doubles) = 0; info[0] = iterations, info[1] = converged.
cudaError_t launch_proj_cg(int p, const Tabs &T, const ElemData &E, const double *jac, double2 *phi,
                           double2 *r, double2 *z, double2 *pp, double2 *q, double *scal, int m,
                           long n0, int max_it, double rel_tol, int *info, cudaStream_t s);
// Device-side Rayleigh-Ritz (rr_device.cu): for each of K k-points the lowest mb eigenpairs of the pencil
// (GA[k], GM[k]) (kc x kc, row-major, both triangles present) restricted to the selected basis columns - all of the
// first mb, column i >= mb only if act[k][i % mb] != 0 and GM_ii > 0, the third group only if usep[k] != 0.
// Outputs C[k] (kc x mb, zero rows for unselected columns), lam[k][mb], info[k] (0 ok, 1 = basis had to be shrunk,
// -1 = failed), and usep[k] for the next call (0 after a shrunk basis).  kc <= 63, mb <= 32.  One CTA per k-point.
cudaError_t launch_rr_solve(const double2 *GA, const double2 *GM, int kc, int mb, int K, const unsigned char *act,
                            unsigned char *usep, double2 *C, double *lam, int *info, cudaStream_t s);

// fill with deterministic pseudo-random complex numbers in (-1,1)
cudaError_t launch_fill_random(double2 *X, long total, unsigned long long seed, cudaStream_t s);

// measured fp64 FMA throughput of the current device (TFLOP/s, best of 5)
cudaError_t measure_fp64_peak(double *tflops, cudaStream_t s);

}  // namespace bloch_b200
