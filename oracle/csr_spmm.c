/* TEST INFRASTRUCTURE ONLY (CPU baseline / reference arm of bench.py): threaded kernels of the CPU
 * restatement of the reference's solve path (oracle/cpu_solver.py).
 *   csr_spmm_z   - complex CSR x dense-block product, the kernel behind the reference's BlockOperator::Mult over
 *                  hypre ParCSR matrices (maxwell/maxwell_bloch.cpp:445-454).  Y[n][m] = A X, X/Y row-major complex.
 *   cheb_step_z  - fused vector update of the Chebyshev-Jacobi preconditioner (stands in for HypreAMS,
 *                  maxwell_bloch.cpp:492-517).
 *   pcg_jacobi_z - block Jacobi-PCG on the projector's inner system S0 phi = rhs (the reference: one MINRES per
 *                  vector per LOBPCG iteration, maxwell_bloch.cpp:2146-2152, 2280-2290), whole loop in C.
 * Complex numbers are interleaved (re, im) doubles; arithmetic written out in real form so that gcc does not
 * route every product through the NaN-recovering __muldc3. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int csr_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void csr_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

#define MAXM 64

void csr_spmm_z(int64_t n, int m, const int32_t *indptr, const int32_t *indices, const double *data,
                const double *X, double *Y) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; i++) {
    double acc[2 * MAXM];
    for (int j = 0; j < 2 * m; j++) acc[j] = 0.0;
    for (int32_t k = indptr[i]; k < indptr[i + 1]; k++) {
      const double ar = data[2 * (int64_t)k], ai = data[2 * (int64_t)k + 1];
      const double *x = X + (int64_t)indices[k] * 2 * m;
      for (int j = 0; j < m; j++) {
        acc[2 * j] += ar * x[2 * j] - ai * x[2 * j + 1];
        acc[2 * j + 1] += ar * x[2 * j + 1] + ai * x[2 * j];
      }
    }
    memcpy(Y + i * 2 * m, acc, sizeof(double) * 2 * m);
  }
}

/* r -= q ; d = a d + b jac .* r ; x += d      (all n x m complex, jac real per row) */
void cheb_step_z(int64_t n, int m, const double *jac, const double *q, double *r, double *d, double *x, double a,
                 double b) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    const double s = b * jac[i];
    const int64_t o = i * 2 * m;
    for (int j = 0; j < 2 * m; j++) {
      const double rr = r[o + j] - q[o + j];
      r[o + j] = rr;
      const double dd = a * d[o + j] + s * rr;
      d[o + j] = dd;
      x[o + j] += dd;
    }
  }
}
/* d = c0 jac .* r ; x = d */
void cheb_first_z(int64_t n, int m, const double *jac, const double *r, double *d, double *x, double c0) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    const double s = c0 * jac[i];
    const int64_t o = i * 2 * m;
    for (int j = 0; j < 2 * m; j++) { d[o + j] = s * r[o + j]; x[o + j] = d[o + j]; }
  }
}

/* column-wise real dots  out[j] = sum_i Re(conj(A[i][j]) B[i][j]) */
static void col_dots(int64_t n, int m, const double *A, const double *B, double *out) {
  double tot[MAXM];
  for (int j = 0; j < m; j++) tot[j] = 0.0;
#pragma omp parallel
  {
    double loc[MAXM];
    for (int j = 0; j < m; j++) loc[j] = 0.0;
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < n; i++) {
      const int64_t o = i * 2 * m;
      for (int j = 0; j < m; j++) loc[j] += A[o + 2 * j] * B[o + 2 * j] + A[o + 2 * j + 1] * B[o + 2 * j + 1];
    }
#pragma omp critical
    for (int j = 0; j < m; j++) tot[j] += loc[j];
  }
  for (int j = 0; j < m; j++) out[j] = tot[j];
}

/* Block Jacobi-PCG: solves S0 phi_j = rhs_j for every column j to |r_j| <= rel |rhs_j|.
 * rhs is overwritten by the residual; returns the iteration count. */
int pcg_jacobi_z(int64_t n, int m, const int32_t *indptr, const int32_t *indices, const double *data,
                 const double *jac, double *rhs, double *phi, double rel, int max_it) {
  if (m > MAXM) return -1;
  const size_t sz = (size_t)n * 2 * m;
  double *z = malloc(sizeof(double) * sz), *p = malloc(sizeof(double) * sz), *q = malloc(sizeof(double) * sz);
  double rr0[MAXM], rz[MAXM], rzn[MAXM], pq[MAXM], rr[MAXM], alpha[MAXM], beta[MAXM];
  memset(phi, 0, sizeof(double) * sz);
  col_dots(n, m, rhs, rhs, rr0);
  double mx = 0.0;
  for (int j = 0; j < m; j++) mx = fmax(mx, rr0[j]);
  int it = 0;
  if (mx > 0.0) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++)
      for (int j = 0; j < 2 * m; j++) { z[i * 2 * m + j] = jac[i] * rhs[i * 2 * m + j]; p[i * 2 * m + j] = z[i * 2 * m + j]; }
    col_dots(n, m, rhs, z, rz);
    for (it = 1; it <= max_it; it++) {
      csr_spmm_z(n, m, indptr, indices, data, p, q);
      col_dots(n, m, p, q, pq);
      for (int j = 0; j < m; j++) alpha[j] = pq[j] != 0.0 ? rz[j] / pq[j] : 0.0;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; i++) {
        const int64_t o = i * 2 * m;
        for (int j = 0; j < 2 * m; j++) {
          const double a = alpha[j >> 1];
          phi[o + j] += a * p[o + j];
          rhs[o + j] -= a * q[o + j];
          z[o + j] = jac[i] * rhs[o + j];
        }
      }
      col_dots(n, m, rhs, rhs, rr);
      int done = 1;
      for (int j = 0; j < m; j++) if (rr[j] > rel * rel * rr0[j]) done = 0;
      if (done) break;
      col_dots(n, m, rhs, z, rzn);
      for (int j = 0; j < m; j++) { beta[j] = rz[j] != 0.0 ? rzn[j] / rz[j] : 0.0; rz[j] = rzn[j]; }
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; i++) {
        const int64_t o = i * 2 * m;
        for (int j = 0; j < 2 * m; j++) p[o + j] = z[o + j] + beta[j >> 1] * p[o + j];
      }
    }
  }
  free(z); free(p); free(q);
  return it > max_it ? max_it : it;
}
