/* TEST INFRASTRUCTURE ONLY (CPU baseline of bench.py): threaded complex CSR x dense-block product,
 * the kernel behind the reference's BlockOperator::Mult over hypre ParCSR matrices
 * (maxwell/maxwell_bloch.cpp:445-454).  Y[n][m] = A X, A complex CSR, X/Y row-major complex. */
#include <complex.h>
#include <stdint.h>
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

void csr_spmm_z(int64_t n, int m, const int32_t *indptr, const int32_t *indices,
                const double complex *data, const double complex *X, double complex *Y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double complex *y = Y + i * m;
    for (int j = 0; j < m; j++) y[j] = 0.0;
    for (int32_t k = indptr[i]; k < indptr[i + 1]; k++) {
      const double complex a = data[k];
      const double complex *x = X + (int64_t)indices[k] * m;
      for (int j = 0; j < m; j++) y[j] += a * x[j];
    }
  }
}
