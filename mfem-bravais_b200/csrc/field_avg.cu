// Cell averages of the Bloch fields of an eigenmode: MaxwellBlochWaveEquation::GetFieldAverages
// (maxwell/maxwell_bloch.cpp:1550-1632) with the linear forms of SetKappa (:211-279).
//
// The reference assembles, per Cartesian direction i, eight linear forms
//     L(w) = int coef(x) {cos, sin}(kappa.x) e_i . w(x) dx ,   coef in {1, eps} on ND, {1, mu^-1} on RT
// and dots them with (Er, Ei), (Br, Bi).  Summing the same quadrature element by element gives
//     avg_i = sum_e coef_e sum_q w_q detJ e^{i kappa.x_q} F_i(x_q),  F = J^-T Ehat (ND) or J Bhat / detJ (RT)
// without ever forming the global linear forms.  Quadrature = MFEM's default for
// VectorFEDomainLFIntegrator on these elements (order 2 * el.GetOrder() -> p + 1 Gauss points per direction).
//
// One item = (element, vector); lane <-> item, warps split the quadrature points; the per-item sums go to
// a [n_items][12] scratch array and a second kernel adds them in a fixed order (deterministic result).
#include "kernels.hpp"
#include "elem_device.cuh"

namespace bloch_b200 {
namespace {

using namespace dev;

constexpr int FA_WARPS = 4;

template <int P>
__global__ void __launch_bounds__(FA_WARPS * 32)
k_field_avg(const __grid_constant__ AvgTabs A, const ElemData E, const double *__restrict__ x0,
            const double *__restrict__ geom, double kx, double ky, double kz,
            const double2 *__restrict__ X, int ldx, const double2 *__restrict__ Y, int ldy, int m,
            long n_items, double2 *__restrict__ part) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sND = reinterpret_cast<double2 *>(smem_raw);
  double2 *sRT = sND + D::LND * 32;
  double2 *sAcc = sRT + D::LRT * 32;                 // [FA_WARPS][6][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long item = (long)blockIdx.x * 32 + lane;
  const bool active = item < n_items;
  const int e = active ? (int)(item / m) : 0;
  const int v = active ? (int)(item - (long)e * m) : 0;
  const int32_t *mp = E.map_nd + (long)e * D::LND;
  for (int j = warp; j < D::LND; j += FA_WARPS) {
    double2 val = make_double2(0.0, 0.0);
    if (active) {
      const int s = __ldg(mp + j);
      const long g = (s < 0 ? -s : s) - 1;
      val = X[g * ldx + v];
      if (s < 0) { val.x = -val.x; val.y = -val.y; }
    }
    sND[j * 32 + lane] = val;
  }
  const int32_t *mr = E.map_rt + (long)e * D::LRT;
  for (int j = warp; j < D::LRT; j += FA_WARPS) {
    double2 val = make_double2(0.0, 0.0);
    if (active) {
      const int s = __ldg(mr + j);
      const long g = (s < 0 ? -s : s) - 1;
      val = Y[g * ldy + v];
      if (s < 0) { val.x = -val.x; val.y = -val.y; }
    }
    sRT[j * 32 + lane] = val;
  }
  __syncthreads();
  const double *gm = geom + 18 * E.cls[e];           // J (row-major), J^-1 (row-major)
  double J[3][3], Ji[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { J[i][j] = gm[3 * i + j]; Ji[i][j] = gm[9 + 3 * i + j]; }
  const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  const double kap[3] = {kx, ky, kz};
  double kh[3];
#pragma unroll
  for (int d = 0; d < 3; d++) kh[d] = J[0][d] * kap[0] + J[1][d] * kap[1] + J[2][d] * kap[2];
  const double ph0 = kap[0] * x0[3 * e] + kap[1] * x0[3 * e + 1] + kap[2] * x0[3 * e + 2];
  double2 acc[6];
#pragma unroll
  for (int k = 0; k < 6; k++) acc[k] = make_double2(0.0, 0.0);
  for (int t = warp; t < Q * Q * Q; t += FA_WARPS) {
    int a[3];
    a[0] = t % Q; a[1] = (t / Q) % Q; a[2] = t / (Q * Q);
    double2 Eh[3], Bh[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      double2 se = make_double2(0.0, 0.0), sb = make_double2(0.0, 0.0);
      for (int o = 0; o < P; o++)
        for (int j1 = 0; j1 < Q; j1++) {
          const double w1 = A.bo[a[c]][o] * A.bc[a[c1]][j1];
          for (int j2 = 0; j2 < Q; j2++) CFMA(se, w1 * A.bc[a[c2]][j2], sND[D::nd(c, o, j1, j2) * 32 + lane]);
        }
      for (int j = 0; j < Q; j++)
        for (int o1 = 0; o1 < P; o1++) {
          const double w1 = A.bc[a[c]][j] * A.bo[a[c1]][o1];
          for (int o2 = 0; o2 < P; o2++) CFMA(sb, w1 * A.bo[a[c2]][o2], sRT[D::rt(c, j, o1, o2) * 32 + lane]);
        }
      Eh[c] = se;
      Bh[c] = sb;
    }
    double sn, cs;
    sincos(ph0 + kh[0] * A.xq[a[0]] + kh[1] * A.xq[a[1]] + kh[2] * A.xq[a[2]], &sn, &cs);
    const double w = A.wq[a[0]] * A.wq[a[1]] * A.wq[a[2]] * det;
    const double pr = w * cs, pi = w * sn;           // w e^{i phase}
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double2 Ep, Bp;                                 // covariant / contravariant Piola maps
      Ep.x = Ji[0][i] * Eh[0].x + Ji[1][i] * Eh[1].x + Ji[2][i] * Eh[2].x;
      Ep.y = Ji[0][i] * Eh[0].y + Ji[1][i] * Eh[1].y + Ji[2][i] * Eh[2].y;
      Bp.x = (J[i][0] * Bh[0].x + J[i][1] * Bh[1].x + J[i][2] * Bh[2].x) / det;
      Bp.y = (J[i][0] * Bh[0].y + J[i][1] * Bh[1].y + J[i][2] * Bh[2].y) / det;
      acc[i].x += pr * Ep.x - pi * Ep.y;
      acc[i].y += pr * Ep.y + pi * Ep.x;
      acc[3 + i].x += pr * Bp.x - pi * Bp.y;
      acc[3 + i].y += pr * Bp.y + pi * Bp.x;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; k++) sAcc[(warp * 6 + k) * 32 + lane] = acc[k];
  __syncthreads();
  if (warp == 0 && active) {
    const double ee = E.eps[e], mu = E.muinv[e];
    double2 *dst = part + item * 12;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double2 sum = make_double2(0.0, 0.0);
      for (int w = 0; w < FA_WARPS; w++) { sum.x += sAcc[(w * 6 + k) * 32 + lane].x; sum.y += sAcc[(w * 6 + k) * 32 + lane].y; }
      const double cf = k < 3 ? ee : mu;
      dst[k] = sum;                                    // E (0..2), C E (3..5)
      dst[6 + k] = make_double2(cf * sum.x, cf * sum.y);   // eps E (6..8), mu^-1 C E (9..11)
    }
  }
}

// out[v][k] = sum_e part[(e*m + v)*12 + k], one CTA per (v, k), fixed summation order
__global__ void k_sum_parts(const double2 *__restrict__ part, int n_elem, int m, double2 *__restrict__ out) {
  __shared__ double2 red[256];
  const int v = blockIdx.x / 12, k = blockIdx.x - v * 12;
  double2 s = make_double2(0.0, 0.0);
  for (int e = threadIdx.x; e < n_elem; e += 256) {
    const double2 x = part[((long)e * m + v) * 12 + k];
    s.x += x.x; s.y += x.y;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) { red[threadIdx.x].x += red[threadIdx.x + w].x; red[threadIdx.x].y += red[threadIdx.x + w].y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

template <int P>
cudaError_t field_avg_t(const AvgTabs &A, const ElemData &E, const double *x0, const double *geom, const double kappa[3],
                        const double2 *X, int ldx, const double2 *Y, int ldy, int nvec, double2 *part, double2 *out,
                        cudaStream_t s) {
  using D = Dim<P>;
  const size_t smem = (size_t)(D::LND + D::LRT + FA_WARPS * 6) * 32 * sizeof(double2);
  static bool attr_set_of[kMaxDevices] = {};
  bool &attr_set = attr_set_of[current_device_slot()];
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(k_field_avg<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  const long n_items = (long)E.n_elem * nvec;
  k_field_avg<P><<<(unsigned)((n_items + 31) / 32), FA_WARPS * 32, smem, s>>>(A, E, x0, geom, kappa[0], kappa[1], kappa[2], X,
                                                                           ldx, Y, ldy, nvec, n_items, part);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  k_sum_parts<<<nvec * 12, 256, 0, s>>>(part, E.n_elem, nvec, out);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_field_avg(int p, const AvgTabs &A, const ElemData &E, const double *x0, const double *geom,
                             const double kappa[3], const double2 *X, int ldx, const double2 *Y, int ldy, int nvec,
                             double2 *part, double2 *out, cudaStream_t s) {
  switch (p) {
    case 1: return field_avg_t<1>(A, E, x0, geom, kappa, X, ldx, Y, ldy, nvec, part, out, s);
    case 2: return field_avg_t<2>(A, E, x0, geom, kappa, X, ldx, Y, ldy, nvec, part, out, s);
    case 3: return field_avg_t<3>(A, E, x0, geom, kappa, X, ldx, Y, ldy, nvec, part, out, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace bloch_b200
