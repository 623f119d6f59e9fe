// Persistent cooperative kernel: the whole block Jacobi-PCG solve of  S0 phi = rhs,
// S0 = G^H M1(eps) G  (the inner solve of the reference's divergence projector,
// MaxwellBlochWaveProjector::Mult, maxwell/maxwell_bloch.cpp:2280-2290, there a MINRES at 1e-13).
//
// One launch runs every CG iteration: the matrix-free S0 application (same mode-space element
// code as k_h1_op), the fused local dot p^H S0 p, the vector updates and the per-column scalar
// reductions, separated by grid-wide barriers (3 per iteration).  Columns are independent CG
// solves; the loop ends when every column reached ||r|| <= rel_tol ||rhs||.  This removes ~8
// launches and all host round trips per iteration of the latency-bound small-N regime.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "elem_device.cuh"
#include "kernels.hpp"

namespace cg = cooperative_groups;

namespace bloch_b200 {

using namespace dev;

namespace {

template <int P, int NW>
__global__ void __launch_bounds__(NW * 32)
k_proj_cg(const __grid_constant__ Tabs T, const ElemData E, const double *__restrict__ jac,
          double2 *__restrict__ phi, double2 *__restrict__ r, double2 *__restrict__ z,
          double2 *__restrict__ p, double2 *__restrict__ q, double *__restrict__ scal, int m,
          long n_items, long n0, int max_it, double rel_tol2, int *__restrict__ info) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  constexpr int NT = NW * 32;
  constexpr int KEEP = (D::LH1 + NW - 1) / NW;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sND = reinterpret_cast<double2 *>(smem_raw);
  constexpr int ELEM_D2 = (P <= 2) ? D::LH1 * NT : (D::LND + D::LH1) * 32;
  double2 *sH = sND + D::LND * 32;   // tile path only
  double *sCP = reinterpret_cast<double *>(sND + ELEM_D2);
  double *sRed = sCP + E.n_class * kClassParDoubles;    // [2][m]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < E.n_class * kClassParDoubles; i += NT) sCP[i] = E.cpar[i];

  // scal layout: rr0[m] | pq[2][m] | rzn[2][m] | rr[2][m] | rz0[m]
  double *rr0 = scal, *pq = scal + m, *rzn = scal + 3 * m, *rr = scal + 5 * m, *rz0 = scal + 7 * m;
  const long total = n0 * m;
  const long nthreads = (long)gridDim.x * NT;
  const long usable = (nthreads / m) * m;
  const long start = (long)blockIdx.x * NT + threadIdx.x;
  const bool vec_active = start < usable;
  const int mycol = (int)(start % m);
  const long ntiles = (n_items + 31) / 32;

  // ---- init: z = jac r ; p = z ; rz0 = <r,z> ; rr0 = <r,r>   (phi = 0 and scal = 0 on entry) ----
  for (int j = threadIdx.x; j < 2 * m; j += NT) sRed[j] = 0.0;
  __syncthreads();
  if (vec_active) {
    double a_rz = 0.0, a_rr = 0.0;
    for (long t = start; t < total; t += usable) {
      const double2 rv = r[t];
      const double s = jac[t / m];
      const double2 zz = make_double2(s * rv.x, s * rv.y);
      z[t] = zz;
      p[t] = zz;
      a_rz = fma(rv.x, zz.x, a_rz); a_rz = fma(rv.y, zz.y, a_rz);
      a_rr = fma(rv.x, rv.x, a_rr); a_rr = fma(rv.y, rv.y, a_rr);
    }
    atomicAdd(&sRed[mycol], a_rz);
    atomicAdd(&sRed[m + mycol], a_rr);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < m; j += NT) {
    atomicAdd(rz0 + j, sRed[j]);
    atomicAdd(rr0 + j, sRed[m + j]);
  }
  grid.sync();
  double rz_cur = rz0[mycol];
  {
    bool allzero = true;
    for (int j = 0; j < m; j++) allzero = allzero && (rr0[j] == 0.0);
    if (allzero) {
      if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = 0; info[1] = 1; }
      return;
    }
  }

  int it = 0;
  bool done = false;
  long long tA = 0, tB = 0, tS1 = 0, tC = 0, tS2 = 0, tD = 0, tS0 = 0, c0, c1;
  for (it = 1; it <= max_it; it++) {
    const int b = it & 1;
    c0 = clock64();
    // ---- phase A: q = 0 ----
    for (long t = start; t < total; t += nthreads) q[t] = make_double2(0.0, 0.0);
    c1 = clock64(); tA += c1 - c0; c0 = c1;
    grid.sync();
    c1 = clock64(); tS0 += c1 - c0; c0 = c1;
    // ---- phase B: q += S0 p (element tiles), pq[b] += p^H S0 p ; block 0 clears parity 1-b ----
    if (blockIdx.x == 0)
      for (int j = threadIdx.x; j < m; j += NT) { pq[(1 - b) * m + j] = 0.0; rzn[(1 - b) * m + j] = 0.0; rr[(1 - b) * m + j] = 0.0; }
    for (int j = threadIdx.x; j < m; j += NT) sRed[j] = 0.0;
    if constexpr (P <= 2) {
      // one thread per (element, vector) item, element tensors in registers / a private column
      __syncthreads();
      double2 *scol = sND + threadIdx.x;      // private column, stride NT
      for (long item = start; item < n_items; item += nthreads) {
        const int e = (int)(item / m);
        const int v = (int)(item - (long)e * m);
        const int32_t *mh = E.map_h1 + (long)e * D::LH1;
#pragma unroll
        for (int k = 0; k < D::LH1; k++) scol[k * NT] = p[(long)(__ldg(mh + k) - 1) * m + v];
        double2 out[D::LH1];
        const double dot = s0_item<P, NT>(T, sCP + kClassParDoubles * E.cls[e], E.eps[e], scol, out);
        double *Qd = reinterpret_cast<double *>(q);
#pragma unroll
        for (int k = 0; k < D::LH1; k++) {
          const long g = (long)(__ldg(mh + k) - 1) * m + v;
          atomicAdd(Qd + 2 * g, out[k].x);
          atomicAdd(Qd + 2 * g + 1, out[k].y);
        }
        atomicAdd(&sRed[v], dot);
      }
    } else {
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        const long item = tile * 32 + lane;
        const bool active = item < n_items;
        const int e = active ? (int)(item / m) : 0;
        const int v = active ? (int)(item - (long)e * m) : 0;
        const int32_t *mh = E.map_h1 + (long)e * D::LH1;
        double2 keep[KEEP];
  #pragma unroll
        for (int k = 0; k < KEEP; k++) {
          const int j = warp + k * NW;
          double2 val = make_double2(0.0, 0.0);
          if (j < D::LH1) {
            if (active) val = p[(long)(__ldg(mh + j) - 1) * m + v];
            sH[j * 32 + lane] = val;
          }
          keep[k] = val;
        }
        __syncthreads();
        const double *cp = sCP + kClassParDoubles * (active ? E.cls[e] : 0);
        const double eps = active ? E.eps[e] : 0.0;
        for (int d = 0; d < 3; d++) {
          h1_transform_dir<P, NW, false>(sH, T, d, warp, lane);
          __syncthreads();
        }
        for (int t = warp; t < 3 * Q * Q; t += NW) {
          const int c = t / (Q * Q), rr_ = t - c * Q * Q;
          const int j1 = rr_ / Q, j2 = rr_ - j1 * Q;
          const int sc = c == 0 ? Q * Q : (c == 1 ? Q : 1);
          int idx[3];
          idx[c] = 0; idx[(c + 1) % 3] = j1; idx[(c + 2) % 3] = j2;
          const int base = h1_idx<P>(idx[0], idx[1], idx[2]);
          const double kc = cp[c];
          double2 in[Q];
  #pragma unroll
          for (int qq = 0; qq < Q; qq++) in[qq] = sH[(base + qq * sc) * 32 + lane];
  #pragma unroll
          for (int o = 0; o < P; o++) {
            double2 acc;
            acc.x = kc * in[o].y;
            acc.y = -kc * in[o].x;
  #pragma unroll
            for (int qq = 0; qq < Q; qq++) CFMA(acc, T.Dt[o][qq], in[qq]);
            sND[D::nd(c, o, j1, j2) * 32 + lane] = acc;
          }
        }
        __syncthreads();
        nd_mass_pointwise<P, NW>(sND, T, cp, eps, warp, lane);
        __syncthreads();
        for (int c = 0; c < 3; c++) {
          const int sc = c == 0 ? Q * Q : (c == 1 ? Q : 1);
          const double kc = cp[c];
          for (int t = warp; t < Q * Q; t += NW) {
            const int j1 = t / Q, j2 = t - j1 * Q;
            int idx[3];
            idx[c] = 0; idx[(c + 1) % 3] = j1; idx[(c + 2) % 3] = j2;
            const int base = h1_idx<P>(idx[0], idx[1], idx[2]);
            double2 in[P];
  #pragma unroll
            for (int o = 0; o < P; o++) in[o] = sND[D::nd(c, o, j1, j2) * 32 + lane];
  #pragma unroll
            for (int qq = 0; qq < Q; qq++) {
              double2 acc = (c == 0) ? make_double2(0.0, 0.0) : sH[(base + qq * sc) * 32 + lane];
  #pragma unroll
              for (int o = 0; o < P; o++) CFMA(acc, T.Dt[o][qq], in[o]);
              if (qq < P) { acc.x -= kc * in[qq].y; acc.y += kc * in[qq].x; }
              sH[(base + qq * sc) * 32 + lane] = acc;
            }
          }
          __syncthreads();
        }
        for (int d = 0; d < 3; d++) {
          h1_transform_dir<P, NW, true>(sH, T, d, warp, lane);
          __syncthreads();
        }
        if (active) {
          double *Qd = reinterpret_cast<double *>(q);
          double dot = 0.0;
  #pragma unroll
          for (int k = 0; k < KEEP; k++) {
            const int j = warp + k * NW;
            if (j < D::LH1) {
              const long g = __ldg(mh + j) - 1;
              const double2 val = sH[j * 32 + lane];
              double *dst = Qd + 2 * (g * m + v);
              atomicAdd(dst, val.x);
              atomicAdd(dst + 1, val.y);
              dot = fma(keep[k].x, val.x, dot);
              dot = fma(keep[k].y, val.y, dot);
            }
          }
          atomicAdd(&sRed[v], dot);
        }
      }
      }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += NT)
      if (sRed[j] != 0.0) atomicAdd(pq + b * m + j, sRed[j]);
    c1 = clock64(); tB += c1 - c0; c0 = c1;
    grid.sync();
    c1 = clock64(); tS1 += c1 - c0; c0 = c1;
    // ---- phase C: alpha ; phi += alpha p ; r -= alpha q ; z = jac r ; rzn[b], rr[b] ----
    for (int j = threadIdx.x; j < 2 * m; j += NT) sRed[j] = 0.0;
    __syncthreads();
    if (vec_active) {
      const double pqv = pq[b * m + mycol];
      const double alpha = pqv != 0.0 ? rz_cur / pqv : 0.0;
      double a_rz = 0.0, a_rr = 0.0;
#pragma unroll 4
      for (long t = start; t < total; t += usable) {
        const double2 pp = p[t], qq = q[t];
        double2 f = phi[t], rv = r[t];
        f.x = fma(alpha, pp.x, f.x); f.y = fma(alpha, pp.y, f.y);
        rv.x = fma(-alpha, qq.x, rv.x); rv.y = fma(-alpha, qq.y, rv.y);
        phi[t] = f;
        r[t] = rv;
        const double s = jac[t / m];
        const double2 zz = make_double2(s * rv.x, s * rv.y);
        z[t] = zz;
        a_rz = fma(rv.x, zz.x, a_rz); a_rz = fma(rv.y, zz.y, a_rz);
        a_rr = fma(rv.x, rv.x, a_rr); a_rr = fma(rv.y, rv.y, a_rr);
      }
      atomicAdd(&sRed[mycol], a_rz);
      atomicAdd(&sRed[m + mycol], a_rr);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += NT) {
      atomicAdd(rzn + b * m + j, sRed[j]);
      atomicAdd(rr + b * m + j, sRed[m + j]);
    }
    c1 = clock64(); tC += c1 - c0; c0 = c1;
    grid.sync();
    c1 = clock64(); tS2 += c1 - c0; c0 = c1;
    // ---- phase D: convergence test (uniform over the grid), beta, p = z + beta p ----
    done = true;
    for (int j = 0; j < m; j++) done = done && (rr[b * m + j] <= rel_tol2 * rr0[j]);
    if (done) break;
    if (vec_active) {
      const double rznew = rzn[b * m + mycol];
      const double beta = rz_cur != 0.0 ? rznew / rz_cur : 0.0;
      rz_cur = rznew;
#pragma unroll 4
      for (long t = start; t < total; t += usable) {
        const double2 zz = z[t], pp = p[t];
        p[t] = make_double2(fma(beta, pp.x, zz.x), fma(beta, pp.y, zz.y));
      }
    }
    // the barrier after phase A of the next iteration orders p before its gather
    c1 = clock64(); tD += c1 - c0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && info[2] == 12345) {
    printf("[proj_cg] its %d cycles/it: A %lld S0 %lld B %lld S1 %lld C %lld S2 %lld D %lld\n", it, tA / it, tS0 / it, tB / it, tS1 / it, tC / it, tS2 / it, tD / it);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = it > max_it ? max_it : it; info[1] = done ? 1 : 0; }
}

template <int P, int NW>
cudaError_t proj_cg_t(const Tabs &T, const ElemData &E, const double *jac, double2 *phi, double2 *r,
                      double2 *z, double2 *p, double2 *q, double *scal, int m, long n0, int max_it,
                      double rel_tol, int *info, cudaStream_t s) {
  using D = Dim<P>;
  const size_t elem_d2 = (P <= 2) ? (size_t)D::LH1 * NW * 32 : (size_t)(D::LND + D::LH1) * 32;
  const size_t smem = elem_d2 * sizeof(double2) +
                      (size_t)(E.n_class * kClassParDoubles + 2 * m) * sizeof(double);
  static int max_blocks_of[kMaxDevices] = {};
  int &max_blocks = max_blocks_of[current_device_slot()];
  cudaError_t err;
  if (max_blocks == 0) {
    err = cudaFuncSetAttribute(k_proj_cg<P, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_proj_cg<P, NW>, NW * 32, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    int cap = 2;
    if (const char *e = std::getenv("BLOCH_CG_CTAS_PER_SM")) cap = std::atoi(e);
    if (cap < 1) cap = 1;
    max_blocks = sms * (per_sm > cap ? cap : per_sm);
  }
  long n_items = (long)E.n_elem * m;
  long ntiles = (P <= 2) ? (n_items + NW * 32 - 1) / (NW * 32) : (n_items + 31) / 32;
  long nvec_blocks = (n0 * m + NW * 32 - 1) / (NW * 32);
  if (nvec_blocks > ntiles) ntiles = nvec_blocks;
  int grid = (int)(ntiles < max_blocks ? ntiles : max_blocks);
  if (grid < 1) grid = 1;
  double rel_tol2 = rel_tol * rel_tol;
  Tabs Tc = T;
  ElemData Ec = E;
  void *args[] = {&Tc, &Ec, &jac, &phi, &r, &z, &p, &q, &scal, &m, &n_items, &n0, &max_it, &rel_tol2, &info};
  return cudaLaunchCooperativeKernel((void *)k_proj_cg<P, NW>, dim3(grid), dim3(NW * 32), args, smem, s);
}

}  // namespace

cudaError_t launch_proj_cg(int p, const Tabs &T, const ElemData &E, const double *jac, double2 *phi,
                           double2 *r, double2 *z, double2 *pp, double2 *q, double *scal, int m,
                           long n0, int max_it, double rel_tol, int *info, cudaStream_t s) {
  switch (p) {
    case 1: return proj_cg_t<1, 8>(T, E, jac, phi, r, z, pp, q, scal, m, n0, max_it, rel_tol, info, s);
    case 2: return proj_cg_t<2, 8>(T, E, jac, phi, r, z, pp, q, scal, m, n0, max_it, rel_tol, info, s);
    case 3: return proj_cg_t<3, 12>(T, E, jac, phi, r, z, pp, q, scal, m, n0, max_it, rel_tol, info, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace bloch_b200
