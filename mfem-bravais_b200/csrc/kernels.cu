// sm_100a kernels of the Bloch Maxwell path: matrix-free, sum-factorised element operators in
// fp64 with the element tensors staged in shared memory, signed periodic gather/scatter, and the
// tall-skinny block algebra of the eigensolver.
//
// Element kernels ("mode space" formulation):
//   * one CTA processes 32 work items (element, vector); lane <-> item, so every shared-memory
//     access is a conflict-free 16-byte-per-lane row and all tensor indices are compile time;
//   * warps split the independent slabs / pencils / grid points of each phase;
//   * closed 1-D directions are mapped nodal -> mode (values at the p Gauss points + top
//     Legendre coefficient), where all 1-D mass matrices are diagonal and the Bloch shift
//     -i kappa_hat is pointwise, so A = (C - i Z)^H M2 (C - i Z) costs 24 real 1-D contractions.
// Local storage is "cyclic": ND component c is [open dir c][closed c+1][closed c+2], RT component
// c is [closed dir c][open c+1][open c+2], which makes the three components share one code path.
#include "kernels.hpp"

#include <cstdio>
#include <cstdlib>

#include "elem_device.cuh"

namespace bloch_b200 {

namespace {

using namespace dev;

// ------------------------------------------------------------------------------------------
// y = ca * A x + cm * M x on ND block vectors
// ------------------------------------------------------------------------------------------
// compute phases of y_e = ca A_e x_e + cm M_e x_e for one tile of 32 items: on entry sND holds the
// gathered nodal values (after a block barrier), on exit the nodal result (after a block barrier)
template <int P, int NW>
__device__ __forceinline__ void nd_compute_tile(double2 *sND, double2 *sRT, const Tabs &T, const double *cp,
                                                double eps, double mui, double ca, double cm, int warp, int lane) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  // ---- nodal -> mode in the closed directions ----
  nd_transform_all<P, NW, 0>(sND, T, warp, lane);
  __syncthreads();
  // ---- Bloch curl: R_c = K_{c+1} F_{c+2} - K_{c+2} F_{c+1},  K_d = Dt - i kh_d (pointwise) ----
  if (ca != 0.0) {
    for (int t = warp; t < 3 * Q; t += NW) {
      const int c = t / Q, j = t - c * Q;
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      const double k1 = cp[c1], k2 = cp[c2];
      double2 A[P][Q], B[P][Q];
#pragma unroll
      for (int o = 0; o < P; o++)
#pragma unroll
        for (int r = 0; r < Q; r++) {
          A[o][r] = sND[D::nd(c2, o, j, r) * 32 + lane];   // F_{c+2}[o2=o, j1=j, j2=r]
          B[o][r] = sND[D::nd(c1, o, r, j) * 32 + lane];   // F_{c+1}[o1=o, j1=r, j2=j]
        }
#pragma unroll
      for (int o1 = 0; o1 < P; o1++)
#pragma unroll
        for (int o2 = 0; o2 < P; o2++) {
          double2 acc;
          acc.x = k1 * A[o2][o1].y - k2 * B[o1][o2].y;
          acc.y = -k1 * A[o2][o1].x + k2 * B[o1][o2].x;
#pragma unroll
          for (int r = 0; r < Q; r++) {
            CFMA(acc, T.Dt[o1][r], A[o2][r]);
            CFMA(acc, -T.Dt[o2][r], B[o1][r]);
          }
          sRT[D::rt(c, j, o1, o2) * 32 + lane] = acc;
        }
    }
  }
  __syncthreads();
  // ---- pointwise RT mass (scaled by ca * muinv) and ND mass (scaled by cm * eps) ----
  if (ca != 0.0) {
    const double *G = cp + 3;
    for (int g = warp; g < Q * Q * Q; g += NW) {
      int i[3];
      i[0] = g / (Q * Q);
      i[1] = (g / Q) % Q;
      i[2] = g % Q;
      double w = ca * mui;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        double o = T.om[0];
#pragma unroll
        for (int r = 1; r < Q; r++) o = (i[d] == r) ? T.om[r] : o;
        w *= o;
      }
      double2 f[3];
      int loc[3];
      bool ex[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        ex[c] = (i[(c + 1) % 3] < P) && (i[(c + 2) % 3] < P);
        loc[c] = ex[c] ? D::rt(c, i[c], i[(c + 1) % 3], i[(c + 2) % 3]) : 0;
        f[c] = ex[c] ? sRT[loc[c] * 32 + lane] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        if (ex[c]) {
          double2 r;
          r.x = w * (G[3 * c] * f[0].x + G[3 * c + 1] * f[1].x + G[3 * c + 2] * f[2].x);
          r.y = w * (G[3 * c] * f[0].y + G[3 * c + 1] * f[1].y + G[3 * c + 2] * f[2].y);
          sRT[loc[c] * 32 + lane] = r;
        }
      }
    }
  }
  if (cm != 0.0) nd_mass_pointwise<P, NW>(sND, T, cp, cm * eps, warp, lane);
  __syncthreads();
  // ---- adjoint curl accumulated onto the mass part, then mode -> nodal (adjoint) ----
  for (int t = warp; t < 3 * P; t += NW) {
    const int c = t / P, o = t - c * P;
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    double2 f[Q][Q];
    const int base = D::nd(c, o, 0, 0);
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++)
        f[a][b] = (cm != 0.0) ? sND[(base + a * Q + b) * 32 + lane] : make_double2(0.0, 0.0);
    if (ca != 0.0) {
      const double k1 = cp[c1], k2 = cp[c2];
      double2 Y1[Q][P], Y2[Q][P];
#pragma unroll
      for (int a = 0; a < Q; a++)
#pragma unroll
        for (int q = 0; q < P; q++) {
          Y1[a][q] = sRT[D::rt(c1, a, q, o) * 32 + lane];   // Y_{c+1}[j=j1, o1, o2=o]
          Y2[a][q] = sRT[D::rt(c2, a, o, q) * 32 + lane];   // Y_{c+2}[j=j2, o1=o, o2]
        }
#pragma unroll
      for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) {
          double2 acc = f[j1][j2];
#pragma unroll
          for (int q = 0; q < P; q++) {
            CFMA(acc, T.Dt[q][j2], Y1[j1][q]);
            CFMA(acc, -T.Dt[q][j1], Y2[j2][q]);
          }
          if (j2 < P) { acc.x -= k2 * Y1[j1][j2].y; acc.y += k2 * Y1[j1][j2].x; }
          if (j1 < P) { acc.x += k1 * Y2[j2][j1].y; acc.y -= k1 * Y2[j2][j1].x; }
          f[j1][j2] = acc;
        }
    }
    slab_transform<P, true>(f, T.TI);
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) sND[(base + a * Q + b) * 32 + lane] = f[a][b];
  }
  __syncthreads();
}

// Persistent, software-pipelined version: each CTA loops over tiles; the gather of tile t+1 is
// issued with cp.async (LDGSTS, 16 B per request) into the second ND buffer while tile t is being
// computed, so the two dependent global round trips (index, value) of the signed gather are off
// the critical path.  Signs are applied in place by the thread that issued the copy.
template <int P, int NW>
__global__ void __launch_bounds__(NW * 32, (P == 2 ? 3 : 1))
k_nd_apply(const __grid_constant__ Tabs T, const ElemData E, const double2 *__restrict__ X,
           double2 *__restrict__ Y, double2 *__restrict__ Z, int m, int ldx, int ldy, long n_items, double ca,
           double cm) {
  using D = Dim<P>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sBuf0 = reinterpret_cast<double2 *>(smem_raw);
  double2 *sBuf1 = sBuf0 + D::LND * 32;
  double2 *sRT = sBuf1 + D::LND * 32;
  double *sCP = reinterpret_cast<double *>(sRT + D::LRT * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < cpar_doubles(E); i += NW * 32) sCP[i] = E.cpar[i];
  constexpr int KG = (D::LND + NW - 1) / NW;
  const long ntiles = (n_items + 31) / 32;
  const double2 zero2 = make_double2(0.0, 0.0);

  auto load_idx = [&](long tile, int (&sidx)[KG], int &cls_, double &eps_, double &mui_) {
    const long item = tile * 32 + lane;
    const bool act = tile < ntiles && item < n_items;
    const int e = act ? (int)(item / m) : 0;
    cls_ = act ? vclass(E, __ldg(E.cls + e), (int)(item - (long)e * m)) : 0;
    eps_ = act ? __ldg(E.eps + e) : 0.0;
    mui_ = act ? __ldg(E.muinv + e) : 0.0;
    const int32_t *mp = E.map_nd + (long)e * D::LND;
#pragma unroll
    for (int k = 0; k < KG; k++) {
      const int j = warp + k * NW;
      sidx[k] = (act && j < D::LND) ? __ldg(mp + j) : 0;
    }
  };
  auto issue_gather = [&](long tile, const int (&sidx)[KG], double2 *buf) {
    const long item = tile * 32 + lane;
    const int e = (int)(item / m);
    const int v = (int)(item - (long)e * m);
#pragma unroll
    for (int k = 0; k < KG; k++) {
      const int j = warp + k * NW;
      if (j < D::LND) {
        double2 *dst = buf + j * 32 + lane;
        const int s = sidx[k];
        if (s != 0) {
          const long g = (s < 0 ? -s : s) - 1;
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(X + g * ldx + v));
        } else {
          *dst = zero2;
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  int idx_cur[KG], idx_nxt[KG];
  int cls_cur, cls_nxt;
  double eps_cur, eps_nxt, mui_cur, mui_nxt;
  long tile = blockIdx.x;
  load_idx(tile, idx_cur, cls_cur, eps_cur, mui_cur);
  if (tile < ntiles) issue_gather(tile, idx_cur, sBuf0);
  load_idx(tile + gridDim.x, idx_nxt, cls_nxt, eps_nxt, mui_nxt);
  int parity = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    double2 *sND = parity ? sBuf1 : sBuf0;
    double2 *sOther = parity ? sBuf0 : sBuf1;
    const long item = tile * 32 + lane;
    const bool active = item < n_items;
    const int e = active ? (int)(item / m) : 0;
    const int v = active ? (int)(item - (long)e * m) : 0;
    // wait for this tile's gather, apply the orientation signs to the entries this thread copied
    asm volatile("cp.async.wait_group 0;\n" ::);
#pragma unroll
    for (int k = 0; k < KG; k++) {
      if (idx_cur[k] < 0) {
        const int j = warp + k * NW;
        double2 val = sND[j * 32 + lane];
        val.x = -val.x; val.y = -val.y;
        sND[j * 32 + lane] = val;
      }
    }
    __syncthreads();     // also orders the previous tile's scatter reads of sOther before the refill
    const long next = tile + gridDim.x;
    if (next < ntiles) issue_gather(next, idx_nxt, sOther);
    int idx_nn[KG], cls_nn;
    double eps_nn, mui_nn;
    load_idx(next + gridDim.x, idx_nn, cls_nn, eps_nn, mui_nn);
    const double *cp = sCP + kClassParDoubles * cls_cur;
    nd_compute_tile<P, NW>(sND, sRT, T, cp, eps_cur, mui_cur, ca, cm, warp, lane);
    if (Z != nullptr) {
      // ---- element-local result to the E-vector (coalesced plain stores; summed by k_nd_reduce) ----
      if (active) {
        double2 *zp = Z + ((long)e * D::LND) * m + v;
#pragma unroll
        for (int k = 0; k < KG; k++) {
          const int j = warp + k * NW;
          if (j < D::LND) zp[(long)j * m] = sND[j * 32 + lane];
        }
      }
    } else {
      // ---- signed scatter-add (indices kept in registers since the gather) ----
      double *Yd = reinterpret_cast<double *>(Y);
#pragma unroll
      for (int k = 0; k < KG; k++) {
        const int s = idx_cur[k];
        if (s != 0) {
          const int j = warp + k * NW;
          const long g = (s < 0 ? -s : s) - 1;
          double2 val = sND[j * 32 + lane];
          if (s < 0) { val.x = -val.x; val.y = -val.y; }
          double *dst = Yd + 2 * (g * ldy + v);
          atomicAdd(dst, val.x);
          atomicAdd(dst + 1, val.y);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KG; k++) { idx_cur[k] = idx_nxt[k]; idx_nxt[k] = idx_nn[k]; }
    cls_cur = cls_nxt; eps_cur = eps_nxt; mui_cur = mui_nxt;
    cls_nxt = cls_nn; eps_nxt = eps_nn; mui_nxt = mui_nn;
    parity ^= 1;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
}

// ------------------------------------------------------------------------------------------
// H1 <-> ND operators: mode 0: y_h1 += G^H M1 G x_h1 ; 1: y_nd = G x_h1 ; 2: y_h1 += G^H M1 x_nd
// G = T01 - i Z01_kappa.  In mode space: F_c = (Dt - i kh_c) along direction c of Phi~.
// ------------------------------------------------------------------------------------------
template <int P, int NW, int MODE>
__global__ void __launch_bounds__(NW * 32)
k_h1_op(const __grid_constant__ Tabs T, const ElemData E, const double2 *__restrict__ X,
        double2 *__restrict__ Y, int m, int ldx, int ldy, long n_items, double ca, double cm) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sND = reinterpret_cast<double2 *>(smem_raw);
  double2 *sH = sND + D::LND * 32;
  double *sCP = reinterpret_cast<double *>(sH + D::LH1 * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < cpar_doubles(E); i += NW * 32) sCP[i] = E.cpar[i];
  const long item = (long)blockIdx.x * 32 + lane;
  const bool active = item < n_items;
  const int e = active ? (int)(item / m) : 0;
  const int v = active ? (int)(item - (long)e * m) : 0;
  const int32_t *mh = E.map_h1 + (long)e * D::LH1;
  const int32_t *mn = E.map_nd + (long)e * D::LND;
  if (MODE == 2) {
    for (int j = warp; j < D::LND; j += NW) {
      double2 val = make_double2(0.0, 0.0);
      if (active) {
        const int s = __ldg(mn + j);
        const long g = (s < 0 ? -s : s) - 1;
        val = ld2(X + g * ldx + v);
        if (s < 0) { val.x = -val.x; val.y = -val.y; }
      }
      sND[j * 32 + lane] = val;
    }
  } else {
    for (int j = warp; j < D::LH1; j += NW) {
      double2 val = make_double2(0.0, 0.0);
      if (active) val = ld2(X + (long)(__ldg(mh + j) - 1) * ldx + v);
      sH[j * 32 + lane] = val;
    }
  }
  __syncthreads();
  const double *cp = sCP + kClassParDoubles * (active ? vclass(E, E.cls[e], v) : 0);
  const double eps = active ? E.eps[e] : 0.0;
  if (MODE == 2) {
    nd_transform_all<P, NW, 0>(sND, T, warp, lane);
    __syncthreads();
  } else {
    for (int d = 0; d < 3; d++) {
      h1_transform_dir<P, NW, false>(sH, T, d, warp, lane);
      __syncthreads();
    }
    // gradient in mode space: F_c[o][j1][j2] = sum_t Dt[o][t] Phi[i_c = t] - i kh_c Phi[i_c = o]
    for (int t = warp; t < 3 * Q * Q; t += NW) {
      const int c = t / (Q * Q), r = t - c * Q * Q;
      const int j1 = r / Q, j2 = r - j1 * Q;
      const int sc = c == 0 ? Q * Q : (c == 1 ? Q : 1);
      int idx[3];
      idx[c] = 0; idx[(c + 1) % 3] = j1; idx[(c + 2) % 3] = j2;
      const int base = h1_idx<P>(idx[0], idx[1], idx[2]);
      const double kc = cp[c];
      double2 in[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) in[q] = sH[(base + q * sc) * 32 + lane];
#pragma unroll
      for (int o = 0; o < P; o++) {
        double2 acc;
        acc.x = kc * in[o].y;
        acc.y = -kc * in[o].x;
#pragma unroll
        for (int q = 0; q < Q; q++) CFMA(acc, T.Dt[o][q], in[q]);
        sND[D::nd(c, o, j1, j2) * 32 + lane] = acc;
      }
    }
    __syncthreads();
  }
  if (MODE == 1) {
    // mode -> nodal ND dofs (inverse transform), plain signed stores
    nd_transform_all<P, NW, 2>(sND, T, warp, lane);
    __syncthreads();
    if (active) {
      for (int j = warp; j < D::LND; j += NW) {
        const int s = __ldg(mn + j);
        const long g = (s < 0 ? -s : s) - 1;
        double2 val = sND[j * 32 + lane];
        if (s < 0) { val.x = -val.x; val.y = -val.y; }
        Y[g * ldy + v] = val;
      }
    }
    return;
  }
  nd_mass_pointwise<P, NW>(sND, T, cp, (MODE == 3) ? ca * eps : eps, warp, lane);
  __syncthreads();
  // scalar variant (MODE 3): + cm * M0(m) x, pointwise in mode space: m_e detJ w(i0) w(i1) w(i2)
  const double mcoef = (MODE == 3 && active) ? cm * E.muinv[e] * cp[21] : 0.0;
  // adjoint gradient: Phi'[i_c = t] (+)= sum_o Dt[o][t] F_c[o] + i kh_c F_c[t] (t < P)
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
      double wj = 1.0;
      if (MODE == 3) {   // c == 0 pencils run along direction 0: (j1, j2) = (i1, i2)
        double o1 = T.om[0], o2 = T.om[0];
#pragma unroll
        for (int r = 1; r < Q; r++) { o1 = (j1 == r) ? T.om[r] : o1; o2 = (j2 == r) ? T.om[r] : o2; }
        wj = mcoef * o1 * o2;
      }
#pragma unroll
      for (int q = 0; q < Q; q++) {
        double2 acc = make_double2(0.0, 0.0);
        if (c != 0 || MODE == 3) acc = sH[(base + q * sc) * 32 + lane];
        if (c == 0 && MODE == 3) { const double w = wj * T.om[q]; acc.x *= w; acc.y *= w; }
#pragma unroll
        for (int o = 0; o < P; o++) CFMA(acc, T.Dt[o][q], in[o]);
        if (q < P) { acc.x -= kc * in[q].y; acc.y += kc * in[q].x; }
        sH[(base + q * sc) * 32 + lane] = acc;
      }
    }
    __syncthreads();
  }
  for (int d = 0; d < 3; d++) {
    h1_transform_dir<P, NW, true>(sH, T, d, warp, lane);
    __syncthreads();
  }
  if (active) {
    double *Yd = reinterpret_cast<double *>(Y);
    for (int j = warp; j < D::LH1; j += NW) {
      const long g = __ldg(mh + j) - 1;
      const double2 val = sH[j * 32 + lane];
      double *dst = Yd + 2 * (g * ldy + v);
      atomicAdd(dst, val.x);
      atomicAdd(dst + 1, val.y);
    }
  }
}

// ------------------------------------------------------------------------------------------
// y_rt = (C - i Z_kappa) x_nd : nodal RT dofs of the Bloch curl (B field), plain stores
// ------------------------------------------------------------------------------------------
template <int P, int NW>
__global__ void __launch_bounds__(NW * 32)
k_curl(const __grid_constant__ Tabs T, const ElemData E, const double2 *__restrict__ X,
       double2 *__restrict__ Y, int m, int ldx, int ldy, long n_items) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sND = reinterpret_cast<double2 *>(smem_raw);
  double2 *sRT = sND + D::LND * 32;
  double *sCP = reinterpret_cast<double *>(sRT + D::LRT * 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < cpar_doubles(E); i += NW * 32) sCP[i] = E.cpar[i];
  const long item = (long)blockIdx.x * 32 + lane;
  const bool active = item < n_items;
  const int e = active ? (int)(item / m) : 0;
  const int v = active ? (int)(item - (long)e * m) : 0;
  const int32_t *mp = E.map_nd + (long)e * D::LND;
  for (int j = warp; j < D::LND; j += NW) {
    double2 val = make_double2(0.0, 0.0);
    if (active) {
      const int s = __ldg(mp + j);
      const long g = (s < 0 ? -s : s) - 1;
      val = ld2(X + g * ldx + v);
      if (s < 0) { val.x = -val.x; val.y = -val.y; }
    }
    sND[j * 32 + lane] = val;
  }
  __syncthreads();
  const double *cp = sCP + kClassParDoubles * (active ? vclass(E, E.cls[e], v) : 0);
  nd_transform_all<P, NW, 0>(sND, T, warp, lane);
  __syncthreads();
  for (int t = warp; t < 3 * Q; t += NW) {
    const int c = t / Q, j = t - c * Q;
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    const double k1 = cp[c1], k2 = cp[c2];
    double2 A[P][Q], B[P][Q];
#pragma unroll
    for (int o = 0; o < P; o++)
#pragma unroll
      for (int r = 0; r < Q; r++) {
        A[o][r] = sND[D::nd(c2, o, j, r) * 32 + lane];
        B[o][r] = sND[D::nd(c1, o, r, j) * 32 + lane];
      }
#pragma unroll
    for (int o1 = 0; o1 < P; o1++)
#pragma unroll
      for (int o2 = 0; o2 < P; o2++) {
        double2 acc;
        acc.x = k1 * A[o2][o1].y - k2 * B[o1][o2].y;
        acc.y = -k1 * A[o2][o1].x + k2 * B[o1][o2].x;
#pragma unroll
        for (int r = 0; r < Q; r++) {
          CFMA(acc, T.Dt[o1][r], A[o2][r]);
          CFMA(acc, -T.Dt[o2][r], B[o1][r]);
        }
        sRT[D::rt(c, j, o1, o2) * 32 + lane] = acc;
      }
  }
  __syncthreads();
  // mode -> nodal along the single closed direction of each RT component (pencils)
  for (int t = warp; t < 3 * P * P; t += NW) {
    const int c = t / (P * P), r = t - c * P * P;
    const int o1 = r / P, o2 = r - o1 * P;
    double2 in[Q], out[Q];
#pragma unroll
    for (int j = 0; j < Q; j++) in[j] = sRT[D::rt(c, j, o1, o2) * 32 + lane];
#pragma unroll
    for (int k = 0; k < Q; k++) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < Q; j++) CFMA(acc, T.TIinv[k][j], in[j]);
      out[k] = acc;
    }
#pragma unroll
    for (int j = 0; j < Q; j++) sRT[D::rt(c, j, o1, o2) * 32 + lane] = out[j];
  }
  __syncthreads();
  if (active) {
    const int32_t *mr = E.map_rt + (long)e * D::LRT;
    for (int j = warp; j < D::LRT; j += NW) {
      const int s = __ldg(mr + j);
      const long g = (s < 0 ? -s : s) - 1;
      double2 val = sRT[j * 32 + lane];
      if (s < 0) { val.x = -val.x; val.y = -val.y; }
      Y[g * ldy + v] = val;
    }
  }
}

template <int P, int NW>
cudaError_t nd_apply_t(const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy,
                       int nvec, double ca, double cm, double2 *z, cudaStream_t s) {
  using D = Dim<P>;
  const size_t smem = (size_t)(2 * D::LND + D::LRT) * 32 * sizeof(double2) +
                      (size_t)E.nk * E.n_class * kClassParDoubles * sizeof(double);
  static int max_ctas_of[kMaxDevices] = {};
  int &max_ctas = max_ctas_of[current_device_slot()];
  if (max_ctas == 0) {
    cudaError_t err = cudaFuncSetAttribute(k_nd_apply<P, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nd_apply<P, NW>, NW * 32, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    max_ctas = sms * per_sm;     // persistent: one wave of co-resident CTAs looping over the tiles
  }
  const long n_items = (long)E.n_elem * nvec;
  const long ntiles = (n_items + 31) / 32;
  const unsigned grid = (unsigned)(ntiles < max_ctas ? ntiles : max_ctas);
  k_nd_apply<P, NW><<<grid, NW * 32, smem, s>>>(T, with_cpk(E, nvec), x, y, z, nvec, ldx, ldy, n_items, ca, cm);
  return cudaGetLastError();
}

// y += ca * eps_e * S_c x_e with the dense element matrix S_c of the element's affine class (L x L complex,
// row-major S[l][k], from the Setup probe).  Same flops as the sum-factorised S0 at p <= 2 but only one
// barrier and ~L dependent steps per thread: the low-LATENCY variant for the coarse multigrid levels,
// where a launch has a few dozen CTAs and the thread-per-item kernel's serial chain is the whole cost.
// lane <-> item, warp <-> a chunk of output rows; S is read with (half-)warp-uniform addresses.
template <int P, int NWD>
__global__ void __launch_bounds__(NWD * 32)
k_h1_dense(const ElemData E, const double2 *__restrict__ S, const double2 *__restrict__ X,
           double2 *__restrict__ Y, int m, int ldx, int ldy, long n_items, double ca) {
  using D = Dim<P>;
  constexpr int L = D::LH1;
  constexpr int RPW = (L + NWD - 1) / NWD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sX = reinterpret_cast<double2 *>(smem_raw);      // [L][32] gathered inputs
  double2 *sS = sX + L * 32;                                  // [n_class][L][L] class matrices (single kappa only)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // batched k-points: nk * n_class matrices do not fit in shared memory; they are read through L1 instead (a tile
  // of 32 items spans at most a few (k-point, class) pairs, so the loads stay nearly warp-uniform)
  const bool staged = E.nk == 1;
  if (staged)
    for (int i = threadIdx.x; i < E.n_class * L * L; i += NWD * 32) sS[i] = S[i];
  const long item = (long)blockIdx.x * 32 + lane;
  const bool active = item < n_items;
  const int e = active ? (int)(item / m) : 0;
  const int v = active ? (int)(item - (long)e * m) : 0;
  const int32_t *mh = E.map_h1 + (long)e * L;
  for (int k = warp; k < L; k += NWD)
    sX[k * 32 + lane] = active ? X[(long)(__ldg(mh + k) - 1) * ldx + v] : make_double2(0.0, 0.0);
  __syncthreads();
  if (!active) return;
  const double2 *Sc = staged ? sS + E.cls[e] * L * L : S + (size_t)vclass(E, E.cls[e], v) * L * L;
  const int row0 = warp * RPW;
  double2 acc[RPW];
#pragma unroll
  for (int r = 0; r < RPW; r++) acc[r] = make_double2(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < L; k++) {
    const double2 xk = sX[k * 32 + lane];
#pragma unroll
    for (int r = 0; r < RPW; r++) {
      if (row0 + r < L) {
        const double2 a = Sc[(row0 + r) * L + k];
        acc[r].x = fma(a.x, xk.x, acc[r].x); acc[r].x = fma(-a.y, xk.y, acc[r].x);
        acc[r].y = fma(a.x, xk.y, acc[r].y); acc[r].y = fma(a.y, xk.x, acc[r].y);
      }
    }
  }
  const double cf = ca * E.eps[e];
  double *Yd = reinterpret_cast<double *>(Y);
#pragma unroll
  for (int r = 0; r < RPW; r++) {
    if (row0 + r < L) {
      const long o = 2 * ((long)(__ldg(mh + row0 + r) - 1) * ldy + v);
      atomicAdd(Yd + o, cf * acc[r].x);
      atomicAdd(Yd + o + 1, cf * acc[r].y);
    }
  }
}

// y += ca * S0_e x with S0 = G^H M1(eps) G: one THREAD per (element, vector) item, the whole element
// operator in registers (s0_item, elem_device.cuh), no block-level synchronisation after the class
// parameters are staged.  Used for the stiffness-only scalar applies of the multigrid V-cycle (p <= 2),
// where the cooperative tile kernel is bound by its barriers rather than by arithmetic.
// EVEC: instead of the atomic scatter, the element-local result goes to the E-vector Y[(e*L + k)*m + v]
// with plain coalesced stores; launch_h1_reduce then sums the copies of every dof.
template <int P, int NT, bool EVEC>
__global__ void __launch_bounds__(NT)
k_h1_s0_item(const __grid_constant__ Tabs T, const ElemData E, const double2 *__restrict__ X,
             double2 *__restrict__ Y, int m, int ldx, int ldy, long n_items, double ca) {
  using D = Dim<P>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *scol = reinterpret_cast<double2 *>(smem_raw) + threadIdx.x;
  double *sCP = reinterpret_cast<double *>(reinterpret_cast<double2 *>(smem_raw) + D::LH1 * NT);
  for (int i = threadIdx.x; i < cpar_doubles(E); i += NT) sCP[i] = E.cpar[i];
  __syncthreads();
  const long item = (long)blockIdx.x * NT + threadIdx.x;
  if (item >= n_items) return;
  const int e = (int)(item / m);
  const int v = (int)(item - (long)e * m);
  const int32_t *mh = E.map_h1 + (long)e * D::LH1;
#pragma unroll
  for (int k = 0; k < D::LH1; k++) scol[k * NT] = X[(long)(__ldg(mh + k) - 1) * ldx + v];
  double2 out[D::LH1];
  s0_item<P, NT>(T, sCP + kClassParDoubles * vclass(E, E.cls[e], v), ca * E.eps[e], scol, out);
  if (EVEC) {
#pragma unroll
    for (int k = 0; k < D::LH1; k++) Y[((long)e * D::LH1 + k) * m + v] = out[k];
    return;
  }
  double *Yd = reinterpret_cast<double *>(Y);
#pragma unroll
  for (int k = 0; k < D::LH1; k++) {
    const long o = 2 * ((long)(__ldg(mh + k) - 1) * ldy + v);
    atomicAdd(Yd + o, out[k].x);
    atomicAdd(Yd + o + 1, out[k].y);
  }
}

// The same operator with one LANE PAIR per (element, vector) item (s0_pair: real part in the even lane, imaginary part
// in the odd lane, Q^3 real accumulators per lane): the register-resident form for order 3, where the complex item
// of k_h1_s0_item (2 Q^3 = 128 accumulators) does not fit and the cooperative tile kernel k_h1_op<3> reaches 20 % of
// the fp64 pipe.  Gather / scatter addresses of a pair are adjacent doubles (coalesced 8-byte accesses).
// index load that the compiler neither hoists nor shares between the gather and the scatter (as __ldg it keeps all
// Q^3 indices of the element live across the whole item: 64 registers at order 3)
__device__ __forceinline__ int ld_index_once(const int32_t *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <int P, int NT>
__global__ void __launch_bounds__(NT)
k_h1_s0_pair(const __grid_constant__ Tabs T, const ElemData E, const double *__restrict__ X, double *__restrict__ Y,
             int m, int ldx, int ldy, long n_items, double ca) {
  using D = Dim<P>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *scol = reinterpret_cast<double *>(smem_raw) + threadIdx.x;
  PairTabs *sT = reinterpret_cast<PairTabs *>(reinterpret_cast<double *>(smem_raw) + D::LH1 * NT);
  double *sCP = reinterpret_cast<double *>(sT + 1);
  for (int i = threadIdx.x; i < cpar_doubles(E); i += NT) sCP[i] = E.cpar[i];
  for (int i = threadIdx.x; i < (kMaxP + 1) * (kMaxP + 1); i += NT) (&sT->TI[0][0])[i] = (&T.TI[0][0])[i];
  for (int i = threadIdx.x; i < kMaxP * (kMaxP + 1); i += NT) (&sT->Dt[0][0])[i] = (&T.Dt[0][0])[i];
  for (int i = threadIdx.x; i < kMaxP + 1; i += NT) sT->om[i] = T.om[i];
  __syncthreads();
  const long t = (long)blockIdx.x * NT + threadIdx.x;
  const bool live = t < 2 * n_items;
  const long item = live ? t >> 1 : 0;              // idle lanes of the last warp shadow item 0 and store nothing
  const int part = (int)(t & 1);
  const int e = (int)(item / m);
  const int v = (int)(item - (long)e * m);
  const int32_t *mh = E.map_h1 + (long)e * D::LH1;
  // gather / scatter in groups of 16 entries behind compiler fences: fully hoisted, the 64 index + address + value
  // registers of order 3 would not fit next to the accumulators
  constexpr int GR = 16;
#pragma unroll
  for (int k0 = 0; k0 < D::LH1; k0 += GR) {
#pragma unroll
    for (int k = k0; k < (k0 + GR < D::LH1 ? k0 + GR : D::LH1); k++)
      scol[k * NT] = X[2 * ((long)(ld_index_once(mh + k) - 1) * ldx + v) + part];
    asm volatile("" ::: "memory");
  }
  double out[D::LH1];
  s0_pair<P, NT>(*sT, sCP + kClassParDoubles * vclass(E, E.cls[e], v), ca * E.eps[e], part ? -1.0 : 1.0, scol, out);
  if (!live) return;
#pragma unroll
  for (int k0 = 0; k0 < D::LH1; k0 += GR) {
#pragma unroll
    for (int k = k0; k < (k0 + GR < D::LH1 ? k0 + GR : D::LH1); k++)
      atomicAdd(Y + 2 * ((long)(ld_index_once(mh + k) - 1) * ldy + v) + part, out[k]);
    asm volatile("" ::: "memory");
  }
}

// y[g][v] = sum over the local copies of H1 dof g of the E-vector Z (ptr/loc: dof -> 1-based positions e*L + k)
__global__ void k_h1_reduce(const int *__restrict__ ptr, const int32_t *__restrict__ loc,
                            const double2 *__restrict__ Z, double2 *__restrict__ Y, long n, int m) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long g = t / m;
    const int v = (int)(t - g * m);
    const int b = __ldg(ptr + g), e = __ldg(ptr + g + 1);
    double2 acc = make_double2(0.0, 0.0);
    for (int q = b; q < e; q++) {
      const double2 z = Z[(long)(__ldg(loc + q) - 1) * m + v];
      acc.x += z.x; acc.y += z.y;
    }
    Y[t] = acc;
  }
}

template <int P, int NW>
cudaError_t h1_op_t(int mode, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y,
                    int ldy, int nvec, double ca, double cm, cudaStream_t s) {
  using D = Dim<P>;
  const size_t smem = (size_t)(D::LND + D::LH1) * 32 * sizeof(double2) +
                      (size_t)E.nk * E.n_class * kClassParDoubles * sizeof(double);
  static bool attr_set_of[kMaxDevices] = {};
  bool &attr_set = attr_set_of[current_device_slot()];
  if (!attr_set) {
    cudaError_t err;
    if constexpr (P <= 3) {
      err = cudaFuncSetAttribute(k_h1_op<P, NW, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
      if (err != cudaSuccess) return err;
      err = cudaFuncSetAttribute(k_h1_op<P, NW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
      if (err != cudaSuccess) return err;
      err = cudaFuncSetAttribute(k_h1_op<P, NW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
      if (err != cudaSuccess) return err;
    }
    err = cudaFuncSetAttribute(k_h1_op<P, NW, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  const long n_items = (long)E.n_elem * nvec;
  const unsigned grid = (unsigned)((n_items + 31) / 32);
  const ElemData Ek = with_cpk(E, nvec);
  if constexpr (P <= 2) {
    static const int item_kernel = [] { const char *e = std::getenv("BLOCH_H1_ITEM"); return e ? std::atoi(e) : 1; }();
    const bool mode_evec = mode == 4;
    if ((mode == 3 || mode == 4) && cm == 0.0 && item_kernel) {
      constexpr int NT = 64;
      const size_t sm = (size_t)D::LH1 * NT * sizeof(double2) + (size_t)E.nk * E.n_class * kClassParDoubles * sizeof(double);
      static bool attr2_of[kMaxDevices] = {};
      bool &attr2 = attr2_of[current_device_slot()];
      if (!attr2) {
        cudaError_t err = cudaFuncSetAttribute(k_h1_s0_item<P, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(k_h1_s0_item<P, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (err != cudaSuccess) return err;
        attr2 = true;
      }
      if (mode_evec)
        k_h1_s0_item<P, NT, true><<<(unsigned)((n_items + NT - 1) / NT), NT, sm, s>>>(T, Ek, x, y, nvec, ldx, ldy, n_items, ca);
      else
        k_h1_s0_item<P, NT, false><<<(unsigned)((n_items + NT - 1) / NT), NT, sm, s>>>(T, Ek, x, y, nvec, ldx, ldy, n_items, ca);
      return cudaGetLastError();
    }
  }
  if constexpr (P == 3) {
    static const int pair_kernel = [] { const char *e = std::getenv("BLOCH_H1_PAIR"); return e ? std::atoi(e) : 1; }();
    if (mode == 3 && cm == 0.0 && pair_kernel) {
      constexpr int NT = 64;
      const size_t sm = (size_t)D::LH1 * NT * sizeof(double) + sizeof(PairTabs) +
                        (size_t)E.nk * E.n_class * kClassParDoubles * sizeof(double);
      static bool attr3_of[kMaxDevices] = {};
      bool &attr3 = attr3_of[current_device_slot()];
      if (!attr3) {
        cudaError_t err = cudaFuncSetAttribute(k_h1_s0_pair<P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (err != cudaSuccess) return err;
        attr3 = true;
      }
      k_h1_s0_pair<P, NT><<<(unsigned)((2 * n_items + NT - 1) / NT), NT, sm, s>>>(
          T, Ek, reinterpret_cast<const double *>(x), reinterpret_cast<double *>(y), nvec, ldx, ldy, n_items, ca);
      return cudaGetLastError();
    }
  }
  if (mode == 4) return cudaErrorInvalidValue;   // E-vector variant exists for the item kernel only
  if (mode == 3) {
    k_h1_op<P, NW, 3><<<grid, NW * 32, smem, s>>>(T, Ek, x, y, nvec, ldx, ldy, n_items, ca, cm);
    return cudaGetLastError();
  }
  if constexpr (P <= 3) {
    if (mode == 0) k_h1_op<P, NW, 0><<<grid, NW * 32, smem, s>>>(T, Ek, x, y, nvec, ldx, ldy, n_items, 1.0, 0.0);
    else if (mode == 1) k_h1_op<P, NW, 1><<<grid, NW * 32, smem, s>>>(T, Ek, x, y, nvec, ldx, ldy, n_items, 1.0, 0.0);
    else k_h1_op<P, NW, 2><<<grid, NW * 32, smem, s>>>(T, Ek, x, y, nvec, ldx, ldy, n_items, 1.0, 0.0);
    return cudaGetLastError();
  }
  return cudaErrorInvalidValue;
}

template <int P, int NW>
cudaError_t curl_t(const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy,
                   int nvec, cudaStream_t s) {
  using D = Dim<P>;
  const size_t smem = (size_t)(D::LND + D::LRT) * 32 * sizeof(double2) +
                      (size_t)E.nk * E.n_class * kClassParDoubles * sizeof(double);
  static bool attr_set_of[kMaxDevices] = {};
  bool &attr_set = attr_set_of[current_device_slot()];
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(k_curl<P, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  const long n_items = (long)E.n_elem * nvec;
  k_curl<P, NW><<<(unsigned)((n_items + 31) / 32), NW * 32, smem, s>>>(T, with_cpk(E, nvec), x, y, nvec, ldx, ldy, n_items);
  return cudaGetLastError();
}

}  // namespace

namespace {
__global__ void k_clear_rows(double2 *__restrict__ y, int ldy, int nvec, const int32_t *__restrict__ rows, long n_rows) {
  const long total = n_rows * nvec;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long r = t / nvec;
    y[(long)__ldg(rows + r) * ldy + (int)(t - r * nvec)] = make_double2(0.0, 0.0);
  }
}
}  // namespace
cudaError_t launch_clear_rows(double2 *y, int ldy, int nvec, const int32_t *rows, long n_rows, cudaStream_t s) {
  const long total = n_rows * nvec;
  long g = (total + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  k_clear_rows<<<(unsigned)g, 256, 0, s>>>(y, ldy, nvec, rows, n_rows);
  return cudaGetLastError();
}

cudaError_t launch_nd_apply(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx,
                            double2 *y, int ldy, int nvec, double ca, double cm, cudaStream_t s,
                            double2 *z) {
  if (z == nullptr) {
    bool launched = false;
    cudaError_t err = launch_nd_comp(p, T, E, x, ldx, y, ldy, nvec, ca, cm, s, &launched);
    if (err != cudaSuccess || launched) return err;
    err = launch_nd_item(p, T, E, x, ldx, y, ldy, nvec, ca, cm, s, &launched);
    if (err != cudaSuccess || launched) return err;
    if (E.fresh_y && E.partial_clear) {   // the kernel below reduces into every row: clear the interior rows as well
      err = cudaMemset2DAsync(y, sizeof(double2) * ldy, 0, sizeof(double2) * nvec, (size_t)E.n_rows_y, s);
      if (err != cudaSuccess) return err;
    }
  }
  static int variant = -1;
  if (variant < 0) { const char *e = std::getenv("BLOCH_ND_WARPS"); variant = e ? std::atoi(e) : 0; }
  if (variant == 1) {
    switch (p) {
      case 1: return nd_apply_t<1, 1>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
      case 2: return nd_apply_t<2, 1>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
      case 3: return nd_apply_t<3, 4>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
    }
  } else if (variant == 2) {
    switch (p) {
      case 1: return nd_apply_t<1, 2>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
      case 2: return nd_apply_t<2, 2>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
      case 3: return nd_apply_t<3, 6>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
    }
  } else if (variant == 3) {
    switch (p) {
      case 1: return nd_apply_t<1, 3>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
      case 2: return nd_apply_t<2, 3>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
      case 3: return nd_apply_t<3, 8>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
    }
  }
  switch (p) {
    case 1: return nd_apply_t<1, 6>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
    case 2: return nd_apply_t<2, 9>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
    case 3: return nd_apply_t<3, 12>(T, E, x, ldx, y, ldy, nvec, ca, cm, z, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_h1_op(int p, int mode, const Tabs &T, const ElemData &E, const double2 *x, int ldx,
                         double2 *y, int ldy, int nvec, cudaStream_t s, double ca, double cm) {
  switch (p) {
    case 1: return h1_op_t<1, 4>(mode, T, E, x, ldx, y, ldy, nvec, ca, cm, s);
    case 2: return h1_op_t<2, 9>(mode, T, E, x, ldx, y, ldy, nvec, ca, cm, s);
    case 3: return h1_op_t<3, 12>(mode, T, E, x, ldx, y, ldy, nvec, ca, cm, s);
    case 4: return h1_op_t<4, 15>(mode, T, E, x, ldx, y, ldy, nvec, ca, cm, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_h1_dense(int p, const ElemData &E, const double2 *S, const double2 *x, int ldx, double2 *y,
                            int ldy, int nvec, double ca, cudaStream_t s) {
  const long n_items = (long)E.n_elem * nvec;
  const unsigned grid = (unsigned)((n_items + 31) / 32);
  const int L = (p + 1) * (p + 1) * (p + 1);
  const size_t smem = ((size_t)L * 32 + (E.nk == 1 ? (size_t)E.n_class * L * L : 0)) * sizeof(double2);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static bool attr_set_of[kMaxDevices] = {};
  bool &attr_set = attr_set_of[current_device_slot()];
  if (!attr_set) {
    cudaError_t err = cudaFuncSetAttribute(k_h1_dense<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(k_h1_dense<2, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (err != cudaSuccess) return err;
    attr_set = true;
  }
  switch (p) {
    case 1: k_h1_dense<1, 4><<<grid, 4 * 32, smem, s>>>(with_cpk(E, nvec), S, x, y, nvec, ldx, ldy, n_items, ca); break;
    case 2: k_h1_dense<2, 9><<<grid, 9 * 32, smem, s>>>(with_cpk(E, nvec), S, x, y, nvec, ldx, ldy, n_items, ca); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_h1_reduce(const int *ptr, const int32_t *loc, const double2 *Z, double2 *Y, long n, int m,
                             cudaStream_t s) {
  long g = (n * m + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  k_h1_reduce<<<(unsigned)g, 256, 0, s>>>(ptr, loc, Z, Y, n, m);
  return cudaGetLastError();
}

cudaError_t launch_curl(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y,
                        int ldy, int nvec, cudaStream_t s) {
  switch (p) {
    case 1: return curl_t<1, 6>(T, E, x, ldx, y, ldy, nvec, s);
    case 2: return curl_t<2, 9>(T, E, x, ldx, y, ldy, nvec, s);
    case 3: return curl_t<3, 12>(T, E, x, ldx, y, ldy, nvec, s);
    default: return cudaErrorInvalidValue;
  }
}

// ==========================================================================================
// layout conversion and block algebra
// ==========================================================================================
namespace {

__global__ void k_pack(const double *__restrict__ reim, double2 *__restrict__ blk, long n, int nvec) {
  const long total = n * nvec;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long i = t / nvec;
    const int v = (int)(t - i * nvec);
    const double *src = reim + (long)v * 2 * n;
    blk[t] = make_double2(src[i], src[n + i]);
  }
}
__global__ void k_unpack(const double2 *__restrict__ blk, double *__restrict__ reim, long n, int nvec) {
  const long total = n * nvec;
  // iterate in output order for coalesced stores
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t / n);
    const long i = t - (long)v * n;
    const double2 z = blk[i * nvec + v];
    double *dst = reim + (long)v * 2 * n;
    dst[i] = z.x;
    dst[n + i] = z.y;
  }
}

// Gram: each CTA reduces a row chunk into a register tile, then atomically adds into C.
// thread (ti, tj) owns C entries (i, j) with i = ti + k*TI... simple generic version.
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_ROWS = 16;   // rows staged per iteration: 16 x (64 + 64) x 16 B = 32 KB
// Register-tiled: the 256 threads form a 16 x 16 grid, thread (ti, tj) owns the 4 x 4 outputs
// C[ti + 16 a][tj + 16 b]; per staged row it reads 4 + 4 shared values for 16 complex MACs.
__global__ void __launch_bounds__(GRAM_THREADS)
k_gram(const double2 *__restrict__ A, int ma, int lda, const double2 *__restrict__ B, int mb,
       int ldb, long n, double2 *__restrict__ C, long rows_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sA = reinterpret_cast<double2 *>(smem_raw);   // [GRAM_ROWS][64] (zero padded)
  double2 *sB = sA + GRAM_ROWS * 64;                      // [GRAM_ROWS][64]
  const long r0 = blockIdx.x * rows_per_cta;
  const long r1 = min(n, r0 + rows_per_cta);
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  double2 acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = make_double2(0.0, 0.0);
  for (long r = r0; r < r1; r += GRAM_ROWS) {
    const int nr = (int)min((long)GRAM_ROWS, r1 - r);
    for (int t = threadIdx.x; t < GRAM_ROWS * 64; t += GRAM_THREADS) {
      const int q = t >> 6, c = t & 63;
      sA[t] = (q < nr && c < ma) ? A[(r + q) * lda + c] : make_double2(0.0, 0.0);
      sB[t] = (q < nr && c < mb) ? B[(r + q) * ldb + c] : make_double2(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll 4
    for (int q = 0; q < GRAM_ROWS; q++) {
      double2 x[4], y[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { x[a] = sA[q * 64 + ti + 16 * a]; y[a] = sB[q * 64 + tj + 16 * a]; }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          acc[a][b].x = fma(x[a].x, y[b].x, acc[a][b].x); acc[a][b].x = fma(x[a].y, y[b].y, acc[a][b].x);
          acc[a][b].y = fma(x[a].x, y[b].y, acc[a][b].y); acc[a][b].y = fma(-x[a].y, y[b].x, acc[a][b].y);
        }
    }
    __syncthreads();
  }
  double *Cd = reinterpret_cast<double *>(C);
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int i = ti + 16 * a, j = tj + 16 * b;
      if (i < ma && j < mb) {
        atomicAdd(Cd + 2 * (i * mb + j), acc[a][b].x);
        atomicAdd(Cd + 2 * (i * mb + j) + 1, acc[a][b].y);
      }
    }
}

// Y = beta*Y + X*C : one thread per (row, output column)
__global__ void k_block_mult(const double2 *__restrict__ X, int k, const double2 *__restrict__ C, int m,
                             double2 *__restrict__ Y, double beta, long n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *sC = reinterpret_cast<double2 *>(smem_raw);
  for (int t = threadIdx.x; t < k * m; t += blockDim.x) sC[t] = C[t];
  __syncthreads();
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long r = t / m;
    const int j = (int)(t - r * m);
    double2 acc = make_double2(0.0, 0.0);
    if (beta != 0.0) { const double2 y = Y[t]; acc.x = beta * y.x; acc.y = beta * y.y; }
    const double2 *xr = X + r * k;
    for (int i = 0; i < k; i++) {
      const double2 x = xr[i], c = sC[i * m + j];
      acc.x = fma(x.x, c.x, acc.x); acc.x = fma(-x.y, c.y, acc.x);
      acc.y = fma(x.x, c.y, acc.y); acc.y = fma(x.y, c.x, acc.y);
    }
    Y[t] = acc;
  }
}

__global__ void k_residual(const double2 *__restrict__ AX, const double2 *__restrict__ MX,
                           const double *__restrict__ lam, double2 *__restrict__ R, long n, int m) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int j = (int)(t % m);
    const double l = lam[j];
    const double2 a = AX[t], b = MX[t];
    R[t] = make_double2(a.x - l * b.x, a.y - l * b.y);
  }
}
__global__ void k_axpby(double a, const double2 *__restrict__ X, double b, double2 *__restrict__ Y, long total) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double2 x = X[t];
    double2 y = make_double2(0.0, 0.0);
    if (b != 0.0) y = Y[t];
    Y[t] = make_double2(a * x.x + b * y.x, a * x.y + b * y.y);
  }
}
__global__ void k_diag_scale(const double *__restrict__ d, const double2 *__restrict__ X,
                             double2 *__restrict__ Y, long n, int m) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double s = d[t / m];
    const double2 x = X[t];
    Y[t] = make_double2(s * x.x, s * x.y);
  }
}
__global__ void k_col_axpy(const double *__restrict__ alpha, double sign, const double2 *__restrict__ X,
                           double2 *__restrict__ Y, long n, int m) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double a = sign * alpha[t % m];
    const double2 x = X[t];
    double2 y = Y[t];
    y.x = fma(a, x.x, y.x); y.y = fma(a, x.y, y.y);
    Y[t] = y;
  }
}
__global__ void k_col_xpby(const double2 *__restrict__ Z, const double *__restrict__ beta,
                           double2 *__restrict__ Pp, long n, int m) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double b = beta[t % m];
    const double2 z = Z[t], p = Pp[t];
    Pp[t] = make_double2(fma(b, p.x, z.x), fma(b, p.y, z.y));
  }
}
// column dots: CTA-level reduction per column then atomicAdd
__global__ void k_col_dot(const double2 *__restrict__ A, const double2 *__restrict__ B, long n, int m,
                          double *__restrict__ d) {
  // only the first `usable` threads work, usable a multiple of m, so that the grid stride keeps
  // every thread on one fixed column
  extern __shared__ double sred[];
  const int tid = threadIdx.x;
  const long total = n * m;
  const long nthreads = (long)gridDim.x * blockDim.x;
  const long usable = (nthreads / m) * m;
  double acc = 0.0;
  const long start = blockIdx.x * (long)blockDim.x + tid;
  if (start < usable) {
    for (long t = start; t < total; t += usable) {
      const double2 a = A[t], b = B[t];
      acc = fma(a.x, b.x, acc);
      acc = fma(a.y, b.y, acc);
    }
  }
  for (int j = tid; j < m; j += blockDim.x) sred[j] = 0.0;
  __syncthreads();
  if (start < usable) atomicAdd(&sred[start % m], acc);
  __syncthreads();
  for (int j = tid; j < m; j += blockDim.x) atomicAdd(d + j, sred[j]);
}
__global__ void k_scalar_div(const double *num, const double *den, double *out, int m) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[j] = (den[j] != 0.0) ? num[j] / den[j] : 0.0;
}
__global__ void k_scatter_diag(const int32_t *__restrict__ map, int L, const int *__restrict__ cls,
                               const double *__restrict__ coef, const double *__restrict__ dloc,
                               int n_elem, double *__restrict__ d, int nk, int n_class) {
  const long total = (long)n_elem * L;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int e = (int)(t / L), l = (int)(t - (long)e * L);
    const int s = map[t];
    const long g = (s < 0 ? -s : s) - 1;
    for (int k = 0; k < nk; k++) atomicAdd(d + g * nk + k, coef[e] * dloc[((long)k * n_class + cls[e]) * L + l]);
  }
}
__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_fill_random(double2 *X, long total, unsigned long long seed) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const unsigned long long a = splitmix(seed + 2 * (unsigned long long)t);
    const unsigned long long b = splitmix(seed + 2 * (unsigned long long)t + 1);
    X[t] = make_double2((double)(a >> 11) * (2.0 / 9007199254740992.0) - 1.0,
                        (double)(b >> 11) * (2.0 / 9007199254740992.0) - 1.0);
  }
}

inline unsigned grid_for(long total, int threads) {
  long g = (total + threads - 1) / threads;
  const long cap = 148L * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t launch_pack(const double *reim, double2 *blk, long n, int nvec, cudaStream_t s) {
  k_pack<<<grid_for(n * nvec, 256), 256, 0, s>>>(reim, blk, n, nvec);
  return cudaGetLastError();
}
cudaError_t launch_unpack(const double2 *blk, double *reim, long n, int nvec, cudaStream_t s) {
  k_unpack<<<grid_for(n * nvec, 256), 256, 0, s>>>(blk, reim, n, nvec);
  return cudaGetLastError();
}
cudaError_t launch_gram(const double2 *A, int ma, int lda, const double2 *B, int mb, int ldb, long n,
                        double2 *C, cudaStream_t s) {
  if (ma > 64 || mb > 64) return cudaErrorInvalidValue;
  cudaError_t err = cudaMemsetAsync(C, 0, sizeof(double2) * ma * mb, s);
  if (err != cudaSuccess) return err;
  long ctas = 148 * 2;
  long rows = (n + ctas - 1) / ctas;
  rows = ((rows + GRAM_ROWS - 1) / GRAM_ROWS) * GRAM_ROWS;
  ctas = (n + rows - 1) / rows;
  const size_t smem = (size_t)GRAM_ROWS * 128 * sizeof(double2);
  k_gram<<<(unsigned)ctas, GRAM_THREADS, smem, s>>>(A, ma, lda, B, mb, ldb, n, C, rows);
  return cudaGetLastError();
}
cudaError_t launch_block_mult(const double2 *X, int k, const double2 *C, int m, double2 *Y,
                              double beta, long n, cudaStream_t s) {
  k_block_mult<<<grid_for(n * m, 256), 256, (size_t)k * m * sizeof(double2), s>>>(X, k, C, m, Y, beta, n);
  return cudaGetLastError();
}
cudaError_t launch_residual(const double2 *AX, const double2 *MX, const double *lambda, double2 *R,
                            long n, int m, cudaStream_t s) {
  k_residual<<<grid_for(n * m, 256), 256, 0, s>>>(AX, MX, lambda, R, n, m);
  return cudaGetLastError();
}
cudaError_t launch_axpby(double a, const double2 *X, double b, double2 *Y, long total, cudaStream_t s) {
  k_axpby<<<grid_for(total, 256), 256, 0, s>>>(a, X, b, Y, total);
  return cudaGetLastError();
}
cudaError_t launch_diag_scale(const double *d, const double2 *X, double2 *Y, long n, int m, cudaStream_t s) {
  k_diag_scale<<<grid_for(n * m, 256), 256, 0, s>>>(d, X, Y, n, m);
  return cudaGetLastError();
}
cudaError_t launch_col_axpy(const double *alpha, double sign, const double2 *X, double2 *Y, long n,
                            int m, cudaStream_t s) {
  k_col_axpy<<<grid_for(n * m, 256), 256, 0, s>>>(alpha, sign, X, Y, n, m);
  return cudaGetLastError();
}
cudaError_t launch_col_xpby(const double2 *Z, const double *beta, double2 *P, long n, int m,
                            cudaStream_t s) {
  k_col_xpby<<<grid_for(n * m, 256), 256, 0, s>>>(Z, beta, P, n, m);
  return cudaGetLastError();
}
cudaError_t launch_col_dot(const double2 *A, const double2 *B, long n, int m, double *d,
                           cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(d, 0, sizeof(double) * m, s);
  if (err != cudaSuccess) return err;
  k_col_dot<<<grid_for(n * m, 256) > 296 ? 296 : grid_for(n * m, 256), 256, sizeof(double) * m, s>>>(A, B, n, m, d);
  return cudaGetLastError();
}
cudaError_t launch_scalar_div(const double *num, const double *den, double *out, int m, cudaStream_t s) {
  k_scalar_div<<<(m + 63) / 64, 64, 0, s>>>(num, den, out, m);
  return cudaGetLastError();
}
cudaError_t launch_scatter_diag(const int32_t *map, int L, const int *cls, const double *coef,
                                const double *dloc, int n_elem, double *d, cudaStream_t s, int nk, int n_class) {
  k_scatter_diag<<<grid_for((long)n_elem * L, 256), 256, 0, s>>>(map, L, cls, coef, dloc, n_elem, d, nk, n_class);
  return cudaGetLastError();
}
cudaError_t launch_fill_random(double2 *X, long total, unsigned long long seed, cudaStream_t s) {
  k_fill_random<<<grid_for(total, 256), 256, 0, s>>>(X, total, seed);
  return cudaGetLastError();
}

// ---- second pass of the atomic-free apply: y[g][v] = sum over the local copies of dof g ----
namespace {
__global__ void k_nd_reduce(const int *__restrict__ ptr, const int32_t *__restrict__ loc,
                            const double2 *__restrict__ Z, double2 *__restrict__ Y, long n, int m, int ldy) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long g = t / m;
    const int v = (int)(t - g * m);
    const int b = __ldg(ptr + g), e = __ldg(ptr + g + 1);
    double2 acc = make_double2(0.0, 0.0);
    for (int k = b; k < e; k++) {
      const int s = __ldg(loc + k);
      const double2 z = Z[(long)((s < 0 ? -s : s) - 1) * m + v];
      if (s < 0) { acc.x -= z.x; acc.y -= z.y; } else { acc.x += z.x; acc.y += z.y; }
    }
    Y[g * ldy + v] = acc;
  }
}
}  // namespace

cudaError_t launch_nd_reduce(const int *ptr, const int32_t *loc, const double2 *z, double2 *y, long n,
                             int m, int ldy, cudaStream_t s) {
  const long total = n * m;
  long g = (total + 255) / 256;
  if (g > 148L * 32) g = 148L * 32;
  k_nd_reduce<<<(unsigned)(g < 1 ? 1 : g), 256, 0, s>>>(ptr, loc, z, y, n, m, ldy);
  return cudaGetLastError();
}

// ---- fp64 FMA throughput probe (roofline denominator for the flop-bound side) ----
namespace {
__global__ void k_fp64_peak(double *out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

cudaError_t measure_fp64_peak(double *tflops, cudaStream_t s) {
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  double *d = nullptr;
  cudaError_t err = cudaMalloc(&d, sizeof(double) * blocks * threads);
  if (err != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fp64_peak<<<blocks, threads, 0, s>>>(d, 1000);
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0, s);
    k_fp64_peak<<<blocks, threads, 0, s>>>(d, iters);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  err = cudaGetLastError();
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return err;
}

}  // namespace bloch_b200
