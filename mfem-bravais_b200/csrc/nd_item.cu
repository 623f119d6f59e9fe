// y += ca * A x + cm * M x on ND block vectors, orders 1 and 2: barrier-free "lane pair per item" kernel.
//
// Same mode-space element operator as k_nd_apply (kernels.cu), different mapping to the machine:
//   * one (element, vector) item is handled by a PAIR of lanes: the even lane carries the real parts, the odd
//     lane the imaginary parts.  Every 1-D contraction is complex-by-real, so the two lanes run the same real
//     code; only the pointwise Bloch terms -i kappa_hat couple them (partner value from the partner's
//     shared-memory column in the curl, from a shuffle in the adjoint curl);
//   * the ND element tensor of a lane lives in a lane-private shared-memory column (entry k at col[32 k]:
//     conflict free), the RT tensor (3 p^2 (p+1) reals) in registers; all tensor indices are compile time,
//     there is no task loop, no index arithmetic and no block barrier (two __syncwarp per item);
//   * a warp works through its own tiles of 16 items, 16 warps per SM, so gathers (long scoreboard) of one warp
//     overlap the fp64 phases of the others without any explicit pipelining.
// Per item and real part at p = 2: 1188 DFMA of contractions + ~350 pointwise, 54 gathers, 54 RED.ADD.F64.
#include "kernels.hpp"

#include <cmath>
#include <cstdlib>

#include "elem_device.cuh"

namespace bloch_b200 {

namespace {

using namespace dev;

// Tables of this kernel (orders <= 2), SCALED so that every mode coefficient carries the square root of its diagonal
// 1-D mass weight:  TI = diag(sqrt(om)) TI,  Dt[a][r] = sqrt(om[a]) Dt[a][r] / sqrt(om[r]),  TIo[o] = sqrt(om[o]) TI
// (the first pass of slab o also applies the weight of the slab's open direction; open Gauss-Lagrange directions
// have the weights om[a], a < p, of the first p closed modes).  Both mode-space mass matrices are then the plain
// pointwise 3x3 couplings G and H, and the Bloch shift -i kappa_hat stays the identity between mode a and open
// point a: no weight products per grid point.
constexpr int kItemMaxP = 2;
struct ItemTabs {
  double TI[kItemMaxP + 1][kItemMaxP + 1];
  double TIo[kItemMaxP][kItemMaxP + 1][kItemMaxP + 1];
  double Dt[kItemMaxP][kItemMaxP + 1];
};

ItemTabs item_tabs(const Tabs &T, int p) {
  ItemTabs S = {};
  double sq[kItemMaxP + 1];
  for (int r = 0; r <= p; r++) sq[r] = std::sqrt(T.om[r]);
  for (int r = 0; r <= p; r++)
    for (int j = 0; j <= p; j++) {
      S.TI[r][j] = sq[r] * T.TI[r][j];
      for (int o = 0; o < p; o++) S.TIo[o][r][j] = sq[o] * sq[r] * T.TI[r][j];
    }
  for (int a = 0; a < p; a++)
    for (int r = 0; r <= p; r++) S.Dt[a][r] = sq[a] * T.Dt[a][r] / sq[r];
  return S;
}

// s[a][b] (Q x Q): apply M1 along b, then M2 along a.  FWD: out[r] = sum_j M[r][j] in[j]; ADJ: out[j] = sum_r M[r][j] in[r]
template <int P, bool ADJ>
__device__ __forceinline__ void slab_tf(double (&s)[P + 1][P + 1], const double (&M1)[kItemMaxP + 1][kItemMaxP + 1],
                                        const double (&M2)[kItemMaxP + 1][kItemMaxP + 1]) {
  constexpr int Q = P + 1;
  double u[Q][Q];
#pragma unroll
  for (int a = 0; a < Q; a++)
#pragma unroll
    for (int r = 0; r < Q; r++) {
      double acc = (ADJ ? M1[0][r] : M1[r][0]) * s[a][0];
#pragma unroll
      for (int j = 1; j < Q; j++) acc = fma(ADJ ? M1[j][r] : M1[r][j], s[a][j], acc);
      u[a][r] = acc;
    }
#pragma unroll
  for (int r = 0; r < Q; r++)
#pragma unroll
    for (int b = 0; b < Q; b++) {
      double acc = (ADJ ? M2[0][r] : M2[r][0]) * u[0][b];
#pragma unroll
      for (int j = 1; j < Q; j++) acc = fma(ADJ ? M2[j][r] : M2[r][j], u[j][b], acc);
      s[r][b] = acc;
    }
}

// x with the sign of the map entry s applied (one LOP3 on the high word)
__device__ __forceinline__ double flip(double x, int s) {
  return __hiloint2double(__double2hiint(x) ^ (s & (int)0x80000000), __double2loint(x));
}


// partner lane's value; volatile so that the two uses of one RT entry are not merged into one long-lived register pair
__device__ __forceinline__ double partner(double x) {
  int lo = __double2loint(x), hi = __double2hiint(x);
  asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(lo));
  asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(hi));
  return __hiloint2double(hi, lo);
}

// NT = threads per block = the register budget (65536 / NT per thread): 512 -> 128 registers, 16 warps per SM
// GV (gather variant): 0 = component by component (index loads, value loads, transform; compiler fences in between
// keep the register pressure of the 168-register build down), 1 = all indices first (8-byte loads), no fences.
template <int P, bool HAS_A, bool HAS_M, int GV>
__device__ __forceinline__ void nd_item_body(const ItemTabs &T, const ElemData &E, const double *__restrict__ X,
                                             double *__restrict__ Y, int m, int ldx, int ldy, long n_items, double ca,
                                             double cm) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sCP = reinterpret_cast<double *>(smem_raw);
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: no divergent-path code
  const int ncp = cpar_doubles(E);
  for (int i = threadIdx.x; i < ncp; i += blockDim.x) sCP[i] = E.cpar[i];
  __syncthreads();
  double *col = sCP + ((ncp + 1) & ~1) + (size_t)warp * (D::LND * 32) + lane;   // own column
  const double *pcol = col - lane + (lane ^ 1);                                  // partner's column
  const int part = lane & 1;
  const double sg = part ? -1.0 : 1.0;
  const long ntiles = (n_items + 15) >> 4;
  const bool small = n_items < 0x7fffffffL;
  // 32-bit offsets in doubles (the launcher checks 2 * n_dofs * ld < 2^32):  (|s| - 1) * 2 ld + 2 v + part
  const unsigned xstep = 2u * (unsigned)ldx, ystep = 2u * (unsigned)ldy;
  const long tstride = (long)gridDim.x * nwarps;

  // element / vector of this lane's item in a tile; idle lanes of the last tile shadow a valid item and add zeros
  auto locate = [&](long tile, int &e, int &v) -> bool {
    long item = tile * 16 + (lane >> 1);
    const bool act = item < n_items;
    if (!act) item = n_items - 1;
    e = small ? (int)((unsigned)item / (unsigned)m) : (int)(item / m);
    v = (int)(item - (long)e * m);
    return act;
  };
  long tile = (long)warp * gridDim.x + blockIdx.x;
  for (; tile < ntiles; tile += tstride) {
    int e, v;
    const bool active = locate(tile, e, v);
    const int32_t *mp = E.map_nd + (long)e * D::LND;
    const double *cp = sCP + kClassParDoubles * vclass(E, __ldg(E.cls + e), v);
    const unsigned xoff = 2u * (unsigned)v + (unsigned)part - xstep, yoff = 2u * (unsigned)v + (unsigned)part - ystep;

    // ---- signed gather + nodal -> mode in the closed directions, one component at a time ----
    int aidx[GV == 1 ? D::LND : 1];
    if (GV == 1) {
      static_assert(D::LND % 2 == 0, "8-byte index loads");
#pragma unroll
      for (int k = 0; k < D::LND / 2; k++) {
        const int2 t = __ldg(reinterpret_cast<const int2 *>(mp) + k);
        aidx[2 * k] = t.x; aidx[2 * k + 1] = t.y;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      int sidx[D::NB];
#pragma unroll
      for (int k = 0; k < D::NB; k++) sidx[k] = GV == 1 ? aidx[c * D::NB + k] : __ldg(mp + c * D::NB + k);
      double xv[D::NB];
#pragma unroll
      for (int k = 0; k < D::NB; k++) {
        const int s = sidx[k];
        const unsigned off = (unsigned)(s < 0 ? -s : s) * xstep + xoff;
        xv[k] = flip(__ldg(X + off), s);
      }
#pragma unroll
      for (int o = 0; o < P; o++) {
        double s[Q][Q];
#pragma unroll
        for (int a = 0; a < Q; a++)
#pragma unroll
          for (int b = 0; b < Q; b++) {
            const int k = (o * Q + a) * Q + b;
            s[a][b] = xv[k];
          }
        slab_tf<P, false>(s, T.TIo[o], T.TI);
#pragma unroll
        for (int a = 0; a < Q; a++)
#pragma unroll
          for (int b = 0; b < Q; b++) col[D::nd(c, o, a, b) * 32] = s[a][b];
      }
      if (GV == 0) asm volatile("" ::: "memory");   // keep the three components' gathers apart (register pressure)
    }
    __syncwarp();

    double R[HAS_A ? D::LRT : 1];
    double ks[3];
    if (HAS_A) {
#pragma unroll
      for (int d = 0; d < 3; d++) ks[d] = sg * cp[d];
      // ---- Bloch curl: R_c = K_{c+1} F_{c+2} - K_{c+2} F_{c+1},  K_d = Dt - i kh_d (pointwise) ----
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
#pragma unroll
        for (int j = 0; j < Q; j++) {
          double A[P][Q], B[P][Q];
#pragma unroll
          for (int o = 0; o < P; o++)
#pragma unroll
            for (int r = 0; r < Q; r++) {
              A[o][r] = col[D::nd(c2, o, j, r) * 32];   // F_{c+2}[o2=o, j1=j, j2=r]
              B[o][r] = col[D::nd(c1, o, r, j) * 32];   // F_{c+1}[o1=o, j1=r, j2=j]
            }
#pragma unroll
          for (int o1 = 0; o1 < P; o1++)
#pragma unroll
            for (int o2 = 0; o2 < P; o2++) {
              const double pa = pcol[D::nd(c2, o2, j, o1) * 32];
              const double pb = pcol[D::nd(c1, o1, o2, j) * 32];
              double acc = ks[c1] * pa;
              acc = fma(-ks[c2], pb, acc);
#pragma unroll
              for (int r = 0; r < Q; r++) {
                acc = fma(T.Dt[o1][r], A[o2][r], acc);
                acc = fma(-T.Dt[o2][r], B[o1][r], acc);
              }
              R[D::rt(c, j, o1, o2)] = acc;
            }
        }
      }
      // ---- pointwise RT mass, scaled by ca * muinv (mode-space weights are 1 with the scaled tables) ----
      const double wA = active ? ca * __ldg(E.muinv + e) : 0.0;   // idle lanes produce exact zeros
      double G[9];
#pragma unroll
      for (int k = 0; k < 9; k++) G[k] = wA * cp[3 + k];
#pragma unroll
      for (int i0 = 0; i0 < Q; i0++)
#pragma unroll
        for (int i1 = 0; i1 < Q; i1++)
#pragma unroll
          for (int i2 = 0; i2 < Q; i2++) {
            const int i[3] = {i0, i1, i2};
            const int nopen = (i0 < P) + (i1 < P) + (i2 < P);
            if (nopen < 2) continue;
            if (nopen == 3) {
              const int l0 = D::rt(0, i0, i1, i2), l1 = D::rt(1, i1, i2, i0), l2 = D::rt(2, i2, i0, i1);
              const double f0 = R[l0], f1 = R[l1], f2 = R[l2];
              R[l0] = fma(G[0], f0, fma(G[1], f1, G[2] * f2));
              R[l1] = fma(G[3], f0, fma(G[4], f1, G[5] * f2));
              R[l2] = fma(G[6], f0, fma(G[7], f1, G[8] * f2));
            } else {
              const int c = i0 == P ? 0 : (i1 == P ? 1 : 2);     // the only component living at this point
              const int l = D::rt(c, i[c], i[(c + 1) % 3], i[(c + 2) % 3]);
              R[l] *= G[4 * c];
            }
          }
    }
    __syncwarp();   // partner reads of the columns are done

    if (HAS_M) {
      // ---- pointwise ND mass in mode space, in place, scaled by cm * eps ----
      const double wM = active ? cm * __ldg(E.eps + e) : 0.0;
      double H[9];
#pragma unroll
      for (int k = 0; k < 9; k++) H[k] = wM * cp[12 + k];
#pragma unroll
      for (int i0 = 0; i0 < Q; i0++)
#pragma unroll
        for (int i1 = 0; i1 < Q; i1++)
#pragma unroll
          for (int i2 = 0; i2 < Q; i2++) {
            const int i[3] = {i0, i1, i2};
            const int nopen = (i0 < P) + (i1 < P) + (i2 < P);
            if (nopen == 0) continue;
            double f[3];
            int loc[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
              loc[c] = D::nd(c, i[c] < P ? i[c] : 0, i[(c + 1) % 3], i[(c + 2) % 3]);
              f[c] = i[c] < P ? col[loc[c] * 32] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 3; c++)
              if (i[c] < P) {
                double r = 0.0;
                bool first = true;
#pragma unroll
                for (int d = 0; d < 3; d++)
                  if (i[d] < P) { r = first ? H[3 * c + d] * f[d] : fma(H[3 * c + d], f[d], r); first = false; }
                col[loc[c] * 32] = r;
              }
          }
    }

    // ---- adjoint curl on top of the mass part, mode -> nodal (adjoint), signed scatter-add ----
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
#pragma unroll
      for (int o = 0; o < P; o++) {
        double f[Q][Q];
#pragma unroll
        for (int a = 0; a < Q; a++)
#pragma unroll
          for (int b = 0; b < Q; b++) f[a][b] = HAS_M ? col[D::nd(c, o, a, b) * 32] : 0.0;
        if (HAS_A) {
#pragma unroll
          for (int j1 = 0; j1 < Q; j1++)
#pragma unroll
            for (int j2 = 0; j2 < Q; j2++) {
              double acc = f[j1][j2];
#pragma unroll
              for (int q = 0; q < P; q++) {
                acc = fma(T.Dt[q][j2], R[D::rt(c1, j1, q, o)], acc);     // Y_{c+1}[j=j1, o1=q, o2=o]
                acc = fma(-T.Dt[q][j1], R[D::rt(c2, j2, o, q)], acc);    // Y_{c+2}[j=j2, o1=o, o2=q]
              }
              if (j2 < P) acc = fma(-ks[c2], partner(R[D::rt(c1, j1, j2 < P ? j2 : 0, o)]), acc);
              if (j1 < P) acc = fma(ks[c1], partner(R[D::rt(c2, j2, o, j1 < P ? j1 : 0)]), acc);
              f[j1][j2] = acc;
            }
        }
        slab_tf<P, true>(f, T.TIo[o], T.TI);
#pragma unroll
        for (int a = 0; a < Q; a++)
#pragma unroll
          for (int b = 0; b < Q; b++) {
            const int s = __ldg(mp + D::nd(c, o, a, b));
            const unsigned off = (unsigned)(s < 0 ? -s : s) * ystep + yoff;
            const bool interior = a > 0 && a < P && b > 0 && b < P;      // compile time: dof of no other element
            if (interior && E.fresh_y) { if (active) Y[off] = flip(f[a][b], s); }
            else atomicAdd(Y + off, flip(f[a][b], s));
          }
        asm volatile("" ::: "memory");
      }
    }
    __syncwarp();   // all lanes are done with this tile's columns before the next tile overwrites them
  }
}

template <int P, bool HAS_A, bool HAS_M, int NT, int GV>
__global__ void __launch_bounds__(NT, 1)
k_nd_item(const __grid_constant__ ItemTabs T, const ElemData E, const double *__restrict__ X, double *__restrict__ Y,
          int m, int ldx, int ldy, long n_items, double ca, double cm) {
  nd_item_body<P, HAS_A, HAS_M, GV>(T, E, X, Y, m, ldx, ldy, n_items, ca, cm);
}

template <int P, bool HAS_A, bool HAS_M, int NT, int GV>
cudaError_t nd_item_t(const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy, int nvec,
                      double ca, double cm, cudaStream_t s, bool *fits) {
  using D = Dim<P>;
  auto kern = k_nd_item<P, HAS_A, HAS_M, NT, GV>;
  // per-device one-time setup (a process may hold handles on several devices)
  constexpr int kMaxDev = 64;
  static int sms_of[kMaxDev] = {};
  static size_t cap_of[kMaxDev] = {};
  const size_t cp_bytes = (size_t)((E.nk * E.n_class * kClassParDoubles + 1) & ~1) * sizeof(double);
  const size_t per_warp = (size_t)D::LND * 32 * sizeof(double);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) { *fits = false; return cudaSuccess; }
  if (sms_of[dev] == 0) {
    int optin = 0, n_sm = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (err != cudaSuccess) return err;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cap_of[dev] = (size_t)optin;
    sms_of[dev] = n_sm;
  }
  const int sms = sms_of[dev];
  const size_t smem_cap = cap_of[dev];
  int nwarps = (int)((smem_cap - cp_bytes) / per_warp);
  if (nwarps > NT / 32) nwarps = NT / 32;
  const double max_off = 2.0 * (double)E.n_elem * D::LND * (double)(ldx > ldy ? ldx : ldy);   // n_dofs <= n_elem * LND
  *fits = nwarps >= 4 && max_off < 4.0e9;
  if (!*fits) return cudaSuccess;
  const long n_items = (long)E.n_elem * nvec;
  if (n_items == 0) return cudaSuccess;
  const long ntiles = (n_items + 15) / 16;
  long blocks = (ntiles + nwarps - 1) / nwarps;
  if (blocks > sms) blocks = sms;
  // few tiles: spread them over all SMs with fewer warps per block (each warp is independent)
  int nw = nwarps;
  if (ntiles < (long)sms * nwarps) {
    blocks = ntiles < sms ? ntiles : sms;
    nw = (int)((ntiles + blocks - 1) / blocks);
  }
  const size_t smem = cp_bytes + per_warp * nw;
  kern<<<(unsigned)blocks, nw * 32, smem, s>>>(
      item_tabs(T, P), with_cpk(E, nvec), reinterpret_cast<const double *>(x), reinterpret_cast<double *>(y), nvec, ldx, ldy, n_items, ca, cm);
  return cudaGetLastError();
}

template <int P, int NT, int GV>
cudaError_t nd_item_p(const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy, int nvec,
                      double ca, double cm, cudaStream_t s, bool *fits) {
  if (ca != 0.0 && cm != 0.0) return nd_item_t<P, true, true, NT, GV>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, fits);
  if (ca != 0.0) return nd_item_t<P, true, false, NT, GV>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, fits);
  return nd_item_t<P, false, true, NT, GV>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, fits);
}

}  // namespace

// Returns cudaSuccess with *launched = false when this variant does not apply (order > 2, class table too
// large for shared memory, or switched off with BLOCH_ND_ITEM=0); the caller then uses k_nd_apply.
cudaError_t launch_nd_item(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy,
                           int nvec, double ca, double cm, cudaStream_t s, bool *launched) {
  static int enabled = -1;
  if (enabled < 0) { const char *e = std::getenv("BLOCH_ND_ITEM"); enabled = e ? std::atoi(e) : 1; }
  *launched = false;
  if (!enabled || p > 2 || (ca == 0.0 && cm == 0.0)) return cudaSuccess;
  // BLOCH_ND_ITEM_MODE: 2 = all indices first (default), 1 = gather component by component behind compiler fences.
  // Order 2 runs 12 warps per SM at 168 registers: 16 warps (128 registers) spill 700 B per lane and measured 1.75x
  // slower, 14 warps at 144 registers do not launch (register file granularity), 8 warps at 255 measured equal.
  static int mode = -1;
  if (mode < 0) { const char *e = std::getenv("BLOCH_ND_ITEM_MODE"); mode = e ? std::atoi(e) : 2; }
  if (p == 1) return mode == 1 ? nd_item_p<1, 512, 0>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched)
                               : nd_item_p<1, 512, 1>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
  if (mode == 2) return nd_item_p<2, 384, 1>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
  return nd_item_p<2, 384, 0>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
}

}  // namespace bloch_b200
