// y += ca * A x + cm * M x on ND block vectors, orders 1..3: barrier-free "six lanes per item" kernel.
//
// Same mode-space element operator as k_nd_apply (kernels.cu) and k_nd_item (nd_item.cu).  Mapping:
//   * one (element, vector) item is handled by SIX lanes = 3 vector components x (real, imaginary part); a warp
//     works on 5 items (lanes 30, 31 shadow lanes 24, 25 and never write).  The kernel's local storage is cyclic
//     (ND component c = [open c][closed c+1][closed c+2], RT component c = [closed c][open c+1][open c+2]), so
//     the six lanes run ONE instruction stream with compile-time tensor indices; what differs per lane is only
//     which lanes hold components c+1, c+2 (two lane offsets) and five scalars (kappa_hat, one row of G, H);
//   * a lane keeps its ND component (p (p+1)^2 reals) in a lane-private shared-memory column and its RT component
//     (p^2 (p+1) reals) in registers.  The curl reads the other components from the neighbours' columns, the
//     pointwise RT mass (3 x 3 coupling G at a grid point) fetches them with shuffles (cyclic orbits of the grid
//     point, so that old values are still in place), the adjoint curl reads the neighbours' RT columns and the ND
//     mass part is folded into the adjoint stage (no in-place update, hence no ordering hazards);
//   * no block barrier, three __syncwarp per tile; registers ~ 1/3 of the lane-pair kernel, so order 3 fits
//     (RT component = 36 reals) and orders 1, 2 run with many more warps per SM.
// Tables are the sqrt-weight scaled ones of nd_item.cu (all mode-space mass weights equal 1).
#include "kernels.hpp"

#include <cmath>
#include <cstdlib>

#include "elem_device.cuh"

namespace bloch_b200 {

namespace {

using namespace dev;

struct CompTabs {
  double TI[kMaxP][kMaxP];             // orders <= 3: (p+1) x (p+1)
  double TIo[kMaxP - 1][kMaxP][kMaxP];
  double Dt[kMaxP - 1][kMaxP];
};

CompTabs comp_tabs(const Tabs &T, int p) {
  CompTabs S = {};
  double sq[kMaxP];
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
__device__ __forceinline__ void slab_tf(double (&s)[P + 1][P + 1], const double (&M1)[kMaxP][kMaxP],
                                        const double (&M2)[kMaxP][kMaxP]) {
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

__device__ __forceinline__ double flip(double x, int s) {
  return __hiloint2double(__double2hiint(x) ^ (s & (int)0x80000000), __double2loint(x));
}

// 16-byte shared-memory load of the (re, im) pair of one entry from the even lane's column; own = this lane's part
__device__ __forceinline__ void ld_pair(const double *even_lane_entry, int part, double &own, double &other) {
  const double2 t = *reinterpret_cast<const double2 *>(even_lane_entry);
  own = part ? t.y : t.x;
  other = part ? t.x : t.y;
}

// predicated fp64 reduction (no branch around the RED)
__device__ __forceinline__ void red_add_if(double *p, double v, bool on) {
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q red.global.add.f64 [%0], %1; }" ::"l"(p), "d"(v), "r"((int)on)
               : "memory");
}


// IPW = items per warp: 5 (lanes 0..29; the item on lanes 12..17 straddles the two half-warps that serve a 64-bit
// shared-memory access, which costs bank conflicts on the neighbour-column reads) or 4 (two items per half-warp,
// lanes 12..15 and 28..31 idle: conflict-free but 20 % fewer items per instruction)
template <int P, bool HAS_A, bool HAS_M, int NT, int IPW>
__global__ void __launch_bounds__(NT, 1)
k_nd_comp(const __grid_constant__ CompTabs T, const ElemData E, const double *__restrict__ X, double *__restrict__ Y,
          int m, int ldx, int ldy, long n_items, double ca, double cm) {
  using D = Dim<P>;
  constexpr int Q = P + 1, NB = D::NB, RB = D::RB;
  // A only: the ND columns are dead once every lane has finished the curl, so R' is published over them (smaller
  // columns -> 12 instead of 10 warps per SM at order 3; 16 warps at 128 registers also fit but measured 4 % slower)
  constexpr bool OVERLAY = HAS_A && !HAS_M;
  constexpr int WCOLS = OVERLAY ? (NB > RB ? NB : RB) : NB + (HAS_A ? RB : 0);        // column entries per lane
  auto nd0 = [](int o, int j1, int j2) { return (o * Q + j1) * Q + j2; };    // local index in the own ND component
  auto rt0 = [](int j, int o1, int o2) { return (j * P + o1) * P + o2; };    // ... in the own RT component
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sCP = reinterpret_cast<double *>(smem_raw);
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int ncp = cpar_doubles(E);
  for (int i = threadIdx.x; i < ncp; i += blockDim.x) sCP[i] = E.cpar[i];
  __syncthreads();
  // lane roles: lane = base(slot) + 2 c + part; spare lanes shadow the lane 6 below them and never write
  constexpr int kItemsPerWarp = IPW;
  const int hb = IPW == 4 ? (lane & 16) : 0;                 // IPW 4: each half-warp holds two items on its lanes 0..11
  const int hl = IPW == 4 ? (lane & 15) : lane;
  const bool spare = IPW == 4 ? hl >= 12 : lane >= 30;
  const int rl = spare ? hl - 6 : hl;
  const int s6 = rl / 6, c = (rl - 6 * s6) >> 1, part = rl & 1;
  const int slot = IPW == 4 ? (hb >> 3) + s6 : s6;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const int lane1 = hb + 6 * s6 + 2 * c1 + part, lane2 = hb + 6 * s6 + 2 * c2 + part;
  double *wbase = sCP + ((ncp + 1) & ~1) + (size_t)warp * (WCOLS * 32);
  double *Fc = wbase + lane;                                          // own ND component
  const double *F1 = wbase + lane1, *F2 = wbase + lane2;              // components c+1, c+2 of the same part
  const double *F1e = wbase + (lane1 - part), *F2e = wbase + (lane2 - part);   // even lane of each pair (16-byte aligned)
  const int rofs = OVERLAY ? 0 : NB * 32;                             // RT columns behind (or over) the ND columns
  const double sg = part ? -1.0 : 1.0;
  const long ntiles = (n_items + kItemsPerWarp - 1) / kItemsPerWarp;
  const bool small = n_items < 0x7fffffffL;
  const unsigned xstep = 2u * (unsigned)ldx, ystep = 2u * (unsigned)ldy;
  const long tstride = (long)gridDim.x * nwarps;

  for (long tile = (long)warp * gridDim.x + blockIdx.x; tile < ntiles; tile += tstride) {
    long item = tile * kItemsPerWarp + slot;
    const bool active = item < n_items;
    if (!active) item = n_items - 1;            // idle slots of the last tile shadow a valid item with zero weights
    const bool writes = active && !spare;
    const int e = small ? (int)((unsigned)item / (unsigned)m) : (int)(item / m);
    const int v = (int)(item - (long)e * m);
    const int32_t *mp = E.map_nd + (long)e * D::LND + c * NB;
    const double *cp = sCP + kClassParDoubles * vclass(E, __ldg(E.cls + e), v);
    // 32-bit offsets in doubles (the launcher checks 2 * n_dofs * ld < 2^32):  (|s| - 1) * 2 ld + 2 v + part
    const unsigned xoff = 2u * (unsigned)v + (unsigned)part - xstep, yoff = 2u * (unsigned)v + (unsigned)part - ystep;

    // ---- signed gather of the own component + nodal -> mode in its two closed directions ----
    // All indices of the component first (8-byte loads), then slab by slab with the NEXT slab's values in flight
    // while the current one is transformed; the compiler fences keep it from hoisting every load to the top
    // (register pressure).
    {
      static_assert(NB % 2 == 0, "8-byte index loads");
      int sidx[NB];
#pragma unroll
      for (int k = 0; k < NB / 2; k++) {
        const int2 t = __ldg(reinterpret_cast<const int2 *>(mp) + k);
        sidx[2 * k] = t.x; sidx[2 * k + 1] = t.y;
      }
      double xa[Q * Q], xb[Q * Q];
#pragma unroll
      for (int k = 0; k < Q * Q; k++) {
        const int s = sidx[k];
        xa[k] = flip(__ldg(X + ((unsigned)(s < 0 ? -s : s) * xstep + xoff)), s);
      }
#pragma unroll
      for (int o = 0; o < P; o++) {
        if (o + 1 < P) {
#pragma unroll
          for (int k = 0; k < Q * Q; k++) {
            const int s = sidx[(o + 1) * Q * Q + k];
            xb[k] = flip(__ldg(X + ((unsigned)(s < 0 ? -s : s) * xstep + xoff)), s);
          }
        }
        double s[Q][Q];
#pragma unroll
        for (int a = 0; a < Q; a++)
#pragma unroll
          for (int b = 0; b < Q; b++) s[a][b] = xa[a * Q + b];
        slab_tf<P, false>(s, T.TIo[o], T.TI);
#pragma unroll
        for (int a = 0; a < Q; a++)
#pragma unroll
          for (int b = 0; b < Q; b++) Fc[nd0(o, a, b) * 32] = s[a][b];
        if (P == 3) asm volatile("" ::: "memory");
#pragma unroll
        for (int k = 0; k < Q * Q; k++) xa[k] = xb[k];
      }
    }
    __syncwarp();

    double ks1 = 0.0, ks2 = 0.0;
    if (HAS_A) {
      ks1 = sg * cp[c1];
      ks2 = sg * cp[c2];
      double R[RB];
      // ---- Bloch curl: R_c = K_{c+1} F_{c+2} - K_{c+2} F_{c+1},  K_d = Dt - i kh_d (pointwise) ----
      // (re, im) of an entry sit in adjacent lanes' columns, i.e. 16 bytes apart from the even lane: wherever both the
      // own and the other part of a neighbour's entry are needed (the -i kappa_hat terms), ONE 16-byte load fetches
      // the pair.  The two halves of the curl are accumulated one neighbour row at a time (a row of F_{c+2} for all
      // o1, then a row of F_{c+1} for all o2): 24 instead of 42 shared-memory loads per j and a quarter of the live
      // registers.
#pragma unroll
      for (int j = 0; j < Q; j++) {
#pragma unroll
        for (int o2 = 0; o2 < P; o2++) {
          double a[Q], pa[P];
#pragma unroll
          for (int r = 0; r < P; r++) ld_pair(F2e + nd0(o2, j, r) * 32, part, a[r], pa[r]);   // F_{c+2}[o2, j1=j, j2=r]
          a[P] = F2[nd0(o2, j, P) * 32];
#pragma unroll
          for (int o1 = 0; o1 < P; o1++) {
            double acc = ks1 * pa[o1];
#pragma unroll
            for (int r = 0; r < Q; r++) acc = fma(T.Dt[o1][r], a[r], acc);
            R[rt0(j, o1, o2)] = acc;
          }
        }
#pragma unroll
        for (int o1 = 0; o1 < P; o1++) {
          double bq[Q], pb[P];
#pragma unroll
          for (int r = 0; r < P; r++) ld_pair(F1e + nd0(o1, r, j) * 32, part, bq[r], pb[r]);  // F_{c+1}[o1, j1=r, j2=j]
          bq[P] = F1[nd0(o1, P, j) * 32];
#pragma unroll
          for (int o2 = 0; o2 < P; o2++) {
            double acc = fma(-ks2, pb[o2], R[rt0(j, o1, o2)]);
#pragma unroll
            for (int r = 0; r < Q; r++) acc = fma(-T.Dt[o2][r], bq[r], acc);
            R[rt0(j, o1, o2)] = acc;
          }
        }
      }
      // ---- pointwise RT mass, scaled by ca * muinv: R'_c = G[c][c] R_c + G[c][c+1] R_{c+1} + G[c][c+2] R_{c+2} ----
      // The same grid point is (x,y,z) in the own frame, (y,z,x) in the frame of component c+1 and (z,x,y) in
      // that of c+2, so the three rotations of a point are fetched (shuffles) before any of them is updated.
      const double wA = active ? ca * __ldg(E.muinv + e) : 0.0;   // idle slots produce exact zeros
      const double g0 = wA * cp[3 + 3 * c + c], g1 = wA * cp[3 + 3 * c + c1], g2 = wA * cp[3 + 3 * c + c2];
#pragma unroll
      for (int o1 = 0; o1 < P; o1++)
#pragma unroll
        for (int o2 = 0; o2 < P; o2++) R[rt0(P, o1, o2)] *= g0;      // closed end mode: only this component lives there
#pragma unroll
      for (int x = 0; x < P; x++)
#pragma unroll
        for (int y = 0; y < P; y++)
#pragma unroll
          for (int z = 0; z < P; z++) {
            const int k0 = rt0(x, y, z), k1 = rt0(y, z, x), k2 = rt0(z, x, y);
            if (x == y && y == z) {
              const double a0 = R[k0];
              const double n1 = __shfl_sync(0xffffffffu, a0, lane1), n2 = __shfl_sync(0xffffffffu, a0, lane2);
              R[k0] = fma(g0, a0, fma(g1, n1, g2 * n2));
            } else if (k0 < k1 && k0 < k2) {
              const double a0 = R[k0], a1 = R[k1], a2 = R[k2];
              const double n1_0 = __shfl_sync(0xffffffffu, a0, lane1), n2_0 = __shfl_sync(0xffffffffu, a0, lane2);
              const double n1_1 = __shfl_sync(0xffffffffu, a1, lane1), n2_1 = __shfl_sync(0xffffffffu, a1, lane2);
              const double n1_2 = __shfl_sync(0xffffffffu, a2, lane1), n2_2 = __shfl_sync(0xffffffffu, a2, lane2);
              R[k0] = fma(g0, a0, fma(g1, n1_1, g2 * n2_2));
              R[k1] = fma(g0, a1, fma(g1, n1_2, g2 * n2_0));
              R[k2] = fma(g0, a2, fma(g1, n1_0, g2 * n2_1));
            }
          }
      // publish R' for the adjoint curl of the other components
      if (OVERLAY) __syncwarp();          // every lane is done reading the ND columns
#pragma unroll
      for (int k = 0; k < RB; k++) Fc[rofs + k * 32] = R[k];
    }
    __syncwarp();

    // ---- ND mass part + adjoint curl, mode -> nodal (adjoint), signed scatter-add; one slab at a time ----
    double h0 = 0.0, h1 = 0.0, h2 = 0.0;
    if (HAS_M) {
      const double wM = active ? cm * __ldg(E.eps + e) : 0.0;
      h0 = wM * cp[12 + 3 * c + c]; h1 = wM * cp[12 + 3 * c + c1]; h2 = wM * cp[12 + 3 * c + c2];
    }
#pragma unroll(P < 3 ? P : 1)
    for (int o = 0; o < P; o++) {
      int sc[Q * Q];                      // this slab's scatter indices, requested before the arithmetic needs them
#pragma unroll
      for (int k = 0; k < Q * Q; k++) sc[k] = __ldg(mp + o * Q * Q + k);
      double f[Q][Q];
#pragma unroll
      for (int a = 0; a < Q; a++)
#pragma unroll
        for (int b = 0; b < Q; b++) {
          double acc = 0.0;
          if (HAS_M) {
            // F'_c(o,a,b) = H[c][c] F_c(o,a,b) + H[c][c+1] F_{c+1}(a,b,o) + H[c][c+2] F_{c+2}(b,o,a)
            acc = h0 * Fc[nd0(o, a, b) * 32];
            if (a < P) acc = fma(h1, F1[nd0(a < P ? a : 0, b, o) * 32], acc);
            if (b < P) acc = fma(h2, F2[nd0(b < P ? b : 0, o, a) * 32], acc);
          }
          f[a][b] = acc;
        }
      if (HAS_A) {
        const double *R1e = F1e + rofs, *R2e = F2e + rofs;
#pragma unroll
        for (int j1 = 0; j1 < Q; j1++) {
          double y[P], py[P];
#pragma unroll
          for (int q = 0; q < P; q++) ld_pair(R1e + rt0(j1, q, o) * 32, part, y[q], py[q]);   // R'_{c+1}[j=j1, o1=q, o2=o]
#pragma unroll
          for (int j2 = 0; j2 < Q; j2++) {
            double acc = f[j1][j2];
#pragma unroll
            for (int q = 0; q < P; q++) acc = fma(T.Dt[q][j2], y[q], acc);
            if (j2 < P) acc = fma(-ks2, py[j2 < P ? j2 : 0], acc);
            f[j1][j2] = acc;
          }
        }
#pragma unroll
        for (int j2 = 0; j2 < Q; j2++) {
          double y[P], py[P];
#pragma unroll
          for (int q = 0; q < P; q++) ld_pair(R2e + rt0(j2, o, q) * 32, part, y[q], py[q]);   // R'_{c+2}[j=j2, o1=o, o2=q]
#pragma unroll
          for (int j1 = 0; j1 < Q; j1++) {
            double acc = f[j1][j2];
#pragma unroll
            for (int q = 0; q < P; q++) acc = fma(-T.Dt[q][j1], y[q], acc);
            if (j1 < P) acc = fma(ks1, py[j1 < P ? j1 : 0], acc);
            f[j1][j2] = acc;
          }
        }
      }
      slab_tf<P, true>(f, T.TIo[o], T.TI);
#pragma unroll
      for (int a = 0; a < Q; a++)
#pragma unroll
        for (int b = 0; b < Q; b++) {
          const int s = sc[a * Q + b];
          const unsigned off = (unsigned)(s < 0 ? -s : s) * ystep + yoff;
          const bool interior = a > 0 && a < P && b > 0 && b < P;        // compile time: dof of no other element
          if (interior && E.fresh_y) { if (writes) Y[off] = flip(f[a][b], s); }
          else red_add_if(Y + off, flip(f[a][b], s), writes);
        }
    }
    __syncwarp();   // all lanes are done with this tile's columns before the next tile overwrites them
  }
}

template <int P, bool HAS_A, bool HAS_M, int NT, int IPW>
cudaError_t nd_comp_t(const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy, int nvec,
                      double ca, double cm, cudaStream_t s, bool *fits) {
  using D = Dim<P>;
  // per-device one-time setup (a process may hold handles on several devices)
  constexpr int kMaxDev = 64;
  static int sms_of[kMaxDev] = {};
  static size_t cap_of[kMaxDev] = {};
  const size_t cp_bytes = (size_t)((E.nk * E.n_class * kClassParDoubles + 1) & ~1) * sizeof(double);
  const size_t per_warp = (size_t)((HAS_A && !HAS_M) ? (D::NB > D::RB ? D::NB : D::RB) : D::NB + (HAS_A ? D::RB : 0)) * 32 * sizeof(double);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) { *fits = false; return cudaSuccess; }
  if (sms_of[dev] == 0) {
    int optin = 0, n_sm = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaError_t err = cudaFuncSetAttribute(k_nd_comp<P, HAS_A, HAS_M, NT, IPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
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
  const long ntiles = (n_items + IPW - 1) / IPW;
  long blocks = (ntiles + nwarps - 1) / nwarps;
  if (blocks > sms) blocks = sms;
  int nw = nwarps;
  if (ntiles < (long)sms * nwarps) {      // few tiles: spread them over all SMs with fewer warps per block
    blocks = ntiles < sms ? ntiles : sms;
    nw = (int)((ntiles + blocks - 1) / blocks);
  }
  const size_t smem = cp_bytes + per_warp * nw;
  k_nd_comp<P, HAS_A, HAS_M, NT, IPW><<<(unsigned)blocks, nw * 32, smem, s>>>(
      comp_tabs(T, P), with_cpk(E, nvec), reinterpret_cast<const double *>(x), reinterpret_cast<double *>(y), nvec, ldx, ldy, n_items,
      ca, cm);
  return cudaGetLastError();
}

template <int P, int NT, int IPW>
cudaError_t nd_comp_p(const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy, int nvec,
                      double ca, double cm, cudaStream_t s, bool *fits) {
  if (ca != 0.0 && cm != 0.0) return nd_comp_t<P, true, true, NT, IPW>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, fits);
  // pure A apply at order 3: 12 warps of 168 registers (16 warps of 128 registers measured 48.6 vs 50.4 GDOF/s)
  if (ca != 0.0) return nd_comp_t<P, true, false, (P == 3 ? 384 : NT), IPW>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, fits);
  return nd_comp_t<P, false, true, NT, IPW>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, fits);
}

}  // namespace

// Returns cudaSuccess with *launched = false when this variant does not apply or is not selected
// (BLOCH_ND_COMP: bit mask of the orders that use it, default = order 3 only).
cudaError_t launch_nd_comp(int p, const Tabs &T, const ElemData &E, const double2 *x, int ldx, double2 *y, int ldy,
                           int nvec, double ca, double cm, cudaStream_t s, bool *launched) {
  static int mask = -1;
  if (mask < 0) { const char *e = std::getenv("BLOCH_ND_COMP"); mask = e ? std::atoi(e) : 4; }
  *launched = false;
  if (p < 1 || p > 3 || !(mask & (1 << (p - 1))) || (ca == 0.0 && cm == 0.0)) return cudaSuccess;
  static int ipw = -1;
  if (ipw < 0) { const char *e = std::getenv("BLOCH_ND_COMP_ITEMS"); ipw = e ? std::atoi(e) : 5; }
  if (p == 1) return nd_comp_p<1, 1024, 5>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
  if (p == 2) return nd_comp_p<2, 768, 5>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
  if (ipw == 4) return nd_comp_p<3, 320, 4>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
  return nd_comp_p<3, 320, 5>(T, E, x, ldx, y, ldy, nvec, ca, cm, s, launched);
}

}  // namespace bloch_b200
