// Device-side Rayleigh-Ritz of the block eigensolver: the small dense generalised Hermitian problems
//     GA c = lambda GM c      (GA = S^H A S, GM = S^H M S, S = [X W P] of one k-point, <= 63 columns)
// that the reference delegates to hypre's LOBPCG / LAPACK (maxwell_bloch.cpp:537-551; dsygv in
// meta_material_solver.cpp:3285-3299), solved by ONE CTA per k-point without leaving the GPU:
//   basis selection (soft locking: W_j / P_j of converged columns and numerically zero columns leave the basis),
//   diagonal scaling, Cholesky of GM in shared memory (pivot test like dense.hpp; on failure the P block, then half of
//   the W block are dropped and the factorisation repeated), reduction to standard form by two triangular solves,
//   parallel cyclic Jacobi (round-robin pairs, complex rotations, columns then rows) for the Hermitian eigenproblem,
//   rank sort, back substitution of the lowest m vectors, un-scaling.
// The host solver of dense.hpp (Householder + QL) stays as the test oracle of this kernel and as the fallback
// (BLOCH_RR_DEVICE=0).  All matrices live in shared memory (3 x 63 x 65 complex doubles = 197 KB).
#include "kernels.hpp"

#include <cstdio>

namespace bloch_b200 {

namespace {

using D2 = double2;
constexpr int RR_MAXN = 63;
constexpr int RR_LD = 65;      // odd pitch: a column walk cycles through all banks
constexpr int RR_THREADS = 512;

__device__ __forceinline__ D2 cmul(D2 a, D2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ D2 cmulc(D2 a, D2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a conj(b)
__device__ __forceinline__ D2 cconj(D2 a) { return make_double2(a.x, -a.y); }

struct RRShared {
  int keep[64];
  double sc[64], w[64];
  int rank[64];
  double rcs[32], rsn[32];
  D2 rph[32];
  int rp[32], rq[32];
  int kk, fail, dropped, done;
  double red[RR_THREADS / 32];
};

__global__ void __launch_bounds__(RR_THREADS)
k_rr_solve(const D2 *__restrict__ GA_all, const D2 *__restrict__ GM_all, int kc, int mb,
           const unsigned char *__restrict__ act, unsigned char *__restrict__ usep, D2 *__restrict__ C_all,
           double *__restrict__ lam_all, int *__restrict__ info, double chol_tol, int max_sweeps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  D2 *A = reinterpret_cast<D2 *>(smem_raw);                 // [RR_MAXN][RR_LD]: GA -> standard form -> diagonalised
  D2 *L = A + RR_MAXN * RR_LD;                              // Cholesky factor of the scaled GM (lower)
  D2 *V = L + RR_MAXN * RR_LD;                              // accumulated Jacobi rotations (eigenvectors in columns)
  RRShared &sh = *reinterpret_cast<RRShared *>(V + RR_MAXN * RR_LD);
  const int b = blockIdx.x, tid = threadIdx.x;
  const D2 *GA = GA_all + (size_t)b * kc * kc, *GM = GM_all + (size_t)b * kc * kc;
  D2 *C = C_all + (size_t)b * kc * mb;
  const unsigned char *ab = act + (size_t)b * mb;

  if (tid == 0) {
    // basis columns that take part: all of X; W_j / P_j only for unconverged j and only if not numerically zero
    int kk = 0;
    const bool up = usep[b] != 0;
    for (int i = 0; i < kc; i++) {
      const int j = i % mb, grp = i / mb;
      const double dii = GM[i * kc + i].x;
      if (grp == 0 || (ab[j] && dii > 1e-26 && !(grp == 2 && !up))) sh.keep[kk++] = i;
    }
    sh.kk = kk;
    sh.dropped = 0;
    sh.done = 0;
  }
  for (int t = tid; t < kc * mb; t += RR_THREADS) C[t] = make_double2(0.0, 0.0);
  __syncthreads();

  for (int attempt = 0; attempt < 4; attempt++) {
    const int kk = sh.kk;
    if (tid == 0) sh.fail = 0;
    __syncthreads();
    // ---- diagonal scaling, Hermitian parts of the two Gram matrices ----
    if (tid < kk) {
      const double dii = GM[sh.keep[tid] * kc + sh.keep[tid]].x;
      if (!(dii > 0.0)) sh.fail = 1;
      sh.sc[tid] = dii > 0.0 ? 1.0 / sqrt(dii) : 0.0;
    }
    __syncthreads();
    for (int t = tid; t < kk * kk; t += RR_THREADS) {
      const int i = t / kk, j = t - i * kk;
      const int gi = sh.keep[i], gj = sh.keep[j];
      const double s = sh.sc[i] * sh.sc[j];
      const D2 x = GA[gi * kc + gj], xt = GA[gj * kc + gi], y = GM[gi * kc + gj], yt = GM[gj * kc + gi];
      A[i * RR_LD + j] = make_double2(0.5 * s * (x.x + xt.x), 0.5 * s * (x.y - xt.y));
      L[i * RR_LD + j] = make_double2(0.5 * s * (y.x + yt.x), 0.5 * s * (y.y - yt.y));
    }
    __syncthreads();
    // ---- Cholesky L L^H of the scaled GM (max diagonal = 1): right-looking, lower triangle ----
    for (int j = 0; j < kk && !sh.fail; j++) {
      if (tid == 0) {
        const double d = L[j * RR_LD + j].x;
        if (!(d > chol_tol)) sh.fail = 1;
        else L[j * RR_LD + j] = make_double2(sqrt(d), 0.0);
      }
      __syncthreads();
      if (sh.fail) break;
      const double inv = 1.0 / L[j * RR_LD + j].x;
      for (int i = j + 1 + tid; i < kk; i += RR_THREADS) {
        D2 v = L[i * RR_LD + j];
        v.x *= inv; v.y *= inv;
        L[i * RR_LD + j] = v;
      }
      __syncthreads();
      const int rem = kk - j - 1;
      for (int t = tid; t < rem * rem; t += RR_THREADS) {
        const int i = j + 1 + t / rem, k = j + 1 + t % rem;
        if (k <= i) {
          const D2 u = cmulc(L[i * RR_LD + j], L[k * RR_LD + j]);
          D2 v = L[i * RR_LD + k];
          v.x -= u.x; v.y -= u.y;
          if (k == i) v.y = 0.0;
          L[i * RR_LD + k] = v;
        }
      }
      __syncthreads();
    }
    __syncthreads();
    if (sh.fail) {
      // drop the P block first, then halve what is left of the W block, then give up (solver.cu: rr_one)
      if (tid == 0) {
        int k2 = sh.kk;
        if (k2 > mb && sh.keep[k2 - 1] >= 2 * mb) {
          while (k2 > 0 && sh.keep[k2 - 1] >= 2 * mb) k2--;
        } else if (k2 > mb) {
          k2 -= (k2 - mb + 1) / 2;
        } else {
          sh.done = -1;
        }
        sh.kk = k2;
        sh.dropped = 1;
      }
      __syncthreads();
      if (sh.done < 0) break;
      continue;
    }
    // ---- A <- L^-1 A L^-H ----
    if (tid < kk) {                       // column tid of X = L^-1 A (forward substitution down the column)
      const int c = tid;
      for (int i = 0; i < kk; i++) {
        D2 s = A[i * RR_LD + c];
        for (int k = 0; k < i; k++) {
          const D2 u = cmul(L[i * RR_LD + k], A[k * RR_LD + c]);
          s.x -= u.x; s.y -= u.y;
        }
        const double inv = 1.0 / L[i * RR_LD + i].x;
        A[i * RR_LD + c] = make_double2(s.x * inv, s.y * inv);
      }
    }
    __syncthreads();
    if (tid < kk) {                       // row tid of Y = X L^-H
      const int r = tid;
      for (int i = 0; i < kk; i++) {
        D2 s = A[r * RR_LD + i];
        for (int k = 0; k < i; k++) {
          const D2 u = cmulc(A[r * RR_LD + k], L[i * RR_LD + k]);
          s.x -= u.x; s.y -= u.y;
        }
        const double inv = 1.0 / L[i * RR_LD + i].x;
        A[r * RR_LD + i] = make_double2(s.x * inv, s.y * inv);
      }
    }
    __syncthreads();
    for (int t = tid; t < kk * kk; t += RR_THREADS) {      // symmetrise, V = I
      const int i = t / kk, j = t - i * kk;
      if (i < j) {
        const D2 x = A[i * RR_LD + j], xt = A[j * RR_LD + i];
        const D2 h = make_double2(0.5 * (x.x + xt.x), 0.5 * (x.y - xt.y));
        A[i * RR_LD + j] = h;
        A[j * RR_LD + i] = cconj(h);
      } else if (i == j) {
        A[i * RR_LD + i].y = 0.0;
      }
      V[i * RR_LD + j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    // ---- parallel cyclic Jacobi: round-robin schedule of disjoint pairs, all rotations of a round at once ----
    const int n2 = kk + (kk & 1), nr = n2 - 1, npairs = n2 / 2;
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
      // convergence: off-diagonal mass against the diagonal
      double off = 0.0, dg = 0.0;
      for (int t = tid; t < kk * kk; t += RR_THREADS) {
        const int i = t / kk, j = t - i * kk;
        const D2 v = A[i * RR_LD + j];
        if (i < j) off += v.x * v.x + v.y * v.y;
        else if (i == j) dg += v.x * v.x;
      }
      for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
      __syncthreads();
      if ((tid & 31) == 0) sh.red[tid >> 5] = off;
      __syncthreads();
      double offt = 0.0;
      for (int wv = 0; wv < RR_THREADS / 32; wv++) offt += sh.red[wv];
      __syncthreads();
      if ((tid & 31) == 0) sh.red[tid >> 5] = dg;
      __syncthreads();
      double dgt = 0.0;
      for (int wv = 0; wv < RR_THREADS / 32; wv++) dgt += sh.red[wv];
      __syncthreads();
      if (offt <= 1e-30 * dgt || offt == 0.0) break;
      for (int r = 0; r < nr; r++) {
        if (tid < npairs) {
          int p, q;
          if (tid == 0) { p = n2 - 1; q = r; }
          else { p = (r + tid) % nr; q = (r - tid + nr) % nr; }
          if (p > q) { const int tmp = p; p = q; q = tmp; }
          double cs = 1.0, sn = 0.0;
          D2 ph = make_double2(1.0, 0.0);
          if (q < kk) {
            const D2 c = A[p * RR_LD + q];
            const double g = sqrt(c.x * c.x + c.y * c.y);
            const double a = A[p * RR_LD + p].x, bb = A[q * RR_LD + q].x;
            if (g > 1e-300 && g > 1e-18 * (fabs(a) + fabs(bb))) {
              const double tau = (bb - a) / (2.0 * g);
              const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
              cs = 1.0 / sqrt(1.0 + tt * tt);
              sn = tt * cs;
              ph = make_double2(c.x / g, c.y / g);
            }
          }
          sh.rp[tid] = p; sh.rq[tid] = q; sh.rcs[tid] = cs; sh.rsn[tid] = sn; sh.rph[tid] = ph;
        }
        __syncthreads();
        // columns of A and V:  col p' = cs col p - sn e^{-i phi} col q ;  col q' = sn col p + cs e^{-i phi} col q
        for (int t = tid; t < npairs * kk; t += RR_THREADS) {
          const int pr = t / kk, k = t - pr * kk;
          const double sn = sh.rsn[pr];
          if (sn == 0.0) continue;
          const int p = sh.rp[pr], q = sh.rq[pr];
          const double cs = sh.rcs[pr];
          const D2 em = cconj(sh.rph[pr]);
          {
            const D2 x = A[k * RR_LD + p], y = cmul(em, A[k * RR_LD + q]);
            A[k * RR_LD + p] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
            A[k * RR_LD + q] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
          }
          {
            const D2 x = V[k * RR_LD + p], y = cmul(em, V[k * RR_LD + q]);
            V[k * RR_LD + p] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
            V[k * RR_LD + q] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
          }
        }
        __syncthreads();
        // rows of A:  row p' = cs row p - sn e^{i phi} row q ;  row q' = sn row p + cs e^{i phi} row q
        for (int t = tid; t < npairs * kk; t += RR_THREADS) {
          const int pr = t / kk, k = t - pr * kk;
          const double sn = sh.rsn[pr];
          if (sn == 0.0) continue;
          const int p = sh.rp[pr], q = sh.rq[pr];
          const double cs = sh.rcs[pr];
          const D2 x = A[p * RR_LD + k], y = cmul(sh.rph[pr], A[q * RR_LD + k]);
          A[p * RR_LD + k] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
          A[q * RR_LD + k] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
        }
        __syncthreads();
      }
    }
    // ---- lowest mb eigenpairs: rank sort, back substitution c = L^-H v, un-scaling ----
    if (tid < kk) sh.w[tid] = A[tid * RR_LD + tid].x;
    __syncthreads();
    if (tid < kk) {
      int rk = 0;
      const double wi = sh.w[tid];
      for (int j = 0; j < kk; j++) {
        const double wj = sh.w[j];
        if (wj < wi || (wj == wi && j < tid)) rk++;
      }
      sh.rank[tid] = rk;
    }
    __syncthreads();
    if (tid < kk && sh.rank[tid] < mb) {
      const int col = tid, out = sh.rank[tid];
      // in place in column `col` of V (each thread owns its column)
      for (int i = kk - 1; i >= 0; i--) {
        D2 s = V[i * RR_LD + col];
        for (int k = i + 1; k < kk; k++) {
          const D2 u = cmul(cconj(L[k * RR_LD + i]), V[k * RR_LD + col]);
          s.x -= u.x; s.y -= u.y;
        }
        const double inv = 1.0 / L[i * RR_LD + i].x;
        V[i * RR_LD + col] = make_double2(s.x * inv, s.y * inv);
      }
      for (int i = 0; i < kk; i++) {
        const D2 v = V[i * RR_LD + col];
        C[sh.keep[i] * mb + out] = make_double2(v.x * sh.sc[i], v.y * sh.sc[i]);
      }
      lam_all[(size_t)b * mb + out] = sh.w[col];
    }
    if (tid == 0) sh.done = 1;
    __syncthreads();
    break;
  }
  if (tid == 0) {
    info[b] = sh.done == 1 ? sh.dropped : -1;
    usep[b] = sh.dropped ? 0 : 1;      // after a degenerate Rayleigh-Ritz the search directions are discarded once
  }
}

}  // namespace

size_t rr_solve_smem_bytes() { return (size_t)3 * RR_MAXN * RR_LD * sizeof(D2) + sizeof(RRShared); }

cudaError_t launch_rr_solve(const double2 *GA, const double2 *GM, int kc, int mb, int K, const unsigned char *act,
                            unsigned char *usep, double2 *C, double *lam, int *info, cudaStream_t s) {
  if (kc > RR_MAXN || mb > kc || mb > 32) return cudaErrorInvalidValue;
  static bool attr_of[kMaxDevices] = {};
  bool &attr = attr_of[current_device_slot()];
  const size_t smem = rr_solve_smem_bytes();
  if (!attr) {
    cudaError_t err = cudaFuncSetAttribute(k_rr_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    attr = true;
  }
  k_rr_solve<<<K, RR_THREADS, smem, s>>>(GA, GM, kc, mb, act, usep, C, lam, info, 1e-9, 30);
  return cudaGetLastError();
}

}  // namespace bloch_b200
