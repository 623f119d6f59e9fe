// Native complex block eigensolver behind Solve() (maxwell/maxwell_bloch.cpp:809-825): what the
// reference delegates to hypre's LOBPCG with preconditioner = Projector o blockdiag(AMS) and a
// MINRES-based divergence projector (maxwell_bloch.cpp:531-551, 1923-1937, 2280-2290).
//
//   * block LOBPCG on m complex vectors (the reference's 2m real ones), basis S = [X W P] kept in
//     ONE row-major array [N][3m] so both Gram matrices are a single tall-skinny contraction each;
//   * preconditioner T = auxiliary-space cycle for A + sigma M (Chebyshev-Jacobi smoother + nodal (H1)^3
//     correction Pi B Pi^H by multigrid V-cycles, aux.cu; precondition_aux below), or a degree-24 Chebyshev
//     polynomial in D^-1 (A + sigma M) (BLOCH_PRECOND=cheb, odd n_sub); the scalar H1 problem uses one V-cycle
//     of its own hierarchy.  Eigenvalues do not depend on T, only the iteration count does;
//   * constraint G^H M x = 0 imposed by projecting W (and the initial block) with a block PCG on
//     S0 = G^H M G preconditioned by a geometric-multigrid V-cycle (mg.cu; Jacobi-PCG fallback, proj_cg.cu),
//     relaxed by the gradient lift once a k-point is warm.
// On the affine WS meshes (C - iZ)(G - iZ0) = 0 holds exactly, so the projected residual equals
// the plain residual and the convergence test is || A x - lambda M x ||_2 <= atol like hypre's.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "core.hpp"
#include "dense.hpp"

using namespace bloch_b200;
using D2 = double2;

namespace {

constexpr int TPB = 256;
inline unsigned grid_for(long total) {
  long g = (total + TPB - 1) / TPB;
  const long cap = 148L * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// R = AS[:, :m] - lam .* MS[:, :m]  (AS, MS with pitch ld);  rn2[j] += |R[:, j]|^2
__global__ void k_resid_norm(const D2 *__restrict__ AS, const D2 *__restrict__ MS, int ld,
                             const double *__restrict__ lam, D2 *__restrict__ R, long n, int m,
                             double *__restrict__ rn2) {
  extern __shared__ double sred[];
  for (int j = threadIdx.x; j < m; j += blockDim.x) sred[j] = 0.0;
  __syncthreads();
  const long nthreads = (long)gridDim.x * blockDim.x;
  const long usable = (nthreads / m) * m;
  const long start = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (start < usable) {
    const int j = (int)(start % m);
    const double l = lam[j];
    double acc = 0.0;
    for (long t = start; t < n * m; t += usable) {
      const long r = t / m;
      const D2 a = AS[r * ld + j], b = MS[r * ld + j];
      const D2 v = make_double2(a.x - l * b.x, a.y - l * b.y);
      R[t] = v;
      acc = fma(v.x, v.x, acc);
      acc = fma(v.y, v.y, acc);
    }
    atomicAdd(&sred[j], acc);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < m; j += blockDim.x) atomicAdd(rn2 + j, sred[j]);
}

// Basis layout with nk k-points batched: S = [X | W | P] is ONE array [N][3 * gs], gs = nk * m; each of the three
// groups holds the m columns of every k-point (column k * m + j), so the X (or W) block of ALL k-points is a
// contiguous column range - the operator / vector kernels see one block vector of nk * m columns - while the basis
// of k-point b is the strided set col(b, i) = (i / m) * gs + b * m + i % m, i < 3 m.
__device__ __forceinline__ int basis_col(int i, int m, int gs, int b) { return (i / m) * gs + b * m + (i % m); }

// Rayleigh-Ritz rotation, in place on the basis arrays (pitch ld = 3 gs), k-point b = blockIdx.y:
//   P_new = sum_{i>=m} S_i C[i][:],  X_new = sum_{i<m} S_i C[i][:] + P_new        (C = C[b], k x m)
// A CTA stages RP rows of one array in shared memory (coalesced), then thread (row, column) forms its two
// sums; CW = 16 or 32 column slots per row (m <= 21), k = number of active basis columns (m, 2m or 3m).
template <int CW>
__global__ void __launch_bounds__(256)
k_rr_update(D2 *__restrict__ S, D2 *__restrict__ AS, D2 *__restrict__ MS, int ld, int k, int m, int gs,
            const D2 *__restrict__ Call, long n) {
  // thread (rr, j) owns column j of RB = 4 consecutive rows: every coefficient C[i][j] fetched from shared memory is
  // used for four rows (5 instead of 8 shared loads per 4 complex MACs) and the four accumulation chains are independent
  constexpr int RB = 4, RG = 256 / CW, RP = RG * RB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  D2 *sC = reinterpret_cast<D2 *>(smem_raw);     // [k][m]
  D2 *sT = sC + k * m;                           // [RP][k]
  const int b = blockIdx.y;
  const D2 *C = Call + (size_t)b * k * m;
  for (int t = threadIdx.x; t < k * m; t += 256) sC[t] = C[t];
  const int rr = threadIdx.x / CW, j = threadIdx.x % CW;
  const long ntiles = (n + RP - 1) / RP;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long r0 = tile * RP;
    const int nr = (int)min((long)RP, n - r0);
#pragma unroll
    for (int a = 0; a < 3; a++) {
      D2 *base = (a == 0 ? S : (a == 1 ? AS : MS)) + r0 * ld;
      __syncthreads();
      for (int t = threadIdx.x; t < nr * k; t += 256) sT[t] = base[(long)(t / k) * ld + basis_col(t % k, m, gs, b)];
      __syncthreads();
      if (j < m) {
        D2 x[RB], pn[RB];
#pragma unroll
        for (int u = 0; u < RB; u++) { x[u] = make_double2(0.0, 0.0); pn[u] = make_double2(0.0, 0.0); }
        const D2 *row = sT + (rr * RB) * k;
        for (int i = 0; i < m; i++) {
          const D2 c = sC[i * m + j];
#pragma unroll
          for (int u = 0; u < RB; u++) {
            const D2 s = row[u * k + i];
            x[u].x = fma(s.x, c.x, x[u].x); x[u].x = fma(-s.y, c.y, x[u].x);
            x[u].y = fma(s.x, c.y, x[u].y); x[u].y = fma(s.y, c.x, x[u].y);
          }
        }
        for (int i = m; i < k; i++) {
          const D2 c = sC[i * m + j];
#pragma unroll
          for (int u = 0; u < RB; u++) {
            const D2 s = row[u * k + i];
            pn[u].x = fma(s.x, c.x, pn[u].x); pn[u].x = fma(-s.y, c.y, pn[u].x);
            pn[u].y = fma(s.x, c.y, pn[u].y); pn[u].y = fma(s.y, c.x, pn[u].y);
          }
        }
#pragma unroll
        for (int u = 0; u < RB; u++) {
          const int r = rr * RB + u;
          if (r < nr) {
            D2 *out = base + (long)r * ld + b * m;
            out[j] = make_double2(x[u].x + pn[u].x, x[u].y + pn[u].y);
            out[2 * gs + j] = pn[u];
          }
        }
      }
    }
  }
}

// the rotation kernel stages 64 (CW = 16) / 32 (CW = 32) rows: up to ~62 KB of dynamic shared memory, opt-in per device
static cudaError_t rr_update_attr() {
  static bool attr_of[kMaxDevices] = {};
  bool &attr = attr_of[current_device_slot()];
  if (attr) return cudaSuccess;
  cudaError_t err = cudaFuncSetAttribute(k_rr_update<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (err != cudaSuccess) return err;
  err = cudaFuncSetAttribute(k_rr_update<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (err != cudaSuccess) return err;
  attr = true;
  return cudaSuccess;
}

// Both Gram matrices of every k-point's basis in one pass over the rows:
//   GA[b][i][j] = sum_r conj(S[r][col(b,i)]) AS[r][col(b,j)],   GM[b] likewise with MS,   i, j < k <= 16 TS.
// 16 x 16 threads, thread (ti, tj) owns the TS x TS outputs (ti + 16 a, tj + 16 c); both matrices are Hermitian up
// to rounding, so only the block pairs a <= c are accumulated and mirrored in the epilogue (2/3 of the flops at
// TS = 3).  Rows are staged 16 at a time with cp.async into a double buffer (the global round trip of the next
// stage overlaps the current one); the column map of a thread's staging slots is computed once.  blockIdx.y = b.
constexpr int GB_ROWS = 16;
template <int TS>
__global__ void __launch_bounds__(256)
k_gram2_basis(const D2 *__restrict__ S, const D2 *__restrict__ AS, const D2 *__restrict__ MS, int ld, int k, int m,
              int gs, long n, D2 *__restrict__ GAall, D2 *__restrict__ GMall, long rows_per_cta) {
  constexpr int W = 16 * TS, SLOTS = (GB_ROWS * W) / 256, NP = TS * (TS + 1) / 2;
  static_assert((GB_ROWS * W) % 256 == 0, "staging slots");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  D2 *buf = reinterpret_cast<D2 *>(smem_raw);          // [2 stages][3 arrays][GB_ROWS][W]
  const int b = blockIdx.y;
  const long r0 = blockIdx.x * rows_per_cta;
  const long r1 = min(n, r0 + rows_per_cta);
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  // staging slots of this thread: (row q, padded column c) -> global column offset (or -1: zero padding)
  int sq[SLOTS], sofs[SLOTS];
#pragma unroll
  for (int u = 0; u < SLOTS; u++) {
    const int t = threadIdx.x + 256 * u;
    sq[u] = t / W;
    const int c = t - sq[u] * W;
    sofs[u] = c < k ? basis_col(c, m, gs, b) : -1;
  }
  auto issue = [&](int stage, long r) {
    D2 *dst = buf + (size_t)stage * 3 * GB_ROWS * W;
#pragma unroll
    for (int u = 0; u < SLOTS; u++) {
      const int t = threadIdx.x + 256 * u;
      const long row = r + sq[u];
      if (sofs[u] >= 0 && row < r1) {
        const long off = row * ld + sofs[u];
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst + t);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d0), "l"(S + off));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d0 + (unsigned)(GB_ROWS * W * sizeof(D2))), "l"(AS + off));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d0 + (unsigned)(2 * GB_ROWS * W * sizeof(D2))), "l"(MS + off));
      } else {
        dst[t] = make_double2(0.0, 0.0);
        dst[GB_ROWS * W + t] = make_double2(0.0, 0.0);
        dst[2 * GB_ROWS * W + t] = make_double2(0.0, 0.0);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  D2 accA[NP], accM[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) { accA[q] = make_double2(0.0, 0.0); accM[q] = make_double2(0.0, 0.0); }
  int stage = 0;
  if (r0 < r1) issue(0, r0);
  for (long r = r0; r < r1; r += GB_ROWS) {
    const bool more = r + GB_ROWS < r1;
    if (more) issue(stage ^ 1, r + GB_ROWS);
    if (more) asm volatile("cp.async.wait_group 1;\n" ::); else asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    const D2 *sS = buf + (size_t)stage * 3 * GB_ROWS * W, *sA = sS + GB_ROWS * W, *sM = sA + GB_ROWS * W;
#pragma unroll 4
    for (int q = 0; q < GB_ROWS; q++) {
      D2 x[TS], ya[TS], ym[TS];
#pragma unroll
      for (int a = 0; a < TS; a++) {
        x[a] = sS[q * W + ti + 16 * a];
        ya[a] = sA[q * W + tj + 16 * a];
        ym[a] = sM[q * W + tj + 16 * a];
      }
      int pidx = 0;
#pragma unroll
      for (int a = 0; a < TS; a++)
#pragma unroll
        for (int c = a; c < TS; c++) {
          accA[pidx].x = fma(x[a].x, ya[c].x, accA[pidx].x); accA[pidx].x = fma(x[a].y, ya[c].y, accA[pidx].x);
          accA[pidx].y = fma(x[a].x, ya[c].y, accA[pidx].y); accA[pidx].y = fma(-x[a].y, ya[c].x, accA[pidx].y);
          accM[pidx].x = fma(x[a].x, ym[c].x, accM[pidx].x); accM[pidx].x = fma(x[a].y, ym[c].y, accM[pidx].x);
          accM[pidx].y = fma(x[a].x, ym[c].y, accM[pidx].y); accM[pidx].y = fma(-x[a].y, ym[c].x, accM[pidx].y);
          pidx++;
        }
    }
    __syncthreads();
    stage ^= 1;
  }
  double *GA = reinterpret_cast<double *>(GAall + (size_t)b * k * k);
  double *GM = reinterpret_cast<double *>(GMall + (size_t)b * k * k);
  int pidx = 0;
#pragma unroll
  for (int a = 0; a < TS; a++)
#pragma unroll
    for (int c = a; c < TS; c++) {
      const int i = ti + 16 * a, j = tj + 16 * c;
      if (i < k && j < k) {
        atomicAdd(GA + 2 * (i * k + j), accA[pidx].x); atomicAdd(GA + 2 * (i * k + j) + 1, accA[pidx].y);
        atomicAdd(GM + 2 * (i * k + j), accM[pidx].x); atomicAdd(GM + 2 * (i * k + j) + 1, accM[pidx].y);
        if (a != c) {   // mirror block: C[j][i] = conj(C[i][j])
          atomicAdd(GA + 2 * (j * k + i), accA[pidx].x); atomicAdd(GA + 2 * (j * k + i) + 1, -accA[pidx].y);
          atomicAdd(GM + 2 * (j * k + i), accM[pidx].x); atomicAdd(GM + 2 * (j * k + i) + 1, -accM[pidx].y);
        }
      }
      pidx++;
    }
}

template <int TS>
static cudaError_t launch_gram2(const D2 *S, const D2 *AS, const D2 *MS, int ld, int kc, int mb, int gs, long n, int K,
                                D2 *GA, D2 *GM, cudaStream_t s) {
  const size_t smem = (size_t)2 * 3 * GB_ROWS * 16 * TS * sizeof(D2);
  static bool attr_of[kMaxDevices] = {};
  bool &attr = attr_of[current_device_slot()];
  if (!attr) {
    cudaError_t err = cudaFuncSetAttribute(k_gram2_basis<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    attr = true;
  }
  long ctas = std::max(1L, 148L * 2 / K);
  long rows = (n + ctas - 1) / ctas;
  rows = ((rows + GB_ROWS - 1) / GB_ROWS) * GB_ROWS;
  ctas = (n + rows - 1) / rows;
  k_gram2_basis<TS><<<dim3((unsigned)ctas, K), 256, smem, s>>>(S, AS, MS, ld, kc, mb, gs, n, GA, GM, rows);
  return cudaGetLastError();
}

// Jacobi diagonals are [n][nk] (one value per dof and k-point); column j of an m-column block vector belongs to
// k-point j / cpk, cpk = m / nk.
// Chebyshev: d = c0 * jac .* r ; x = d
__global__ void k_cheb_first(const double *__restrict__ jac, const D2 *__restrict__ r, D2 *__restrict__ d,
                             D2 *__restrict__ x, double c0, long n, int m, int nk, int cpk) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long row = t / m;
    const double s = c0 * jac[row * nk + (int)(t - row * m) / cpk];
    const D2 v = r[t];
    const D2 o = make_double2(s * v.x, s * v.y);
    d[t] = o;
    x[t] = o;
  }
}
// r -= q ; d = a*d + b * jac .* r ; x += d
__global__ void k_cheb_step(const double *__restrict__ jac, const D2 *__restrict__ q, D2 *__restrict__ r,
                            D2 *__restrict__ d, D2 *__restrict__ x, double a, double b, long n, int m, int nk, int cpk) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long row = t / m;
    const double s = b * jac[row * nk + (int)(t - row * m) / cpk];
    const D2 qq = q[t];
    D2 rr = r[t];
    rr.x -= qq.x; rr.y -= qq.y;
    r[t] = rr;
    D2 dd = d[t];
    dd.x = a * dd.x + s * rr.x;
    dd.y = a * dd.y + s * rr.y;
    d[t] = dd;
    D2 xx = x[t];
    xx.x += dd.x; xx.y += dd.y;
    x[t] = xx;
  }
}

// Chebyshev as a three-term recurrence on the iterate (x_0 = 0):
//   x_next = x + a (x - x_prev) + b jac .* (r - q),   q = (A + sigma M) x
// 4 reads + 1 write per entry (x_next overwrites x_prev), against 4 reads + 3 writes of the (r, d, x) form above, and
// the right-hand side r is never modified (no working copy).  first != 0: x_prev is the zero vector (not read).
__global__ void k_cheb3(const double *__restrict__ jac, const D2 *__restrict__ q, const D2 *__restrict__ r,
                        const D2 *__restrict__ x, D2 *__restrict__ xprev_next, double a, double b, long n, int m, int nk,
                        int cpk, int first) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long row = t / m;
    const double s = b * jac[row * nk + (int)(t - row * m) / cpk];
    const D2 qq = q[t], rr = r[t], xx = x[t];
    D2 xp = make_double2(0.0, 0.0);
    if (!first) xp = xprev_next[t];
    D2 o;
    o.x = xx.x + a * (xx.x - xp.x) + s * (rr.x - qq.x);
    o.y = xx.y + a * (xx.y - xp.y) + s * (rr.y - qq.y);
    xprev_next[t] = o;
  }
}
// x = c0 * jac .* r
__global__ void k_cheb3_first(const double *__restrict__ jac, const D2 *__restrict__ r, D2 *__restrict__ x, double c0,
                              long n, int m, int nk, int cpk) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long row = t / m;
    const double s = c0 * jac[row * nk + (int)(t - row * m) / cpk];
    const D2 v = r[t];
    x[t] = make_double2(s * v.x, s * v.y);
  }
}

// smoothing step of the auxiliary-space preconditioner: rr = r - q (q may be null: rr = r), d = c0 * jac .* rr,
// x = d (acc == 0) or x += d; the residual rr is stored when r_out is given (Chebyshev smoothers of degree >= 2
// continue with k_cheb_step on it)
__global__ void k_sm_first(const double *__restrict__ jac, const D2 *__restrict__ r, const D2 *__restrict__ q,
                           D2 *__restrict__ r_out, D2 *__restrict__ d, D2 *__restrict__ x, double c0, long n, int m,
                           int nk, int cpk, int acc) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long row = t / m;
    const double s = c0 * jac[row * nk + (int)(t - row * m) / cpk];
    D2 rr = r[t];
    if (q) { const D2 qq = q[t]; rr.x -= qq.x; rr.y -= qq.y; }
    if (r_out) r_out[t] = rr;
    const D2 o = make_double2(s * rr.x, s * rr.y);
    if (d) d[t] = o;
    if (acc) { D2 xx = x[t]; xx.x += o.x; xx.y += o.y; x[t] = xx; }
    else x[t] = o;
  }
}
// rr = r - q
__global__ void k_sub(const D2 *__restrict__ r, const D2 *__restrict__ q, D2 *__restrict__ rr, long total) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const D2 a = r[t], b = q[t];
    rr[t] = make_double2(a.x - b.x, a.y - b.y);
  }
}

// jac[i] = 1 / (dA[i] + sigma * dM[i])   (elementwise over the [n][nk] tables)
__global__ void k_make_jacobi(const double *dA, const double *dM, double sigma, double *jac, long n) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    const double d = dA[t] + sigma * dM[t];
    jac[t] = d > 0 ? 1.0 / d : 0.0;
  }
}

// column sums of a contiguous n x m block: sums[j] += sum_r X[r][j]  (complex)
__global__ void k_col_sum(const D2 *__restrict__ X, long n, int m, double *__restrict__ sums) {
  extern __shared__ double sred[];   // [2m]
  for (int j = threadIdx.x; j < 2 * m; j += blockDim.x) sred[j] = 0.0;
  __syncthreads();
  const long nthreads = (long)gridDim.x * blockDim.x;
  const long usable = (nthreads / m) * m;
  const long start = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (start < usable) {
    double ax = 0.0, ay = 0.0;
    for (long t = start; t < n * m; t += usable) { const D2 v = X[t]; ax += v.x; ay += v.y; }
    atomicAdd(&sred[2 * (start % m)], ax);
    atomicAdd(&sred[2 * (start % m) + 1], ay);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * m; j += blockDim.x) atomicAdd(sums + j, sred[j]);
}
// X[r][j] -= sums[j] / n for the columns of the k-points flagged in gflag (kappa == 0: S0 singular on constants)
__global__ void k_col_shift(D2 *__restrict__ X, long n, int m, const double *__restrict__ sums,
                            const int *__restrict__ gflag, int cpk) {
  const long total = n * m;
  const double inv = 1.0 / (double)n;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int j = (int)(t % m);
    if (!gflag[j / cpk]) continue;
    D2 v = X[t];
    v.x -= sums[2 * j] * inv; v.y -= sums[2 * j + 1] * inv;
    X[t] = v;
  }
}
// X[:, :m] (pitch ld) -= G (contiguous n x m)
__global__ void k_sub_strided(D2 *__restrict__ X, int ld, const D2 *__restrict__ G, long n, int m) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long r = t / m;
    const int j = (int)(t - r * m);
    D2 x = X[r * ld + j];
    const D2 g = G[t];
    x.x -= g.x; x.y -= g.y;
    X[r * ld + j] = x;
  }
}
// X[r][j] *= sc[j / cpk]   (per-k-point real factor, e.g. the lift parameter tau)
__global__ void k_col_scale_k(D2 *__restrict__ X, long n, int m, const double *__restrict__ sc, int cpk) {
  const long total = n * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const double f = sc[(int)(t % m) / cpk];
    D2 v = X[t];
    v.x *= f; v.y *= f;
    X[t] = v;
  }
}

double env_double(const char *name, double dflt) {
  const char *s = std::getenv(name);
  return s ? std::atof(s) : dflt;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// phase timing (bloch_set_profile): CUDA event pairs on the handle's stream, summed per category at the end of
// the solve.  Off by default - the events themselves are cheap but the bookkeeping is not free.
// ------------------------------------------------------------------------------------------
void bloch_handle_s::prof_begin(int cat) {
  if (!profile) return;
  cudaEvent_t a = nullptr, b = nullptr;
  if (prof_pool.size() >= 2) {
    a = prof_pool.back(); prof_pool.pop_back();
    b = prof_pool.back(); prof_pool.pop_back();
  } else {
    BLOCH_CUDA(cudaEventCreate(&a));
    BLOCH_CUDA(cudaEventCreate(&b));
  }
  BLOCH_CUDA(cudaEventRecord(a, stream));
  prof_open[cat] = {a, b};
}
void bloch_handle_s::prof_end(int cat) {
  if (!profile || !prof_open[cat].first) return;
  BLOCH_CUDA(cudaEventRecord(prof_open[cat].second, stream));
  prof_done[cat].push_back(prof_open[cat]);
  prof_open[cat] = {nullptr, nullptr};
}
void bloch_handle_s::prof_collect() {
  for (int c = 0; c < 8; c++) {
    double tot = 0.0;
    for (auto &pr : prof_done[c]) {
      float ms = 0.f;
      cudaEventSynchronize(pr.second);
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) tot += ms;
      prof_pool.push_back(pr.first);
      prof_pool.push_back(pr.second);
    }
    prof_done[c].clear();
    if (profile) stats.prof_ms[c] = tot;
  }
}

// ------------------------------------------------------------------------------------------
// x (n x nvec, pitch ldx) <- x - G S0^-1 G^H M x      (maxwell_bloch.cpp:2280-2290)
// rel_tol_k: relative tolerance of the inner S0 solve per k-point (host array of h->nk entries)
// ------------------------------------------------------------------------------------------
using ProjWork = bloch_handle_s::ProjWork;
static ProjWork &proj_work(bloch_handle_s *h) { return h->pw; }

static void remove_gamma_means(bloch_handle_s *h, D2 *v, long n, int m, double *scratch /* >= 2 m doubles */) {
  if (!h->any_gamma) return;
  cudaStream_t s = h->stream;
  BLOCH_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * m, s));
  const unsigned g1 = std::min<unsigned>(grid_for(n * m), 148);
  k_col_sum<<<g1, TPB, sizeof(double) * 2 * m, s>>>(v, n, m, scratch);
  k_col_shift<<<grid_for(n * m), TPB, 0, s>>>(v, n, m, scratch, h->d_gflag.p, std::max(1, m / h->nk));
  h->count_launch(2);
}

static void project_ld(bloch_handle_s *h, D2 *x, int ldx, int nvec, const double *rel_tol_k, int max_it, int *iters) {
  cudaStream_t s = h->stream;
  const long N = h->N, N0 = h->N0;
  const int m = nvec;
  if (m % h->nk != 0) throw std::invalid_argument("block width must be a multiple of the k-point batch size");
  h->prof_begin(2);
  ProjWork &w = proj_work(h);
  w.rhs.alloc((size_t)N0 * m); w.phi.alloc((size_t)N0 * m); w.z.alloc((size_t)N0 * m);
  w.p.alloc((size_t)N0 * m); w.q.alloc((size_t)N0 * m); w.g.alloc((size_t)N * m);
  w.scal.alloc(8 * m + 2);
  static const bool dbg = std::getenv("BLOCH_CG_DEBUG") != nullptr;
  int *d_info = reinterpret_cast<int *>(w.scal.p + 8 * m);
  // rhs = G^H M x  (becomes the CG residual), phi = 0
  BLOCH_CUDA(cudaMemsetAsync(w.rhs.p, 0, sizeof(D2) * N0 * m, s));
  BLOCH_CUDA(launch_h1_op(h->p, 2, h->tabs, h->E, x, ldx, w.rhs.p, m, m, s));
  BLOCH_CUDA(cudaMemsetAsync(w.phi.p, 0, sizeof(D2) * N0 * m, s));
  BLOCH_CUDA(cudaMemsetAsync(w.scal.p, 0, sizeof(double) * (8 * m + 2), s));
  if (h->any_gamma) {
    // Gamma point: S0 = G^T M G is singular (constants).  The exact right-hand side is orthogonal
    // to the constants; rounding is not, and CG amplifies that component without bound - remove it.
    remove_gamma_means(h, w.rhs.p, N0, m, w.scal.p);
    BLOCH_CUDA(cudaMemsetAsync(w.scal.p, 0, sizeof(double) * (8 * m + 2), s));
  }
  if (dbg) { int magic = 12345; BLOCH_CUDA(cudaMemcpyAsync(d_info + 2, &magic, sizeof(int), cudaMemcpyHostToDevice, s)); }
  int mg_its = -1;
  if (h->mg && h->use_mg) {
    // V-cycle preconditioned block PCG (mesh-independent iteration count)
    mg_its = mg_solve(h->mg, h, w.rhs.p, w.phi.p, m, rel_tol_k, 200);
  } else {
    if (h->nk != 1)
      throw std::invalid_argument("batched k-points need the multigrid projector (even n_sub, BLOCH_MG != 0)");
    // Jacobi-PCG, the whole block solve in one cooperative launch
    BLOCH_CUDA(launch_proj_cg(h->p, h->tabs, h->E, h->d_jac0.p, w.phi.p, w.rhs.p, w.z.p, w.p.p, w.q.p,
                              w.scal.p, m, N0, max_it, rel_tol_k[0], d_info, s));
  }
  // x -= G phi
  BLOCH_CUDA(launch_h1_op(h->p, 1, h->tabs, h->E, w.phi.p, m, w.g.p, m, m, s));
  k_sub_strided<<<grid_for(N * m), TPB, 0, s>>>(x, ldx, w.g.p, N, m);
  h->count_launch(4);
  int info[2] = {0, 0};
  if (mg_its < 0) {
    BLOCH_CUDA(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, s));
    h_sync(s);
  } else {
    info[0] = mg_its;
  }
  h->stats.inner_iterations += info[0];
  if (iters) *iters = info[0];
  h->prof_end(2);
}
static void project_ld(bloch_handle_s *h, D2 *x, int ldx, int nvec, double rel_tol, int max_it, int *iters) {
  std::vector<double> tol(h->nk, rel_tol);
  project_ld(h, x, ldx, nvec, tol.data(), max_it, iters);
}

void bloch_handle_s::project(D2 *x, int nvec, double rel_tol, int *iters) {
  d_jac0.alloc((size_t)N0 * nk);
  k_make_jacobi<<<grid_for(N0 * nk), TPB, 0, stream>>>(d_diagS0.p, d_diagS0.p, 0.0, d_jac0.p, N0 * nk);
  count_launch();
  project_ld(this, x, nvec, nvec, rel_tol, 3000, iters);
}

// ------------------------------------------------------------------------------------------
// Solve(): block LOBPCG
// ------------------------------------------------------------------------------------------
void bloch_handle_s::solve() {
  bloch_b200::EigProblem prob;
  prob.n = N;
  prob.apply = [this](const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm) {
    apply_nd_ld(x, ldx, y, ldy, nvec, ca, cm);
  };
  prob.diagA = d_diagA.p;
  prob.diagM = d_diagM.p;
  prob.lmax_local = lmax_local;
  prob.constrained = true;
  prob.X = &d_X;
  prob.evals = &eigenvalues;
  prob.have = &have_vectors;
  prob.blk = &block;
  prob.nbands = nbands;
  prob.use_init = true;
  lobpcg(prob);
}

// scalar H1 Bloch Helmholtz variant (misc/scalar3d.cpp:662-818): no null space, no projector
void bloch_handle_s::solve_scalar() {
  bloch_b200::EigProblem prob;
  prob.n = N0;
  prob.apply = [this](const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm) {
    apply_scalar_ld(x, ldx, y, ldy, nvec, ca, cm);
  };
  prob.diagA = d_diagS0.p;
  prob.diagM = d_diagM0.p;
  prob.lmax_local = lmax_local_h1;
  prob.constrained = false;
  if (use_mg && mg_scalar && env_double("BLOCH_SCALAR_MG", 1.0) != 0.0) prob.mg_precond = mg_scalar;   // built by setup()
  prob.X = &d_Xs;
  prob.evals = &eigenvalues_s;
  prob.have = &have_vectors_s;
  prob.blk = &block_s;
  prob.nbands = nbands_s;
  prob.use_init = false;
  lobpcg(prob);
}

// All nk k-points of the handle are iterated TOGETHER: they share mesh, maps and coefficients, so every kernel of
// the iteration (operator applies, preconditioner, projector multigrid, Gram, rotation) runs once on block vectors
// of nk * mb columns; only the class tables, Jacobi diagonals, Ritz values and Rayleigh-Ritz rotations are per
// k-point.  The eigenproblems stay independent (maxwell_dispersion.cpp:475-531): per-k Gram matrices, per-k
// Rayleigh-Ritz, per-k convergence; a k-point that has converged is frozen (its W / P columns leave the basis).
void bloch_handle_s::lobpcg(bloch_b200::EigProblem &prob) {
  using dense::cplx;
  using dense::Mat;
  cudaStream_t s = stream;
  const long N = prob.n;
  const int nb = prob.nbands;
  const int K = nk;
  int &block = *prob.blk;
  int &have_vectors = *prob.have;
  bloch_b200::DevBuf<D2> &d_X = *prob.X;
  std::vector<double> &eigenvalues = *prob.evals;
  const double lmax_local = prob.lmax_local;
  int mb = nb + std::max((int)env_double("BLOCH_GUARD", 6.0), nb / 4);
  if (mb > 21) mb = 21;      // 3 mb <= 64 basis columns (Gram kernel), lanes of k_rr_update
  if (3L * mb > N) mb = (int)(N / 3);
  if (nb > mb) throw std::invalid_argument("problem too small for the requested number of bands");
  if (block != mb) { have_vectors = 0; block = mb; if (prob.constrained) have_hist = 0; }
  const int gs = K * mb;          // columns of one basis group (X, W or P) over all k-points
  const int ld = 3 * gs;
  const long Nl = N;

  cudaEvent_t ev0, ev1;
  BLOCH_CUDA(cudaEventCreate(&ev0));
  BLOCH_CUDA(cudaEventCreate(&ev1));
  BLOCH_CUDA(cudaEventRecord(ev0, s));
  prof_begin(0);
  stats.iterations = 0;
  stats.converged = 0;
  stats.inner_iterations = 0;

  // ---- preconditioner data: sigma, Jacobi, lambda_max estimate ----
  const double vol23 = std::cbrt(mesh.volume) * std::cbrt(mesh.volume);
  sigma = env_double("BLOCH_SIGMA_SCALE", 1.0) / vol23 + env_double("BLOCH_SIGMA_BETA", 0.0) * beta * beta;
  cheb_degree = (int)env_double("BLOCH_CHEB_DEGREE", 24);
  const double cheb_ratio = env_double("BLOCH_CHEB_RATIO", 300.0);
  const double proj_tol = env_double("BLOCH_PROJ_TOL_FACTOR", 1e-2) * tol;
  const bool warm = env_double("BLOCH_WARM_START", 1.0) != 0.0;
  const int refresh_every = (int)env_double("BLOCH_REFRESH_EVERY", 1.0);
  const bool verbose = env_double("BLOCH_VERBOSE", 0.0) != 0.0;
  d_jac.alloc((size_t)Nl * K);
  d_jac0.alloc((size_t)N0 * K);
  k_make_jacobi<<<grid_for(Nl * K), TPB, 0, s>>>(prob.diagA, prob.diagM, sigma, d_jac.p, Nl * K);
  if (prob.constrained) k_make_jacobi<<<grid_for(N0 * K), TPB, 0, s>>>(d_diagS0.p, d_diagS0.p, 0.0, d_jac0.p, N0 * K);
  count_launch(2);

  // workspace lives in the handle and only grows: cudaMalloc/cudaFree per solve cost random
  // 0.5-1.5 s stalls on the shared boxes (measured)
  DevBuf<D2> &S = lw.S, &AS = lw.AS, &MS = lw.MS, &R = lw.R, &Wc = lw.Wc, &Dd = lw.Dd, &Tq = lw.Tq, &Qb = lw.Qb;
  DevBuf<D2> &dC = lw.dC, &dGA = lw.dGA, &dGM = lw.dGM;
  DevBuf<double> &dlam = lw.dlam, &drn = lw.drn;
  S.alloc((size_t)Nl * ld); AS.alloc((size_t)Nl * ld); MS.alloc((size_t)Nl * ld);
  R.alloc((size_t)Nl * gs); Wc.alloc((size_t)Nl * gs); Dd.alloc((size_t)Nl * gs);
  Tq.alloc((size_t)Nl * gs); Qb.alloc((size_t)Nl * gs);
  const int kmax = 3 * mb;
  dC.alloc((size_t)K * kmax * mb); dGA.alloc((size_t)K * kmax * kmax); dGM.alloc((size_t)K * kmax * kmax);
  dlam.alloc(gs); drn.alloc(gs); lw.dtau.alloc(K);
  // Rayleigh-Ritz on the device (rr_device.cu) unless BLOCH_RR_DEVICE=0: no host dense algebra, no copies of the
  // Gram matrices, no synchronisation between the Gram kernel and the rotation
  static const bool rr_device = env_double("BLOCH_RR_DEVICE", 1.0) != 0.0;
  lw.dact.alloc((size_t)gs); lw.dusep.alloc(K); lw.dinfo.alloc(K);
  std::vector<unsigned char> act8((size_t)gs, 1);
  std::vector<int> hinfo(K, 0);
  if (rr_device) {
    BLOCH_CUDA(cudaMemsetAsync(lw.dusep.p, 0, K, s));
    BLOCH_CUDA(cudaMemsetAsync(lw.dinfo.p, 0, sizeof(int) * K, s));
  }

  auto op = [&](const D2 *x, int ldx, D2 *y, int ldy, int nvec, double ca, double cm) {
    prob.apply(x, ldx, y, ldy, nvec, ca, cm);
  };
  // Gradient lift.  The constrained pencil is solved as the UNCONSTRAINED pencil (A_tau, M) with
  //     A_tau = A + tau M G B G^H M,   B = one multigrid V-cycle for S0 (fixed SPD linear operator).
  // Divergence-free vectors do not see the extra term (G^H M x = 0), so every physical eigenpair is an exact
  // eigenpair of (A_tau, M); gradients G phi, the kernel of A, are moved from 0 to tau * eig(B S0) in
  // [~0.25 tau, tau], above the wanted bands.  Gradient contamination of the search directions then shows up
  // as HIGH Ritz values that Rayleigh-Ritz ignores instead of spurious zeros, so W needs only a rough
  // projection (relative 1e-2, a few PCG iterations) instead of one to 1e-2 * tol (16 iterations) per outer
  // iteration; the price is one V-cycle per lifted operator application.
  // (only with a warm start: tau must sit a modest factor above the wanted bands, and the Ritz values of a
  // random block say nothing about them).  tau is per k-point; a k-point that is not lifted yet has tau = 0.
  const bool warm_block = warm && !(prob.use_init && n_init > 0) && have_vectors == mb && d_X.n >= (size_t)Nl * gs;
  const bool lift_allowed = prob.constrained && mg && use_mg && !two_pass && env_double("BLOCH_LIFT", 1.0) != 0.0;
  std::vector<char> lift(K, 0);
  std::vector<double> tau(K, 0.0);
  bool any_lift = false;
  const double lift_factor = env_double("BLOCH_LIFT_TAU", 8.0);
  const double lift_ptol = env_double("BLOCH_LIFT_PROJ_TOL", 1e-1);
  const double lift_xtol = env_double("BLOCH_LIFT_X_TOL", 1e-1);
  if (lift_allowed) { lw.Lu.alloc((size_t)N0 * gs); lw.Lphi.alloc((size_t)N0 * gs); lw.Lg.alloc((size_t)Nl * gs); }
  auto opA = [&](const D2 *x, int ldx, D2 *y, int ldy, int nvec) {   // y = A_tau x   (nvec = gs)
    op(x, ldx, y, ldy, nvec, 1.0, 0.0);
    if (!any_lift) return;
    prof_begin(6);
    BLOCH_CUDA(cudaMemsetAsync(lw.Lu.p, 0, sizeof(D2) * N0 * nvec, s));
    BLOCH_CUDA(launch_h1_op(p, 2, tabs, E, x, ldx, lw.Lu.p, nvec, nvec, s));           // u = G^H M x
    if (any_gamma) {   // S0 singular on constants at Gamma: u is orthogonal to them up to rounding
      ProjWork &w = proj_work(this);
      w.scal.alloc(8 * nvec + 2);
      remove_gamma_means(this, lw.Lu.p, N0, nvec, w.scal.p);
    }
    mg_vcycle(mg, this, lw.Lu.p, lw.Lphi.p, nvec);                                      // phi = B u
    k_col_scale_k<<<grid_for(N0 * nvec), TPB, 0, s>>>(lw.Lphi.p, N0, nvec, lw.dtau.p, nvec / K);   // phi *= tau_k
    BLOCH_CUDA(launch_h1_op(p, 1, tabs, E, lw.Lphi.p, nvec, lw.Lg.p, nvec, nvec, s));  // g = G phi
    BLOCH_CUDA(launch_nd_apply(p, tabs, E, lw.Lg.p, nvec, y, ldy, nvec, 0.0, 1.0, s));  // y += M g
    count_launch(4);
    prof_end(6);
  };

  // lambda_max(D^-1 (A + sigma M)) <= max over element classes of the local scaled spectra
  // (x^H A x = sum_e x_e^H A_e x_e <= mu sum_e x_e^H diag(A_e) x_e); computed once per handle from
  // the probe launch, 5% margin for the weak kappa dependence.  A power iteration on the global
  // operator UNDER-estimates the clustered top of the spectrum and makes the Chebyshev
  // preconditioner indefinite.
  lmaxA = env_double("BLOCH_LMAX_SCALE", 1.05) * lmax_local;
  if (verbose) std::printf("[lobpcg] lambda_max bound %.4f sigma %.3f cheb degree %d ratio %.0f, %d k-point(s)\n", lmaxA, sigma, cheb_degree, cheb_ratio, K);
  const double lmax = lmaxA, lmin = lmaxA / cheb_ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;

  // out = T r_in (both contiguous N x gs, r_in preserved): Chebyshev iteration on A + sigma M
  auto precondition = [&](const D2 *r_in, D2 *out) {
    prof_begin(3);
    prof_in_precond = true;
    const unsigned g = grid_for(Nl * gs);
    static const bool three_term = env_double("BLOCH_CHEB_THREE_TERM", 1.0) != 0.0;
    if (three_term) {
      // iterates alternate between Dd and `out`; with x_1 placed so that the last one lands in `out`
      D2 *xa = (cheb_degree % 2 == 0) ? Dd.p : out, *xb = (cheb_degree % 2 == 0) ? out : Dd.p;   // x_1 in xa
      k_cheb3_first<<<g, TPB, 0, s>>>(d_jac.p, r_in, xa, 1.0 / theta, Nl, gs, K, mb);
      count_launch();
      double rho = 1.0 / sigma1;
      D2 *xcur = xa, *xoth = xb;
      for (int k = 1; k < cheb_degree; k++) {
        op(xcur, gs, Qb.p, gs, gs, 1.0, sigma);
        const double rho_n = 1.0 / (2.0 * sigma1 - rho);
        k_cheb3<<<g, TPB, 0, s>>>(d_jac.p, Qb.p, r_in, xcur, xoth, rho_n * rho, 2.0 * rho_n / delta, Nl, gs, K, mb, k == 1 ? 1 : 0);
        count_launch();
        rho = rho_n;
        std::swap(xcur, xoth);
      }
      if (xcur != out) BLOCH_CUDA(cudaMemcpyAsync(out, xcur, sizeof(D2) * Nl * gs, cudaMemcpyDeviceToDevice, s));
    } else {
      BLOCH_CUDA(cudaMemcpyAsync(Tq.p, r_in, sizeof(D2) * Nl * gs, cudaMemcpyDeviceToDevice, s));
      k_cheb_first<<<g, TPB, 0, s>>>(d_jac.p, Tq.p, Dd.p, out, 1.0 / theta, Nl, gs, K, mb);
      count_launch();
      double rho = 1.0 / sigma1;
      for (int k = 1; k < cheb_degree; k++) {
        op(Dd.p, gs, Qb.p, gs, gs, 1.0, sigma);
        const double rho_n = 1.0 / (2.0 * sigma1 - rho);
        k_cheb_step<<<g, TPB, 0, s>>>(d_jac.p, Qb.p, Tq.p, Dd.p, out, rho_n * rho, 2.0 * rho_n / delta, Nl, gs, K, mb);
        count_launch();
        rho = rho_n;
      }
    }
    prof_in_precond = false;
    prof_end(3);
  };

  // Auxiliary-space preconditioner (what AMS does for the reference, maxwell_bloch.cpp:492-517): symmetric
  // multiplicative combination of a Chebyshev-Jacobi smoother on A + sigma M with the nodal-space correction
  // Pi B Pi^H (aux.cu: one H1 multigrid V-cycle per Cartesian component):
  //     x = S r;  x += Pi B Pi^H (r - (A + sigma M) x);  x += S (r - (A + sigma M) x)
  // The gradient part of AMS is not needed: the projector that follows removes gradients exactly.  Outer iteration
  // counts no longer grow with n_sub / order, and one application costs 2 + 2 (s - 1) ND applies + 3 scalar V-cycles
  // instead of 23 ND applies.
  const bool aux_on = prob.constrained && aux && use_aux && mg && use_mg;
  const int sm_degree = std::max(1, (int)env_double("BLOCH_AUX_SMOOTH_DEGREE", 2.0));
  const double sm_ratio = env_double("BLOCH_AUX_SMOOTH_RATIO", 4.0);
  auto precondition_aux = [&](const D2 *r_in, D2 *out) {
    prof_begin(3);
    prof_in_precond = true;
    const unsigned g = grid_for(Nl * gs);
    const double lo = lmax / sm_ratio;
    const double th = 0.5 * (lmax + lo), de = 0.5 * (lmax - lo), s1 = th / de;
    // x (+)= S b with b = r_in - q (q == null: b = r_in); uses Tq (residual), Dd (direction), Qb (operator output)
    auto smooth = [&](const D2 *q, int acc) {
      if (sm_degree == 1) {
        k_sm_first<<<g, TPB, 0, s>>>(d_jac.p, r_in, q, nullptr, nullptr, out, 1.0 / th, Nl, gs, K, mb, acc);
        count_launch();
        return;
      }
      k_sm_first<<<g, TPB, 0, s>>>(d_jac.p, r_in, q, Tq.p, Dd.p, out, 1.0 / th, Nl, gs, K, mb, acc);
      count_launch();
      double rho = 1.0 / s1;
      for (int k = 1; k < sm_degree; k++) {
        op(Dd.p, gs, Qb.p, gs, gs, 1.0, sigma);
        const double rho_n = 1.0 / (2.0 * s1 - rho);
        k_cheb_step<<<g, TPB, 0, s>>>(d_jac.p, Qb.p, Tq.p, Dd.p, out, rho_n * rho, 2.0 * rho_n / de, Nl, gs, K, mb);
        count_launch();
        rho = rho_n;
      }
    };
    smooth(nullptr, 0);
    op(out, gs, Qb.p, gs, gs, 1.0, sigma);
    k_sub<<<g, TPB, 0, s>>>(r_in, Qb.p, Tq.p, Nl * gs);
    count_launch();
    prof_in_precond = false;
    aux_correct(aux, this, Tq.p, out, gs);
    prof_in_precond = true;
    op(out, gs, Qb.p, gs, gs, 1.0, sigma);
    smooth(Qb.p, 1);
    prof_in_precond = false;
    prof_end(3);
  };

  // Scalar H1 variant (misc/scalar3d.cpp:70-85: the reference preconditions with BoomerAMG sweeps): the stiffness operator
  // A = (grad + i kappa)^H k (grad + i kappa) IS the level-0 operator of the handle's H1 multigrid (coefficient k in the
  // eps slot), so T = one V-cycle of a hierarchy built on it (kind 2: with the sigma * mass shift on the constant mode, so
  // that T does not blow up the near-constant fields at small kappa); BCC order 4 n_sub 8, 20 modes: 1.71 -> 0.84 s.
  const bool mgpre_on = !prob.constrained && prob.mg_precond != nullptr;
  auto precondition_mg = [&](const D2 *r_in, D2 *out) {
    prof_begin(3);
    mg_vcycle(prob.mg_precond, this, r_in, out, gs);
    prof_end(3);
  };

  // ---- Rayleigh-Ritz of every k-point on its first kc basis columns ----
  std::vector<D2> hGA((size_t)K * kmax * kmax), hGM((size_t)K * kmax * kmax), hC((size_t)K * kmax * mb);
  std::vector<double> lam((size_t)gs, 0.0), rn((size_t)gs, 0.0);
  std::vector<char> active((size_t)gs, 1);          // column still iterating (soft locking)
  std::vector<char> use_p(K, 0), frozen(K, 0), refresh_k(K, 0);
  double t_host_rr = 0.0;
  // one k-point: false if even the X block alone is not positive definite
  auto rr_one = [&](int b, int kc, bool &dropped) -> bool {
    const D2 *gA = hGA.data() + (size_t)b * kc * kc, *gM = hGM.data() + (size_t)b * kc * kc;
    // basis columns that take part: all of X; W_j / P_j only for unconverged j (soft locking) and
    // only if they are not numerically zero (X is M-orthonormal, so diag(GM) of X is ~1)
    std::vector<int> keep;
    for (int i = 0; i < kc; i++) {
      const int j = i % mb, grp = i / mb;
      const double dii = gM[i * kc + i].x;
      if (grp == 0 || (active[(size_t)b * mb + j] && dii > 1e-26 && !(grp == 2 && !use_p[b]))) keep.push_back(i);
    }
    int kk = (int)keep.size();
    std::vector<double> l;
    Mat C;
    bool ok = false;
    while (!ok) {
      Mat GA((size_t)kk * kk), GM((size_t)kk * kk), Cs;
      for (int a = 0; a < kk; a++)
        for (int c = 0; c < kk; c++) {
          // Hermitian part (the two triangles are computed independently on the device)
          const int i = keep[a], j = keep[c];
          const D2 x = gA[i * kc + j], xt = gA[j * kc + i], y = gM[i * kc + j], yt = gM[j * kc + i];
          GA[a * kk + c] = 0.5 * cplx(x.x + xt.x, x.y - xt.y);
          GM[a * kk + c] = 0.5 * cplx(y.x + yt.x, y.y - yt.y);
        }
      ok = dense::hegv_lowest(kk, mb, GA, GM, l, Cs);
      if (ok) {
        C.assign((size_t)kc * mb, cplx(0));
        for (int a = 0; a < kk; a++)
          for (int j = 0; j < mb; j++) C[(size_t)keep[a] * mb + j] = Cs[(size_t)a * mb + j];
      } else {
        // drop the P block first, then halve what is left of the W block, then give up
        if (kk > mb && keep.back() >= 2 * mb) {
          while (!keep.empty() && keep.back() >= 2 * mb) keep.pop_back();
        } else if (kk > mb) {
          const int nw = kk - mb;
          for (int d = 0; d < (nw + 1) / 2; d++) keep.pop_back();
        } else {
          return false;
        }
        kk = (int)keep.size();
        dropped = true;
      }
    }
    D2 *hc = hC.data() + (size_t)b * kc * mb;
    for (size_t i = 0; i < (size_t)kc * mb; i++) hc[i] = make_double2(C[i].real(), C[i].imag());
    for (int j = 0; j < mb; j++) lam[(size_t)b * mb + j] = l[j];
    return true;
  };
  auto rayleigh_ritz = [&](int kc) {
    BLOCH_CUDA(rr_update_attr());
    prof_begin(4);
    BLOCH_CUDA(cudaMemsetAsync(dGA.p, 0, sizeof(D2) * K * kc * kc, s));
    BLOCH_CUDA(cudaMemsetAsync(dGM.p, 0, sizeof(D2) * K * kc * kc, s));
    if (kc <= 32) BLOCH_CUDA(launch_gram2<2>(S.p, AS.p, MS.p, ld, kc, mb, gs, Nl, K, dGA.p, dGM.p, s));
    else if (kc <= 48) BLOCH_CUDA(launch_gram2<3>(S.p, AS.p, MS.p, ld, kc, mb, gs, Nl, K, dGA.p, dGM.p, s));
    else BLOCH_CUDA(launch_gram2<4>(S.p, AS.p, MS.p, ld, kc, mb, gs, Nl, K, dGA.p, dGM.p, s));
    count_launch();
    if (rr_device) {
      for (size_t t = 0; t < (size_t)gs; t++) act8[t] = active[t] ? 1 : 0;
      BLOCH_CUDA(cudaMemcpyAsync(lw.dact.p, act8.data(), (size_t)gs, cudaMemcpyHostToDevice, s));
      BLOCH_CUDA(launch_rr_solve(dGA.p, dGM.p, kc, mb, K, lw.dact.p, lw.dusep.p, dC.p, dlam.p, lw.dinfo.p, s));
      if (mb <= 16) {
        const unsigned g = (unsigned)std::max(1L, std::min<long>((Nl + 63) / 64, 148L * 4 / K));
        k_rr_update<16><<<dim3(g, K), 256, sizeof(D2) * (kc * mb + 64 * kc), s>>>(S.p, AS.p, MS.p, ld, kc, mb, gs, dC.p, Nl);
      } else {
        const unsigned g = (unsigned)std::max(1L, std::min<long>((Nl + 31) / 32, 148L * 4 / K));
        k_rr_update<32><<<dim3(g, K), 256, sizeof(D2) * (kc * mb + 32 * kc), s>>>(S.p, AS.p, MS.p, ld, kc, mb, gs, dC.p, Nl);
      }
      count_launch(2);
      prof_end(4);
      return;
    }
    prof_end(4);
    BLOCH_CUDA(cudaMemcpyAsync(hGA.data(), dGA.p, sizeof(D2) * K * kc * kc, cudaMemcpyDeviceToHost, s));
    BLOCH_CUDA(cudaMemcpyAsync(hGM.data(), dGM.p, sizeof(D2) * K * kc * kc, cudaMemcpyDeviceToHost, s));
    h_sync(s);
    auto th0 = std::chrono::steady_clock::now();     // the GPU is idle from here until the rotation is launched
    // the K dense problems are independent: spread them over host threads when there are several
    std::vector<char> okk(K, 1), dropped(K, 0);
    auto work = [&](int b0, int b1) {
      for (int b = b0; b < b1; b++) { bool d = false; okk[b] = rr_one(b, kc, d) ? 1 : 0; dropped[b] = d ? 1 : 0; }
    };
    static const int rr_threads = std::max(1, (int)env_double("BLOCH_RR_THREADS", 4.0));
    const int nt = std::min(K, rr_threads);
    if (nt <= 1) {
      work(0, K);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(work, (int)((long)K * t / nt), (int)((long)K * (t + 1) / nt));
      for (auto &t : th) t.join();
    }
    for (int b = 0; b < K; b++) {
      if (!okk[b]) throw std::runtime_error("Rayleigh-Ritz failed (basis numerically rank deficient)");
      if (dropped[b]) refresh_k[b] = 1;
    }
    BLOCH_CUDA(cudaMemcpyAsync(dC.p, hC.data(), sizeof(D2) * K * kc * mb, cudaMemcpyHostToDevice, s));
    BLOCH_CUDA(cudaMemcpyAsync(dlam.p, lam.data(), sizeof(double) * gs, cudaMemcpyHostToDevice, s));
    t_host_rr += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th0).count();
    prof_begin(4);
    if (mb <= 16) {
      const unsigned g = (unsigned)std::max(1L, std::min<long>((Nl + 63) / 64, 148L * 4 / K));
      k_rr_update<16><<<dim3(g, K), 256, sizeof(D2) * (kc * mb + 64 * kc), s>>>(S.p, AS.p, MS.p, ld, kc, mb, gs, dC.p, Nl);
    } else {
      const unsigned g = (unsigned)std::max(1L, std::min<long>((Nl + 31) / 32, 148L * 4 / K));
      k_rr_update<32><<<dim3(g, K), 256, sizeof(D2) * (kc * mb + 32 * kc), s>>>(S.p, AS.p, MS.p, ld, kc, mb, gs, dC.p, Nl);
    }
    count_launch();
    prof_end(4);
    h_sync(s);   // hC / lam are reused by the next call
  };

  // ---- initial block ----
  if (prob.use_init && n_init > 0) {
    if (K != 1) throw std::invalid_argument("user initial vectors are supported for single-kappa solves only");
    const int mi = std::min(n_init, mb);
    DevBuf<double> tmp;
    tmp.alloc((size_t)2 * Nl * mi);
    BLOCH_CUDA(cudaMemcpyAsync(tmp.p, init_vecs.data(), sizeof(double) * 2 * Nl * mi, cudaMemcpyHostToDevice, s));
    BLOCH_CUDA(launch_fill_random(Wc.p, Nl * mb, 0xB10C4ULL, s));
    BLOCH_CUDA(launch_pack(tmp.p, Dd.p, Nl, mi, s));
    BLOCH_CUDA(cudaMemcpy2DAsync(Wc.p, sizeof(D2) * mb, Dd.p, sizeof(D2) * mi, sizeof(D2) * mi, Nl, cudaMemcpyDeviceToDevice, s));
    h_sync(s);
    count_launch(2);
  } else if (warm_block) {
    BLOCH_CUDA(cudaMemcpyAsync(Wc.p, d_X.p, sizeof(D2) * Nl * gs, cudaMemcpyDeviceToDevice, s));
  } else {
    BLOCH_CUDA(launch_fill_random(Wc.p, Nl * gs, 0xB10C4ULL, s));
    count_launch();
  }
  {
    int its = 0;
    if (prob.constrained) project_ld(this, Wc.p, gs, gs, std::min(proj_tol, 1e-10), 3000, &its);
    BLOCH_CUDA(cudaMemcpy2DAsync(S.p, sizeof(D2) * ld, Wc.p, sizeof(D2) * gs, sizeof(D2) * gs, Nl, cudaMemcpyDeviceToDevice, s));
    op(S.p, ld, AS.p, ld, gs, 1.0, 0.0);    // X is divergence-free to 1e-10 here: A_tau X = A X
    op(S.p, ld, MS.p, ld, gs, 0.0, 1.0);
    try {
      rayleigh_ritz(mb);
    } catch (const std::runtime_error &) {
      throw std::runtime_error("initial block is rank deficient");
    }
    std::fill(refresh_k.begin(), refresh_k.end(), 0);
  }
  auto fetch_ritz = [&]() {   // device Rayleigh-Ritz: Ritz values and status of the last call (caller synchronises)
    BLOCH_CUDA(cudaMemcpyAsync(lam.data(), dlam.p, sizeof(double) * gs, cudaMemcpyDeviceToHost, s));
    BLOCH_CUDA(cudaMemcpyAsync(hinfo.data(), lw.dinfo.p, sizeof(int) * K, cudaMemcpyDeviceToHost, s));
  };
  auto check_ritz = [&](const char *what) {
    for (int b = 0; b < K; b++)
      if (hinfo[b] < 0) throw std::runtime_error(what);
  };
  if (rr_device) {
    fetch_ritz();
    h_sync(s);
    check_ritz("initial block is rank deficient");
  }
  std::vector<double> top_prev(K, 0.0);
  auto top_ritz = [&](int b) {
    double top = 0.0;
    for (int j = 0; j < mb; j++) top = std::max(top, std::fabs(lam[(size_t)b * mb + j]));
    return top;
  };
  auto enable_lift = [&](int b) {   // X (and P) of k-point b are exactly projected at this point: A_tau = A on them
    lift[b] = 1;
    any_lift = true;
    tau[b] = lift_factor * std::max(top_ritz(b), 1e-3 / vol23);
    BLOCH_CUDA(cudaMemcpyAsync(lw.dtau.p, tau.data(), sizeof(double) * K, cudaMemcpyHostToDevice, s));
    h_sync(s);
    if (verbose) std::printf("[lobpcg] k %d: gradient lift on, tau = %.4g (%.1f x top Ritz value)\n", b, tau[b], lift_factor);
  };
  BLOCH_CUDA(cudaMemsetAsync(lw.dtau.p, 0, sizeof(double) * K, s));
  for (int b = 0; b < K; b++) {
    if (lift_allowed && warm_block) enable_lift(b);
    top_prev[b] = top_ritz(b);
  }

  double t_pre = 0, t_proj = 0, t_op = 0, t_rr = 0, t_res = 0;
  auto tick = [&]() {
    if (verbose) cudaStreamSynchronize(s);
    return std::chrono::steady_clock::now();
  };
  auto since = [&](std::chrono::steady_clock::time_point t0) {
    if (verbose) cudaStreamSynchronize(s);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  bool have_P = false;
  // Sweep history (core.hpp): per k-point, X(kappa_prev2) may serve as the P block of the first Rayleigh-Ritz when this
  // k-point continues a straight walk.  The direction test excludes a walk that turns back, so the history can never
  // contain the solution of the k-point being solved (kappa == kappa_prev2 is impossible with same-direction steps).
  std::vector<char> hist_k(K, 0);
  bool any_hist = false;
  if (prob.constrained && warm_block && lift_allowed && have_hist == mb && d_Xhist.n >= (size_t)Nl * gs &&
      kappas_prev.size() == (size_t)3 * K && kappas_prev2.size() == (size_t)3 * K && env_double("BLOCH_HISTORY", 1.0) != 0.0) {
    for (int b = 0; b < K; b++) {
      double d1[3], d2[3], n1 = 0, n2 = 0, dot = 0;
      for (int d = 0; d < 3; d++) {
        d1[d] = kappas[3 * b + d] - kappas_prev[3 * b + d];
        d2[d] = kappas_prev[3 * b + d] - kappas_prev2[3 * b + d];
        n1 += d1[d] * d1[d]; n2 += d2[d] * d2[d]; dot += d1[d] * d2[d];
      }
      // only for finely sampled walks: the extrapolation error is O(step^2) against O(step) of the plain warm start, and
      // the extra block costs about a third of an iteration - measured on FCC order 2: steps of 1 % / 3 % / 6 % / 12 %
      // of Gamma-X change the warm iteration count by -2 / -1 / 0 / +1 (profiles/hist_probe_r2.log)
      const double step_max = env_double("BLOCH_HISTORY_MAX_STEP", 0.03) * 2.0 * M_PI / std::cbrt(mesh.volume);
      hist_k[b] = n1 > 0 && n2 > 0 && dot > 0.7 * std::sqrt(n1 * n2) && n1 < 9.0 * n2 && n2 < 9.0 * n1 &&
                  n1 < step_max * step_max;
      any_hist = any_hist || hist_k[b];
    }
  }
  int it = 0;
  std::vector<int> nconv(K, 0), its_k(K, 0);
  std::vector<double> maxres(K, 0.0);
  std::vector<double> ptol(K, proj_tol), xtol(K, 1e30);
  for (it = 0; it < max_iter; it++) {
    // residuals and their norms
    auto t0 = tick();
    BLOCH_CUDA(cudaMemsetAsync(drn.p, 0, sizeof(double) * gs, s));
    k_resid_norm<<<grid_for(Nl * gs), TPB, sizeof(double) * gs, s>>>(AS.p, MS.p, ld, dlam.p, R.p, Nl, gs, drn.p);
    count_launch();
    BLOCH_CUDA(cudaMemcpyAsync(rn.data(), drn.p, sizeof(double) * gs, cudaMemcpyDeviceToHost, s));
    if (rr_device && it > 0) fetch_ritz();       // the only synchronisation of the outer loop besides the inner PCG checks
    h_sync(s);
    t_res += since(t0);
    if (rr_device && it > 0) {
      check_ritz("Rayleigh-Ritz failed (basis numerically rank deficient)");
      if (lift_allowed)       // the check the host path makes right after its Rayleigh-Ritz (one synchronisation later here)
        for (int b = 0; b < K; b++) {
          if (lift[b] || frozen[b]) continue;
          const double top = top_ritz(b);
          if (it >= 2 && std::fabs(top - top_prev[b]) <= 0.1 * top) enable_lift(b);
          top_prev[b] = top;
        }
    }
    bool all_done = true;
    for (int b = 0; b < K; b++) {
      if (frozen[b]) continue;
      nconv[b] = 0;
      maxres[b] = 0;
      for (int j = 0; j < mb; j++) active[(size_t)b * mb + j] = std::sqrt(rn[(size_t)b * mb + j]) > 0.1 * tol;
      for (int j = 0; j < nb; j++) {
        const double r = std::sqrt(rn[(size_t)b * mb + j]);
        maxres[b] = std::max(maxres[b], r);
        if (r <= tol) nconv[b]++;
      }
      its_k[b] = it;
      if (nconv[b] == nb) {   // this k-point is done: freeze it (its W / P columns leave the basis)
        frozen[b] = 1;
        for (int j = 0; j < mb; j++) active[(size_t)b * mb + j] = 0;
      } else {
        all_done = false;
      }
    }
    if (verbose) {
      for (int b = 0; b < K; b++) {
        std::printf("[lobpcg] it %3d k %d conv %2d maxres %.3e lam:", it, b, nconv[b], maxres[b]);
        for (int j = 0; j < std::min(nb, 6); j++) std::printf(" %.8f", lam[(size_t)b * mb + j]);
        std::printf("\n");
      }
    }
    if (all_done) break;
    // W = P_proj T R
    t0 = tick();
    if (aux_on) precondition_aux(R.p, Wc.p);
    else if (mgpre_on) precondition_mg(R.p, Wc.p);
    else precondition(R.p, Wc.p);
    t_pre += since(t0);
    t0 = tick();
    int its = 0;
    // inexact inner solves: the gradient content left in W only has to stay well below the
    // current eigen-residual level (it enters X scaled by the size of the update)
    for (int b = 0; b < K; b++) {
      ptol[b] = frozen[b] ? 1e30 : (lift[b] ? lift_ptol : proj_tol);
      xtol[b] = (lift[b] && !frozen[b]) ? lift_xtol : 1e30;
    }
    if (prob.constrained) project_ld(this, Wc.p, gs, gs, ptol.data(), 3000, &its);
    t_proj += since(t0);
    t0 = tick();
    BLOCH_CUDA(cudaMemcpy2DAsync(S.p + gs, sizeof(D2) * ld, Wc.p, sizeof(D2) * gs, sizeof(D2) * gs, Nl, cudaMemcpyDeviceToDevice, s));
    opA(S.p + gs, ld, AS.p + gs, ld, gs);
    op(S.p + gs, ld, MS.p + gs, ld, gs, 0.0, 1.0);
    if (it == 0 && any_hist) {
      // P block of the first iteration = the eigenvectors of the last-but-one k-point of the walk, projected like W
      // (the lift is on: what is left of their gradient content shows up as high Ritz values)
      BLOCH_CUDA(cudaMemcpyAsync(Wc.p, d_Xhist.p, sizeof(D2) * Nl * gs, cudaMemcpyDeviceToDevice, s));
      std::vector<double> htol(K);
      for (int b = 0; b < K; b++) htol[b] = (hist_k[b] && !frozen[b]) ? lift_ptol : 1e30;
      int its3 = 0;
      project_ld(this, Wc.p, gs, gs, htol.data(), 3000, &its3);
      BLOCH_CUDA(cudaMemcpy2DAsync(S.p + 2 * gs, sizeof(D2) * ld, Wc.p, sizeof(D2) * gs, sizeof(D2) * gs, Nl, cudaMemcpyDeviceToDevice, s));
      opA(S.p + 2 * gs, ld, AS.p + 2 * gs, ld, gs);
      op(S.p + 2 * gs, ld, MS.p + 2 * gs, ld, gs, 0.0, 1.0);
      std::vector<unsigned char> up8(K);
      for (int b = 0; b < K; b++) { up8[b] = hist_k[b] ? 1 : 0; use_p[b] = hist_k[b]; }
      if (rr_device) BLOCH_CUDA(cudaMemcpyAsync(lw.dusep.p, up8.data(), K, cudaMemcpyHostToDevice, s));
      h_sync(s);   // up8 / htol are stack temporaries
      have_P = true;
    }
    t_op += since(t0);
    t0 = tick();
    rayleigh_ritz(have_P ? 3 * mb : 2 * mb);
    have_P = true;
    bool any_refresh = false;
    for (int b = 0; b < K; b++) {
      use_p[b] = refresh_k[b] ? 0 : 1;       // after a degenerate Rayleigh-Ritz the search directions are discarded
      any_refresh = any_refresh || refresh_k[b];
      refresh_k[b] = 0;
    }
    if (any_refresh || refresh_every <= 1 || (it % refresh_every) == refresh_every - 1) {
      // recompute A X and M X from X instead of carrying them by recurrence
      // lifted mode: the gradient content X picked up from the roughly projected W / P is removed here
      // (relative to its own small size), before A_tau X and M X are rebuilt from X
      if (any_lift) { int its2 = 0; project_ld(this, S.p, ld, gs, xtol.data(), 3000, &its2); }
      opA(S.p, ld, AS.p, ld, gs);
      op(S.p, ld, MS.p, ld, gs, 0.0, 1.0);
    }
    if (lift_allowed && !rr_device) {
      for (int b = 0; b < K; b++) {
        if (lift[b] || frozen[b]) continue;
        const double top = top_ritz(b);
        if (it >= 1 && std::fabs(top - top_prev[b]) <= 0.1 * top) enable_lift(b);
        top_prev[b] = top;
      }
    }
    t_rr += since(t0);
  }
  if (verbose)
    std::printf("[lobpcg] %d its; ms: precond %.2f project %.2f (%d cg its) AW/MW %.2f RR %.2f resid %.2f\n", it, t_pre,
                t_proj, stats.inner_iterations, t_op, t_rr, t_res);
  stats.iterations = it;
  stats.k_iterations = its_k;
  stats.k_converged = nconv;
  stats.k_max_residual = maxres;
  stats.converged = *std::min_element(nconv.begin(), nconv.end());
  stats.max_residual = *std::max_element(maxres.begin(), maxres.end());
  eigenvalues.assign((size_t)K * nb, 0.0);
  for (int b = 0; b < K; b++)
    for (int j = 0; j < nb; j++) eigenvalues[(size_t)b * nb + j] = lam[(size_t)b * mb + j];
  if (prob.constrained) {   // sweep history: the previous solution moves to d_Xhist (pointer swap), kappas shift
    if (have_vectors == mb && d_X.n >= (size_t)Nl * gs) {
      std::swap(d_X.p, d_Xhist.p);
      std::swap(d_X.n, d_Xhist.n);
      have_hist = mb;
    } else {
      have_hist = 0;
    }
    kappas_prev2 = kappas_prev;
    kappas_prev = kappas;
  }
  d_X.alloc((size_t)Nl * gs);
  BLOCH_CUDA(cudaMemcpy2DAsync(d_X.p, sizeof(D2) * gs, S.p, sizeof(D2) * ld, sizeof(D2) * gs, Nl, cudaMemcpyDeviceToDevice, s));
  have_vectors = mb;
  prof_end(0);
  BLOCH_CUDA(cudaEventRecord(ev1, s));
  BLOCH_CUDA(cudaEventSynchronize(ev1));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0, ev1);
  stats.seconds = 1e-3 * ms;
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  prof_collect();
  if (profile) stats.prof_ms[5] = t_host_rr;
}

// ------------------------------------------------------------------------------------------
// Reduced-basis k-sweep (MaxwellDispersion::buildRawBasis / approxEigenfrequencies,
// meta-material/meta_material_solver.cpp:3132-3305): full solves only at symmetry (+ mid) points;
// at any other kappa the raw eigenvectors are projected with that kappa's divergence projector
// and the pencil (A, M) is solved in their span (the reference: dsygv on the real 2N form; here
// the equivalent complex Hermitian reduced problem).
// ------------------------------------------------------------------------------------------
void bloch_handle_s::rb_append() {
  if (have_vectors <= 0) throw std::invalid_argument("no eigenvectors to append (call bloch_solve first)");
  if (nk != 1) throw std::invalid_argument("the reduced-basis sweep works on single-kappa handles");
  const int nb = nbands;
  if (rb_size + nb > rb_cap) {
    const int ncap = std::max(2 * rb_cap, rb_size + nb + 64);
    DevBuf<D2> grown;
    grown.alloc((size_t)N * ncap);
    BLOCH_CUDA(cudaMemsetAsync(grown.p, 0, sizeof(D2) * (size_t)N * ncap, stream));
    if (rb_size > 0)
      BLOCH_CUDA(cudaMemcpy2DAsync(grown.p, sizeof(D2) * ncap, d_rb.p, sizeof(D2) * rb_cap, sizeof(D2) * rb_size, N,
                                   cudaMemcpyDeviceToDevice, stream));
    h_sync(stream);
    std::swap(d_rb.p, grown.p);
    std::swap(d_rb.n, grown.n);
    rb_cap = ncap;
  }
  BLOCH_CUDA(cudaMemcpy2DAsync(d_rb.p + rb_size, sizeof(D2) * rb_cap, d_X.p, sizeof(D2) * block, sizeof(D2) * nb, N,
                               cudaMemcpyDeviceToDevice, stream));
  h_sync(stream);
  rb_size += nb;
}

void bloch_handle_s::rb_approx(double *lambda, int n) {
  using dense::cplx;
  using dense::Mat;
  const int K = rb_size;
  if (K < n || n < 1) throw std::invalid_argument("reduced basis smaller than the number of requested eigenvalues");
  cudaStream_t s = stream;
  static const bool verbose = std::getenv("BLOCH_VERBOSE") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    h_sync(s);
    auto t1 = std::chrono::steady_clock::now();
    std::printf("[rb] %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  };
  k_make_jacobi<<<grid_for(N0 * nk), TPB, 0, s>>>(d_diagS0.p, d_diagS0.p, 0.0, (d_jac0.alloc((size_t)N0 * nk), d_jac0.p), N0 * nk);
  d_rb_p.alloc((size_t)N * K); d_rb_ap.alloc((size_t)N * K); d_rb_mp.alloc((size_t)N * K);
  static const int CH = std::getenv("BLOCH_RB_CHUNK") ? std::max(1, std::atoi(std::getenv("BLOCH_RB_CHUNK"))) : 64;
  d_rb_tmp.alloc((size_t)N * CH);
  for (int c0 = 0; c0 < K; c0 += CH) {
    const int mc = std::min(CH, K - c0);
    BLOCH_CUDA(cudaMemcpy2DAsync(d_rb_tmp.p, sizeof(D2) * mc, d_rb.p + c0, sizeof(D2) * rb_cap, sizeof(D2) * mc, N,
                                 cudaMemcpyDeviceToDevice, s));
    project_ld(this, d_rb_tmp.p, mc, mc, 1e-8, 3000, nullptr);           // projector of the CURRENT kappa
    BLOCH_CUDA(cudaMemcpy2DAsync(d_rb_p.p + c0, sizeof(D2) * K, d_rb_tmp.p, sizeof(D2) * mc, sizeof(D2) * mc, N,
                                 cudaMemcpyDeviceToDevice, s));
    apply_nd_ld(d_rb_p.p + c0, K, d_rb_ap.p + c0, K, mc, 1.0, 0.0);
    apply_nd_ld(d_rb_p.p + c0, K, d_rb_mp.p + c0, K, mc, 0.0, 1.0);
  }
  lap("project + A/M applies");
  // Gram matrices block by block (the Gram kernel handles up to 64 x 64 outputs per launch); upper
  // triangle of blocks only, all blocks of both matrices copied back behind one synchronisation
  const int GB = 32, nblk = (K + GB - 1) / GB;
  Mat GA((size_t)K * K), GM((size_t)K * K);
  const size_t per = (size_t)GB * GB, nout = (size_t)2 * nblk * nblk * per;
  DevBuf<D2> &dG = lw.dGA;
  dG.alloc(nout);
  std::vector<D2> hG(nout);
  for (int which = 0; which < 2; which++) {
    const D2 *Y = which == 0 ? d_rb_ap.p : d_rb_mp.p;
    for (int bi = 0; bi < nblk; bi++)
      for (int bj = bi; bj < nblk; bj++) {
        const int i0 = bi * GB, j0 = bj * GB, mi = std::min(GB, K - i0), mj = std::min(GB, K - j0);
        BLOCH_CUDA(launch_gram(d_rb_p.p + i0, mi, K, Y + j0, mj, K, N, dG.p + ((size_t)(which * nblk + bi) * nblk + bj) * per, s));
        count_launch();
      }
  }
  BLOCH_CUDA(cudaMemcpyAsync(hG.data(), dG.p, sizeof(D2) * nout, cudaMemcpyDeviceToHost, s));
  h_sync(s);
  for (int which = 0; which < 2; which++) {
    Mat &G = which == 0 ? GA : GM;
    for (int bi = 0; bi < nblk; bi++)
      for (int bj = bi; bj < nblk; bj++) {
        const int i0 = bi * GB, j0 = bj * GB, mi = std::min(GB, K - i0), mj = std::min(GB, K - j0);
        const D2 *blk = hG.data() + ((size_t)(which * nblk + bi) * nblk + bj) * per;
        for (int a = 0; a < mi; a++)
          for (int b = 0; b < mj; b++) {
            const cplx v(blk[a * mj + b].x, blk[a * mj + b].y);
            G[(size_t)(i0 + a) * K + j0 + b] = v;
            if (bi != bj) G[(size_t)(j0 + b) * K + i0 + a] = std::conj(v);
          }
      }
    for (int a = 0; a < K; a++)
      for (int b = a; b < K; b++) {
        const cplx v = 0.5 * (G[(size_t)a * K + b] + std::conj(G[(size_t)b * K + a]));
        G[(size_t)a * K + b] = v;
        G[(size_t)b * K + a] = std::conj(v);
      }
  }
  lap("gram");
  std::vector<double> lam;
  // vectors taken from neighbouring k-points are nearly dependent: pivoted-Cholesky subset (dense.hpp)
  int rank = 0;
  if (!dense::hegv_lowest_values(K, n, GA, GM, lam, 1e-10, &rank))
    throw std::runtime_error("reduced-basis eigenproblem failed");
  if (verbose) std::printf("[rb] basis %d -> rank %d\n", K, rank);
  lap("dense hegv");
  for (int i = 0; i < n; i++) lambda[i] = lam[i];
}
