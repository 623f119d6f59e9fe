// Device-side building blocks shared by the element kernels (kernels.cu) and the persistent
// projector CG kernel (proj_cg.cu).  See kernels.cu for the formulation.
#pragma once
#include "kernels.hpp"

namespace bloch_b200 {
namespace dev {

__device__ __forceinline__ double2 ld2(const double2 *p) { return *p; }

// row of the (k-point, class) parameter table used by column v of an element of class cls (kernels.hpp: ElemData)
__device__ __forceinline__ int vclass(const ElemData &E, int cls, int v) { return cls + E.n_class * (v / E.cpk); }
// number of doubles of the whole table
__device__ __forceinline__ int cpar_doubles(const ElemData &E) { return E.nk * E.n_class * kClassParDoubles; }

#define CFMA(acc, a, z)            \
  {                                \
    (acc).x = fma((a), (z).x, (acc).x); \
    (acc).y = fma((a), (z).y, (acc).y); \
  }

template <int P>
struct Dim {
  static constexpr int Q = P + 1;
  static constexpr int NB = P * Q * Q;   // ND dofs per component
  static constexpr int RB = Q * P * P;   // RT dofs per component
  static constexpr int LND = 3 * NB, LRT = 3 * RB, LH1 = Q * Q * Q;
  __host__ __device__ static constexpr int nd(int c, int o, int j1, int j2) {
    return c * NB + (o * Q + j1) * Q + j2;
  }
  __host__ __device__ static constexpr int rt(int c, int j, int o1, int o2) {
    return c * RB + (j * P + o1) * P + o2;
  }
};

// ---- in-register slab transforms: s[a][b] (Q x Q), apply matrix along both indices ----
// FWD: out[r] = sum_j Mx[r][j] in[j] ; ADJ: out[j] = sum_r Mx[r][j] in[r]
template <int P, bool ADJ>
__device__ __forceinline__ void slab_transform(double2 (&s)[P + 1][P + 1], const double (&Mx)[kMaxP + 1][kMaxP + 1]) {
  constexpr int Q = P + 1;
  double2 u[Q][Q];
#pragma unroll
  for (int a = 0; a < Q; a++)
#pragma unroll
    for (int r = 0; r < Q; r++) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const double m = ADJ ? Mx[j][r] : Mx[r][j];
        CFMA(acc, m, s[a][j]);
      }
      u[a][r] = acc;
    }
#pragma unroll
  for (int r = 0; r < Q; r++)
#pragma unroll
    for (int b = 0; b < Q; b++) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const double m = ADJ ? Mx[j][r] : Mx[r][j];
        CFMA(acc, m, u[j][b]);
      }
      s[r][b] = acc;
    }
}

// ND buffer: transform every (c, o) slab along its two closed directions
template <int P, int NW, int WHICH>   // WHICH: 0 = TI fwd, 1 = TI adjoint, 2 = TIinv fwd
__device__ __forceinline__ void nd_transform_all(double2 *sND, const Tabs &T, int warp, int lane) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  for (int t = warp; t < 3 * P; t += NW) {
    const int c = t / P, o = t - c * P;
    double2 s[Q][Q];
    const int base = D::nd(c, o, 0, 0);
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) s[a][b] = sND[(base + a * Q + b) * 32 + lane];
    if (WHICH == 0) slab_transform<P, false>(s, T.TI);
    else if (WHICH == 1) slab_transform<P, true>(s, T.TI);
    else slab_transform<P, false>(s, T.TIinv);
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) sND[(base + a * Q + b) * 32 + lane] = s[a][b];
  }
}

// pointwise ND mass in mode space, in place:  F_c(i) <- coef * Omega(i) * sum_d H[c][d] F_d(i)
template <int P, int NW>
__device__ __forceinline__ void nd_mass_pointwise(double2 *sND, const Tabs &T, const double *cp,
                                                  double coef, int warp, int lane) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  const double *H = cp + 12;
  for (int g = warp; g < Q * Q * Q; g += NW) {
    int i[3];
    i[0] = g / (Q * Q);
    i[1] = (g / Q) % Q;
    i[2] = g % Q;
    double w = coef;
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
      ex[c] = i[c] < P;
      loc[c] = D::nd(c, i[c], i[(c + 1) % 3], i[(c + 2) % 3]);
      f[c] = ex[c] ? sND[loc[c] * 32 + lane] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (ex[c]) {
        double2 r;
        r.x = w * (H[3 * c] * f[0].x + H[3 * c + 1] * f[1].x + H[3 * c + 2] * f[2].x);
        r.y = w * (H[3 * c] * f[0].y + H[3 * c + 1] * f[1].y + H[3 * c + 2] * f[2].y);
        sND[loc[c] * 32 + lane] = r;
      }
    }
  }
}

template <int P>
__device__ __forceinline__ int h1_idx(int i0, int i1, int i2) {
  return (i0 * (P + 1) + i1) * (P + 1) + i2;
}

template <int P, int NW, bool ADJ>
__device__ __forceinline__ void h1_transform_dir(double2 *sH, const Tabs &T, int dir, int warp, int lane) {
  constexpr int Q = P + 1;
  const int sd = dir == 0 ? Q * Q : (dir == 1 ? Q : 1);
  const int s1 = dir == 0 ? Q : Q * Q, s2 = dir == 2 ? Q : 1;   // strides of the other two dirs
  for (int t = warp; t < Q * Q; t += NW) {
    const int a = t / Q, b = t - a * Q;
    const int base = a * s1 + b * s2;
    double2 in[Q], out[Q];
#pragma unroll
    for (int j = 0; j < Q; j++) in[j] = sH[(base + j * sd) * 32 + lane];
#pragma unroll
    for (int r = 0; r < Q; r++) {
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const double mm = ADJ ? T.TI[j][r] : T.TI[r][j];
        CFMA(acc, mm, in[j]);
      }
      out[r] = acc;
    }
#pragma unroll
    for (int j = 0; j < Q; j++) sH[(base + j * sd) * 32 + lane] = out[j];
  }
}


// ---- register-resident S0 element operator for one (element, vector) item, P <= 2 ----
// sin: thread-private shared-memory column (element k at sin[k * STRIDE]) holding the gathered
// nodal values on entry (overwritten by their mode-space transform); out: S0_e applied, nodal.
// Returns Re(phi^H S0_e phi) (evaluated in mode space).
template <int P, int STRIDE>
__device__ __forceinline__ double s0_item(const Tabs &T, const double *cp, double eps, double2 *sin,
                                          double2 (&out)[(P + 1) * (P + 1) * (P + 1)]) {
  constexpr int Q = P + 1;
  auto IDX = [](int i0, int i1, int i2) { return (i0 * Q + i1) * Q + i2; };
  // nodal -> mode, direction by direction, in place in the private column
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int sd = d == 0 ? Q * Q : (d == 1 ? Q : 1);
    const int s1 = d == 0 ? Q : Q * Q, s2 = d == 2 ? Q : 1;
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) {
        const int base = a * s1 + b * s2;
        double2 in[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) in[j] = sin[(base + j * sd) * STRIDE];
#pragma unroll
        for (int r = 0; r < Q; r++) {
          double2 acc = make_double2(0.0, 0.0);
#pragma unroll
          for (int j = 0; j < Q; j++) CFMA(acc, T.TI[r][j], in[j]);
          sin[(base + r * sd) * STRIDE] = acc;
        }
      }
  }
  const double kh[3] = {cp[0], cp[1], cp[2]};
  const double *H = cp + 12;
#pragma unroll
  for (int k = 0; k < Q * Q * Q; k++) out[k] = make_double2(0.0, 0.0);
  double dot = 0.0;
#pragma unroll
  for (int i0 = 0; i0 < Q; i0++)
#pragma unroll
    for (int i1 = 0; i1 < Q; i1++)
#pragma unroll
      for (int i2 = 0; i2 < Q; i2++) {
        const int i[3] = {i0, i1, i2};
        const double w = eps * T.om[i0] * T.om[i1] * T.om[i2];
        const double2 centre = sin[IDX(i0, i1, i2) * STRIDE];
        double2 F[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          F[d] = make_double2(0.0, 0.0);
          if (i[d] < P) {
            F[d].x = kh[d] * centre.y;
            F[d].y = -kh[d] * centre.x;
#pragma unroll
            for (int t = 0; t < Q; t++) {
              int j[3] = {i0, i1, i2};
              j[d] = t;
              const double2 val = sin[IDX(j[0], j[1], j[2]) * STRIDE];
              CFMA(F[d], T.Dt[i[d]][t], val);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (i[c] < P) {
            double2 mf;
            mf.x = w * (H[3 * c] * F[0].x + H[3 * c + 1] * F[1].x + H[3 * c + 2] * F[2].x);
            mf.y = w * (H[3 * c] * F[0].y + H[3 * c + 1] * F[1].y + H[3 * c + 2] * F[2].y);
#pragma unroll
            for (int t = 0; t < Q; t++) {
              int j[3] = {i0, i1, i2};
              j[c] = t;
              CFMA(out[IDX(j[0], j[1], j[2])], T.Dt[i[c]][t], mf);
            }
            out[IDX(i0, i1, i2)].x -= kh[c] * mf.y;
            out[IDX(i0, i1, i2)].y += kh[c] * mf.x;
          }
        }
      }
  // phi^H S0 phi in mode space
#pragma unroll
  for (int k = 0; k < Q * Q * Q; k++) {
    const double2 a = sin[k * STRIDE];
    dot = fma(a.x, out[k].x, dot);
    dot = fma(a.y, out[k].y, dot);
  }
  // mode -> nodal (adjoint transform) in registers
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int sd = d == 0 ? Q * Q : (d == 1 ? Q : 1);
    const int s1 = d == 0 ? Q : Q * Q, s2 = d == 2 ? Q : 1;
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) {
        const int base = a * s1 + b * s2;
        double2 in[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) in[j] = out[base + j * sd];
#pragma unroll
        for (int r = 0; r < Q; r++) {
          double2 acc = make_double2(0.0, 0.0);
#pragma unroll
          for (int j = 0; j < Q; j++) CFMA(acc, T.TI[j][r], in[j]);
          out[base + r * sd] = acc;
        }
      }
  }
  return dot;
}

// ---- S0 element operator with the real and the imaginary part of an (element, vector) item in ADJACENT LANES ----
// Every tensor contraction of S0_e is complex-by-real, so both parts run the same real code; only the Bloch terms
// -i kh couple them (partner value by SHFL.BFLY 1).  Per lane: the Q^3 nodal values of its part in a lane-private
// shared-memory column (entry k at sin[k * STRIDE], overwritten by the mode-space transform) and Q^3 accumulators in
// registers - half the registers of s0_item, which is what makes order 3 fit (64 accumulators).
// sgn = +1 for the real lane, -1 for the imaginary lane.  All 32 lanes of the warp must call it.
// T: a SHARED-MEMORY copy of the 1-D tables (st_TI[r][j], st_Dt[a][t], st_om[i], rows of kMaxP + 1): as kernel
// parameters they arrive through LDC into registers, ptxas hoists those loads and order 3 spills; shared-memory
// loads stay behind the compiler fences below.
struct PairTabs { double TI[kMaxP + 1][kMaxP + 1], Dt[kMaxP][kMaxP + 1], om[kMaxP + 1]; };
template <int P, int STRIDE>
__device__ __forceinline__ void s0_pair(const PairTabs &T, const double *cp, double eps, double sgn, double *sin,
                                        double (&out)[(P + 1) * (P + 1) * (P + 1)]) {
  constexpr int Q = P + 1;
  auto IDX = [](int i0, int i1, int i2) { return (i0 * Q + i1) * Q + i2; };
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int sd = d == 0 ? Q * Q : (d == 1 ? Q : 1);
    const int s1 = d == 0 ? Q : Q * Q, s2 = d == 2 ? Q : 1;
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) {
        const int base = a * s1 + b * s2;
        double in[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) in[j] = sin[(base + j * sd) * STRIDE];
#pragma unroll
        for (int r = 0; r < Q; r++) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < Q; j++) acc = fma(T.TI[r][j], in[j], acc);
          sin[(base + r * sd) * STRIDE] = acc;
        }
        asm volatile("" ::: "memory");
      }
  }
  const double kh[3] = {sgn * cp[0], sgn * cp[1], sgn * cp[2]};
  const double *H = cp + 12;
#pragma unroll
  for (int k = 0; k < Q * Q * Q; k++) out[k] = 0.0;
#pragma unroll
  for (int i0 = 0; i0 < Q; i0++)
#pragma unroll
    for (int i1 = 0; i1 < Q; i1++)
#pragma unroll
      for (int i2 = 0; i2 < Q; i2++) {
        const int i[3] = {i0, i1, i2};
        double eps_here = eps;
        asm volatile("" : "+d"(eps_here));     // opaque per point: the Q^3 weights are not precomputed into registers
        const double w = eps_here * T.om[i0] * T.om[i1] * T.om[i2];
        const double centre_other = __shfl_xor_sync(0xffffffffu, sin[IDX(i0, i1, i2) * STRIDE], 1);
        double F[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          F[d] = 0.0;
          if (i[d] < P) {
            F[d] = kh[d] * centre_other;          // (-i kh u): re <- +kh im, im <- -kh re
#pragma unroll
            for (int t = 0; t < Q; t++) {
              int j[3] = {i0, i1, i2};
              j[d] = t;
              F[d] = fma(T.Dt[i[d]][t], sin[IDX(j[0], j[1], j[2]) * STRIDE], F[d]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (i[c] < P) {
            const double mf = w * (H[3 * c] * F[0] + H[3 * c + 1] * F[1] + H[3 * c + 2] * F[2]);
            const double mf_other = __shfl_xor_sync(0xffffffffu, mf, 1);
#pragma unroll
            for (int t = 0; t < Q; t++) {
              int j[3] = {i0, i1, i2};
              j[c] = t;
              out[IDX(j[0], j[1], j[2])] = fma(T.Dt[i[c]][t], mf, out[IDX(j[0], j[1], j[2])]);
            }
            out[IDX(i0, i1, i2)] = fma(-kh[c], mf_other, out[IDX(i0, i1, i2)]);   // (+i kh mf): re <- -kh im, im <- +kh re
          }
        }
        // keeps ptxas from hoisting the shared-memory loads of later points above this one (at order 3 it otherwise
        // pulls hundreds of them to the top: 255 registers and spills)
        asm volatile("" ::: "memory");
      }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int sd = d == 0 ? Q * Q : (d == 1 ? Q : 1);
    const int s1 = d == 0 ? Q : Q * Q, s2 = d == 2 ? Q : 1;
#pragma unroll
    for (int a = 0; a < Q; a++)
#pragma unroll
      for (int b = 0; b < Q; b++) {
        const int base = a * s1 + b * s2;
        double in[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) in[j] = out[base + j * sd];
#pragma unroll
        for (int r = 0; r < Q; r++) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < Q; j++) acc = fma(T.TI[j][r], in[j], acc);
          out[base + r * sd] = acc;
        }
        asm volatile("" ::: "memory");
      }
  }
}

}  // namespace dev
}  // namespace bloch_b200
