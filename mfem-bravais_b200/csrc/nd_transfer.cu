// ND prolongation between nested meshes (n_sub -> 2 n_sub, same order): the refinement operator the
// reference applies to eigenvectors in its multilevel warm start (MaxwellBlochWaveSolver::
// GetEigenfrequencies, meta-material/meta_material_solver.cpp:2829-2849: GetUpdateOperator()->Mult).
//
// Nodal interpolation of the coarse ND field at the fine dofs.  Both meshes are affine with
// J_fine = J_coarse / 2, so each component interpolates separately with 1-D tables:
//   closed direction : Pc[a][i][j] = c_j((a + l_i) / 2)         (value of the coarse GLL-Lagrange basis)
//   open direction   : Po[a][o][k] = o_k((a + g_o) / 2) / 2     (dof = t . J^T E carries the edge length)
// a = which half of the parent the child occupies.  Tangential continuity makes the copies written
// by neighbouring children agree, so the result is stored, not accumulated.
#include "kernels.hpp"
#include "elem_device.cuh"

namespace bloch_b200 {
namespace {

using namespace dev;

template <int P>
__global__ void k_nd_prolong(const __grid_constant__ NdTransfer1D T, const int32_t *__restrict__ map_f,
                             const int32_t *__restrict__ map_c, int n_elem_f, int n_f,
                             const double2 *__restrict__ xc, int ldc, double2 *__restrict__ xf, int ldf, int m) {
  using D = Dim<P>;
  constexpr int Q = P + 1;
  const long total = (long)n_elem_f * D::LND * m;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t % m);
    const long r = t / m;
    const int i = (int)(r % D::LND), e = (int)(r / D::LND);
    const int c = i / D::NB, rem = i - c * D::NB;
    const int o = rem / (Q * Q), j1 = (rem / Q) % Q, j2 = rem % Q;
    // parent element and the child's position inside it (element order of build_mesh)
    const int n3 = n_f * n_f * n_f, nc = n_f / 2;
    const int blk = e / n3, er = e - blk * n3;
    const int idx[3] = {er % n_f, (er / n_f) % n_f, er / (n_f * n_f)};
    const int par = blk * nc * nc * nc + ((idx[2] / 2) * nc + idx[1] / 2) * nc + idx[0] / 2;
    const int a0 = idx[c] & 1, a1 = idx[(c + 1) % 3] & 1, a2 = idx[(c + 2) % 3] & 1;
    const int32_t *mc = map_c + (long)par * D::LND;
    double2 acc = make_double2(0.0, 0.0);
    for (int k = 0; k < P; k++) {
      const double w0 = T.Po[a0][o][k];
      for (int k1 = 0; k1 < Q; k1++) {
        const double w1 = w0 * T.Pc[a1][j1][k1];
        if (w1 == 0.0) continue;
        for (int k2 = 0; k2 < Q; k2++) {
          const double w = w1 * T.Pc[a2][j2][k2];
          if (w == 0.0) continue;
          const int s = __ldg(mc + D::nd(c, k, k1, k2));
          const double2 z = xc[(long)((s < 0 ? -s : s) - 1) * ldc + v];
          const double ws = s < 0 ? -w : w;
          acc.x = fma(ws, z.x, acc.x);
          acc.y = fma(ws, z.y, acc.y);
        }
      }
    }
    const int sf = __ldg(map_f + (long)e * D::LND + i);
    if (sf < 0) { acc.x = -acc.x; acc.y = -acc.y; }
    xf[(long)((sf < 0 ? -sf : sf) - 1) * ldf + v] = acc;
  }
}

}  // namespace

cudaError_t launch_nd_prolong(int p, const NdTransfer1D &T, const int32_t *map_f, const int32_t *map_c, int n_elem_f,
                              int n_f, const double2 *xc, int ldc, double2 *xf, int ldf, int m, cudaStream_t s) {
  const long total = (long)n_elem_f * 3 * p * (p + 1) * (p + 1) * m;
  long g = (total + 255) / 256;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  switch (p) {
    case 1: k_nd_prolong<1><<<(unsigned)g, 256, 0, s>>>(T, map_f, map_c, n_elem_f, n_f, xc, ldc, xf, ldf, m); break;
    case 2: k_nd_prolong<2><<<(unsigned)g, 256, 0, s>>>(T, map_f, map_c, n_elem_f, n_f, xc, ldc, xf, ldf, m); break;
    case 3: k_nd_prolong<3><<<(unsigned)g, 256, 0, s>>>(T, map_f, map_c, n_elem_f, n_f, xc, ldc, xf, ldf, m); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace bloch_b200
