// Auxiliary-space part of the ND preconditioner: what HypreAMS contributes to the reference's
// Precond = Projector o blockdiag(AMS) (maxwell/maxwell_bloch.cpp:492-517, 2280-2290).
//
// AMS (Hiptmair-Xu) preconditions the definite curl-curl problem A + sigma M by
//     smoother  +  G B_0 G^H  +  Pi B_vec Pi^H,
// Pi: (H1_p)^3 -> ND_p the nodal interpolation of continuous vector fields, B_vec one multigrid cycle per
// Cartesian component for the scalar operator  -div mu^-1 grad  (here its Bloch form (grad + i kappa)^H mu^-1
// (grad + i kappa)).  In this solver the gradient part is redundant - the divergence projector that follows the
// preconditioner removes every gradient - so only the Pi part is built:
//   * Pi and Pi^T as explicit sparse matrices (real weights, kappa independent, built once per handle): the ND dof
//     functional of MFEM's nodal hexahedral basis is  v -> (J e_c) . v(x_node), so row (dof) holds
//     J[d][c] * c_a(g_i) for the p + 1 closed 1-D basis functions along the dof's open direction c and the three
//     components d: 3 (p + 1) entries, gathered by the same k_csr_apply kernel as the multigrid transfers;
//   * the periodic parts are interpolated (the Bloch phase lives in the operators), so Pi carries
//     no phase factor;
//   * B_vec = one V-cycle of a second H1Multigrid instance with mu^-1 as coefficient (mg_create kind 1), applied to
//     the three component blocks [N0][m] of a [3 N0][m] vector - geometric levels n -> n/2 -> ..., element-wise
//     matrix-free level operators, dense inverse on the coarsest level: iteration counts independent of n_sub.
#include <cmath>
#include <vector>

#include "aux.hpp"
#include "core.hpp"
#include "mg.hpp"

using namespace bloch_b200;
using D2 = double2;

namespace bloch_b200 {

struct AuxSpace {
  H1Multigrid *mg = nullptr;
  long N = 0, N0 = 0;
  DevBuf<int> pi_ptr, pit_ptr;          // Pi: rows = ND dofs, columns d * N0 + node; Pi^T: rows = d * N0 + node
  DevBuf<int32_t> pi_col, pit_col;
  DevBuf<double> pi_val, pit_val;
  DevBuf<D2> r3, z3;                    // [3 N0][m]
  ~AuxSpace() { if (mg) mg_destroy(mg); }
};

// Pi in CSR form (rows = ND dofs, column d * N0 + node): pure host code on the handle's mesh, maps and 1-D tables, so
// the matrix can be checked against the oracle without a device (bloch_debug_pi_matrix, tests/test_host.py)
void aux_build_pi(const bloch_handle_s *h, std::vector<int> &ptr, std::vector<int32_t> &col, std::vector<double> &val) {
  const int p = h->p, Q = p + 1, nb = p * Q * Q, LN = h->L_nd, LH = h->L_h1;
  const long N = h->N, N0 = h->N0;
  const std::vector<int32_t> &nd = h->maps.nd, &h1 = h->maps.h1;
  // representative (element, local index) of every ND dof
  std::vector<long> rep(N, -1);
  for (long e = 0; e < h->mesh.n_elem; e++)
    for (int j = 0; j < LN; j++) {
      const long g = std::labs((long)nd[(size_t)e * LN + j]) - 1;
      if (rep[g] < 0) rep[g] = e * LN + j;
    }
  ptr.assign(N + 1, 0);
  col.clear();
  val.clear();
  col.reserve((size_t)N * 3 * Q);
  val.reserve((size_t)N * 3 * Q);
  for (long g = 0; g < N; g++) {
    const long e = rep[g] / LN;
    const int j = (int)(rep[g] % LN);
    const double sgn = nd[(size_t)e * LN + j] < 0 ? -1.0 : 1.0;
    const int c = j / nb, r = j % nb;
    // natural local order: component c has p nodes along direction c and Q along the others, first index fastest
    const int n0 = c == 0 ? p : Q, n1 = c == 1 ? p : Q;
    int idx[3] = {r % n0, (r / n0) % n1, r / (n0 * n1)};
    const int io = idx[c];
    const double *J = &h->mesh.J[9 * h->mesh.cls[e]];
    for (int d = 0; d < 3; d++) {
      const double Jdc = J[3 * d + c];
      if (Jdc == 0.0) continue;
      for (int a = 0; a < Q; a++) {
        const double w = sgn * Jdc * h->basis.I[(size_t)io * Q + a];
        if (std::fabs(w) < 1e-15 * std::fabs(Jdc)) continue;
        int k[3] = {idx[0], idx[1], idx[2]};
        k[c] = a;
        const long node = (long)h1[(size_t)e * LH + k[0] + Q * (k[1] + Q * k[2])] - 1;
        col.push_back((int32_t)(d * N0 + node));
        val.push_back(w);
      }
    }
    ptr[g + 1] = (int)col.size();
  }
}

AuxSpace *aux_create(bloch_handle_s *h) {
  if (h->p > 3) return nullptr;
  H1Multigrid *mg = mg_create(h, 1);
  if (!mg) return nullptr;
  AuxSpace *ax = new AuxSpace();
  ax->mg = mg;
  ax->N = h->N;
  ax->N0 = h->N0;
  const long N = h->N, N0 = h->N0;
  std::vector<int> ptr;
  std::vector<int32_t> col;
  std::vector<double> val;
  aux_build_pi(h, ptr, col, val);
  // transpose by counting sort
  std::vector<int> tptr(3 * N0 + 1, 0);
  for (int32_t cidx : col) tptr[cidx + 1]++;
  for (long i = 0; i < 3 * N0; i++) tptr[i + 1] += tptr[i];
  std::vector<int32_t> tcol(col.size());
  std::vector<double> tval(col.size());
  std::vector<int> fill(tptr.begin(), tptr.end() - 1);
  for (long g = 0; g < N; g++)
    for (int k = ptr[g]; k < ptr[g + 1]; k++) {
      const int pos = fill[col[k]]++;
      tcol[pos] = (int32_t)g;
      tval[pos] = val[k];
    }
  cudaStream_t s = h->stream;
  ax->pi_ptr.upload(ptr, s); ax->pi_col.upload(col, s); ax->pi_val.upload(val, s);
  ax->pit_ptr.upload(tptr, s); ax->pit_col.upload(tcol, s); ax->pit_val.upload(tval, s);
  h_sync(s);
  return ax;
}

void aux_destroy(AuxSpace *ax) { delete ax; }

void aux_setup(AuxSpace *ax, bloch_handle_s *h) { mg_setup(ax->mg, h); }

void aux_apply_pi(AuxSpace *ax, bloch_handle_s *h, const D2 *u3, D2 *y, int m, bool accumulate) {
  launch_csr_apply(ax->pi_ptr.p, ax->pi_col.p, ax->pi_val.p, u3, y, ax->N, m, accumulate ? 1 : 0, h->stream);
  h->count_launch();
}

void aux_apply_pit(AuxSpace *ax, bloch_handle_s *h, const D2 *x, D2 *u3, int m) {
  launch_csr_apply(ax->pit_ptr.p, ax->pit_col.p, ax->pit_val.p, x, u3, 3 * ax->N0, m, 0, h->stream);
  h->count_launch();
}

void aux_vcycles(AuxSpace *ax, bloch_handle_s *h, const D2 *b3, D2 *z3, int m) {
  for (int d = 0; d < 3; d++) mg_vcycle(ax->mg, h, b3 + (size_t)d * ax->N0 * m, z3 + (size_t)d * ax->N0 * m, m);
}

void aux_correct(AuxSpace *ax, bloch_handle_s *h, const D2 *r, D2 *x, int m) {
  const size_t sz = (size_t)3 * ax->N0 * m;
  ax->r3.alloc(sz);
  ax->z3.alloc(sz);
  aux_apply_pit(ax, h, r, ax->r3.p, m);
  aux_vcycles(ax, h, ax->r3.p, ax->z3.p, m);
  aux_apply_pi(ax, h, ax->z3.p, x, m, true);
}

}  // namespace bloch_b200
