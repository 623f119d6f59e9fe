// Geometric h-multigrid preconditioned block PCG for S0 phi = rhs (see mg.cu).
#pragma once
#include <cuda_runtime.h>

struct bloch_handle_s;

namespace bloch_b200 {
struct H1Multigrid;
// builds the nested level hierarchy (topology, transfer tables); nullptr if n_sub is odd
H1Multigrid *mg_create(bloch_handle_s *h);
void mg_destroy(H1Multigrid *mg);
// per (kappa, coefficients): class tables, restricted coefficients, Jacobi diagonals, coarse inverse
void mg_setup(H1Multigrid *mg, bloch_handle_s *h);
// rhs (N0 x m contiguous) is overwritten by the final residual; returns the PCG iteration count.
// rel_tol_k: relative residual tolerance per k-point of the handle's batch (host array, h->nk entries; column j of
// the block belongs to k-point j / (m / nk))
int mg_solve(H1Multigrid *mg, bloch_handle_s *h, double2 *rhs, double2 *phi, int m, const double *rel_tol_k, int max_it);
// x = B b: ONE V-cycle (a fixed symmetric positive definite linear operator, spectrally equivalent to S0^-1);
// b (N0 x m contiguous) is preserved
void mg_vcycle(H1Multigrid *mg, bloch_handle_s *h, const double2 *b, double2 *x, int m);
}  // namespace bloch_b200
