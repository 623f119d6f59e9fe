// Geometric h-multigrid preconditioned block PCG for S0 phi = rhs (see mg.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

struct bloch_handle_s;

namespace bloch_b200 {
struct H1Multigrid;
// builds the nested level hierarchy (topology, transfer tables); nullptr if n_sub is odd.
// kind 0: level operators S0 = G^H M1(eps) G (the projector's problem); kind 1: the same Bloch Laplacian with
// mu^-1 as coefficient - the scalar operator of the auxiliary nodal space (H1)^3 of the ND preconditioner (aux.cu);
// kind 2: the kind-0 operator as PRECONDITIONER of the scalar H1 eigenproblem (misc/scalar3d.cpp), i.e. with the
// sigma * mass shift on the constant mode that kind 1 has as well (the projector needs the unshifted S0)
H1Multigrid *mg_create(bloch_handle_s *h, int kind = 0);
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
// number of H1 dofs of a level (0 = the handle's mesh), -1 if the level does not exist
long mg_level_size(H1Multigrid *mg, int level);
// test hook: level 0 <-> level 1 transfer, variant 0 CSR / 1 sum-factorised / 2 element-wise; dir 0 prolongation, 1 restriction
void mg_debug_transfer(H1Multigrid *mg, bloch_handle_s *h, int variant, int dir, const double2 *x, double2 *y, int m);
// Y[row][v] (+)= sum_k val[k] X[col[k]][v], rows in CSR form (real weights, complex block vectors of m columns)
void launch_csr_apply(const int *ptr, const int32_t *col, const double *val, const double2 *X, double2 *Y, long nrows,
                      int m, int accumulate, cudaStream_t s);
}  // namespace bloch_b200
