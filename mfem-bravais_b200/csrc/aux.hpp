// Auxiliary nodal space (H1_p)^3 of the ND preconditioner (see aux.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

struct bloch_handle_s;

namespace bloch_b200 {
struct AuxSpace;
// Pi / Pi^T matrices + the mu^-1 multigrid hierarchy; nullptr when the mesh has no nested coarser level (odd n_sub)
AuxSpace *aux_create(bloch_handle_s *h);
void aux_destroy(AuxSpace *ax);
// the nodal interpolation Pi: (H1_p)^3 -> ND_p in CSR form (host only; works on topology-only handles)
void aux_build_pi(const bloch_handle_s *h, std::vector<int> &ptr, std::vector<int32_t> &col, std::vector<double> &val);
// per (kappa batch, coefficients): level operators of the auxiliary multigrid
void aux_setup(AuxSpace *ax, bloch_handle_s *h);
// y (+)= Pi u3: u3 is [3 N0][m] (component blocks), y is [N][m]
void aux_apply_pi(AuxSpace *ax, bloch_handle_s *h, const double2 *u3, double2 *y, int m, bool accumulate);
// u3 = Pi^T x
void aux_apply_pit(AuxSpace *ax, bloch_handle_s *h, const double2 *x, double2 *u3, int m);
// z3 = B b3: one V-cycle per component block ([3 N0][m] vectors)
void aux_vcycles(AuxSpace *ax, bloch_handle_s *h, const double2 *b3, double2 *z3, int m);
// x += Pi B Pi^T r, B = one V-cycle per component; r, x contiguous [N][m]
void aux_correct(AuxSpace *ax, bloch_handle_s *h, const double2 *r, double2 *x, int m);
}  // namespace bloch_b200
