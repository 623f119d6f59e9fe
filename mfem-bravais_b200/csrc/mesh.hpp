// Periodic hexahedral mesh of a Wigner-Seitz cell and the element -> global DOF maps of the
// H1_p / ND_p / RT_{p-1} spaces the reference builds in
// MaxwellBlochWaveEquation::MaxwellBlochWaveEquation (maxwell/maxwell_bloch.cpp:34-140) on the
// mesh produced by lib/bravais.cpp:249-355 (WS mesh -> uniform refinement -> MakePeriodicMesh).
//
// All coarse hexes of the CUB/FCC/BCC dissections are parallelepipeds, so every element is
// affine: x = x0 + J xhat, with J taken from a tiny set of "classes" (one per coarse hex).
// Periodic identification is done on entities (vertices, edges, faces) through their centre
// modulo the lattice, never through vertex pairs (small periodic meshes).
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "bravais.hpp"

namespace bloch_b200 {

struct HexMesh {
  int n_sub = 0;                  // subdivisions per coarse-hex edge
  int n_elem = 0, n_class = 0;
  std::vector<double> x0;         // [n_elem][3] element origin
  std::vector<int> cls;           // [n_elem]
  std::vector<double> J;          // [n_class][3][3], J[i][j] = d x_i / d xhat_j (already / n_sub)
  std::vector<double> rec;        // [3][3] reciprocal vectors (rows)
  double volume = 0;
  int n_vert = 0, n_edge = 0, n_face = 0;   // periodic entity counts (Euler: V - E + F - Ne = 0)

  void centers(std::vector<double> &c) const;   // [n_elem][3]
};

// Local DOF orderings.
//  "natural" (exported through the C ABI, identical to the oracle's):
//     ND : [x-comp (p,p+1,p+1)] [y-comp (p+1,p,p+1)] [z-comp (p+1,p+1,p)]   i fastest
//     RT : [x-comp (p+1,p,p)]   [y-comp (p,p+1,p)]   [z-comp (p,p,p+1)]
//     H1 : (p+1)^3 i fastest
//  "cyclic" (used by the kernels): component c is stored [open dir c][closed c+1][closed c+2]
//     for ND and [closed dir c][open c+1][open c+2] for RT, last index fastest... see kernels.
struct DofMaps {
  int p = 0;
  long n_h1 = 0, n_nd = 0, n_rt = 0;
  int l_h1 = 0, l_nd = 0, l_rt = 0;   // local sizes
  // signed 1-based global ids in natural local order: s = +-(gid+1)
  std::vector<int32_t> h1, nd, rt;    // [n_elem][l_*]
};

// Builds the mesh: every coarse WS hex subdivided n x n x n.  Throws std::runtime_error if a
// coarse hex is not affine.
void build_ws_mesh(const bravais::BravaisLattice &lat, int n_sub, HexMesh &mesh);

// Generic entry: affine hexes given by 8 vertices (MFEM order) + lattice reciprocal vectors.
void build_mesh(const std::vector<std::array<double, 3>> &vert,
                const std::vector<std::array<int, 8>> &hex, const double rec[9], int n_sub,
                HexMesh &mesh);

void build_dofmaps(HexMesh &mesh, int p, DofMaps &maps);

}  // namespace bloch_b200
