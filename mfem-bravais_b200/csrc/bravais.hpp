// MFEM-free restatement of the lattice / k-path API of the reference's lib/bravais.hpp:64-181
// (class BravaisLattice) and the factory lib/bravais.hpp:1175-1239, for the lattices the
// dispersion hot path is quoted on: CUB, FCC, BCC.
// Vectors are plain std::array<double,3>; reciprocal vectors satisfy a_i . b_j = delta_ij
// (no 2 pi), GetSymmetryPoint returns 2 pi * sp exactly like lib/bravais.cpp:201-206.
#pragma once
#include <array>
#include <map>
#include <string>
#include <vector>

namespace bloch_b200 {
namespace bravais {

using Vec3 = std::array<double, 3>;

// same enumerators and values as lib/bravais.hpp:24-50
enum BRAVAIS_LATTICE_TYPE {
  INVALID_TYPE = 0,
  PRIMITIVE_SEGMENT,
  PRIMITIVE_SQUARE, PRIMITIVE_HEXAGONAL, PRIMITIVE_RECTANGULAR, CENTERED_RECTANGULAR,
  PRIMITIVE_OBLIQUE,
  PRIMITIVE_CUBIC, FACE_CENTERED_CUBIC, BODY_CENTERED_CUBIC, PRIMITIVE_TETRAGONAL,
  BODY_CENTERED_TETRAGONAL, PRIMITIVE_ORTHORHOMBIC, FACE_CENTERED_ORTHORHOMBIC,
  BODY_CENTERED_ORTHORHOMBIC, BASE_CENTERED_ORTHORHOMBIC, PRIMITIVE_HEXAGONAL_PRISM,
  PRIMITIVE_RHOMBOHEDRAL, PRIMITIVE_MONOCLINIC, BASE_CENTERED_MONOCLINIC, PRIMITIVE_TRICLINIC
};

class BravaisLattice {
public:
  BRAVAIS_LATTICE_TYPE GetLatticeType() const { return type_; }
  const std::string &GetLatticeTypeLabel() const { return label_; }
  unsigned int GetDim() const { return 3; }
  double GetUnitCellVolume() const { return vol_; }
  double GetBrillouinZoneVolume() const { return bz_vol_; }

  void GetLatticeVectors(std::vector<Vec3> &a) const { a = lat_vecs_; }
  void GetReciprocalLatticeVectors(std::vector<Vec3> &b) const { b = rec_vecs_; }
  void GetTranslationVectors(std::vector<Vec3> &t) const { t = trn_vecs_; }
  void GetFaceRadii(std::vector<double> &r) const { r = face_radii_; }

  // Returns true if the point required mapping (lib/bravais.cpp:159-199)
  bool MapToPrimitiveCell(const Vec3 &pt, Vec3 &ipt) const;

  unsigned int GetNumberSymmetryPoints() const { return (unsigned)sp_.size(); }
  unsigned int GetNumberPaths() const { return (unsigned)path_.size(); }
  unsigned int GetNumberPathSegments(int i) const { return (unsigned)path_[i].size() - 1; }
  unsigned int GetNumberIntermediatePoints() const;

  void GetSymmetryPoint(int i, Vec3 &pt) const;             // 2 pi * sp_[i]
  const std::string &GetSymmetryPointLabel(int i) const { return sl_[i]; }
  int GetSymmetryPointIndex(const std::string &label) const;
  void GetIntermediatePoint(int p, int s, Vec3 &pt) const;  // 2 pi * midpoint
  const std::string &GetIntermediatePointLabel(int p, int s) const { return il_[p][s]; }
  void GetPathSegmentEndPointIndices(int p, int s, int &e0, int &e1) const {
    e0 = path_[p][s]; e1 = path_[p][s + 1];
  }

  // Coarse Wigner-Seitz hex mesh (GetWignerSeitzCellMesh of the reference):
  // vertices and 8 vertex ids per hex in MFEM ordering.
  const std::vector<Vec3> &WignerSeitzVertices() const { return ws_vert_; }
  const std::vector<std::array<int, 8>> &WignerSeitzHexes() const { return ws_hex_; }

protected:
  friend BravaisLattice *BravaisLatticeFactory(BRAVAIS_LATTICE_TYPE, double, double, double,
                                                double, double, double);
  void Finish();   // volumes, index map, intermediate points
  std::vector<Vec3> lat_vecs_, rec_vecs_, trn_vecs_;
  std::vector<double> face_radii_;
  std::vector<Vec3> sp_;
  std::vector<std::string> sl_;
  std::map<std::string, int> si_;
  std::vector<std::vector<Vec3>> ip_;
  std::vector<std::vector<std::string>> il_;
  std::vector<std::vector<int>> path_;
  std::vector<Vec3> ws_vert_;
  std::vector<std::array<int, 8>> ws_hex_;
  std::string label_;
  BRAVAIS_LATTICE_TYPE type_ = INVALID_TYPE;
  double vol_ = 0, bz_vol_ = 0;
};

// Returns nullptr for a lattice type that is not restated yet.  Caller owns the object.
BravaisLattice *BravaisLatticeFactory(BRAVAIS_LATTICE_TYPE type, double a = 1.0, double b = 1.0,
                                      double c = 1.0, double alpha = 0, double beta = 0,
                                      double gamma = 0);

}  // namespace bravais
}  // namespace bloch_b200
