// Lattice catalogue restated from the reference's lib/bravais.cpp (CUB :1881-1953 and
// :2275-2313, FCC :2361-2526, BCC :2655-2893).  Only data that defines the reference's
// discretisation is kept: lattice / reciprocal / translation vectors, face radii, symmetry
// points and labels, k-paths, intermediate-point labels and the coarse hexahedral
// Wigner-Seitz dissection (1 / 4 / 16 parallelepipeds).
#include "bravais.hpp"

#include <cmath>

namespace bloch_b200 {
namespace bravais {

static inline Vec3 lin(double a, const Vec3 &x, double b = 0, const Vec3 &y = Vec3{0, 0, 0},
                       double c = 0, const Vec3 &z = Vec3{0, 0, 0}) {
  return Vec3{a * x[0] + b * y[0] + c * z[0], a * x[1] + b * y[1] + c * z[1],
              a * x[2] + b * y[2] + c * z[2]};
}
static inline double dot(const Vec3 &x, const Vec3 &y) {
  return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}
static double triple(const std::vector<Vec3> &v) {
  return v[0][0] * (v[1][1] * v[2][2] - v[1][2] * v[2][1]) +
         v[0][1] * (v[1][2] * v[2][0] - v[1][0] * v[2][2]) +
         v[0][2] * (v[1][0] * v[2][1] - v[1][1] * v[2][0]);
}

void BravaisLattice::Finish() {
  vol_ = triple(lat_vecs_);      // lib/bravais.cpp:76-108
  bz_vol_ = triple(rec_vecs_);
  for (size_t i = 0; i < sl_.size(); i++) si_[sl_[i]] = (int)i;
  ip_.resize(path_.size());      // lib/bravais.cpp:59-74: midpoints of the segments
  for (size_t p = 0; p < path_.size(); p++) {
    ip_[p].resize(path_[p].size() - 1);
    for (size_t s = 0; s + 1 < path_[p].size(); s++)
      ip_[p][s] = lin(0.5, sp_[path_[p][s]], 0.5, sp_[path_[p][s + 1]]);
  }
}

unsigned int BravaisLattice::GetNumberIntermediatePoints() const {
  unsigned int n = 0;
  for (auto &p : path_) n += (unsigned)p.size() - 1;
  return n;
}

void BravaisLattice::GetSymmetryPoint(int i, Vec3 &pt) const { pt = lin(2.0 * M_PI, sp_[i]); }

int BravaisLattice::GetSymmetryPointIndex(const std::string &label) const {
  auto it = si_.find(label);
  return it == si_.end() ? -1 : it->second;
}

void BravaisLattice::GetIntermediatePoint(int p, int s, Vec3 &pt) const {
  pt = lin(2.0 * M_PI, ip_[p][s]);
}

// Minimum-norm lattice image of pt.  The reference (lib/bravais.cpp:159-199) searches the
// 2^3 integer neighbours of B^T pt; same search here, with the images pt - A n.
bool BravaisLattice::MapToPrimitiveCell(const Vec3 &pt, Vec3 &ipt) const {
  int lo[3], hi[3];
  for (int i = 0; i < 3; i++) {
    double v = dot(rec_vecs_[i], pt);
    lo[i] = (int)std::floor(v);
    hi[i] = (int)std::ceil(v);
  }
  ipt = pt;
  double pmin = std::sqrt(dot(pt, pt));
  bool mapped = false;
  for (int j = 0; j < 8; j++) {
    Vec3 v = pt;
    bool zero = true;
    for (int i = 0; i < 3; i++) {
      int n = ((j >> i) & 1) ? hi[i] : lo[i];
      for (int d = 0; d < 3; d++) v[d] -= n * lat_vecs_[i][d];
      zero = zero && (n == 0);
    }
    double nrm = std::sqrt(dot(v, v));
    if (nrm < pmin - 1e-14) { ipt = v; pmin = nrm; mapped = !zero; }
  }
  return mapped;
}

// ---------------------------------------------------------------------------------------

static void build_cubic(BravaisLattice *L, double a, std::vector<Vec3> &lat, std::vector<Vec3> &rec,
                        std::vector<Vec3> &trn, std::vector<double> &rad, std::vector<Vec3> &sp,
                        std::vector<std::string> &sl, std::vector<std::vector<int>> &path,
                        std::vector<std::vector<std::string>> &il, std::vector<Vec3> &wv,
                        std::vector<std::array<int, 8>> &wh) {
  (void)L;
  lat = {{a, 0, 0}, {0, a, 0}, {0, 0, a}};
  rec = {{1 / a, 0, 0}, {0, 1 / a, 0}, {0, 0, 1 / a}};
  trn = lat;
  rad = {0.5 * a, 0.5 * a, 0.5 * a};
  sl = {"Gamma", "X", "M", "R"};
  sp = {Vec3{0, 0, 0}, lin(0.5, rec[1]), lin(0.5, rec[0], 0.5, rec[1]),
        lin(0.5, rec[0], 0.5, rec[1], 0.5, rec[2])};
  path = {{0, 1, 2, 0, 3, 1}, {2, 3}};
  il = {{"Delta", "Z", "Sigma", "Lambda", "S"}, {"T"}};
  double h = 0.5 * a;
  wv = {{-h, -h, -h}, {h, -h, -h}, {h, h, -h}, {-h, h, -h},
        {-h, -h, h},  {h, -h, h},  {h, h, h},  {-h, h, h}};
  wh = {{0, 1, 2, 3, 4, 5, 6, 7}};
}

static void build_fcc(double a, std::vector<Vec3> &lat, std::vector<Vec3> &rec,
                      std::vector<Vec3> &trn, std::vector<double> &rad, std::vector<Vec3> &sp,
                      std::vector<std::string> &sl, std::vector<std::vector<int>> &path,
                      std::vector<std::vector<std::string>> &il, std::vector<Vec3> &wv,
                      std::vector<std::array<int, 8>> &wh) {
  double h = 0.5 * a, q = 0.25 * a;
  lat = {{0, h, h}, {h, 0, h}, {h, h, 0}};
  rec = {{-1 / a, 1 / a, 1 / a}, {1 / a, -1 / a, 1 / a}, {1 / a, 1 / a, -1 / a}};
  trn = {{h, h, 0}, {h, -h, 0}, {h, 0, h}, {0, h, h}, {-h, 0, h}, {0, -h, h}};
  rad.assign(6, 0.5 * a / std::sqrt(6.0));
  sl = {"Gamma", "X", "W", "K", "L", "U"};
  sp = {Vec3{0, 0, 0},
        lin(0.5, rec[0], 0.0, rec[1], 0.5, rec[2]),
        lin(0.5, rec[0], 0.25, rec[1], 0.75, rec[2]),
        lin(0.375, rec[0], 0.375, rec[1], 0.75, rec[2]),
        lin(0.5, rec[0], 0.5, rec[1], 0.5, rec[2]),
        lin(0.625, rec[0], 0.25, rec[1], 0.625, rec[2])};
  path = {{0, 1, 2, 3, 0, 4, 5, 2, 4, 3}, {5, 1}};
  il = {{"Delta", "Z", "WK", "Sigma", "Lambda", "LU", "UW", "Q", "LK"}, {"T"}};
  // rhombic dodecahedron: origin, 6 four-fold tips, 8 three-fold corners
  wv.assign(15, Vec3{0, 0, 0});
  for (int d = 0; d < 3; d++) { wv[1 + 2 * d][d] = -h; wv[2 + 2 * d][d] = h; }
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++)
        wv[7 + 4 * i + 2 * j + k] = Vec3{(2 * i - 1) * q, (2 * j - 1) * q, (2 * k - 1) * q};
  wh = {{0, 9, 5, 11, 8, 1, 7, 3},
        {0, 11, 5, 9, 14, 2, 13, 4},
        {0, 8, 6, 14, 9, 1, 10, 4},
        {0, 14, 6, 8, 11, 2, 12, 3}};
}

static void build_bcc(double a, std::vector<Vec3> &lat, std::vector<Vec3> &rec,
                      std::vector<Vec3> &trn, std::vector<double> &rad, std::vector<Vec3> &sp,
                      std::vector<std::string> &sl, std::vector<std::vector<int>> &path,
                      std::vector<std::vector<std::string>> &il, std::vector<Vec3> &wv,
                      std::vector<std::array<int, 8>> &wh) {
  double h = 0.5 * a, q = 0.25 * a;
  lat = {{-h, h, h}, {h, -h, h}, {h, h, -h}};
  rec = {{0, 1 / a, 1 / a}, {1 / a, 0, 1 / a}, {1 / a, 1 / a, 0}};
  trn = {{a, 0, 0}, {0, a, 0}, {0, 0, a}, {h, h, h}, {-h, h, h}, {-h, -h, h}, {h, -h, h}};
  rad = {q / M_SQRT2, q / M_SQRT2, q / M_SQRT2};
  for (int i = 0; i < 4; i++) rad.push_back(q * std::sqrt(1.5));
  sl = {"Gamma", "H", "N", "P"};
  sp = {Vec3{0, 0, 0}, lin(0.5, rec[0], -0.5, rec[1], 0.5, rec[2]), lin(0.5, rec[2]),
        lin(0.25, rec[0], 0.25, rec[1], 0.25, rec[2])};
  path = {{0, 1, 2, 0, 3, 1}, {3, 2}};
  il = {{"Delta", "G", "Sigma", "Lambda", "F"}, {"D"}};
  // truncated octahedron: 24 corners on the six square faces (4 per face, faces in the order
  // -x +x -y +y -z +z), the 8 points (+-q,+-q,+-q) and the 6 points +-q e_d
  wv.clear();
  for (int d = 0; d < 3; d++) {
    int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;   // in-face axes in increasing order
    for (int s = -1; s <= 1; s += 2) {
      double c1[4] = {q, 0, -q, 0}, c2[4] = {0, q, 0, -q};
      for (int k = 0; k < 4; k++) {
        Vec3 v{0, 0, 0};
        v[d] = s * h; v[d1] = c1[k]; v[d2] = c2[k];
        wv.push_back(v);
      }
    }
  }
  for (int i = -1; i <= 1; i += 2)
    for (int j = -1; j <= 1; j += 2)
      for (int k = -1; k <= 1; k += 2) wv.push_back(Vec3{i * q, j * q, k * q});
  for (int d = 0; d < 3; d++)
    for (int s = -1; s <= 1; s += 2) { Vec3 v{0, 0, 0}; v[d] = s * q; wv.push_back(v); }
  wh = {{0, 1, 2, 3, 26, 32, 24, 18},     {26, 32, 24, 18, 15, 35, 36, 17},
        {15, 35, 36, 17, 30, 33, 28, 16}, {30, 33, 28, 16, 4, 5, 6, 7},
        {9, 8, 11, 10, 23, 29, 34, 25},   {23, 29, 34, 25, 22, 37, 32, 1},
        {22, 37, 32, 1, 21, 31, 35, 27},  {21, 31, 35, 27, 13, 12, 15, 14},
        {16, 17, 18, 19, 28, 36, 24, 11}, {28, 36, 24, 11, 33, 35, 32, 34},
        {33, 35, 32, 34, 5, 31, 37, 29},  {5, 31, 37, 29, 20, 21, 22, 23},
        {24, 11, 34, 32, 2, 10, 25, 1},   {11, 28, 33, 34, 8, 6, 5, 29},
        {30, 15, 35, 33, 4, 12, 31, 5},   {15, 26, 32, 35, 14, 0, 1, 27}};
}

// Hexagonal prism lattice (lib/bravais.cpp:6023-6138) with the reference's 6-hex Wigner-Seitz
// layout (rhombus prisms, 3 per layer, 2 layers; vertex/element tables at :6167-6199, which the
// reference keeps commented out in favour of a wedge mesh - hexahedra are what this path supports).
static void build_hex(double a, double c, std::vector<Vec3> &lat, std::vector<Vec3> &rec,
                      std::vector<Vec3> &trn, std::vector<double> &rad, std::vector<Vec3> &sp,
                      std::vector<std::string> &sl, std::vector<std::vector<int>> &path,
                      std::vector<std::vector<std::string>> &il, std::vector<Vec3> &wv,
                      std::vector<std::array<int, 8>> &wh) {
  const double s3 = std::sqrt(3.0);
  lat = {{0.5 * a, -std::sqrt(0.75) * a, 0}, {0.5 * a, std::sqrt(0.75) * a, 0}, {0, 0, c}};
  rec = {{1 / a, -1 / (s3 * a), 0}, {1 / a, 1 / (s3 * a), 0}, {0, 0, 1 / c}};
  trn = {{a, 0, 0}, {0.5 * a, std::sqrt(0.75) * a, 0}, {0.5 * a, -std::sqrt(0.75) * a, 0}, {0, 0, c}};
  const double r0 = std::min(0.5 * a / s3, 0.5 * c);
  rad = {r0, r0, r0, 0.5 * a};
  sl = {"Gamma", "A", "H", "K", "L", "M"};
  sp = {Vec3{0, 0, 0},
        lin(0.5, rec[2]),
        lin(1.0 / 3.0, rec[0], 1.0 / 3.0, rec[1], 0.5, rec[2]),
        lin(1.0 / 3.0, rec[0], 1.0 / 3.0, rec[1]),
        lin(0.5, rec[0], 0.0, rec[1], 0.5, rec[2]),
        lin(0.5, rec[0])};
  path = {{0, 5, 3, 0, 1, 4, 2, 1}, {4, 5}, {3, 2}};
  il = {{"Sigma", "MK", "GammaK", "Delta", "AL", "LH", "AH"}, {"LM"}, {"HK"}};
  const double h = a / s3;
  const double ring[6][2] = {{0, -h}, {0.5 * a, -0.5 * h}, {0.5 * a, 0.5 * h}, {0, h}, {-0.5 * a, 0.5 * h}, {-0.5 * a, -0.5 * h}};
  wv.clear();
  for (int layer = 0; layer < 3; layer++) {
    const double z = 0.5 * c * (layer - 1);
    wv.push_back(Vec3{0, 0, z});
    for (int k = 0; k < 6; k++) wv.push_back(Vec3{ring[k][0], ring[k][1], z});
  }
  wh.clear();
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) {
      const int v = 7 * i;
      std::array<int, 8> e{0 + v, 2 * j + 1 + v, 2 * j + 2 + v, ((2 * j + 2) % 6) + 1 + v, 0, 0, 0, 0};
      for (int k = 0; k < 4; k++) e[4 + k] = e[k] + 7;
      wh.push_back(e);
    }
}

BravaisLattice *BravaisLatticeFactory(BRAVAIS_LATTICE_TYPE type, double a, double, double c, double,
                                      double, double) {
  if (a <= 0.0) a = 1.0;
  if (c <= 0.0) c = 1.0;   // default-parameter rule of the reference factory (lib/bravais.cpp:8662-8691)
  BravaisLattice *L = new BravaisLattice();
  L->type_ = type;
  switch (type) {
    case PRIMITIVE_CUBIC:
      L->label_ = "CUB";
      build_cubic(L, a, L->lat_vecs_, L->rec_vecs_, L->trn_vecs_, L->face_radii_, L->sp_, L->sl_,
                  L->path_, L->il_, L->ws_vert_, L->ws_hex_);
      break;
    case FACE_CENTERED_CUBIC:
      L->label_ = "FCC";
      build_fcc(a, L->lat_vecs_, L->rec_vecs_, L->trn_vecs_, L->face_radii_, L->sp_, L->sl_,
                L->path_, L->il_, L->ws_vert_, L->ws_hex_);
      break;
    case BODY_CENTERED_CUBIC:
      L->label_ = "BCC";
      build_bcc(a, L->lat_vecs_, L->rec_vecs_, L->trn_vecs_, L->face_radii_, L->sp_, L->sl_,
                L->path_, L->il_, L->ws_vert_, L->ws_hex_);
      break;
    case PRIMITIVE_HEXAGONAL_PRISM:
      L->label_ = "HEX";
      build_hex(a, c, L->lat_vecs_, L->rec_vecs_, L->trn_vecs_, L->face_radii_, L->sp_, L->sl_, L->path_,
                L->il_, L->ws_vert_, L->ws_hex_);
      break;
    default:
      delete L;
      return nullptr;
  }
  L->Finish();
  return L;
}

}  // namespace bravais
}  // namespace bloch_b200
