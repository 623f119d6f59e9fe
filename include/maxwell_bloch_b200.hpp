// C++ host-side mirror of mfem::bloch::MaxwellBlochWaveEquation (maxwell/maxwell_bloch.hpp:140-234)
// over the C ABI of bloch_b200.h.  Same member names and argument meaning as the reference; MFEM
// types are flattened: Vector -> std::vector<double>, Coefficient -> one value per element
// (the reference samples its coefficients into an L2 order-0 grid function anyway,
// maxwell/maxwell_dispersion.cpp:398-420), HypreParVector of length 2N -> [re(N); im(N)].
// Header-only; link with -lbloch_b200.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdint>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "bloch_b200.h"

namespace bloch_b200 {

struct Error : std::runtime_error {
  explicit Error(const std::string &what) : std::runtime_error(what + ": " + bloch_last_error()) {}
};
inline void check(int rc, const char *what) { if (rc < 0) throw Error(what); }

// bravais::BravaisLattice (lib/bravais.hpp:64-181) + BravaisLatticeFactory (:1175-1239)
class BravaisLattice {
public:
  BravaisLattice(int lattice_type, double a = 1, double b = 1, double c = 1, double alpha = 0,
                 double beta = 0, double gamma = 0) {
    check(bloch_lattice_create(&h_, lattice_type, a, b, c, alpha, beta, gamma), "bloch_lattice_create");
  }
  ~BravaisLattice() { bloch_lattice_destroy(h_); }
  BravaisLattice(const BravaisLattice &) = delete;
  BravaisLattice &operator=(const BravaisLattice &) = delete;
  bloch_lattice handle() const { return h_; }

  std::string GetLatticeTypeLabel() const { char b[64]; bloch_lattice_label(h_, b, 64); return b; }
  double GetUnitCellVolume() const { return bloch_lattice_volume(h_); }
  void GetLatticeVectors(std::vector<std::vector<double>> &a) const { a = mat(true); }
  void GetReciprocalLatticeVectors(std::vector<std::vector<double>> &b) const { b = mat(false); }
  void GetTranslationVectors(std::vector<std::vector<double>> &t) const {
    const int nt = bloch_lattice_num_translations(h_);
    std::vector<double> trn(3 * nt), rad(nt);
    check(bloch_lattice_translations(h_, trn.data(), rad.data()), "GetTranslationVectors");
    t.assign(nt, std::vector<double>(3));
    for (int i = 0; i < nt; i++) for (int d = 0; d < 3; d++) t[i][d] = trn[3 * i + d];
  }
  unsigned GetNumberSymmetryPoints() const { return bloch_lattice_num_symmetry_points(h_); }
  unsigned GetNumberPaths() const { return bloch_lattice_num_paths(h_); }
  unsigned GetNumberPathSegments(int p) const { return bloch_lattice_num_path_segments(h_, p); }
  void GetSymmetryPoint(int i, std::vector<double> &pt) const {
    pt.assign(3, 0.0);
    check(bloch_lattice_symmetry_point(h_, i, pt.data(), nullptr, 0), "GetSymmetryPoint");
  }
  std::string GetSymmetryPointLabel(int i) const {
    char b[64];
    check(bloch_lattice_symmetry_point(h_, i, nullptr, b, 64), "GetSymmetryPointLabel");
    return b;
  }
  int GetSymmetryPointIndex(const std::string &l) const { return bloch_lattice_symmetry_point_index(h_, l.c_str()); }
  void GetIntermediatePoint(int p, int s, std::vector<double> &pt) const {
    pt.assign(3, 0.0);
    check(bloch_lattice_intermediate_point(h_, p, s, pt.data(), nullptr, 0), "GetIntermediatePoint");
  }
  std::string GetIntermediatePointLabel(int p, int s) const {
    char b[64];
    check(bloch_lattice_intermediate_point(h_, p, s, nullptr, b, 64), "GetIntermediatePointLabel");
    return b;
  }
  void GetPathSegmentEndPointIndices(int p, int s, int &e0, int &e1) const {
    check(bloch_lattice_path_segment(h_, p, s, &e0, &e1), "GetPathSegmentEndPointIndices");
  }

private:
  std::vector<std::vector<double>> mat(bool lat) const {
    double a[9], b[9];
    bloch_lattice_vectors(h_, a, b);
    const double *m = lat ? a : b;
    return {{m[0], m[1], m[2]}, {m[3], m[4], m[5]}, {m[6], m[7], m[8]}};
  }
  bloch_lattice h_ = nullptr;
};

class MaxwellBlochWaveEquation {
public:
  // reference: MaxwellBlochWaveEquation(ParMesh &pmesh, int order); the periodic WS mesh is named
  // by (lattice, n_sub) here, n_sub = 2^(serial + parallel refinements)
  MaxwellBlochWaveEquation(const BravaisLattice &lat, int n_sub, int order, int device = -1) : order_(order) {
    check(bloch_create(&h_, lat.handle(), n_sub, order, device), "bloch_create");
    int64_t ne, n, nrt, nh1;
    int nc;
    bloch_num_elements(h_, &ne, &nc);
    bloch_num_dofs(h_, &n, &nrt, &nh1);
    n_elem_ = ne; N_ = n; Nrt_ = nrt; Nh1_ = nh1;
  }
  ~MaxwellBlochWaveEquation() { bloch_destroy(h_); }
  MaxwellBlochWaveEquation(const MaxwellBlochWaveEquation &) = delete;
  MaxwellBlochWaveEquation &operator=(const MaxwellBlochWaveEquation &) = delete;

  int64_t GetHCurlTrueVSize() const { return N_; }     // GetHCurlFESpace()->GlobalTrueVSize()
  int64_t GetHDivTrueVSize() const { return Nrt_; }
  int64_t GetNE() const { return n_elem_; }
  void GetElementCenters(std::vector<double> &xyz) const {
    xyz.resize(3 * n_elem_);
    check(bloch_element_centers(h_, xyz.data()), "bloch_element_centers");
  }

  void SetKappa(const std::vector<double> &kappa) {
    kappa_.assign(kappa.begin(), kappa.begin() + 3);
    check(bloch_set_kappa(h_, kappa.data()), "SetKappa");
  }
  // k-point batch (bloch_set_kappa_batch): nk Bloch vectors, kappas[3 * nk], iterated together by the next Solve();
  // SelectKPoint chooses the k-point the getters refer to.  GetEigenvaluesBatch = the 4-argument convenience form of
  // GetEigenvalues (maxwell_bloch.cpp:1078-1097) for nk k-points at once, eigenvalues[k] as the reference returns them.
  void SetKappaBatch(const std::vector<double> &kappas) {
    kappa_.assign(kappas.begin(), kappas.begin() + 3);
    check(bloch_set_kappa_batch(h_, (int)(kappas.size() / 3), kappas.data()), "SetKappaBatch");
  }
  void SelectKPoint(int k) { check(bloch_select_kpoint(h_, k), "SelectKPoint"); }
  void GetEigenvaluesBatch(int nev, const std::vector<double> &kappas, std::vector<std::vector<double>> &eigenvalues) {
    SetNumEigs(nev);
    SetKappaBatch(kappas);
    Setup();
    Solve();
    const int nk = (int)(kappas.size() / 3);
    eigenvalues.resize(nk);
    for (int k = 0; k < nk; k++) { SelectKPoint(k); GetEigenvalues(eigenvalues[k]); }
    SelectKPoint(0);
  }
  void SetBeta(double beta) { beta_ = beta; push_beta_zeta(); }
  void SetZeta(const std::vector<double> &zeta) { zeta_ = zeta; push_beta_zeta(); }
  void SetAbsoluteTolerance(double atol) { check(bloch_set_tol(h_, atol, 2000), "SetAbsoluteTolerance"); }
  void SetNumEigs(int nev) { nev_ = nev; check(bloch_set_num_bands(h_, (nev + 1) / 2), "SetNumEigs"); }
  void SetMassCoef(const std::vector<double> &eps_per_elem) { check(bloch_set_eps(h_, eps_per_elem.data()), "SetMassCoef"); }
  void SetStiffnessCoef(const std::vector<double> &muinv_per_elem) { check(bloch_set_muinv(h_, muinv_per_elem.data()), "SetStiffnessCoef"); }
  void Setup() { check(bloch_setup(h_), "Setup"); }
  void SetInitialVectors(int num_vecs, const double *vecs /* num_vecs x 2N */) {
    check(bloch_set_initial_vectors(h_, num_vecs, vecs), "SetInitialVectors");
  }
  void SetBravaisLattice(const BravaisLattice &) {}     // reference stores an unused pointer
  void Solve() { check(bloch_solve(h_), "Solve"); }

  // One block of the assembled operators in hypre's IJ text format, as HypreParMatrix::Print writes it
  // (file <path>.00000: "ilower iupper jlower jupper", then "i j value" lines) - the -wm dump of
  // maxwell_dispersion.cpp:553-590.  which = 0: A (imag = false: Ar, true: Ai incl. its block coefficient), 1: M.
  void WriteMatrix(int which, bool imag, const std::string &path) {
    int64_t nnz = 0;
    check(bloch_assemble_matrix(h_, which, &nnz), "bloch_assemble_matrix");
    int64_t n_nd = 0;
    check(bloch_num_dofs(h_, &n_nd, nullptr, nullptr), "bloch_num_dofs");
    std::vector<int64_t> ptr(n_nd + 1);
    std::vector<int32_t> col(nnz);
    std::vector<double> re(nnz), im(nnz);
    check(bloch_get_matrix(h_, ptr.data(), col.data(), re.data(), im.data()), "bloch_get_matrix");
    FILE *f = std::fopen((path + ".00000").c_str(), "w");
    if (!f) throw Error(("cannot open " + path).c_str());
    std::fprintf(f, "%lld %lld %lld %lld\n", 0LL, (long long)n_nd - 1, 0LL, (long long)n_nd - 1);
    const std::vector<double> &v = imag ? im : re;
    for (int64_t i = 0; i < n_nd; i++)
      for (int64_t q = ptr[i]; q < ptr[i + 1]; q++)
        if (!imag || v[q] != 0.0) std::fprintf(f, "%lld %d %.14e\n", (long long)i, col[q], v[q]);
    std::fclose(f);
  }

  // GetFieldAverages (maxwell_bloch.cpp:1550-1632); i counts REAL modes: mode 2b is the complex band b,
  // mode 2b+1 its partner i*E, i.e. (Er, Ei) -> (-Ei, Er), and likewise for B, D, H
  void GetFieldAverages(unsigned int i, std::vector<double> &Er, std::vector<double> &Ei, std::vector<double> &Br,
                        std::vector<double> &Bi, std::vector<double> &Dr, std::vector<double> &Di,
                        std::vector<double> &Hr, std::vector<double> &Hi) {
    double o[24];
    check(bloch_get_field_averages(h_, (int)(i / 2), o), "GetFieldAverages");
    std::vector<double> *re[4] = {&Er, &Br, &Dr, &Hr}, *im[4] = {&Ei, &Bi, &Di, &Hi};
    for (int f = 0; f < 4; f++) {
      re[f]->assign(3, 0.0);
      im[f]->assign(3, 0.0);
      for (int k = 0; k < 3; k++) {
        const double a = o[6 * f + k], b = o[6 * f + 3 + k];
        (*re[f])[k] = (i % 2) ? -b : a;
        (*im[f])[k] = (i % 2) ? a : b;
      }
    }
  }

  // reduced-basis sweep pieces of MaxwellDispersion (meta-material/meta_material_solver.cpp:3132-3305)
  void ReducedBasisClear() { check(bloch_rb_clear(h_), "ReducedBasisClear"); }
  void ReducedBasisAppend() { check(bloch_rb_append(h_), "ReducedBasisAppend"); }   // bands of the last Solve()
  int ReducedBasisSize() const { return bloch_rb_size(h_); }
  // approxEigenfrequencies(omega): nev values omega = sqrt|lambda|, every complex band twice
  void ApproxEigenfrequencies(const std::vector<double> &kappa, std::vector<double> &omega) {
    const int nb = (nev_ + 1) / 2;
    std::vector<double> lam(nb);
    check(bloch_rb_approx(h_, kappa.data(), lam.data(), nb), "ApproxEigenfrequencies");
    omega.resize(nev_);
    for (int i = 0; i < nev_; i++) omega[i] = std::sqrt(std::fabs(lam[i / 2]));
  }

  // nev values, every complex band twice like the reference's real 2N form (maxwell_bloch.cpp:1052-1076)
  void GetEigenvalues(std::vector<double> &eigenvalues) {
    const int nb = (nev_ + 1) / 2;
    std::vector<double> lam(nb);
    check(bloch_get_eigenvalues(h_, lam.data(), nb), "GetEigenvalues");
    eigenvalues.resize(nev_);
    for (int i = 0; i < nev_; i++) eigenvalues[i] = lam[i / 2];
  }
  // the convenience form (maxwell_bloch.cpp:1078-1100)
  void GetEigenvalues(int nev, const std::vector<double> &kappa, const std::vector<double> *init_vecs,
                      std::vector<double> &eigenvalues) {
    SetNumEigs(nev);
    SetKappa(kappa);
    Setup();
    if (init_vecs && !init_vecs->empty()) SetInitialVectors((int)(init_vecs->size() / (2 * N_)), init_vecs->data());
    Solve();
    GetEigenvalues(eigenvalues);
  }
  // real mode i of the reference's numbering: complex band i/2; odd i is the band times the
  // imaginary unit, (Er, Ei) -> (-Ei, Er), exactly the pairing the real block form produces
  void GetEigenvectorE(unsigned i, std::vector<double> &Er, std::vector<double> &Ei) {
    Er.resize(N_); Ei.resize(N_);
    check(bloch_get_eigenvector_E(h_, i / 2, Er.data(), Ei.data()), "GetEigenvectorE");
    if (i & 1) { for (int64_t k = 0; k < N_; k++) { double t = Er[k]; Er[k] = -Ei[k]; Ei[k] = t; } }
  }
  void GetEigenvectorB(unsigned i, std::vector<double> &Br, std::vector<double> &Bi) {
    Br.resize(Nrt_); Bi.resize(Nrt_);
    check(bloch_get_eigenvector_B(h_, i / 2, Br.data(), Bi.data()), "GetEigenvectorB");
    if (i & 1) { for (int64_t k = 0; k < Nrt_; k++) { double t = Br[k]; Br[k] = -Bi[k]; Bi[k] = t; } }
  }
  // GetEigenvector (maxwell_bloch.hpp:187-190): all four parts of real mode i
  void GetEigenvector(unsigned i, std::vector<double> &Er, std::vector<double> &Ei, std::vector<double> &Br,
                      std::vector<double> &Bi) {
    GetEigenvectorE(i, Er, Ei);
    GetEigenvectorB(i, Br, Bi);
  }
  // DetermineBasis (maxwell_bloch.cpp:1700-1727): right-handed orthonormal frame with e[2] = kappa / |kappa|,
  // e[1] = the part of v1 orthogonal to kappa, e[0] = e[1] x e[2]; the Cartesian frame when |kappa| < 1e-4.  The
  // reference's body subtracts (e2.v1) v1 instead of (e2.v1) e2 and assigns e[0][0] twice (e[0][1] stays unset);
  // this builds the frame those lines are evidently meant to build.
  void DetermineBasis(const std::vector<double> &v1, std::vector<std::vector<double>> &e) const {
    e.assign(3, std::vector<double>(3, 0.0));
    const double kn = kappa_.size() == 3 ? std::sqrt(kappa_[0] * kappa_[0] + kappa_[1] * kappa_[1] + kappa_[2] * kappa_[2]) : 0.0;
    if (kn < 1.0e-4) { for (int i = 0; i < 3; i++) e[i][i] = 1.0; return; }
    for (int i = 0; i < 3; i++) e[2][i] = kappa_[i] / kn;
    const double d = e[2][0] * v1[0] + e[2][1] * v1[1] + e[2][2] * v1[2];
    double nrm = 0;
    for (int i = 0; i < 3; i++) { e[1][i] = v1[i] - d * e[2][i]; nrm += e[1][i] * e[1][i]; }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < 3; i++) e[1][i] /= nrm;
    e[0][0] = e[1][1] * e[2][2] - e[1][2] * e[2][1];
    e[0][1] = e[1][2] * e[2][0] - e[1][0] * e[2][2];
    e[0][2] = e[1][0] * e[2][1] - e[1][1] * e[2][0];
  }
  // empty in the reference as well (maxwell_bloch.cpp:1694-1698)
  void ComputeHomogenizedCoefs() {}

  // WriteVisitFields (maxwell_bloch.cpp:1730-1822): E_r, E_i, B_r, B_i of every mode, cycle = mode number, time =
  // omega.  The reference writes an MFEM VisIt data collection; here one legacy-VTK unstructured grid per mode,
  // <prefix>/<label>_<cycle %06d>.vtk (fields evaluated at the element corners with the covariant / contravariant
  // Piola maps; per-element corner copies, so the discontinuous parts of the FE fields survive) plus
  // <prefix>/<label>.visit listing the files and <label>.times with omega - VisIt and ParaView open these directly.
  void WriteVisitFields(const std::string &prefix, const std::string &label) {
    const int p = order_, q = p + 1;
    std::vector<double> g(p), l(q), x0, J;
    std::vector<int> cls;
    points01(p, g, l);
    geometry(x0, cls, J);
    static const double ref[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    auto lagr = [](const std::vector<double> &nodes, double x, std::vector<double> &v) {
      v.assign(nodes.size(), 1.0);
      for (size_t j = 0; j < nodes.size(); j++)
        for (size_t k = 0; k < nodes.size(); k++)
          if (k != j) v[j] *= (x - nodes[k]) / (nodes[j] - nodes[k]);
    };
    // 1-D basis values at the corner coordinates 0 and 1
    std::vector<double> O[2], Cl[2];
    for (int a = 0; a < 2; a++) { lagr(g, (double)a, O[a]); lagr(l, (double)a, Cl[a]); }
    const int Lnd = bloch_local_size(h_, 1), Lrt = bloch_local_size(h_, 2);
    std::vector<int32_t> mnd((size_t)n_elem_ * Lnd), mrt((size_t)n_elem_ * Lrt);
    check(bloch_get_dofmap(h_, 1, mnd.data()), "bloch_get_dofmap");
    check(bloch_get_dofmap(h_, 2, mrt.data()), "bloch_get_dofmap");
    std::vector<double> ev;
    GetEigenvalues(ev);
    std::ofstream list(prefix + "/" + label + ".visit"), times(prefix + "/" + label + ".times");
    list << "!NBLOCKS 1\n";
    std::vector<double> F[4];
    for (int i = 0; i < nev_; i++) {
      GetEigenvector(i, F[0], F[1], F[2], F[3]);
      char name[64];
      std::snprintf(name, sizeof(name), "_%06d.vtk", i + 1);
      const std::string fn = label + name;
      std::FILE *f = std::fopen((prefix + "/" + fn).c_str(), "w");
      if (!f) throw Error(("cannot write " + prefix + "/" + fn).c_str());
      std::fprintf(f, "# vtk DataFile Version 3.0\nmfem-bravais_b200 Bloch fields\nASCII\nDATASET UNSTRUCTURED_GRID\n");
      std::fprintf(f, "POINTS %lld double\n", (long long)(8 * n_elem_));
      for (int64_t e = 0; e < n_elem_; e++) {
        const double *Je = &J[9 * cls[e]];
        for (int c = 0; c < 8; c++) {
          double x[3];
          for (int d = 0; d < 3; d++) x[d] = x0[3 * e + d] + Je[3 * d] * ref[c][0] + Je[3 * d + 1] * ref[c][1] + Je[3 * d + 2] * ref[c][2];
          std::fprintf(f, "%.12g %.12g %.12g\n", x[0], x[1], x[2]);
        }
      }
      std::fprintf(f, "CELLS %lld %lld\n", (long long)n_elem_, (long long)(9 * n_elem_));
      for (int64_t e = 0; e < n_elem_; e++) {
        std::fprintf(f, "8");
        for (int c = 0; c < 8; c++) std::fprintf(f, " %lld", (long long)(8 * e + c));
        std::fprintf(f, "\n");
      }
      std::fprintf(f, "CELL_TYPES %lld\n", (long long)n_elem_);
      for (int64_t e = 0; e < n_elem_; e++) std::fprintf(f, "12\n");
      std::fprintf(f, "POINT_DATA %lld\n", (long long)(8 * n_elem_));
      static const char *names[4] = {"E_r", "E_i", "B_r", "B_i"};
      for (int w = 0; w < 4; w++) {
        const bool nd = w < 2;
        const int L = nd ? Lnd : Lrt, nbc = nd ? p * q * q : p * p * q;
        const std::vector<int32_t> &map = nd ? mnd : mrt;
        std::fprintf(f, "VECTORS %s double\n", names[w]);
        for (int64_t e = 0; e < n_elem_; e++) {
          const double *Je = &J[9 * cls[e]];
          const double det = Je[0] * (Je[4] * Je[8] - Je[5] * Je[7]) - Je[1] * (Je[3] * Je[8] - Je[5] * Je[6]) + Je[2] * (Je[3] * Je[7] - Je[4] * Je[6]);
          double Ji[9];   // inverse of Je (row-major)
          Ji[0] = (Je[4] * Je[8] - Je[5] * Je[7]) / det; Ji[1] = (Je[2] * Je[7] - Je[1] * Je[8]) / det; Ji[2] = (Je[1] * Je[5] - Je[2] * Je[4]) / det;
          Ji[3] = (Je[5] * Je[6] - Je[3] * Je[8]) / det; Ji[4] = (Je[0] * Je[8] - Je[2] * Je[6]) / det; Ji[5] = (Je[2] * Je[3] - Je[0] * Je[5]) / det;
          Ji[6] = (Je[3] * Je[7] - Je[4] * Je[6]) / det; Ji[7] = (Je[1] * Je[6] - Je[0] * Je[7]) / det; Ji[8] = (Je[0] * Je[4] - Je[1] * Je[3]) / det;
          for (int c = 0; c < 8; c++) {
            const int a[3] = {(int)ref[c][0], (int)ref[c][1], (int)ref[c][2]};
            double vr[3];   // reference-space components
            for (int comp = 0; comp < 3; comp++) {
              // ND component: open along comp, closed otherwise; RT component: closed along comp, open otherwise
              int n[3];
              const std::vector<double> *B[3];
              for (int d = 0; d < 3; d++) {
                const bool open = nd ? d == comp : d != comp;
                n[d] = open ? p : q;
                B[d] = open ? &O[a[d]] : &Cl[a[d]];
              }
              double acc = 0;
              for (int k2 = 0; k2 < n[2]; k2++) for (int k1 = 0; k1 < n[1]; k1++) for (int k0 = 0; k0 < n[0]; k0++) {
                const double wgt = (*B[0])[k0] * (*B[1])[k1] * (*B[2])[k2];
                if (wgt == 0.0) continue;
                const int32_t sg = map[(size_t)e * L + comp * nbc + k0 + n[0] * (k1 + n[1] * k2)];
                const double val = F[w][(sg < 0 ? -sg : sg) - 1];
                acc += wgt * (sg < 0 ? -val : val);
              }
              vr[comp] = acc;
            }
            double v[3];
            for (int d = 0; d < 3; d++)
              v[d] = nd ? Ji[d] * vr[0] + Ji[3 + d] * vr[1] + Ji[6 + d] * vr[2]                       // J^-T v
                        : (Je[3 * d] * vr[0] + Je[3 * d + 1] * vr[1] + Je[3 * d + 2] * vr[2]) / det;   // J v / det
            std::fprintf(f, "%.12g %.12g %.12g\n", v[0], v[1], v[2]);
          }
        }
      }
      std::fclose(f);
      const double om = ev[i] > 0.0 ? std::sqrt(ev[i]) : (ev[i] > -1.0e-6 ? 0.0 : -1.0);
      list << fn << "\n";
      times << (i + 1) << " " << fn << " " << om << "\n";
    }
  }

  // GetAOperator()->Mult / GetMOperator()->Mult / GetSubSpaceProjector()->Mult on 2N vectors
  void MultA(const std::vector<double> &x, std::vector<double> &y) { y.resize(x.size()); check(bloch_apply_A(h_, x.data(), y.data(), (int)(x.size() / (2 * N_))), "MultA"); }
  void MultM(const std::vector<double> &x, std::vector<double> &y) { y.resize(x.size()); check(bloch_apply_M(h_, x.data(), y.data(), (int)(x.size() / (2 * N_))), "MultM"); }
  void MultProjector(const std::vector<double> &x, std::vector<double> &y) { y.resize(x.size()); check(bloch_apply_projector(h_, x.data(), y.data(), (int)(x.size() / (2 * N_))), "MultProjector"); }

  // IdentifyDegeneracies (maxwell_bloch.cpp:1493-1548): groups of eigenvalue indices closer than
  // max(zero_tol, rel_tol * |lambda|)
  void IdentifyDegeneracies(double zero_tol, double rel_tol, std::vector<std::set<int>> &degen) {
    std::vector<double> ev;
    GetEigenvalues(ev);
    degen.clear();
    for (int i = 0; i < (int)ev.size(); i++) {
      if (!degen.empty()) {
        const double prev = ev[*degen.back().rbegin()];
        const double tol = std::max(zero_tol, rel_tol * std::fabs(prev));
        if (std::fabs(ev[i] - prev) <= tol) { degen.back().insert(i); continue; }
      }
      degen.push_back(std::set<int>{i});
    }
  }
  void GetSolverStats(double &meanTime, double &stdDevTime, double &meanIter, double &stdDevIter, int &nSolves) {
    bloch_stats st;
    bloch_get_stats(h_, &st);
    times_.push_back(st.solve_seconds);
    iters_.push_back(st.iterations);
    nSolves = (int)times_.size();
    auto stat = [](const std::vector<double> &v, double &m, double &s) {
      m = 0; s = 0;
      for (double x : v) m += x;
      m /= v.size();
      for (double x : v) s += (x - m) * (x - m);
      s = std::sqrt(s / v.size());
    };
    stat(times_, meanTime, stdDevTime);
    stat(iters_, meanIter, stdDevIter);
  }

  // ---- interchange: the refined Wigner-Seitz cell as a NON-periodic `MFEM mesh v1.0` file (hexahedra in MFEM vertex
  // order, boundary = faces of a single element), `<path>.trans` = the lattice's translation vectors and `<path>.coef`
  // = `attribute eps muinv` for the distinct element-wise coefficient pairs: the input of the reference's
  // MakePeriodicMesh pipeline (lib/bravais.cpp:343-355, 9548-9832), so that an MFEM/hypre run elsewhere works on the
  // very same elements.  Same files as the Python mirror's write_mfem_mesh.
  void WriteMesh(const std::string &path, const BravaisLattice &lat, const std::vector<double> &eps,
                 const std::vector<double> &muinv) const {
    std::vector<double> x0, J;
    std::vector<int> cls;
    geometry(x0, cls, J);
    static const double ref[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int hexf[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
    std::map<std::pair<double, double>, int> pairs;
    std::vector<int> attr(n_elem_);
    for (int64_t e = 0; e < n_elem_; e++) {
      const std::pair<double, double> key(eps.empty() ? 1.0 : eps[e], muinv.empty() ? 1.0 : muinv[e]);
      auto it = pairs.find(key);
      if (it == pairs.end()) it = pairs.emplace(key, (int)pairs.size() + 1).first;
      attr[e] = it->second;
    }
    std::map<std::array<long long, 3>, int> vid;
    std::vector<std::array<double, 3>> verts;
    std::vector<std::array<int, 8>> hexes(n_elem_);
    for (int64_t e = 0; e < n_elem_; e++) {
      const double *Je = &J[9 * cls[e]];
      for (int l = 0; l < 8; l++) {
        std::array<double, 3> p;
        std::array<long long, 3> key;
        for (int i = 0; i < 3; i++) {
          p[i] = x0[3 * e + i] + Je[3 * i + 0] * ref[l][0] + Je[3 * i + 1] * ref[l][1] + Je[3 * i + 2] * ref[l][2];
          key[i] = std::llround(p[i] * 1e10);
        }
        auto it = vid.find(key);
        if (it == vid.end()) { it = vid.emplace(key, (int)verts.size()).first; verts.push_back(p); }
        hexes[e][l] = it->second;
      }
    }
    std::map<std::array<int, 4>, std::pair<int, std::array<int, 4>>> faces;
    for (int64_t e = 0; e < n_elem_; e++)
      for (auto &f : hexf) {
        std::array<int, 4> fv = {hexes[e][f[0]], hexes[e][f[1]], hexes[e][f[2]], hexes[e][f[3]]}, key = fv;
        std::sort(key.begin(), key.end());
        auto &slot = faces[key];
        if (slot.first++ == 0) slot.second = fv;
      }
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("WriteMesh: cannot open " + path);
    size_t nb = 0;
    for (auto &kv : faces) nb += kv.second.first == 1;
    std::fprintf(f, "MFEM mesh v1.0\n\n#\n# Wigner-Seitz cell written by bloch_b200 (non-periodic; translation vectors in the .trans file)\n"
                    "# MFEM geometry types: SQUARE = 3, CUBE = 5\n#\n\ndimension\n3\n\nelements\n%lld\n", (long long)n_elem_);
    for (int64_t e = 0; e < n_elem_; e++) {
      std::fprintf(f, "%d 5", attr[e]);
      for (int l = 0; l < 8; l++) std::fprintf(f, " %d", hexes[e][l]);
      std::fprintf(f, "\n");
    }
    std::fprintf(f, "\nboundary\n%zu\n", nb);
    for (auto &kv : faces)
      if (kv.second.first == 1) { auto &q = kv.second.second; std::fprintf(f, "1 3 %d %d %d %d\n", q[0], q[1], q[2], q[3]); }
    std::fprintf(f, "\nvertices\n%zu\n3\n", verts.size());
    for (auto &v : verts) std::fprintf(f, "%.17g %.17g %.17g\n", v[0] + 0.0, v[1] + 0.0, v[2] + 0.0);
    std::fclose(f);
    f = std::fopen((path + ".coef").c_str(), "w");
    std::vector<std::pair<int, std::pair<double, double>>> rows;
    for (auto &kv : pairs) rows.push_back({kv.second, kv.first});
    std::sort(rows.begin(), rows.end());
    for (auto &r : rows) std::fprintf(f, "%d %.17g %.17g\n", r.first, r.second.first, r.second.second);
    std::fclose(f);
    std::vector<std::vector<double>> t;
    lat.GetTranslationVectors(t);
    f = std::fopen((path + ".trans").c_str(), "w");
    for (auto &v : t) std::fprintf(f, "%.17g %.17g %.17g\n", v[0], v[1], v[2]);
    std::fclose(f);
  }

  // ---- CreateInitialVectors (maxwell/maxwell_dispersion.cpp:735-1060): plane waves E0 exp(i 2 pi sum_j n_j b_j . x)
  // for the lattice's table of shifts n, E0 = two unit vectors orthogonal to k (three Cartesian ones when |k| < 1e-2),
  // nodally interpolated into ND (dof = edge vector . field at the ND node).  One complex vector [re(N); im(N)] per
  // (n, E0) - the reference's second vector of each pair is i times the first -, ordered by |kappa + 2 pi G|, at most
  // `count` (<= 0: all).  literal keeps the reference's k = kappa + sum_j n_j b_j (no 2 pi on b, sic).
  void CreateInitialVectors(const BravaisLattice &lat, const std::vector<double> &kappa, std::vector<double> &vecs,
                            int &num_vecs, int count = 0, bool literal = true) const {
    std::vector<std::vector<double>> b;
    lat.GetReciprocalLatticeVectors(b);
    const std::string label = lat.GetLatticeTypeLabel();
    std::vector<std::array<int, 3>> table = {{0, 0, 0}};
    if (label == "FCC") {
      for (int a : {1, -1}) for (int c : {1, -1}) for (int d : {1, -1}) table.push_back({a, c, d});
    } else if (label == "BCC") {
      for (int c : {1, -1}) for (int d : {1, -1}) table.push_back({0, c, d});
      for (int a : {1, -1}) for (int d : {1, -1}) table.push_back({a, 0, d});
      for (int a : {1, -1}) for (int c : {1, -1}) table.push_back({a, c, 0});
    } else {
      for (int d = 0; d < 3; d++) for (int sgn : {1, -1}) { std::array<int, 3> n = {0, 0, 0}; n[d] = sgn; table.push_back(n); }
    }
    struct Wave { double key; double G[3], e0[3]; };
    std::vector<Wave> waves;
    const double two_pi = 2.0 * M_PI;
    for (auto &n : table) {
      double G[3], k[3], kt[3];
      for (int d = 0; d < 3; d++) {
        G[d] = n[0] * b[0][d] + n[1] * b[1][d] + n[2] * b[2][d];
        kt[d] = kappa[d] + two_pi * G[d];
        k[d] = literal ? kappa[d] + G[d] : kt[d];
      }
      const double nk = std::sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
      const double key = std::sqrt(kt[0] * kt[0] + kt[1] * kt[1] + kt[2] * kt[2]);
      auto push = [&](const double *e) { Wave w; w.key = key; for (int d = 0; d < 3; d++) { w.G[d] = G[d]; w.e0[d] = e[d]; } waves.push_back(w); };
      if (nk < 1e-2) {
        const double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
        push(ex); push(ey); push(ez);
      } else {
        int least = 0;
        for (int d = 1; d < 3; d++) if (std::fabs(k[d]) < std::fabs(k[least])) least = d;
        double u[3] = {0, 0, 0}, v[3];
        u[least] = 1.0;
        const double uk = k[least] / (nk * nk);
        double nu = 0;
        for (int d = 0; d < 3; d++) { u[d] -= uk * k[d]; nu += u[d] * u[d]; }
        nu = std::sqrt(nu);
        for (int d = 0; d < 3; d++) u[d] /= nu;
        v[0] = (k[1] * u[2] - k[2] * u[1]) / nk; v[1] = (k[2] * u[0] - k[0] * u[2]) / nk; v[2] = (k[0] * u[1] - k[1] * u[0]) / nk;
        push(u); push(v);
      }
    }
    std::stable_sort(waves.begin(), waves.end(), [](const Wave &a, const Wave &c) { return a.key < c.key; });
    if (count > 0 && (int)waves.size() > count) waves.resize(count);
    num_vecs = (int)waves.size();
    // ND nodes: open direction on the Gauss-Legendre points, closed directions on the Gauss-Lobatto points of [0,1]
    const int p = order_, q = p + 1, nb = p * q * q;
    std::vector<double> g(p), l(q);
    points01(p, g, l);
    std::vector<double> x0, J;
    std::vector<int> cls;
    geometry(x0, cls, J);
    const int L = bloch_local_size(h_, 1);
    std::vector<int32_t> map((size_t)n_elem_ * L);
    check(bloch_get_dofmap(h_, 1, map.data()), "bloch_get_dofmap");
    vecs.assign((size_t)num_vecs * 2 * N_, 0.0);
    for (int w = 0; w < num_vecs; w++) {
      double *re = &vecs[(size_t)w * 2 * N_], *im = re + N_;
      for (int64_t e = 0; e < n_elem_; e++) {
        const double *Je = &J[9 * cls[e]];
        for (int c = 0; c < 3; c++) {
          const int n0 = c == 0 ? p : q, n1 = c == 1 ? p : q, n2 = c == 2 ? p : q;
          const double t[3] = {Je[c], Je[3 + c], Je[6 + c]};                 // physical edge vector of direction c
          const double amp = t[0] * waves[w].e0[0] + t[1] * waves[w].e0[1] + t[2] * waves[w].e0[2];
          for (int k2 = 0; k2 < n2; k2++) for (int k1 = 0; k1 < n1; k1++) for (int k0 = 0; k0 < n0; k0++) {
            const double xi[3] = {c == 0 ? g[k0] : l[k0], c == 1 ? g[k1] : l[k1], c == 2 ? g[k2] : l[k2]};
            double ph = 0;
            for (int i = 0; i < 3; i++) {
              const double x = x0[3 * e + i] + Je[3 * i] * xi[0] + Je[3 * i + 1] * xi[1] + Je[3 * i + 2] * xi[2];
              ph += waves[w].G[i] * x;
            }
            const int32_t s = map[(size_t)e * L + c * nb + k0 + n0 * (k1 + n1 * k2)];
            const int64_t gid = (s < 0 ? -s : s) - 1;
            const double sg = s < 0 ? -1.0 : 1.0;
            re[gid] = sg * amp * std::cos(two_pi * ph);
            im[gid] = sg * amp * std::sin(two_pi * ph);
          }
        }
      }
    }
  }
  bloch_handle handle() const { return h_; }

private:
  void geometry(std::vector<double> &x0, std::vector<int> &cls, std::vector<double> &J) const {
    int64_t ne;
    int nc;
    bloch_num_elements(h_, &ne, &nc);
    x0.resize(3 * ne); cls.resize(ne); J.resize(9 * (size_t)nc);
    check(bloch_element_geometry(h_, x0.data(), cls.data(), J.data()), "bloch_element_geometry");
  }
  // p Gauss-Legendre and p+1 Gauss-Lobatto points on [0,1] (Newton on the Legendre polynomials)
  static void points01(int p, std::vector<double> &g, std::vector<double> &l) {
    auto leg = [](int n, double x, double &P, double &dP) {
      double p0 = 1.0, p1 = x;
      if (n == 0) { P = 1; dP = 0; return; }
      for (int k = 2; k <= n; k++) { const double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = pk; }
      P = p1; dP = n * (x * p1 - p0) / (x * x - 1.0);
    };
    for (int i = 0; i < p; i++) {
      double z = -std::cos(M_PI * (i + 0.75) / (p + 0.5)), P, dP;
      for (int it = 0; it < 100; it++) { leg(p, z, P, dP); const double dz = P / dP; z -= dz; if (std::fabs(dz) < 1e-16) break; }
      g[i] = 0.5 * (z + 1.0);
    }
    l[0] = 0.0; l[p] = 1.0;
    for (int i = 1; i < p; i++) {
      double z = -std::cos(M_PI * i / p), P, dP;
      for (int it = 0; it < 100; it++) {
        leg(p, z, P, dP);
        const double ddP = (2.0 * z * dP - p * (p + 1) * P) / (1.0 - z * z);
        const double dz = dP / ddP;
        z -= dz;
        if (std::fabs(dz) < 1e-16) break;
      }
      l[i] = 0.5 * (z + 1.0);
    }
  }
  void push_beta_zeta() {
    if (zeta_.size() == 3) {
      double k[3] = {beta_ * zeta_[0], beta_ * zeta_[1], beta_ * zeta_[2]};
      check(bloch_set_kappa(h_, k), "SetBeta/SetZeta");
    }
  }
  bloch_handle h_ = nullptr;
  int64_t n_elem_ = 0, N_ = 0, Nrt_ = 0, Nh1_ = 0;
  int order_ = 1;
  int nev_ = 20;
  double beta_ = 0;
  std::vector<double> zeta_, kappa_;
  std::vector<double> times_, iters_;
};

// Mirror of ScalarFloquetWaveEquation (misc/scalar3d.cpp:99-162, 592-898): the scalar H1 Bloch Helmholtz problem
//     (G - i Z_kappa)^T M1(k) (G - i Z_kappa) u = lambda M0(m) u,      kappa = beta * pi / 180 * zeta,
// with the phase shift beta in DEGREES and the direction zeta from azimuth / inclination in degrees (:665-669),
// element-wise constant coefficients k (stiffness) and m (mass).  The reference hands its operators to HypreLOBPCG
// in the driver (:427-444, preconditioner = BoomerAMG sweeps, :70-85); here Solve() runs the library's block LOBPCG
// with a geometric-multigrid V-cycle as preconditioner.  nev counts REAL modes (2 per complex mode).
class ScalarFloquetWaveEquation {
public:
  ScalarFloquetWaveEquation(const BravaisLattice &lat, int n_sub, int order, int device = -1) {
    if (bloch_create(&h_, lat.handle(), n_sub, order, device) != BLOCH_OK) throw Error("bloch_create");
    int64_t ne, n, nrt, nh1;
    int nc;
    bloch_num_elements(h_, &ne, &nc);
    bloch_num_dofs(h_, &n, &nrt, &nh1);
    n_elem_ = ne; N_ = nh1;
    k_.assign((size_t)ne, 1.0);
    m_.assign((size_t)ne, 1.0);
  }
  ~ScalarFloquetWaveEquation() { bloch_destroy(h_); }
  ScalarFloquetWaveEquation(const ScalarFloquetWaveEquation &) = delete;
  ScalarFloquetWaveEquation &operator=(const ScalarFloquetWaveEquation &) = delete;

  int64_t GetH1TrueVSize() const { return N_; }          // GetFESpace()->GlobalTrueVSize()
  int64_t GetNE() const { return n_elem_; }
  void GetElementCenters(std::vector<double> &xyz) const {
    xyz.resize(3 * (size_t)n_elem_);
    check(bloch_element_centers(h_, xyz.data()), "bloch_element_centers");
  }
  void SetBeta(double beta_degrees) { beta_ = beta_degrees; explicit_kappa_ = false; }
  void SetAzimuth(double alpha_a) { alpha_a_ = alpha_a; explicit_kappa_ = false; }
  void SetInclination(double alpha_i) { alpha_i_ = alpha_i; explicit_kappa_ = false; }
  // convenience beyond the reference: the Bloch vector itself
  void SetKappa(const std::vector<double> &kappa) { kappa_ = kappa; explicit_kappa_ = true; }
  void GetZeta(std::vector<double> &zeta) const {        // scalar3d.cpp:665-669
    const double d = M_PI / 180.0;
    zeta = {std::cos(alpha_i_ * d) * std::cos(alpha_a_ * d), std::cos(alpha_i_ * d) * std::sin(alpha_a_ * d), std::sin(alpha_i_ * d)};
  }
  void GetKappa(std::vector<double> &kappa) const {
    if (explicit_kappa_) { kappa = kappa_; return; }
    GetZeta(kappa);
    for (double &x : kappa) x *= beta_ * M_PI / 180.0;   // :733 (beta^2 pi^2 / 32400 = (beta pi / 180)^2), :784-785
  }
  void SetMassCoef(const std::vector<double> &m_per_elem) { m_ = m_per_elem; }
  void SetStiffnessCoef(const std::vector<double> &k_per_elem) { k_ = k_per_elem; }
  void SetNumEigs(int nev) { nev_ = nev; }
  void SetAbsoluteTolerance(double atol, int max_iter = 1000) { atol_ = atol; max_iter_ = max_iter; }
  void Setup() {
    if ((int64_t)k_.size() != n_elem_ || (int64_t)m_.size() != n_elem_) throw std::invalid_argument("coefficient arrays must have one entry per element");
    check(bloch_scalar_set_coefs(h_, k_.data(), m_.data()), "bloch_scalar_set_coefs");
    check(bloch_scalar_set_num_modes(h_, (nev_ + 1) / 2), "bloch_scalar_set_num_modes");
    check(bloch_set_tol(h_, atol_, max_iter_), "bloch_set_tol");
    std::vector<double> kappa;
    GetKappa(kappa);
    check(bloch_set_kappa(h_, kappa.data()), "bloch_set_kappa");
    check(bloch_setup(h_), "Setup");
  }
  void Solve() { check(bloch_scalar_solve(h_), "Solve"); }
  void GetEigenvalues(std::vector<double> &eigenvalues) {   // nev values, each complex mode twice
    const int nb = (nev_ + 1) / 2;
    std::vector<double> lam(nb);
    check(bloch_scalar_get_eigenvalues(h_, lam.data(), nb), "GetEigenvalues");
    eigenvalues.resize(nev_);
    for (int i = 0; i < nev_; i++) eigenvalues[i] = lam[i / 2];
  }
  // GetAOperator()->Mult / GetMOperator()->Mult on [re; im] vectors of length 2 N
  void MultA(const std::vector<double> &x, std::vector<double> &y) { y.resize(x.size()); check(bloch_scalar_apply(h_, 0, x.data(), y.data(), 1), "MultA"); }
  void MultM(const std::vector<double> &x, std::vector<double> &y) { y.resize(x.size()); check(bloch_scalar_apply(h_, 1, x.data(), y.data(), 1), "MultM"); }
  void GetSolverStats(int &iterations, double &seconds) {
    bloch_stats st;
    bloch_get_stats(h_, &st);
    iterations = st.iterations;
    seconds = st.solve_seconds;
  }

private:
  bloch_handle h_ = nullptr;
  int64_t n_elem_ = 0, N_ = 0;
  int nev_ = 5, max_iter_ = 1000;                  // scalar3d.cpp:291, 431
  double beta_ = 1.0, alpha_a_ = 0.0, alpha_i_ = 90.0, atol_ = 1e-6;   // class defaults of :593-595, tolerance :432
  bool explicit_kappa_ = false;
  std::vector<double> kappa_, k_, m_;
};

}  // namespace bloch_b200
