// C++ host-side mirror of mfem::bloch::MaxwellBlochWaveEquation (maxwell/maxwell_bloch.hpp:140-234)
// over the C ABI of bloch_b200.h.  Same member names and argument meaning as the reference; MFEM
// types are flattened: Vector -> std::vector<double>, Coefficient -> one value per element
// (the reference samples its coefficients into an L2 order-0 grid function anyway,
// maxwell/maxwell_dispersion.cpp:398-420), HypreParVector of length 2N -> [re(N); im(N)].
// Header-only; link with -lbloch_b200.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
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
  MaxwellBlochWaveEquation(const BravaisLattice &lat, int n_sub, int order, int device = -1) {
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

  void SetKappa(const std::vector<double> &kappa) { check(bloch_set_kappa(h_, kappa.data()), "SetKappa"); }
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
  bloch_handle handle() const { return h_; }

private:
  void push_beta_zeta() {
    if (zeta_.size() == 3) {
      double k[3] = {beta_ * zeta_[0], beta_ * zeta_[1], beta_ * zeta_[2]};
      check(bloch_set_kappa(h_, k), "SetBeta/SetZeta");
    }
  }
  bloch_handle h_ = nullptr;
  int64_t n_elem_ = 0, N_ = 0, Nrt_ = 0, Nh1_ = 0;
  int nev_ = 20;
  double beta_ = 0;
  std::vector<double> zeta_;
  std::vector<double> times_, iters_;
};

}  // namespace bloch_b200
