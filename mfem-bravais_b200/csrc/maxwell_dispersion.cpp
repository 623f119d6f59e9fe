// maxwell_dispersion-shaped C++ driver (reference: maxwell/maxwell_dispersion.cpp:114-699) over
// include/maxwell_bloch_b200.hpp: same flag names (-bl -o -sr -pr -p -a -np), same k-path walk
// with the symmetry-point cache (:475-648), same disp.dat format (:1062-1087).
// Extra: -nb <complex bands> (the reference derives nev from its plane-wave initial guess),
// -dev <cuda device>, -out <directory>.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../../include/maxwell_bloch_b200.hpp"

using namespace bloch_b200;

static int prob_ = 2;

// mass_coef (maxwell_dispersion.cpp:1449-1551), cases on the dispersion path
static double mass_coef(const double *x) {
  const double eps1 = 10.0, eps2 = 100.0;
  const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  switch (prob_) {
    case 0: if (std::fabs(x[0]) <= 0.5) return eps1; break;                       // slab
    case 1: if (std::sqrt(x[0] * x[0] + x[1] * x[1]) <= 0.5) return eps1; break;  // cylinder
    case 2: if (r <= 0.25) return eps1; break;                                    // sphere
    case 3: {                                                                      // shell + 3 rods
      const double r1 = 0.14, r2 = 0.36, r3 = 0.105, eps3 = 12.96;
      if (r <= r1) return 1.0;
      if (r <= r2) return eps3;
      if (std::sqrt(x[1] * x[1] + x[2] * x[2]) <= r3) return eps3;
      if (std::sqrt(x[2] * x[2] + x[0] * x[0]) <= r3) return eps3;
      if (std::sqrt(x[0] * x[0] + x[1] * x[1]) <= r3) return eps3;
      break;
    }
    case 7: return (std::fabs(x[0]) <= 0.25 && std::fabs(x[1]) <= 0.25) ? 1.0 : 13.0;
    case 8: if (std::fabs(x[0]) < 0.1 || std::fabs(x[1]) < 0.1 || std::fabs(x[2]) < 0.1) return eps2; break;
    default: return 1.0;
  }
  return 1.0;
}

static void WriteDispersionData(std::ostream &os, int c, const std::string &label,
                                const std::vector<double> &eigenvalues) {
  os << c << "\t" << label;
  for (double ev : eigenvalues) {
    if (ev > 0.0) os << "\t" << std::sqrt(ev);
    else if (ev > -1.0e-6) os << "\t" << 0.0;
    else os << "\t" << -1.0;
  }
  os << std::endl;
}

int main(int argc, char **argv) {
  int bl_type = 1, order = 1, sr = 0, pr = 2, np = 0, nb = 10, dev = -1, kb = 1;
  bool write_mats = false, write_mesh = false, plane_wave_init = false, visit = false;
  double a = -1.0;
  std::string out = ".";
  for (int i = 1; i < argc; i++) {
    auto next = [&](const char *f) -> const char * {
      if (i + 1 >= argc) { std::cerr << "missing value for " << f << std::endl; std::exit(1); }
      return argv[++i];
    };
    std::string f = argv[i];
    if (f == "-bl") bl_type = std::atoi(next("-bl"));
    else if (f == "-o") order = std::atoi(next("-o"));
    else if (f == "-sr") sr = std::atoi(next("-sr"));
    else if (f == "-pr") pr = std::atoi(next("-pr"));
    else if (f == "-p") prob_ = std::atoi(next("-p"));
    else if (f == "-a") a = std::atof(next("-a"));
    else if (f == "-np") np = std::atoi(next("-np"));
    else if (f == "-nb") nb = std::atoi(next("-nb"));
    else if (f == "-dev") dev = std::atoi(next("-dev"));
    else if (f == "-kb" || f == "--k-batch") kb = std::atoi(next("-kb"));   // k-points iterated together (1 = the reference's loop)
    else if (f == "-out") out = next("-out");
    else if (f == "-wm" || f == "--write-mats") write_mats = true;
    else if (f == "-wmesh" || f == "--write-mesh") write_mesh = true;      // ws-cell.mesh (+ .trans, .coef) for MFEM cross-checks
    else if (f == "-iv" || f == "--plane-wave-init") plane_wave_init = true; // CreateInitialVectors block per k-point (:531)
    else if (f == "-visit" || f == "--visit") visit = true;   // field files of the symmetry points (WriteVisitFields, :546-550, 641)
    else if (f == "-no-vis" || f == "-no-visit" || f == "-no-wm" || f == "-mp" || f == "-no-mp") {}
    else { std::cerr << "unknown option " << f << std::endl; return 1; }
  }
  try {
    // the reference maps bl_type + 5 onto the enum of its absent ../common/bravais.hpp
    // (maxwell_dispersion.cpp:290-294); against the enum of lib/bravais.hpp:24-50, which this
    // library mirrors, -bl 1/2/3 = CUB/FCC/BCC is bl_type + 6
    BravaisLattice bravais(bl_type + 6, a);
    const int n_sub = 1 << (sr + pr);
    MaxwellBlochWaveEquation eq(bravais, n_sub, order, dev);
    std::cout << "Lattice " << bravais.GetLatticeTypeLabel() << ", n_sub " << n_sub << ", order " << order
              << ", H(curl) unknowns " << eq.GetHCurlTrueVSize() << std::endl;
    std::vector<double> xyz, eps(eq.GetNE()), mu(eq.GetNE(), 1.0);
    eq.GetElementCenters(xyz);
    for (int64_t e = 0; e < eq.GetNE(); e++) eps[e] = mass_coef(&xyz[3 * e]);
    eq.SetMassCoef(eps);
    eq.SetStiffnessCoef(mu);
    eq.SetAbsoluteTolerance(1e-6);
    if (write_mesh) eq.WriteMesh(out + "/ws-cell.mesh", bravais, eps, mu);
    // the reference rebuilds its plane-wave block for every kappa (CreateInitialVectors, :735-1060) and hands it to
    // GetEigenvalues; the default here is the solver's own guess (seeded random + warm start from the previous point)
    std::vector<double> init;
    auto solve = [&](const std::vector<double> &kap, std::vector<double> &ev) {
      if (plane_wave_init) {
        int nv = 0;
        eq.CreateInitialVectors(bravais, kap, init, nv);
        eq.GetEigenvalues(2 * nb, kap, &init, ev);
      } else {
        eq.GetEigenvalues(2 * nb, kap, nullptr, ev);
      }
    };

    if (kb > 1) {
      // Batched walk: the same rows in the same order, but the unique k-points (symmetry points cached by label,
      // :506, 604-614) are collected first and solved kb at a time with GetEigenvaluesBatch (independent
      // eigenproblems, one set of kernel launches); disp.dat is identical to the sequential walk's.
      struct Row { std::string label; int unique; unsigned path; };
      std::vector<Row> rows;
      std::vector<std::vector<double>> ukappa;
      std::map<std::string, int> by_label;
      auto add = [&](const std::string &label, const std::vector<double> &kap, unsigned p, bool cache) {
        int u;
        if (cache && by_label.count(label)) u = by_label[label];
        else { u = (int)ukappa.size(); ukappa.push_back(kap); if (cache) by_label[label] = u; }
        rows.push_back({label, u, p});
      };
      for (unsigned p = 0; p < bravais.GetNumberPaths(); p++)
        for (unsigned s = 0; s < bravais.GetNumberPathSegments(p); s++) {
          int e0, e1;
          bravais.GetPathSegmentEndPointIndices(p, s, e0, e1);
          std::vector<double> kappa0, kappa1, kappa(3);
          bravais.GetSymmetryPoint(e0, kappa0);
          bravais.GetSymmetryPoint(e1, kappa1);
          for (int i = 0; i <= np; i++) {
            for (int d = 0; d < 3; d++)
              kappa[d] = double(np + 1 - i) / (np + 1) * kappa0[d] + double(i) / (np + 1) * kappa1[d];
            std::string label = "-";
            if (i == 0) label = bravais.GetSymmetryPointLabel(e0);
            else if (np % 2 == 1 && i == (np + 1) / 2) label = bravais.GetIntermediatePointLabel(p, s);
            add(label, kappa, p, i == 0);
          }
          if (s + 1 == bravais.GetNumberPathSegments(p)) add(bravais.GetSymmetryPointLabel(e1), kappa1, p, true);
        }
      std::vector<std::vector<double>> uev(ukappa.size());
      for (size_t u0 = 0; u0 < ukappa.size(); u0 += kb) {
        const size_t u1 = std::min(ukappa.size(), u0 + (size_t)kb);
        std::vector<double> flat;
        for (size_t u = u0; u < u1; u++) flat.insert(flat.end(), ukappa[u].begin(), ukappa[u].end());
        std::vector<std::vector<double>> ev;
        eq.GetEigenvaluesBatch(2 * nb, flat, ev);
        for (size_t u = u0; u < u1; u++) uev[u] = ev[u - u0];
      }
      std::ofstream ofs_disp(out + "/disp.dat");
      int c = 0;
      for (size_t r = 0; r < rows.size(); r++) {
        if (r > 0 && rows[r].path != rows[r - 1].path) ofs_disp << std::endl;
        WriteDispersionData(ofs_disp, c++, rows[r].label, uev[rows[r].unique]);
      }
      ofs_disp << std::endl;
    } else {
      std::ofstream ofs_disp(out + "/disp.dat");
      std::map<std::string, std::vector<double>> sp_eigs;   // symmetry-point cache (:506, 604-614)
      int c = 0;
      for (unsigned p = 0; p < bravais.GetNumberPaths(); p++) {
        for (unsigned s = 0; s < bravais.GetNumberPathSegments(p); s++) {
          int e0, e1;
          bravais.GetPathSegmentEndPointIndices(p, s, e0, e1);
          std::vector<double> kappa0, kappa1, kappa(3), eigenvalues;
          bravais.GetSymmetryPoint(e0, kappa0);
          bravais.GetSymmetryPoint(e1, kappa1);
          for (int i = 0; i <= np; i++) {
            for (int d = 0; d < 3; d++)
              kappa[d] = double(np + 1 - i) / (np + 1) * kappa0[d] + double(i) / (np + 1) * kappa1[d];
            std::string label = "-";
            if (i == 0) label = bravais.GetSymmetryPointLabel(e0);
            else if (np % 2 == 1 && i == (np + 1) / 2) label = bravais.GetIntermediatePointLabel(p, s);
            if (i == 0 && sp_eigs.count(label)) {
              eigenvalues = sp_eigs[label];
            } else {
              solve(kappa, eigenvalues);
              if (i == 0) sp_eigs[label] = eigenvalues;
              if (visit && label != "-") eq.WriteVisitFields(out, "Maxwell-Dispersion-" + label);
              if (write_mats && label != "-") {            // Ar / Ai / M dump (:553-590), hypre IJ text format
                eq.WriteMatrix(0, false, out + "/Ar" + label + ".mat");
                eq.WriteMatrix(0, true, out + "/Ai" + label + ".mat");
                eq.WriteMatrix(1, false, out + "/M" + label + ".mat");
              }
            }
            WriteDispersionData(ofs_disp, c++, label, eigenvalues);
          }
          if (s + 1 == bravais.GetNumberPathSegments(p)) {   // close the path at its last symmetry point
            std::string label = bravais.GetSymmetryPointLabel(e1);
            if (!sp_eigs.count(label)) {
              solve(kappa1, eigenvalues);
              sp_eigs[label] = eigenvalues;
            }
            WriteDispersionData(ofs_disp, c++, label, sp_eigs[label]);
          }
        }
        ofs_disp << std::endl;
      }
    }
    double mt, st, mi, si;
    int ns;
    eq.GetSolverStats(mt, st, mi, si, ns);
    std::ofstream(out + "/stats_0.out") << "Timings: " << mt << " " << st << std::endl;   // (:651-655)
  } catch (const std::exception &e) {
    std::cerr << "maxwell_dispersion_b200: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
