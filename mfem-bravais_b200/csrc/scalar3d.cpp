// scalar3d-shaped C++ driver (reference: misc/scalar3d.cpp:280-557) over include/maxwell_bloch_b200.hpp:
// the scalar H1 Bloch Helmholtz eigenproblem of ScalarFloquetWaveEquation with the reference's flags
// (-o order, -sr / -pr refinements, -nev real modes, -b phase shift in degrees, -az / -inc direction in degrees)
// and its piecewise constant coefficients (mass_coef / stiffness_coef, :560-588).  The reference reads a periodic
// cube mesh (-m); here the periodic cell is a Bravais lattice's Wigner-Seitz cell: -bl 1/2/3 = CUB/FCC/BCC
// (config 5 of BASELINE.json: the truncated octahedron = BCC), n_sub = 2^(sr + pr).  Extra: -dev <cuda device>.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/maxwell_bloch_b200.hpp"

using namespace bloch_b200;

static double mass_coef(const double *x) { return std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) <= 0.5 ? 10.0 : 1.0; }
static double stiffness_coef(const double *x) { return std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) <= 0.5 ? 5.0 : 0.1; }

int main(int argc, char **argv) {
  int bl_type = 1, order = 1, sr = 0, pr = 2, nev = 5, dev = -1;
  double beta = 1.0, alpha_a = 0.0, alpha_i = 0.0;      // driver defaults of scalar3d.cpp:287-293
  for (int i = 1; i < argc; i++) {
    auto next = [&](const char *f) -> const char * {
      if (i + 1 >= argc) { std::cerr << "missing value for " << f << std::endl; std::exit(1); }
      return argv[++i];
    };
    const std::string f = argv[i];
    if (f == "-bl") bl_type = std::atoi(next("-bl"));
    else if (f == "-o") order = std::atoi(next("-o"));
    else if (f == "-sr") sr = std::atoi(next("-sr"));
    else if (f == "-pr") pr = std::atoi(next("-pr"));
    else if (f == "-nev") nev = std::atoi(next("-nev"));
    else if (f == "-b") beta = std::atof(next("-b"));
    else if (f == "-az") alpha_a = std::atof(next("-az"));
    else if (f == "-inc") alpha_i = std::atof(next("-inc"));
    else if (f == "-dev") dev = std::atoi(next("-dev"));
    else if (f == "-no-vis" || f == "-no-visit" || f == "-vis" || f == "-visit") {}
    else { std::cerr << "unknown option " << f << std::endl; return 1; }
  }
  try {
    BravaisLattice bravais(bl_type + 6);
    const int n_sub = 1 << (sr + pr);
    ScalarFloquetWaveEquation eq(bravais, n_sub, order, dev);
    std::cout << "Lattice " << bravais.GetLatticeTypeLabel() << ", n_sub " << n_sub << ", order " << order
              << ", H1 unknowns " << eq.GetH1TrueVSize() << std::endl;
    std::vector<double> xyz, m(eq.GetNE()), k(eq.GetNE());
    eq.GetElementCenters(xyz);
    for (int64_t e = 0; e < eq.GetNE(); e++) { m[e] = mass_coef(&xyz[3 * e]); k[e] = stiffness_coef(&xyz[3 * e]); }
    eq.SetBeta(beta);
    eq.SetAzimuth(alpha_a);
    eq.SetInclination(alpha_i);
    eq.SetMassCoef(m);
    eq.SetStiffnessCoef(k);
    eq.SetNumEigs(nev);
    eq.SetAbsoluteTolerance(1e-6, 1000);      // lobpcg->SetTol(1e-6), SetMaxIter(1000), :431-432
    eq.Setup();
    eq.Solve();
    std::vector<double> eigenvalues;
    eq.GetEigenvalues(eigenvalues);
    int its = 0;
    double secs = 0.0;
    eq.GetSolverStats(its, secs);
    std::cout << "Eigenvalues:";
    std::cout.precision(12);
    for (double ev : eigenvalues) std::cout << " " << ev;
    std::cout << std::endl << "LOBPCG iterations " << its << ", solve time " << secs << " s" << std::endl;
  } catch (const std::exception &ex) {
    std::cerr << "scalar3d_b200: " << ex.what() << std::endl;
    return 2;
  }
  return 0;
}
