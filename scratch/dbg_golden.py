import sys, json, os
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as bloch
cases = json.load(open("tests/golden/bands_small.json"))
for c in cases:
    L = bloch.BravaisLattice(c["lattice"])
    eq = bloch.MaxwellBlochWaveEquation(L, c["n_sub"], c["order"])
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()) if c["sphere"] else np.ones(eq.n_elem))
    nb = len(c["eigenvalues"])
    eq.SetAbsoluteTolerance(1e-7, 300)
    try:
        lam = eq.GetEigenvalues(2 * nb, np.array(c["kappa"]))[0::2]
        print(c["lattice"], c["n_sub"], c["order"], "N", eq.N, "ok", np.abs(lam-np.array(c["eigenvalues"])).max(), eq.GetSolverStats()["iterations"])
    except Exception as e:
        print(c["lattice"], c["n_sub"], c["order"], "N", eq.N, "FAIL", e, eq.GetSolverStats())
