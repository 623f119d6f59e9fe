"""cold solves at small |kappa| (the homogenisation regime): outer iterations with the auxiliary-space preconditioner
and with the Chebyshev polynomial only"""
import sys, os, json
sys.path.insert(0, '.')
import numpy as np
import mfem_bravais_b200 as m
for name, n, p in (("CUB", 16, 1), ("FCC", 8, 2), ("BCC", 4, 3)):
    L = m.BravaisLattice(name)
    for mag in (0.0, 1e-3, 1e-2, 1e-1, 0.5):
        row = {}
        for mode in ("aux", "cheb"):
            os.environ["BLOCH_PRECOND"] = mode
            eq = m.MaxwellBlochWaveEquation(L, n, p)
            eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 400)
            eq.SetKappa(mag * np.array([1.0, 0.7, 0.4])); eq.Setup(); eq.Solve()
            st = eq.GetSolverStats(); row[mode] = (st["iterations"], st["converged_bands"], round(float(eq.band_eigenvalues()[3]), 6))
            del eq
        print(name, n, p, "|kappa| scale", mag, row, flush=True)
