import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import mfem_bravais_b200 as m
L = m.BravaisLattice("BCC"); kH = 0.5 * L.GetSymmetryPoint(1)
for mode in ("1", "0"):
    os.environ["BLOCH_SCALAR_MG"] = mode
    for n, p in ((8, 4), (8, 2)):
        eq = m.ScalarFloquetWaveEquation(L, n, p)
        c = eq.element_centers(); ins = np.linalg.norm(c, axis=1) <= 0.5
        eq.SetStiffnessCoef(np.where(ins, 5.0, 0.1)); eq.SetMassCoef(np.where(ins, 10.0, 1.0))
        eq.SetNumEigs(40); eq.SetAbsoluteTolerance(1e-6, 400)
        eq.SetKappa(kH); eq.Setup(); t0 = time.time(); eq.Solve(); t1 = time.time()
        st = eq.GetSolverStats(); lam = eq.mode_eigenvalues()
        print("BLOCH_SCALAR_MG=%s BCC n=%d p=%d N=%d: %d iterations, %.3f s, converged %d, lam %s" % (mode, n, p, eq.N, st["iterations"], t1 - t0, st["converged_bands"], np.round(lam[:3], 8)), flush=True)
        del eq
