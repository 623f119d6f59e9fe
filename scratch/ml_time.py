import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
L = m.BravaisLattice("FCC")
eps_fn = lambda c: m.sphere_eps(c, 0.25, 10.0, 1.0)
k = L.GetSymmetryPoint(L.GetSymmetryPointIndex("X")) * 0.6
for n0, lv in ((4, 2), (4, 3)):
    ml = m.MaxwellBlochWaveSolver(L, n0, 2, 10, eps_fn=eps_fn, max_lvl=lv, tol=1e-14)
    ml.SetKappa(k); ml.GetEigenfrequencies()          # builds levels, warms up
    t = time.time(); ml.GetEigenfrequencies(); tm = time.time() - t
    nf = n0 * 2 ** (lv - 1)
    d = m.MaxwellBlochWaveEquation(L, nf, 2); d.SetMassCoef(eps_fn(d.element_centers()))
    d.GetEigenvalues(20, k)
    d2 = m.MaxwellBlochWaveEquation(L, nf, 2); d2.SetMassCoef(eps_fn(d2.element_centers()))
    d2.SetNumEigs(20); d2.SetKappa(k); d2.Setup()
    t = time.time(); d2.Solve(); td = time.time() - t
    print("fine n=%d: multilevel %.3f s (iters per level %s)  cold direct %.3f s (%d iters)" %
          (nf, tm, ml.level_iters, td, d2.GetSolverStats()["iterations"]), flush=True)
