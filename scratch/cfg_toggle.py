import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import mfem_bravais_b200 as m
name, n, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
L = m.BravaisLattice(name)
eq = m.MaxwellBlochWaveEquation(L, n, p)
eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 400)
k0 = 0.5 * L.GetSymmetryPoint(1)
ts = []
for f in (1.0, 1.1, 1.2):
    eq.SetKappa(f * k0); t = time.time(); eq.Setup(); eq.Solve(); ts.append(time.time() - t)
st = eq.GetSolverStats()
print(name, n, p, "solve times", [round(x, 3) for x in ts], "iters", st["iterations"], "inner", st["inner_iterations"], flush=True)
