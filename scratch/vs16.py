import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
L = m.BravaisLattice("FCC")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eq = m.MaxwellBlochWaveEquation(L, n, 2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers(), 0.25, 10.0, 1.0))
ks = m.k_path(L, ["Gamma", "X", "W", "L", "Gamma"], 8)
t = time.time()
lam, st = m.dispersion_sweep(eq, ks[:4], 10)
print("total %.3f s, iters %s inner %s" % (time.time() - t, [s["iterations"] for s in st], [s["inner_iterations"] for s in st]))
