import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
L = m.BravaisLattice("FCC")
eq = m.MaxwellBlochWaveEquation(L, 8, 2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers(), 0.25, 10.0, 1.0))
md = m.MaxwellDispersion(eq, L, 10, samp_pow=3, mid_pts=True)
K = md.buildRawBasis()
ks = m.k_path(L, ["Gamma", "X", "W", "L", "Gamma"], 8)
for k in ks:
    t = time.time(); md.approxEigenfrequencies(k); print("py total %.2f ms" % (1e3 * (time.time() - t)), flush=True)
