import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
L = m.BravaisLattice("FCC")
for n in (8, 16):
    eq = m.MaxwellBlochWaveEquation(L, n, 2)
    eq.SetMassCoef(m.sphere_eps(eq.element_centers(), 0.25, 10.0, 1.0))
    ks = m.k_path(L, ["Gamma", "X", "W"], 8)
    eq.SetKappa(ks[0]); eq.Setup()
    t = time.time()
    for k in ks[1:]:
        eq.SetKappa(k); eq.Setup()
    print("n", n, "setup per k: %.2f ms" % (1e3 * (time.time() - t) / (len(ks) - 1)), flush=True)
