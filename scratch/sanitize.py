"""small run touching every kernel, for compute-sanitizer memcheck"""
import sys
sys.path.insert(0, '.')
import numpy as np
import mfem_bravais_b200 as m
rng = np.random.default_rng(0)
for name, n, p in [("FCC", 2, 1), ("CUB", 2, 2), ("BCC", 2, 3)]:
    L = m.BravaisLattice(name); eq = m.MaxwellBlochWaveEquation(L, n, p)
    eq.SetMassCoef(rng.uniform(1, 10, eq.n_elem)); eq.SetKappa([0.3, 0.2, -0.4]); eq.Setup()
    x = rng.uniform(-1, 1, (3, 2 * eq.N))
    eq.MultA(x); eq.MultM(x); eq.MultC(x); eq.MultProjector(x)
    eq.debug_h1op(0, rng.uniform(-1, 1, (2, 2 * eq.N_h1)))
    eq.SetNumEigs(8); eq.SetAbsoluteTolerance(1e-6, 60)
    try:
        eq.Solve(); print(name, p, eq.band_eigenvalues())
        eq.GetEigenvectorE(0); eq.GetEigenvectorB(0)
    except m.BlochError as e:
        print("solve:", e)
L = m.BravaisLattice("CUB"); sc = m.ScalarFloquetWaveEquation(L, 2, 4)
sc.SetStiffnessCoef(np.ones(sc.n_elem)); sc.SetMassCoef(np.ones(sc.n_elem)); sc.SetKappa([0.5, 0.1, 0.2])
sc.SetNumEigs(6); sc.SetAbsoluteTolerance(1e-6, 60); sc.Setup(); sc.Solve(); print("scalar", sc.mode_eigenvalues())
print("sanitize run done")
