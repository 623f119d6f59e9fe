"""Outer iterations / solve time of one cold k-point per config with the auxiliary-space preconditioner (default) and
with the Chebyshev polynomial only (BLOCH_PRECOND=cheb, read at handle creation).  usage: aux_cmp.py tag [cases...]"""
import sys, time, json, os
sys.path.insert(0, '.')
import numpy as np
import mfem_bravais_b200 as m

CASES = {"cub16": ("CUB", 16, 1, "G"), "fcc8": ("FCC", 8, 2, "X"), "fcc16": ("FCC", 16, 2, "X"), "bcc8": ("BCC", 8, 3, "H"),
         "bcc12": ("BCC", 12, 3, "H"), "hex8": ("HEX", 8, 2, "K"), "bcc4": ("BCC", 4, 3, "H"), "fcc4": ("FCC", 4, 2, "X")}
out = []
tag = sys.argv[1]
for case in sys.argv[2:]:
    name, n, p, pt = CASES[case]
    L = m.BravaisLattice(name)
    kap = np.zeros(3) if pt == "G" else 0.5 * L.GetSymmetryPoint(1 if name != "HEX" else 5)
    for mode in os.environ.get("AUX_CMP_MODES", "aux,cheb").split(","):
        os.environ["BLOCH_PRECOND"] = mode
        eq = m.MaxwellBlochWaveEquation(L, n, p)
        eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 400)
        eq.SetKappa(kap); t0 = time.time(); eq.Setup(); t1 = time.time(); eq.Solve(); t2 = time.time()
        lam = eq.band_eigenvalues(); st = eq.GetSolverStats()
        # second, warm solve of a neighbouring k-point (what a sweep does)
        eq.SetKappa(kap * 0.95 + 0.01); eq.Setup(); t3 = time.time(); eq.Solve(); t4 = time.time(); st2 = eq.GetSolverStats()
        rec = dict(case=case, mode=mode, N=int(eq.N), setup_s=round(t1 - t0, 3), solve_s=round(t2 - t1, 3), iterations=st["iterations"],
                   inner=st["inner_iterations"], converged=st["converged_bands"], warm_solve_s=round(t4 - t3, 3), warm_iterations=st2["iterations"],
                   lam=[round(float(x), 9) for x in lam[:4]])
        print(json.dumps(rec), flush=True); out.append(rec)
        del eq
json.dump(out, open("gpurun_out/aux_cmp_%s.json" % tag, "w"), indent=1)
