"""phase timing of one cold solve: BCC p3 n8 Maxwell and scalar H1 p4 BCC n8 (regression hunt); usage: regress.py [maxwell|scalar]"""
import os, sys, time
sys.path.insert(0, ".")
os.environ["BLOCH_VERBOSE"] = "1"
import numpy as np
import mfem_bravais_b200 as m
which = sys.argv[1] if len(sys.argv) > 1 else "maxwell"
L = m.BravaisLattice("BCC"); kH = 0.5 * L.GetSymmetryPoint(1)
if which == "scalar":
    eq = m.ScalarFloquetWaveEquation(L, 8, 4)
    c = eq.element_centers(); ins = np.linalg.norm(c, axis=1) <= 0.5
    eq.SetStiffnessCoef(np.where(ins, 5.0, 0.1)); eq.SetMassCoef(np.where(ins, 10.0, 1.0))
    eq.SetNumEigs(40); eq.SetAbsoluteTolerance(1e-6, 400); eq.SetKappa(kH)
else:
    eq = m.MaxwellBlochWaveEquation(L, 8, 3)
    eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 400); eq.SetKappa(kH)
t0 = time.time(); eq.Setup(); t1 = time.time(); eq.Solve(); t2 = time.time()
print("setup %.3f s solve %.3f s" % (t1 - t0, t2 - t1), eq.GetSolverStats())
