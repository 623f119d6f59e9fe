import sys, json, os
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as bloch
L = bloch.BravaisLattice("BCC"); eq = bloch.MaxwellBlochWaveEquation(L, 1, 1)
eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
eq.SetAbsoluteTolerance(1e-9, 40)
try:
    lam = eq.GetEigenvalues(12, np.array([0.3,0.2,-0.1]))[0::2]; print(lam)
except Exception as e: print(e)
print(eq.GetSolverStats())
