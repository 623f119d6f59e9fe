import sys, json, os
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as bloch
L = bloch.BravaisLattice("CUB"); eq = bloch.MaxwellBlochWaveEquation(L, 2, 2)
eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
eq.SetAbsoluteTolerance(float(sys.argv[1]), 60)
try:
    lam = eq.GetEigenvalues(16, np.zeros(3))[0::2]; print(lam)
except Exception as e: print(e)
print(eq.GetSolverStats())
