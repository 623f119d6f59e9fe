import sys, time; sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as m
# BASELINE config[0]: simple-cubic unit cell with dielectric sphere, ND order 1, Gamma point, 10 bands
for n in (8, 16):
    L=m.BravaisLattice("CUB"); eq=m.MaxwellBlochWaveEquation(L,n,1)
    eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
    eq.SetAbsoluteTolerance(1e-6)
    t=time.time(); lam=eq.GetEigenvalues(20, np.zeros(3)); dt=time.time()-t
    print("n",n,"N",eq.N,"time %.3f"%dt, eq.GetSolverStats()); print(np.round(lam[0::2],6))
