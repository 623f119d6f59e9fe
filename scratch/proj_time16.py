import sys, time, os
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as m
n=int(sys.argv[1])
L=m.BravaisLattice("FCC"); eq=m.MaxwellBlochWaveEquation(L,n,2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6)
for ki in [3,12]:
    t=time.time(); eq.SetKappa(ks[ki]); eq.Setup(); t1=time.time()-t; eq.Solve(); t2=time.time()-t
    print("n",n,"N",eq.N,"k%d"%ki,"setup %.3f solve %.3f"%(t1,t2-t1), eq.GetSolverStats())
