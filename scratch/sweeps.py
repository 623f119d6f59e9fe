import sys, time, os
sys.path.insert(0,'.')
import numpy as np, torch
import mfem_bravais_b200 as m
L=m.BravaisLattice("FCC"); eq=m.MaxwellBlochWaveEquation(L,8,2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6)
for sweep in range(3):
    ts=[]; its=[]
    for k in ks[:16]:
        t=time.time(); eq.SetKappa(k); eq.Setup(); t1=time.time(); eq.Solve(); t2=time.time()
        ts.append((t1-t,t2-t1)); its.append(eq.GetSolverStats()['iterations'])
    print("sweep",sweep,"total %.3f s"%sum(a+b for a,b in ts), "setup ms",[int(a*1e3) for a,b in ts], "solve ms",[int(b*1e3) for a,b in ts], "its",its)
