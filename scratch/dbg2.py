import sys, json, os
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as bloch
for (name,n,p) in [("CUB",2,2),("CUB",4,1),("BCC",1,1)]:
    L = bloch.BravaisLattice(name); eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
    for kap in [np.zeros(3), np.array([0.3,0.2,-0.1])]:
        eq.SetKappa(kap); eq.Setup()
        rng=np.random.default_rng(0)
        x=rng.uniform(-1,1,(4,2*eq.N))
        y=eq.MultProjector(x); s0=eq.GetSolverStats()['inner_iterations']
        y2=eq.MultProjector(y); s1=eq.GetSolverStats()['inner_iterations']
        g0=np.abs(eq.debug_h1op(2,x)).max(); g1=np.abs(eq.debug_h1op(2,y)).max()
        print(name,n,p,"kappa",kap,"|GtMx|",g0,"|GtMPx|",g1,"|P^2x-Px|",np.abs(y2-y).max(),"cg its",s0,s1-s0)
