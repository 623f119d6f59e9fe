import sys, time, os
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as m
L=m.BravaisLattice("FCC"); eq=m.MaxwellBlochWaveEquation(L,8,2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
rng=np.random.default_rng(0)
x=rng.uniform(-1,1,(14,2*eq.N))
for ki in [0,3,7,20,31]:
    eq.SetKappa(ks[ki]); eq.Setup()
    eq.MultProjector(x)
    s0=eq.GetSolverStats()['inner_iterations']
    t=time.time()
    for _ in range(3): eq.MultProjector(x)
    dt=(time.time()-t)/3
    its=(eq.GetSolverStats()['inner_iterations']-s0)/3
    print("k%d |k|=%.3f  proj(1e-13) %.2f ms, %d its, %.1f us/it"%(ki,np.linalg.norm(ks[ki]),dt*1e3,its,dt*1e6/max(its,1)))
