set -x
python -m pytest tests/test_gpu_eigen.py -x -q -m gpu 2>&1 | tail -40
BLOCH_VERBOSE=1 python - <<'PY'
import numpy as np, time, sys
sys.path.insert(0,'.')
import mfem_bravais_b200 as m
for (name,n,p) in [("FCC",8,2)]:
    L=m.BravaisLattice(name); eq=m.MaxwellBlochWaveEquation(L,n,p)
    eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
    ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
    eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6)
    for k in ks[:3]:
        t=time.time(); eq.SetKappa(k); eq.Setup(); t1=time.time()-t; eq.Solve(); t2=time.time()-t
        print("N",eq.N,"k",k,"setup %.3f solve %.3f"%(t1,t2-t1), eq.GetSolverStats()); print(eq.band_eigenvalues())
PY
