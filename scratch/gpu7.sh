BLOCH_VERBOSE=1 python - <<'PY' 2>&1 | cut -c1-200
import numpy as np, time, sys
sys.path.insert(0,'.')
import mfem_bravais_b200 as m
for (name,n,p) in [("FCC",16,2)]:
    L=m.BravaisLattice(name); eq=m.MaxwellBlochWaveEquation(L,n,p)
    eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
    ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
    eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 70)
    for k in ks[[3]]:
        t=time.time(); eq.SetKappa(k); eq.Setup(); t1=time.time()-t
        try: eq.Solve()
        except Exception as e: print(e)
        t2=time.time()-t
        print("N",eq.N,"k",k,"setup %.3f solve %.3f"%(t1,t2-t1), eq.GetSolverStats())
PY
