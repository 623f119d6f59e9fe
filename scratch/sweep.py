import numpy as np, time, sys, os, itertools
sys.path.insert(0,'.')
import mfem_bravais_b200 as m
n=int(sys.argv[1]); 
L=m.BravaisLattice("FCC"); eq=m.MaxwellBlochWaveEquation(L,n,2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 150)
os.environ["BLOCH_WARM_START"]="0"
for deg,ratio,sig in [(10,50,4),(16,100,4),(24,300,4),(32,600,4),(16,100,1),(16,100,16),(24,300,16)]:
    os.environ["BLOCH_CHEB_DEGREE"]=str(deg); os.environ["BLOCH_CHEB_RATIO"]=str(ratio); os.environ["BLOCH_SIGMA_SCALE"]=str(sig)
    for ki in [3,12]:
        eq.SetKappa(ks[ki]); eq.Setup()
        t=time.time()
        try: eq.Solve(); ok=True
        except Exception as e: ok=False
        st=eq.GetSolverStats()
        print("n=%d deg=%d ratio=%d sig=%g k%d: ok=%d its=%d cg=%d t=%.3f"%(n,deg,ratio,sig,ki,ok,st['iterations'],st['inner_iterations'],time.time()-t), flush=True)
