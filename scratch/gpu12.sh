python bench.py --no-cpu-baseline --steps 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'e2e',d['e2e']['value'],'its',d['lobpcg_iterations_mean'],'launches',d['gpu_launches'], d['clocks'])
"
BLOCH_VERBOSE=1 python - <<'PY' 2>&1 | grep -v "lobpcg\] it "
import numpy as np, time, sys
sys.path.insert(0,'.')
import mfem_bravais_b200 as m
L=m.BravaisLattice("FCC"); eq=m.MaxwellBlochWaveEquation(L,8,2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
ks=m.k_path(L,['Gamma','X','W','L','Gamma'],8)
eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6)
for k in ks[[0,1,2,3]]:
    t=time.time(); eq.SetKappa(k); eq.Setup(); t1=time.time()-t; eq.Solve(); t2=time.time()-t
    print("setup %.3f solve %.3f"%(t1,t2-t1), eq.GetSolverStats())
PY
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,utilization.gpu --format=csv
