"""one k-point of each BASELINE.json config at full size (documentation run)"""
import sys, time, json
sys.path.insert(0,'.')
import numpy as np
import mfem_bravais_b200 as m
out=[]
def run(tag, name, n, p, nb, kap, scalar=False):
    L=m.BravaisLattice(name)
    t0=time.time()
    if scalar:
        eq=m.ScalarFloquetWaveEquation(L,n,p)
        c=eq.element_centers(); ins=np.linalg.norm(c,axis=1)<=0.5
        eq.SetStiffnessCoef(np.where(ins,5.0,0.1)); eq.SetMassCoef(np.where(ins,10.0,1.0))
        eq.SetNumEigs(2*nb); eq.SetAbsoluteTolerance(1e-6, 400)
        eq.SetKappa(kap); t1=time.time(); eq.Setup(); eq.Solve(); lam=eq.mode_eigenvalues(); N=eq.N
    else:
        eq=m.MaxwellBlochWaveEquation(L,n,p)
        eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(2*nb); eq.SetAbsoluteTolerance(1e-6, 400)
        eq.SetKappa(kap); t1=time.time(); eq.Setup(); eq.Solve(); lam=eq.band_eigenvalues(); N=eq.N
    t2=time.time(); st=eq.GetSolverStats()
    rec=dict(config=tag,lattice=name,n_sub=n,order=p,N=int(N),bands=nb,create_s=t1-t0,solve_s=t2-t1,iterations=st['iterations'],converged=st['converged_bands'],max_residual=st['max_residual'],lambda_first=[float(x) for x in lam[:4]])
    print(json.dumps(rec), flush=True); out.append(rec)
L=m.BravaisLattice("BCC"); kH=0.5*L.GetSymmetryPoint(1)
run("1 CUB p1 n16 Gamma","CUB",16,1,10,np.zeros(3))
run("2 FCC p2 n16","FCC",16,2,10,0.5*m.BravaisLattice("FCC").GetSymmetryPoint(1))
run("3 BCC p3 n8 (parity size)","BCC",8,3,10,kH)
run("3 BCC p3 n12 (2.24M DOF)","BCC",12,3,10,kH)
run("4 HEX p2 n8","HEX",8,2,10,0.5*m.BravaisLattice("HEX").GetSymmetryPoint(5))
run("5 scalar H1 p4 BCC n8","BCC",8,4,20,kH,scalar=True)
json.dump(out,open("gpurun_out/configs_%s.json" % (sys.argv[1] if len(sys.argv) > 1 else "r2"),"w"),indent=1)
