"""one warm single-kappa solve with the phase profile on.  usage: solve_profile.py LAT n p   (NCU_RANGE=1 under
ncu --profile-from-start off: only the profiled solve is captured)"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
import mfem_bravais_b200 as m
name, n, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lat = m.BravaisLattice(name)
kap = 0.5 * lat.GetSymmetryPoint(1 if name != "HEX" else 5)
eq = m.MaxwellBlochWaveEquation(lat, n, p)
eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 400)
eq.SetKappa(kap); eq.Setup(); eq.Solve()
eq.SetKappa(0.97 * kap + 0.01); eq.Setup(); eq.Solve()
eq.SetProfile(True)
rt = None
if os.environ.get("NCU_RANGE"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12"); rt.cudaProfilerStart()
eq.SetKappa(0.94 * kap + 0.02); t0 = time.time(); eq.Setup(); t1 = time.time(); eq.Solve(); t2 = time.time()
if rt:
    rt.cudaProfilerStop()
st = eq.GetSolverStats(); pr = eq.GetProfile()
print("%s n=%d p=%d N=%d: setup %.1f ms solve %.1f ms, %d iterations, %d inner, %d launches" % (name, n, p, eq.N, 1e3 * (t1 - t0), 1e3 * (t2 - t1), st["iterations"], st["inner_iterations"], st["kernel_launches"]))
for k, v in pr.items():
    print("  %-28s %8.2f ms  %5.1f %%" % (k, v, 100 * v / pr["solve"]))
