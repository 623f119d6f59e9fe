"""one warm batched solve (FCC order 2, nk k-points of the Gamma-X-W-L-Gamma path) with the phase profile on.
usage: batch_profile.py [n_sub] [nk] [verbose]"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
import mfem_bravais_b200 as m
lat = m.BravaisLattice("FCC")
ks = m.k_path(lat, ["Gamma", "X", "W", "L", "Gamma"], 8)
eq = m.MaxwellBlochWaveEquation(lat, nsub, 2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 2000)
chunk = len(ks) // nk
idx = lambda r: [(s * chunk + r) % len(ks) for s in range(nk)]
eq.SolveBatch(ks[idx(0)]); eq.SolveBatch(ks[idx(1)])
eq.SetProfile(True)
if len(sys.argv) > 3:
    os.environ["BLOCH_VERBOSE"] = "1"
rt = None
if os.environ.get("NCU_RANGE"):      # ncu --profile-from-start off: only the third solve is captured
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaProfilerStart()
t0 = time.time()
lam, st = eq.SolveBatch(ks[idx(2)])
dt = time.time() - t0
if rt:
    rt.cudaProfilerStop()
pr = eq.GetProfile()
print("nk %d n_sub %d: wall %.1f ms (%.1f per k-point), iterations %s" % (nk, nsub, 1e3 * dt, 1e3 * dt / nk, [s["iterations"] for s in st]))
tot = pr["solve"]
for k, v in pr.items():
    print("  %-28s %8.2f ms  %5.1f %%" % (k, v, 100 * v / tot))
print("  inner its", st[0]["inner_iterations"], "launches", st[0]["kernel_launches"])
