import time, numpy as np, sys, os
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
L = m.BravaisLattice("FCC")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eq = m.MaxwellBlochWaveEquation(L, n, 2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers(), 0.25, 10.0, 1.0))
ks = m.k_path(L, ["Gamma", "X", "W", "L", "Gamma"], 8)
m.dispersion_sweep(eq, ks[:2], 10)
t = time.time()
lam, st = m.dispersion_sweep(eq, ks, 10)
dt = time.time() - t
tag = os.environ.get("BLOCH_LIFT", "1")
np.save("/tmp/lam_lift%s_%d.npy" % (tag, n), lam)
print("lift %s n %d: %.3f s, %.1f k/s, iters %d, inner %d, maxres %.2e" % (tag, n, dt, len(ks) / dt,
      sum(s["iterations"] for s in st), sum(s["inner_iterations"] for s in st), max(s["max_residual"] for s in st)), flush=True)
ref = "/tmp/lam_lift0_%d.npy" % n
if tag != "0" and os.path.exists(ref):
    print("   max |dlam| vs lift 0: %.2e" % np.abs(np.load(ref) - lam).max())
