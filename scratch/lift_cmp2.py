import time, numpy as np, sys, os
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
name, n, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
L = m.BravaisLattice(name)
eq = m.MaxwellBlochWaveEquation(L, n, p)
eq.SetMassCoef(m.sphere_eps(eq.element_centers(), 0.3, 13.0, 1.0))
pts = [L.GetSymmetryPoint(i) for i in range(min(4, L.GetNumberSymmetryPoints()))]
ks = []
for a, b in zip(pts[:-1], pts[1:]):
    for t in (0.0, 0.25, 0.5, 0.75):
        ks.append((1 - t) * a + t * b)
ks.append(pts[-1])
t = time.time()
lam, st = m.dispersion_sweep(eq, ks, 8, tol=1e-7)
dt = time.time() - t
tag = os.environ.get("BLOCH_LIFT", "1")
fn = "/tmp/l2_%s_%d_%d_%s.npy" % (name, n, p, tag)
np.save(fn, lam)
msg = "%s n%d p%d lift %s: %.2f s iters %d inner %d maxres %.1e" % (name, n, p, tag, dt, sum(s["iterations"] for s in st), sum(s["inner_iterations"] for s in st), max(s["max_residual"] for s in st))
ref = "/tmp/l2_%s_%d_%d_0.npy" % (name, n, p)
if tag != "0" and os.path.exists(ref):
    msg += " | max|dlam| vs exact-projection run %.1e" % np.abs(np.load(ref) - lam).max()
print(msg, flush=True)
