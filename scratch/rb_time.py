import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
import mfem_bravais_b200 as m
L = m.BravaisLattice("FCC")
eq = m.MaxwellBlochWaveEquation(L, 8, 2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers(), 0.25, 10.0, 1.0))
md = m.MaxwellDispersion(eq, L, 10, samp_pow=3, mid_pts=True)
t = time.time(); K = md.buildRawBasis(); print("raw basis", K, "cols in", time.time() - t, "s", flush=True)
ks = m.k_path(L, ["Gamma", "X", "W", "L", "Gamma"], 8)
t = time.time()
app = np.array([md.approxEigenfrequencies(k) for k in ks]); ta = time.time() - t
print("approx: %d k-points in %.3f s = %.1f k/s" % (len(ks), ta, len(ks) / ta), flush=True)
t = time.time()
lam, _ = m.dispersion_sweep(eq, ks, 10); tf = time.time() - t
full = np.sqrt(np.abs(lam))
print("full: %.3f s = %.1f k/s" % (tf, len(ks) / tf))
err = np.abs(app - full) / np.maximum(full, 1e-3)
print("max rel err of approx omega", err.max(), "median", np.median(err))
