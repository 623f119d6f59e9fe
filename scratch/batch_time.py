"""k-points/s of the FCC order-2 path sweep as a function of the batch size (k-points iterated together per handle)
and of the number of concurrent handles (streams).  usage: batch_time.py [n_sub] [steps]"""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np
import mfem_bravais_b200 as m

nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lat = m.BravaisLattice("FCC")
ks = m.k_path(lat, ["Gamma", "X", "W", "L", "Gamma"], 8)

def sweep(eqs, B, idxs):
    """T handles x B slots: the index list is cut into T*B contiguous chunks; handle t walks chunks t*B..t*B+B-1 in lock step"""
    T = len(eqs)
    nch = T * B
    bounds = [len(idxs) * c // nch for c in range(nch + 1)]
    chunks = [idxs[bounds[c]:bounds[c + 1]] for c in range(nch)]
    out, its = {}, []
    def worker(t):
        mine = chunks[t * B:(t + 1) * B]
        for r in range(max(len(c) for c in mine)):
            act = [c[r] if r < len(c) else c[-1] for c in mine if len(c) > 0]
            lam, st = eqs[t].SolveBatch([ks[i % len(ks)] for i in act])
            for j, i in enumerate(act):
                out[i] = lam[j]
            its.append(max(s["iterations"] for s in st))
    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [t.start() for t in th]; [t.join() for t in th]
    return out, its

ref = None
for T, B in [(1, 1), (4, 1), (1, 2), (1, 4), (1, 8), (2, 4), (2, 8), (1, 16), (4, 4)]:
    eqs = [m.MaxwellBlochWaveEquation(lat, nsub, 2) for _ in range(T)]
    eps = m.sphere_eps(eqs[0].element_centers())
    for eq in eqs:
        eq.SetMassCoef(eps); eq.SetNumEigs(20); eq.SetAbsoluteTolerance(1e-6, 2000)
    sweep(eqs, B, list(range(100, 100 + max(4, T * B))))     # warm-up (kernel attributes, graphs, buffers)
    t0 = time.time()
    out, its = sweep(eqs, B, list(range(steps)))
    dt = time.time() - t0
    lam = np.array([out[i] for i in range(steps)])
    if ref is None:
        ref = lam
    err = np.max(np.abs(lam - ref) / np.maximum(np.abs(ref), 1e-3))
    print("streams %d batch %2d: %6.2f k-points/s  (%.1f ms per k-point, mean its %.1f, max rel dev from first row %.1e)"
          % (T, B, steps / dt, 1e3 * dt / steps, np.mean(its), err), flush=True)
    del eqs
