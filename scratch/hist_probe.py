import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import mfem_bravais_b200 as bloch
from test_gpu_history import _walk
L = bloch.BravaisLattice("FCC")
X = L.GetSymmetryPoint(L.GetSymmetryPointIndex("X"))
for step in (0.12, 0.06, 0.03, 0.01):
    ks = [(0.30 + step * i) * X + np.array([0.05, 0.02, 0.0]) for i in range(6)]
    lam1, it1 = _walk(bloch, ks, True)
    lam0, it0 = _walk(bloch, ks, False)
    print("step", step, "history", it1, "none", it0, "max diff", np.abs(lam1 - lam0).max(), flush=True)
