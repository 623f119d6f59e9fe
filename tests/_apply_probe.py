"""Prints invariants of Y = A X and Y = M X for seeded inputs (JSON); run in a subprocess so that the kernel
selection variables (BLOCH_ND_ITEM / BLOCH_ND_COMP / BLOCH_TWO_PASS, read once per process) can differ per run."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfem_bravais_b200 as m  # noqa: E402

name, n, p, nvec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lat = m.BravaisLattice(name)
eq = m.MaxwellBlochWaveEquation(lat, n, p)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
rng = np.random.default_rng(12345)
eq.SetStiffnessCoef(rng.uniform(0.5, 2.0, eq.n_elem))
eq.SetKappa(np.array([0.7, -1.3, 2.1]))
eq.Setup()
x = rng.uniform(-1, 1, (nvec, 2 * eq.N))
w = rng.uniform(-1, 1, 2 * eq.N)
ya, ym = eq.MultA(x), eq.MultM(x)
print(json.dumps({"N": int(eq.N), "normA": [float(np.linalg.norm(r)) for r in ya],
                  "normM": [float(np.linalg.norm(r)) for r in ym],
                  "wA": [float(r @ w) for r in ya], "wM": [float(r @ w) for r in ym],
                  "headA": [float(v) for v in ya[0, :8]]}))
