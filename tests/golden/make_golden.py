"""Generates tests/golden/bands_small.json from the CPU oracle (oracle/bloch_oracle.py).

The reference ships no golden vectors and cannot be built here, so these
fixtures pin the ORACLE (dense constrained-pencil eigen-solve of the assembled matrices) against
regressions and give the GPU tests a fixture that does not need the oracle at run time.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.bloch_oracle import BlochOperators, Lattice, Mesh, Spaces, empty_lattice_eigs  # noqa: E402

CASES = [  # lattice, n, p, kappa (absolute) or ("sym", label, fraction), sphere?, nev
    ("CUB", 3, 1, [0.7, -0.4, 1.1], True, 8),
    ("CUB", 2, 2, ("sym", "X", 1.0), True, 8),
    ("CUB", 2, 2, [0.0, 0.0, 0.0], True, 8),
    ("FCC", 2, 1, ("sym", "L", 0.5), True, 6),
    ("FCC", 2, 2, ("sym", "X", 0.5), True, 10),
    ("FCC", 2, 2, ("sym", "W", 1.0), False, 8),
    ("BCC", 1, 2, ("sym", "H", 0.5), True, 8),
    ("BCC", 1, 1, [0.3, 0.2, -0.1], True, 6),
    ("CUB", 2, 3, [1.0, 0.5, 0.25], True, 10),
]


def main():
    out = []
    for name, n, p, kap, sphere, nev in CASES:
        lat = Lattice(name)
        mesh = Mesh(lat, n)
        kappa = np.array(kap, float) if not isinstance(kap, tuple) else kap[2] * lat.kappa(kap[1])
        eps = mesh.sphere_eps() if sphere else np.ones(mesh.ne)
        ops = BlochOperators(Spaces(mesh, p), eps).set_kappa(kappa)
        w = ops.eig_dense(nev)
        rec = {"lattice": name, "n_sub": n, "order": p, "kappa": kappa.tolist(), "sphere": sphere,
               "n_nd": int(ops.sp_.n_nd), "eigenvalues": w.tolist()}
        if not sphere:
            rec["empty_lattice_exact"] = empty_lattice_eigs(lat, kappa, nev).tolist()
        out.append(rec)
        print(name, n, p, np.round(w, 6))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bands_small.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
