"""Generates tests/golden/bands_baseline.json: band eigenvalues of the CPU oracle at the sizes of the
BASELINE.json configurations the bench and the GPU parity tests run on (FCC order 2 n_sub 8 = the bench
workload, three k-points of Gamma-X-W-L-Gamma; CUB order 1 n_sub 16 at Gamma; HEX order 2 n_sub 8).

Method: ARPACK shift-invert around sigma on the oracle's ASSEMBLED pencil (A_c, M_c) with a sparse LU
(oracle/bloch_oracle.py::eig_shift_invert).  Completeness of the lowest bands is certified per case: the
returned window around sigma must contain members of the gradient null space (lambda ~ 0), so that every
positive eigenvalue below sigma is inside the window.  Takes a few minutes per case on 8 cores.
Run from the repo root:  python tests/golden/make_bands_baseline.py
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.bloch_oracle import BlochOperators, Lattice, Mesh, Spaces  # noqa: E402

PATH = ["Gamma", "X", "W", "L", "Gamma"]
CASES = [  # lattice, n, p, k-points (index into the 32-point bench path, or absolute kappa), bands wanted
    ("FCC", 8, 2, [("path", 3), ("path", 11), ("path", 20)], 10),
    ("CUB", 16, 1, [("abs", [0.0, 0.0, 0.0])], 10),
    ("HEX", 8, 2, [("abs", [1.1, 0.6, 0.9])], 10),
]


def lowest_bands(ops, nev, sigma0):
    """nev lowest eigenvalues above the null space, certified complete (see module docstring)."""
    sigma = sigma0
    for _ in range(8):
        w = ops.eig_shift_invert(nev, sigma, extra=nev + 20)
        scale = max(abs(w).max(), 1.0)
        zeros = w[abs(w) < 1e-7 * scale]
        pos = w[w > 1e-7 * scale]
        below = pos[pos < sigma]
        print("   sigma %.3f: %d null-space members, %d positive (%d below sigma)" % (sigma, len(zeros), len(pos), len(below)))
        if len(zeros) == 0:          # window does not reach down to 0: it may miss low bands
            sigma *= 0.6
            continue
        if len(pos) < nev:           # window too low
            sigma *= 1.5
            continue
        return np.sort(pos)[:nev], sigma
    raise RuntimeError("no certified window found")


def main():
    out = []
    for name, n, p, kpts, nev in CASES:
        lat = Lattice(name)
        mesh = Mesh(lat, n)
        t0 = time.time()
        ops = BlochOperators(Spaces(mesh, p), mesh.sphere_eps())
        path = lat.kpath(PATH, 8) if name == "FCC" else None
        for kind, val in kpts:
            kappa = path[val] if kind == "path" else np.array(val, float)
            ops.set_kappa(kappa)
            # first guess of the window centre: a bit below the 10th empty-lattice-like level
            sigma0 = 0.5 * (2.0 * np.pi / lat.volume ** (1.0 / 3.0)) ** 2
            w, sigma = lowest_bands(ops, nev, sigma0)
            rec = {"lattice": name, "n_sub": n, "order": p, "kappa": kappa.tolist(),
                   "path_index": val if kind == "path" else None, "n_nd": int(ops.sp_.n_nd),
                   "sigma": sigma, "eigenvalues": w.tolist()}
            out.append(rec)
            print(name, n, p, kappa, np.round(w, 6), "%.0f s" % (time.time() - t0), flush=True)
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bands_baseline.json"), "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
