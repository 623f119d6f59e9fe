"""Device-side Rayleigh-Ritz kernel (csrc/rr_device.cu) against the host solver of csrc/dense.hpp (bloch_debug_hegv)
and scipy: lowest m eigenpairs of random Hermitian pencils (A, B) with B positive definite, the sizes the eigensolver
produces (16/32/48/63 columns, m = 16/21), soft-locking masks, degenerate spectra, a nearly dependent basis that forces
the P block out, and the solver run with both Rayleigh-Ritz paths."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ri(z):
    return np.ascontiguousarray(np.stack([z.real, z.imag], axis=-1))


def device_hegv(bloch, GA, GM, m, act=None, use_p=1):
    from mfem_bravais_b200.capi import dptr
    nk, n = GA.shape[0], GA.shape[1]
    lam = np.zeros((nk, m))
    c = np.zeros((nk, n, m, 2))
    info = np.zeros(nk, np.int32)
    a = None if act is None else np.ascontiguousarray(act, np.uint8).ctypes.data_as(C.POINTER(C.c_ubyte))
    rc = bloch.lib().bloch_debug_hegv_device(n, m, nk, dptr(ri(GA)), dptr(ri(GM)), a, use_p, dptr(lam), dptr(c),
                                             info.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return lam, c[..., 0] + 1j * c[..., 1], info


def pencils(rng, nk, n, cond=1e3, degenerate=False):
    GA, GM = [], []
    for _ in range(nk):
        X = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        Q, _ = np.linalg.qr(X)
        w = np.sort(rng.uniform(0.5, 40.0, n))
        if degenerate:
            w[1] = w[0]; w[4] = w[3] = w[2]
        A0 = (Q * w) @ Q.conj().T
        Y = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        U, _ = np.linalg.qr(Y)
        T = (U * np.geomspace(1.0, 1.0 / np.sqrt(cond), n)) @ U.conj().T * rng.uniform(0.3, 3.0)      # B = T^H T
        GA.append(T.conj().T @ A0 @ T)
        GM.append(T.conj().T @ T)
    return np.array(GA), np.array(GM)


@pytest.mark.parametrize("n,m", [(16, 16), (32, 16), (48, 16), (63, 21), (5, 2)])
@pytest.mark.parametrize("degenerate", [False, True])
def test_device_rr_matches_scipy_and_host(bloch, n, m, degenerate):
    if degenerate and n < 8:
        pytest.skip("too small for the degenerate pattern")
    rng = np.random.default_rng(11 + n)
    GA, GM = pencils(rng, 3, n, degenerate=degenerate)
    lam, Cc, info = device_hegv(bloch, GA, GM, m)
    assert (info == 0).all()
    from mfem_bravais_b200.capi import dptr
    for k in range(3):
        w = sla.eigh(GA[k], GM[k], eigvals_only=True)[:m]
        assert np.allclose(lam[k], w, rtol=1e-10, atol=1e-12)
        c = Cc[k]
        assert np.allclose(c.conj().T @ GM[k] @ c, np.eye(m), atol=1e-9)           # GM-orthonormal
        assert np.allclose(c.conj().T @ GA[k] @ c, np.diag(lam[k]), atol=1e-8)     # diagonalises GA
        hl, hc = np.zeros(m), np.zeros((n, m, 2))
        assert bloch.lib().bloch_debug_hegv(n, m, dptr(ri(GA[k])), dptr(ri(GM[k])), dptr(hl), dptr(hc), 0) == 0
        assert np.allclose(lam[k], hl, rtol=1e-10, atol=1e-12)


def test_device_rr_soft_locking_and_rank_deficiency(bloch):
    rng = np.random.default_rng(3)
    m, n = 8, 24
    GA, GM = pencils(rng, 2, n)
    act = np.ones((2, m), np.uint8)
    act[0, [1, 5]] = 0                                   # converged columns: their W and P members leave the basis
    lam, Cc, info = device_hegv(bloch, GA, GM, m, act=act)
    keep = [i for i in range(n) if i < m or act[0, i % m]]
    w = sla.eigh(GA[0][np.ix_(keep, keep)], GM[0][np.ix_(keep, keep)], eigvals_only=True)[:m]
    assert np.allclose(lam[0], w, rtol=1e-10) and info[0] == 0
    dropped = [i for i in range(n) if i not in keep]
    assert np.all(Cc[0][dropped] == 0)
    # without the P block (use_p = 0): the first two groups only
    lam2, C2, info2 = device_hegv(bloch, GA, GM, m, use_p=0)
    w2 = sla.eigh(GA[1][:2 * m, :2 * m], GM[1][:2 * m, :2 * m], eigvals_only=True)[:m]
    assert np.allclose(lam2[1], w2, rtol=1e-10) and np.all(C2[1][2 * m:] == 0)
    # a P column that is (numerically) a copy of a W column: the Cholesky pivot test fails, the P block is dropped
    S = rng.normal(size=(200, n)) + 1j * rng.normal(size=(200, n))
    S[:, 2 * m + 3] = S[:, m + 3] * (1.0 + 1e-13)
    H = rng.normal(size=(200, 200)); H = H @ H.T + np.eye(200)
    GAd, GMd = (S.conj().T @ H @ S)[None], (S.conj().T @ S)[None]
    lam3, C3, info3 = device_hegv(bloch, GAd, GMd, m)
    assert info3[0] == 1 and np.all(C3[0][2 * m:] == 0)
    w3 = sla.eigh(GAd[0][:2 * m, :2 * m], GMd[0][:2 * m, :2 * m], eigvals_only=True)[:m]
    assert np.allclose(lam3[0], w3, rtol=1e-9)


SOLVE = """
import sys, json
sys.path.insert(0, %r)
import numpy as np
import mfem_bravais_b200 as m
lat = m.BravaisLattice("FCC")
eq = m.MaxwellBlochWaveEquation(lat, 4, 2)
eq.SetMassCoef(m.sphere_eps(eq.element_centers())); eq.SetNumEigs(16); eq.SetAbsoluteTolerance(1e-8)
ks = np.array([[0.7, -0.4, 1.1], [0.0, 0.0, 0.0], [0.0, 3.0, 0.0]])
lam, st = eq.SolveBatch(ks)
lam2, st2 = eq.SolveBatch(ks * 1.03)
print(json.dumps({"lam": lam.tolist(), "lam2": lam2.tolist(), "its": [s["iterations"] for s in st + st2],
                  "conv": [s["converged_bands"] for s in st + st2]}))
"""


def test_solver_with_device_and_host_rayleigh_ritz_agree():
    import json
    out = {}
    for mode in ("1", "0"):
        env = dict(os.environ, BLOCH_RR_DEVICE=mode)
        r = subprocess.run([sys.executable, "-c", SOLVE % ROOT], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("lam", "lam2"):
        a, b = np.array(out["1"][key]), np.array(out["0"][key])
        assert np.allclose(a, b, rtol=1e-7, atol=1e-8), (key, a, b)
    assert all(c == 8 for c in out["1"]["conv"]) and all(c == 8 for c in out["0"]["conv"])
    assert max(out["1"]["its"]) <= max(out["0"]["its"]) + 3
