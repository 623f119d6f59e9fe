"""Operator apply at BASELINE.json's full sizes, where the assembled oracle is too slow: size-independent
properties (Hermitian symmetry, linearity, exact sequence A G = 0, M > 0) and agreement of the three ND apply
kernels (lane pair per item, six lanes per item, first-generation cooperative tile; SURVEY.md section 8(a) rows
a3-a7) on the same seeded inputs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FULL = [("CUB", 16, 1), ("FCC", 16, 2), ("BCC", 8, 3)]       # configs 1, 2, 3 (parity size of config 3)


def cdot(a, b, N):
    """<a, b> for vectors stored [re; im]."""
    ac, bc = a[:N] + 1j * a[N:], b[:N] + 1j * b[N:]
    return np.vdot(ac, bc)


@pytest.mark.parametrize("name,n,p", FULL)
def test_full_size_properties(bloch, name, n, p):
    rng = np.random.default_rng(7)
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
    eq.SetStiffnessCoef(rng.uniform(0.5, 2.0, eq.n_elem))
    eq.SetKappa(np.array([0.7, -1.3, 2.1]))
    eq.Setup()
    N = eq.N
    x = rng.uniform(-1, 1, (2, 2 * N))
    ax, mx = eq.MultA(x), eq.MultM(x)
    scale = np.linalg.norm(ax[0]) * np.linalg.norm(x[1])
    # Hermitian: <x1, A x0> = conj <x0, A x1>
    assert abs(cdot(x[1], ax[0], N) - np.conj(cdot(x[0], ax[1], N))) < 1e-12 * scale
    assert abs(cdot(x[1], mx[0], N) - np.conj(cdot(x[0], mx[1], N))) < 1e-12 * np.linalg.norm(mx[0]) * np.linalg.norm(x[1])
    # positive (semi-)definite
    assert cdot(x[0], ax[0], N).real > 0 and abs(cdot(x[0], ax[0], N).imag) < 1e-12 * scale
    assert cdot(x[0], mx[0], N).real > 0
    # linearity, also across the columns of one block launch
    comb = (0.3 * x[0] - 1.7 * x[1])[None, :]
    assert np.linalg.norm(eq.MultA(comb)[0] - (0.3 * ax[0] - 1.7 * ax[1])) < 1e-12 * np.linalg.norm(ax[0])
    # exact sequence on affine meshes: (C - i Z)(T01 - i Z01) = 0, hence A G phi = 0
    phi = rng.uniform(-1, 1, (1, 2 * eq.N_h1))
    g = eq.debug_h1op(1, phi)
    ag = eq.MultA(g)
    assert np.linalg.norm(ag) < 1e-11 * np.linalg.norm(ax[0]) * np.linalg.norm(g) / np.linalg.norm(x[0])


def _probe(env, name, n, p, nvec=3):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, os.path.join(HERE, "_apply_probe.py"), name, str(n), str(p), str(nvec)],
                         env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("name,n,p", [("FCC", 8, 2), ("BCC", 4, 3), ("CUB", 12, 1)])
def test_kernel_generations_agree(name, n, p):
    ref = _probe({"BLOCH_ND_ITEM": "0", "BLOCH_ND_COMP": "0"}, name, n, p)          # first-generation k_nd_apply
    variants = {"default": {}, "six-lane": {"BLOCH_ND_COMP": "7"}, "two-pass": {"BLOCH_TWO_PASS": "1"}}
    if p <= 2:
        variants["lane-pair"] = {"BLOCH_ND_COMP": "0"}
    for tag, env in variants.items():
        got = _probe(env, name, n, p)
        for key in ("normA", "normM", "wA", "wM", "headA"):
            a, b = np.array(got[key]), np.array(ref[key])
            assert np.allclose(a, b, rtol=1e-11, atol=1e-11 * np.abs(b).max()), (tag, key, a, b)
