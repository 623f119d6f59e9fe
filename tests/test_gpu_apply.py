"""Parity of the CUDA element kernels (through the C ABI) with the assembled oracle operators:
north-star criterion 'operator apply to relative 1e-12 against the assembled reference matrix'."""
import numpy as np
import pytest

from helpers import oracle_on_product_maps, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [("CUB", 3, 1), ("CUB", 2, 2), ("CUB", 2, 3), ("FCC", 2, 1), ("FCC", 2, 2), ("FCC", 1, 3),
         ("BCC", 1, 1), ("BCC", 1, 2), ("BCC", 1, 3), ("CUB", 1, 1), ("FCC", 3, 2), ("HEX", 2, 2), ("HEX", 1, 3)]


def _setup(bloch, name, n, p, seed=0):
    rng = np.random.default_rng(seed)
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eps = rng.uniform(1.0, 10.0, eq.n_elem)
    mui = rng.uniform(0.5, 2.0, eq.n_elem)
    eq.SetMassCoef(eps)
    eq.SetStiffnessCoef(mui)
    kappa = rng.normal(size=3) * 2.0
    eq.SetKappa(kappa)
    eq.Setup()
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps, mui)
    ops.set_kappa(kappa)
    return eq, ops, rng


@pytest.mark.parametrize("name,n,p", CASES)
def test_apply_A_and_M_match_assembled(bloch, name, n, p):
    eq, ops, rng = _setup(bloch, name, n, p)
    for nvec in (1, 3, 10):
        x = rng.uniform(-1, 1, (nvec, 2 * eq.N))
        assert rel_err(eq.MultA(x), ops.apply_A(x)) < TOL
        assert rel_err(eq.MultM(x), ops.apply_M(x)) < TOL


@pytest.mark.parametrize("name,n,p", CASES)
def test_gamma_point_branch(bloch, name, n, p):
    """kappa = 0 selects the real branch (fabs(beta) > 0 test, maxwell_bloch.cpp:402)."""
    eq, ops, rng = _setup(bloch, name, n, p, seed=1)
    eq.SetKappa(np.zeros(3))
    eq.Setup()
    ops.set_kappa(np.zeros(3))
    x = rng.uniform(-1, 1, (2, 2 * eq.N))
    ref = ops.apply_A(x)
    # on the one-element periodic meshes the assembled curl vanishes identically (A x = 0 exactly), so the error is
    # measured against max(|A x|, 1e-2 |M x| / h^2) ~ the size of the element-level terms that cancel
    floor = 1e-2 * np.linalg.norm(ops.apply_M(x)) * (n * p) ** 2
    assert np.linalg.norm(eq.MultA(x) - ref) / max(np.linalg.norm(ref), floor) < TOL


@pytest.mark.parametrize("name,n,p", CASES[:9])
def test_projector_pieces(bloch, name, n, p):
    eq, ops, rng = _setup(bloch, name, n, p, seed=2)
    N, N0 = eq.N, eq.N_h1
    G, M1, S0 = ops.G_c(), ops.M_c(), ops.S0_c()
    ph = rng.uniform(-1, 1, (3, 2 * N0))
    phc = ops.to_c(ph, N0)
    assert rel_err(ops.to_c(eq.debug_h1op(1, ph), N), (G @ phc.T).T) < TOL
    assert rel_err(ops.to_c(eq.debug_h1op(0, ph), N0), (S0 @ phc.T).T) < TOL
    x = rng.uniform(-1, 1, (3, 2 * N))
    xc = ops.to_c(x, N)
    assert rel_err(ops.to_c(eq.debug_h1op(2, x), N0), (G.conj().T @ (M1 @ xc.T)).T) < TOL
    # S0 assembled the reference way equals G^H M G
    assert abs(S0 - G.conj().T @ M1 @ G).max() < 1e-11 * abs(S0).max()


@pytest.mark.parametrize("name,n,p", [("CUB", 3, 1), ("FCC", 2, 2), ("BCC", 1, 2)])
def test_curl_operator(bloch, name, n, p):
    eq, ops, rng = _setup(bloch, name, n, p, seed=3)
    x = rng.uniform(-1, 1, (2, 2 * eq.N))
    y = ops.to_c(eq.MultC(x), eq.N_rt)
    ref = (ops.C_c() @ ops.to_c(x, eq.N).T).T
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize("name,n,p", [("CUB", 3, 1), ("FCC", 2, 2), ("BCC", 1, 2), ("CUB", 2, 3)])
def test_projector(bloch, name, n, p):
    """P^2 = P, G^H M P x = 0, and parity with the oracle projector (maxwell_bloch.cpp:2280-2290)."""
    eq, ops, rng = _setup(bloch, name, n, p, seed=4)
    N = eq.N
    x = rng.uniform(-1, 1, (3, 2 * N))
    y = eq.MultProjector(x)
    ref = ops.apply_projector(ops.to_c(x, N))
    assert rel_err(ops.to_c(y, N), ref) < 1e-9
    y2 = eq.MultProjector(y)
    assert rel_err(y2, y) < 1e-9
    gm = eq.debug_h1op(2, y)
    assert np.abs(gm).max() < 1e-9 * np.abs(eq.debug_h1op(2, x)).max()
