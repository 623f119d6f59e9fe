"""Band eigenvalues of the CUDA block LOBPCG (through the C ABI) against the oracle's
independent eigen-solves.  Tolerance: relative 1e-6 on lambda (north star: band frequencies to
relative 1e-6), tested at 1e-7 with solver tolerance 1e-8."""
import numpy as np
import pytest

from helpers import oracle_on_product_maps
from oracle.bloch_oracle import Lattice, empty_lattice_eigs

pytestmark = pytest.mark.gpu


def _eq(bloch, name, n, p, eps_fn=None, device=-1):
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p, device)
    eps = bloch.sphere_eps(eq.element_centers()) if eps_fn is None else eps_fn(eq)
    eq.SetMassCoef(eps)
    return L, eq, eps


@pytest.mark.parametrize("name,n,p,nb", [("CUB", 3, 1, 6), ("CUB", 2, 2, 8), ("FCC", 2, 1, 4),
                                          ("FCC", 2, 2, 10), ("BCC", 1, 2, 8), ("CUB", 2, 3, 10), ("HEX", 2, 2, 8)])
def test_bands_match_dense_constrained_pencil(bloch, name, n, p, nb):
    L, eq, eps = _eq(bloch, name, n, p)
    kappa = np.array([0.7, -0.4, 1.1])
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps)
    ops.set_kappa(kappa)
    ref = ops.eig_dense(nb)
    eq.SetAbsoluteTolerance(1e-9)
    lam = eq.GetEigenvalues(2 * nb, kappa)
    assert lam.shape == (2 * nb,)
    assert np.allclose(lam[0::2], lam[1::2])            # complex bands reported twice
    assert np.allclose(lam[0::2], ref, rtol=1e-7, atol=1e-9), (lam[0::2], ref)
    st = eq.GetSolverStats()
    assert st["converged_bands"] == nb


def test_empty_lattice_first_band_exact(bloch):
    """eps = mu = 1: the lowest pair is |kappa|^2 exactly (constant envelope is in the space)."""
    L, eq, _ = _eq(bloch, "FCC", 2, 2, eps_fn=lambda e: np.ones(e.n_elem))
    kappa = 0.3 * L.GetSymmetryPoint(L.GetSymmetryPointIndex("L"))
    eq.SetAbsoluteTolerance(1e-10)
    lam = eq.GetEigenvalues(8, kappa)[0::2]
    assert abs(lam[0] - kappa @ kappa) < 1e-9 * (kappa @ kappa)
    assert abs(lam[1] - kappa @ kappa) < 1e-9 * (kappa @ kappa)
    ex = empty_lattice_eigs(Lattice("FCC"), kappa, 4)
    assert np.all(lam[2:] > ex[2:] * 0.999)             # conforming: higher bands from above


def test_kappa_symmetry_and_gamma(bloch):
    L, eq, eps = _eq(bloch, "CUB", 4, 1)
    eq.SetAbsoluteTolerance(1e-9)
    k = np.array([0.9, 0.2, -0.5])
    a = eq.GetEigenvalues(12, k)[0::2]
    b = eq.GetEigenvalues(12, -k)[0::2]
    assert np.allclose(a, b, rtol=1e-7)                 # spectrum(kappa) = spectrum(-kappa)
    # cubic symmetry: X along x, y, z
    xs = [eq.GetEigenvalues(12, np.pi * np.eye(3)[d])[0::2] for d in range(3)]
    assert np.allclose(xs[0], xs[1], rtol=1e-7) and np.allclose(xs[0], xs[2], rtol=1e-7)
    # Gamma: three harmonic zero modes, then the real spectrum (maxwell_bloch.cpp:561-582)
    g = eq.GetEigenvalues(16, np.zeros(3))[0::2]
    assert np.all(np.abs(g[:3]) < 1e-8)
    ops, _ = oracle_on_product_maps(eq, "CUB", 4, 1, eps)
    ops.set_kappa(np.zeros(3))
    ref = ops.eig_dense(8)
    assert np.allclose(g[3:], ref[3:], rtol=1e-7), (g, ref)


def test_fcc_sphere_against_shift_invert(bloch):
    """moderate size (FCC p=2 n=4, N = 6144): ARPACK shift-invert + sparse LU on the assembled
    oracle matrices, an algorithm unrelated to LOBPCG"""
    L, eq, eps = _eq(bloch, "FCC", 4, 2)
    kappa = L.GetSymmetryPoint(L.GetSymmetryPointIndex("X")) * 0.5
    eq.SetAbsoluteTolerance(1e-8)
    lam = eq.GetEigenvalues(20, kappa)[0::2]
    ops, _ = oracle_on_product_maps(eq, "FCC", 4, 2, eps)
    ops.set_kappa(kappa)
    w = ops.eig_shift_invert(10, sigma=0.6 * lam[-1], extra=14)
    w = w[w > 1e-6 * lam[-1]]
    # every GPU band must be found by the oracle
    for l in lam:
        assert np.min(np.abs(w - l)) < 1e-6 * l, (l, w)


def test_eigenvectors_satisfy_pencil(bloch):
    L, eq, eps = _eq(bloch, "FCC", 2, 2)
    kappa = np.array([1.0, 0.5, 0.25])
    eq.SetAbsoluteTolerance(1e-9)
    lam = eq.GetEigenvalues(8, kappa)[0::2]
    for i in range(4):
        re, im = eq.GetEigenvectorE(i)
        x = np.concatenate([re, im])
        r = eq.MultA(x) - lam[i] * eq.MultM(x)
        assert np.linalg.norm(r) < 1e-7
        assert abs(x @ eq.MultM(x) - 1.0) < 1e-8       # M-normalised
        assert np.abs(eq.debug_h1op(2, x)).max() < 1e-7  # divergence constraint
        # B = C E / sqrt(lambda), Bi = block 0, Br = -block 1 (maxwell_bloch.cpp:1432-1457)
        br, bi = eq.GetEigenvectorB(i)
        ce = eq.MultC(x)
        assert np.allclose(bi, ce[: eq.N_rt] / np.sqrt(lam[i]), atol=1e-10)
        assert np.allclose(br, -ce[eq.N_rt:] / np.sqrt(lam[i]), atol=1e-10)


def test_golden_fixtures(bloch):
    """committed oracle fixtures (tests/golden/make_golden.py): no oracle code runs here"""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bands_small.json")))
    for c in cases:
        L = bloch.BravaisLattice(c["lattice"])
        eq = bloch.MaxwellBlochWaveEquation(L, c["n_sub"], c["order"])
        assert eq.N == c["n_nd"]
        eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()) if c["sphere"] else np.ones(eq.n_elem))
        nb = len(c["eigenvalues"])
        # residual 1e-7 => eigenvalue error ~1e-14 relative (quadratic); the tiny Gamma-point case
        # (12-fold cluster cut by the block) is not reliably driven below 1e-8, see DESIGN.md section 6
        eq.SetAbsoluteTolerance(1e-7)
        lam = eq.GetEigenvalues(2 * nb, np.array(c["kappa"]))[0::2]
        assert np.allclose(lam, c["eigenvalues"], rtol=1e-7, atol=1e-8), (c["lattice"], c["order"], lam, c["eigenvalues"])


@pytest.mark.parametrize("name,n,p", [("FCC", 2, 2), ("CUB", 4, 1), ("BCC", 2, 1)])
def test_warm_started_path_matches_oracle_at_every_point(bloch, name, n, p):
    """A k-path on ONE handle: from the second point on the solver warm-starts from the previous
    eigenvectors and runs its gradient-lifted iteration (roughly projected search directions, lifted
    pencil A + tau M G B G^H M).  Every point, including Gamma reached from a neighbour, must give the
    eigenvalues of the constrained pencil and divergence-free, M-orthonormal eigenvectors."""
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eps = bloch.sphere_eps(eq.element_centers(), 0.3, 7.0, 1.0)
    eq.SetMassCoef(eps)
    nb = 4
    eq.SetNumEigs(2 * nb)
    eq.SetAbsoluteTolerance(1e-8)
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps)
    X = L.GetSymmetryPoint(1)
    path = [X, 0.7 * X, 0.4 * X + np.array([0.1, 0.0, 0.2]), 0.1 * X, np.zeros(3), 0.5 * X]
    for k in path:
        eq.SetKappa(k)
        eq.Setup()
        eq.Solve()
        lam = eq.band_eigenvalues()
        ref = ops.set_kappa(k).eig_dense(nb)
        assert np.allclose(lam, ref, rtol=1e-7, atol=1e-7), (k, lam, ref)
        G, M = ops.G_c(), ops.M_c()
        for i in range(nb):
            er, ei = eq.GetEigenvectorE(i)
            x = er + 1j * ei
            assert np.abs(G.conj().T @ (M @ x)).max() < 1e-7          # divergence constraint
            r = ops.A_c() @ x - lam[i] * (M @ x)
            assert np.linalg.norm(r) < 5e-7 * max(1.0, np.linalg.norm(M @ x))
