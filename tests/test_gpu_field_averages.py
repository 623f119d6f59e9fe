"""GetFieldAverages (maxwell/maxwell_bloch.cpp:1550-1632, linear forms of SetKappa :211-279): the
device quadrature of the Bloch fields against the oracle's literally assembled linear forms, on the
product's own eigenvectors."""
import numpy as np
import pytest

from helpers import oracle_on_product_maps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n,p", [("CUB", 3, 1), ("FCC", 2, 2), ("BCC", 1, 2), ("CUB", 2, 3)])
def test_field_averages_match_linear_forms(bloch, name, n, p):
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    rng = np.random.default_rng(11)
    eps, mui = rng.uniform(1, 6, eq.n_elem), rng.uniform(0.5, 2.0, eq.n_elem)
    eq.SetMassCoef(eps)
    eq.SetStiffnessCoef(mui)
    k = np.array([0.7, -0.4, 0.2])
    nb = 3
    eq.SetNumEigs(2 * nb)
    eq.SetAbsoluteTolerance(1e-9)
    eq.SetKappa(k)
    eq.Setup()
    eq.Solve()
    lam = eq.band_eigenvalues()
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps, mui)
    ops.set_kappa(k)
    for i in range(nb):
        er, ei = eq.GetEigenvectorE(i)
        ref = ops.field_averages(er + 1j * ei, lam[i])
        got = eq.GetFieldAverages(i)
        for key in "EBDH":
            scale = max(np.abs(ref[key]).max(), 1e-12)
            assert np.abs(got[key] - ref[key]).max() < 1e-10 * max(scale, 1.0), (key, got[key], ref[key])
        # B from the product's own B eigenvector and the B linear forms
        br, bi = eq.GetEigenvectorB(i)
        Lf = ops.field_average_forms()
        assert np.allclose(Lf["B"] @ (br + 1j * bi), got["B"], atol=1e-10)


def test_field_averages_plane_wave_limit(bloch):
    """Empty lattice, lowest band: the full Bloch field is the plane wave e^{i kappa.x} E0 (up to the
    discretisation), so |<E>| ~ V |E0| and <D> = <E>, <H> = <B>, kappa . <E> ~ 0, and <B> ~ kappa x <E> / omega
    up to the reference's sign/phase convention (|<B>| = |kappa x <E>| / omega)."""
    L = bloch.BravaisLattice("CUB")
    eq = bloch.MaxwellBlochWaveEquation(L, 4, 2)
    k = np.array([0.3, 0.2, -0.1])
    eq.SetNumEigs(4)
    eq.SetAbsoluteTolerance(1e-10)
    eq.SetKappa(k)
    eq.Setup()
    eq.Solve()
    lam = eq.band_eigenvalues()
    assert abs(lam[0] - k @ k) < 1e-6
    a = eq.GetFieldAverages(0)
    assert np.allclose(a["D"], a["E"]) and np.allclose(a["H"], a["B"])
    assert abs(k @ a["E"]) < 1e-5 * np.linalg.norm(a["E"])
    kxE = np.cross(k, a["E"]) / np.sqrt(lam[0])
    assert abs(np.linalg.norm(kxE) - np.linalg.norm(a["B"])) < 1e-4 * np.linalg.norm(a["B"])
