"""k-point batching (bloch_set_kappa_batch): nk Bloch vectors iterated together must give exactly what nk separate
GetEigenvalues calls give (the reference solves the k-points one after another, maxwell_dispersion.cpp:475-531).
Compared: operator / projector applies column group by column group against single-kappa handles (1e-13) and band
eigenvalues against single-kappa solves and against the oracle's dense constrained pencil (relative 1e-7), with a
Gamma point (singular S0, real AME branch of the reference) inside the batch."""
import numpy as np
import pytest

from helpers import oracle_on_product_maps, rel_err

pytestmark = pytest.mark.gpu


def _eq(bloch, name, n, p):
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eps = bloch.sphere_eps(eq.element_centers())
    eq.SetMassCoef(eps)
    return L, eq, eps


KAPPAS = np.array([[0.7, -0.4, 1.1], [0.0, 0.0, 0.0], [0.0, 3.0, 0.0], [2.1, 2.1, 2.1], [0.3, 0.0, -0.2]])


@pytest.mark.parametrize("name,n,p", [("FCC", 2, 2), ("CUB", 4, 1), ("BCC", 2, 3)])
def test_batched_applies_match_single_kappa(bloch, name, n, p):
    L, eq, eps = _eq(bloch, name, n, p)
    _, ref, _ = _eq(bloch, name, n, p)
    nk, c = len(KAPPAS), 3
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (nk * c, 2 * eq.N))
    eq.SetKappaBatch(KAPPAS)
    eq.Setup()
    ya, ym, yp = eq.MultA(x), eq.MultM(x), eq.MultProjector(x)
    for k in range(nk):
        ref.SetKappa(KAPPAS[k])
        ref.Setup()
        sl = slice(k * c, (k + 1) * c)
        assert rel_err(ya[sl], ref.MultA(x[sl])) < 1e-13
        assert rel_err(ym[sl], ref.MultM(x[sl])) < 1e-13
        assert rel_err(yp[sl], ref.MultProjector(x[sl])) < 1e-9


@pytest.mark.parametrize("name,n,p,nb", [("FCC", 2, 2, 8), ("CUB", 4, 1, 6)])
def test_batched_solve_matches_single_solves_and_oracle(bloch, name, n, p, nb):
    L, eq, eps = _eq(bloch, name, n, p)
    _, ref, _ = _eq(bloch, name, n, p)
    for e in (eq, ref):
        e.SetNumEigs(2 * nb)
        e.SetAbsoluteTolerance(1e-9)
    lam, stats = eq.SolveBatch(KAPPAS)
    assert lam.shape == (len(KAPPAS), nb)
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps)
    for k, kap in enumerate(KAPPAS):
        assert stats[k]["converged_bands"] == nb
        single = ref.GetEigenvalues(2 * nb, kap)[0::2]
        assert np.allclose(lam[k], single, rtol=1e-7, atol=1e-8), (k, lam[k], single)
        ops.set_kappa(kap)
        if np.linalg.norm(kap) > 0:
            dense = ops.eig_dense(nb)
            assert np.allclose(lam[k], dense, rtol=1e-7, atol=1e-8), (k, lam[k], dense)
        else:
            assert np.all(np.abs(lam[k][:3]) < 1e-7)      # three harmonic zero modes at Gamma
    # a second batch on the same handle warm-starts every k-point from its predecessor in the same slot
    lam2, stats2 = eq.SolveBatch(KAPPAS * 1.05)
    for k, kap in enumerate(KAPPAS * 1.05):
        single = ref.GetEigenvalues(2 * nb, kap)[0::2]
        assert np.allclose(lam2[k], single, rtol=1e-7, atol=1e-8), (k, lam2[k], single)


def test_batched_eigenvectors_and_batch_size_changes(bloch):
    name, n, p, nb = "FCC", 2, 2, 6
    L, eq, eps = _eq(bloch, name, n, p)
    _, ref, _ = _eq(bloch, name, n, p)
    eq.SetNumEigs(2 * nb)
    eq.SetAbsoluteTolerance(1e-9)
    ks = KAPPAS[[0, 2, 3]]
    lam, _ = eq.SolveBatch(ks)
    for k in range(3):
        eq.SelectKPoint(k)
        ref.SetKappa(ks[k])
        ref.Setup()
        for i in (0, nb - 1):
            er, ei = eq.GetEigenvectorE(i)
            x = np.concatenate([er, ei])
            r = ref.MultA(x) - lam[k][i] * ref.MultM(x)
            assert np.linalg.norm(r) < 1e-7, (k, i, np.linalg.norm(r))
            br, bi = eq.GetEigenvectorB(i)
            cx = ref.MultC(x) / np.sqrt(abs(lam[k][i]))
            assert np.allclose(bi, cx[: eq.N_rt], atol=1e-9) and np.allclose(br, -cx[eq.N_rt:], atol=1e-9)
    with pytest.raises(bloch.BlochError):
        eq.SelectKPoint(3)
    # shrinking / growing the batch on the same handle (graphs, probes, work space follow)
    one = eq.GetEigenvalues(2 * nb, ks[1])[0::2]
    assert np.allclose(one, lam[1], rtol=1e-7)
    lam5, _ = eq.SolveBatch(KAPPAS)
    assert np.allclose(lam5[[0, 2, 3]], lam, rtol=1e-7, atol=1e-8)
