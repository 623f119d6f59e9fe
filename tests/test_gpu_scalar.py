"""Scalar H1 Bloch Helmholtz variant (BASELINE config 5, misc/scalar3d.cpp): operator parity 1e-12
against the assembled oracle blocks and eigenvalues against the dense pencil (rel 1e-7)."""
import numpy as np
import pytest

from helpers import rel_err
from oracle.bloch_oracle import Lattice, Mesh, ScalarOperators, Spaces

pytestmark = pytest.mark.gpu


def _coefs(centers, radius=0.5):
    """mass_coef / stiffness_coef of scalar3d.cpp:564-588 sampled at element centres"""
    inside = np.linalg.norm(centers, axis=1) <= radius
    return np.where(inside, 5.0, 0.1), np.where(inside, 10.0, 1.0)      # k, m


def _pair(bloch, name, n, p):
    L = bloch.BravaisLattice(name)
    eq = bloch.ScalarFloquetWaveEquation(L, n, p)
    k, m = _coefs(eq.element_centers(), 0.3)
    eq.SetStiffnessCoef(k)
    eq.SetMassCoef(m)
    mesh = Mesh(Lattice(name), n)
    x0, cls, J = eq.element_geometry()
    assert np.allclose(x0, mesh.x0) and (cls == mesh.cls).all()
    hg = eq.dofmap()
    ng, ns = eq._eq.dofmap("nd") if p <= 3 else (None, None)
    maps = dict(h1_gid=hg)
    if p <= 3:
        maps.update(nd_gid=ng, nd_sign=ns)
        sp = Spaces(mesh, p, dofmaps=maps)
    else:                                   # order 4: the oracle builds its own ND numbering
        sp = Spaces(mesh, p)
        sp.h1_gid = hg
    return eq, ScalarOperators(sp, k, m)


@pytest.mark.parametrize("name,n,p", [("CUB", 3, 1), ("FCC", 2, 2), ("BCC", 1, 3), ("CUB", 2, 4), ("BCC", 1, 4)])
def test_scalar_operator_parity(bloch, name, n, p):
    eq, ops = _pair(bloch, name, n, p)
    zeta = np.array([1.0, 2.0, -0.5]) / np.linalg.norm([1.0, 2.0, -0.5])
    eq.SetZeta(zeta)
    eq.SetBeta(75.0)                                               # degrees
    eq.Setup()
    ops.set_kappa(75.0 * np.pi / 180.0 * zeta)
    x = np.random.default_rng(5).uniform(-1, 1, (3, 2 * eq.N))
    assert rel_err(eq.MultA(x), (ops.A_block() @ x.T).T) < 1e-12
    assert rel_err(eq.MultM(x), (ops.M_block() @ x.T).T) < 1e-12


@pytest.mark.parametrize("name,n,p,nm", [("CUB", 3, 2, 5), ("BCC", 1, 2, 6), ("CUB", 2, 4, 10)])
def test_scalar_eigenvalues(bloch, name, n, p, nm):
    eq, ops = _pair(bloch, name, n, p)
    kappa = np.array([0.8, -0.3, 0.5])
    eq.SetKappa(kappa)
    eq.SetNumEigs(2 * nm)
    eq.SetAbsoluteTolerance(1e-9)
    eq.Setup()
    eq.Solve()
    lam = eq.mode_eigenvalues()
    ref = ops.set_kappa(kappa).eig_dense(nm)
    assert np.allclose(lam, ref, rtol=1e-7, atol=1e-9), (lam, ref)
    ev = eq.GetEigenvalues()
    assert len(ev) == 2 * nm and np.allclose(ev[0::2], ev[1::2])


@pytest.mark.parametrize("name,n,p,nm,kappa", [("CUB", 4, 1, 6, (0.8, -0.3, 0.5)), ("FCC", 2, 2, 6, (0.8, -0.3, 0.5)),
                                               ("CUB", 4, 3, 8, (0.2, 0.1, 0.0)), ("CUB", 4, 2, 6, (0.0, 0.0, 0.0))])
def test_scalar_eigenvalues_with_multigrid_preconditioner(bloch, name, n, p, nm, kappa):
    """Even n_sub: the scalar solve is preconditioned by one V-cycle of the H1 multigrid (the reference: BoomerAMG sweeps,
    misc/scalar3d.cpp:70-85).  Same eigenvalues as with the Chebyshev polynomial (BLOCH_SCALAR_MG=0) and as the oracle's
    dense pencil; the outer iteration count stays in the same range (a V-cycle costs 4 fine-level applies, the polynomial 23)."""
    import os
    res = {}
    for mode in ("1", "0"):
        os.environ["BLOCH_SCALAR_MG"] = mode
        try:
            eq, ops = _pair(bloch, name, n, p)
            eq.SetKappa(np.array(kappa))
            eq.SetNumEigs(2 * nm)
            eq.SetAbsoluteTolerance(1e-9)
            eq.Setup()
            eq.Solve()
            res[mode] = (eq.mode_eigenvalues().copy(), eq.GetSolverStats()["iterations"])
        finally:
            del os.environ["BLOCH_SCALAR_MG"]
    ref = ops.set_kappa(np.array(kappa)).eig_dense(nm)
    assert np.allclose(res["1"][0], ref, rtol=1e-7, atol=1e-8), (res["1"][0], ref)
    assert np.allclose(res["1"][0], res["0"][0], rtol=1e-7, atol=1e-8)
    assert res["1"][1] <= 2 * res["0"][1], (res["1"][1], res["0"][1])
