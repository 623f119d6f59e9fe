"""The auxiliary-space preconditioner (csrc/aux.cu; the job HypreAMS does for the reference,
maxwell/maxwell_bloch.cpp:492-517): Pi / Pi^T against the oracle's assembled nodal interpolation, the component
V-cycles against the assembled Bloch Laplacian, and the eigen-solver with it against the Chebyshev-only solver."""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spl

from helpers import oracle_on_product_maps, rel_err
from oracle.bloch_oracle import BlochOperators, nodal_interpolation

pytestmark = pytest.mark.gpu


def _setup(bloch, name, n, p, seed=0, kappa=None, eps_const=False):
    rng = np.random.default_rng(seed)
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eps = np.ones(eq.n_elem) if eps_const else rng.uniform(1.0, 10.0, eq.n_elem)
    mui = rng.uniform(0.5, 2.0, eq.n_elem)
    eq.SetMassCoef(eps)
    eq.SetStiffnessCoef(mui)
    kappa = rng.normal(size=3) * 2.0 if kappa is None else np.asarray(kappa, float)
    eq.SetKappa(kappa)
    eq.Setup()
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps, mui)
    ops.set_kappa(kappa)
    return eq, ops, rng, mui


@pytest.mark.parametrize("name,n,p", [("CUB", 2, 1), ("FCC", 2, 2), ("BCC", 2, 3), ("HEX", 2, 2), ("FCC", 4, 1)])
def test_pi_and_its_transpose_match_the_assembled_interpolation(bloch, name, n, p):
    eq, ops, rng, _ = _setup(bloch, name, n, p)
    N, N0 = eq.N, eq.N_h1
    Pi = nodal_interpolation(ops.sp_)
    for nvec in (1, 4):
        u = rng.uniform(-1, 1, (nvec, 2 * 3 * N0))
        uc = ops.to_c(u, 3 * N0)
        assert rel_err(ops.to_c(eq.debug_aux(0, u), N), (Pi @ uc.T).T) < 1e-13
        x = rng.uniform(-1, 1, (nvec, 2 * N))
        xc = ops.to_c(x, N)
        assert rel_err(ops.to_c(eq.debug_aux(1, x), 3 * N0), (Pi.T @ xc.T).T) < 1e-13


@pytest.mark.parametrize("name,n,p", [("FCC", 4, 2), ("CUB", 4, 1), ("BCC", 2, 3)])
def test_component_vcycle_is_hermitian_and_spectrally_equivalent(bloch, name, n, p):
    """B = one V-cycle for L = (grad + i kappa)^H mu^-1 (grad + i kappa) per Cartesian component: a fixed Hermitian
    positive definite operator with eig(B L) in a mesh-independent interval below ~1."""
    eq, ops, rng, mui = _setup(bloch, name, n, p, seed=1)
    N0 = eq.N_h1
    # the oracle's S0 with mu^-1 in the place of eps is the auxiliary operator
    opsL = BlochOperators(ops.sp_, mui)
    opsL.set_kappa(ops.kappa)
    Lmat = opsL.S0_c().tocsc()
    m = 6
    u = rng.uniform(-1, 1, (m, 2 * 3 * N0))
    Bu = ops.to_c(eq.debug_aux(2, u), 3 * N0)
    uc = ops.to_c(u, 3 * N0)
    Gm = uc.conj() @ Bu.T                                  # <u_i, B u_j>
    assert np.abs(Gm - Gm.conj().T).max() < 1e-10 * np.abs(Gm).max()
    assert np.linalg.eigvalsh(0.5 * (Gm + Gm.conj().T)).min() > 0
    # the components do not mix and each equals the scalar V-cycle: Rayleigh quotients of B against L^-1
    lu = spl.splu(Lmat)
    for d in range(3):
        ud = uc[:, d * N0:(d + 1) * N0]
        Bd = Bu[:, d * N0:(d + 1) * N0]
        Li = lu.solve(np.ascontiguousarray(ud.T)).T
        q = np.real(np.sum(ud.conj() * Bd, axis=1)) / np.real(np.sum(ud.conj() * Li, axis=1))
        assert q.min() > 0.15 and q.max() < 1.2, q
    # only component d of the input reaches component d of the output
    u1 = np.zeros((1, 2 * 3 * N0))
    u1[0, :N0] = rng.uniform(-1, 1, N0)
    B1 = ops.to_c(eq.debug_aux(2, u1), 3 * N0)
    assert np.abs(B1[0, N0:]).max() == 0.0


@pytest.mark.parametrize("name,n,p,kappa", [("FCC", 4, 2, (1.3, 0.4, -0.7)), ("CUB", 4, 1, (0.0, 0.0, 0.0)),
                                            ("BCC", 2, 3, (2.0, 0.5, 0.3)), ("HEX", 2, 2, (0.9, 0.1, 0.6)),
                                            ("CUB", 2, 3, (0.0, 0.0, 0.0))])
def test_solver_with_auxiliary_space_matches_chebyshev_only(bloch, name, n, p, kappa):
    """Same eigenvalues whichever preconditioner iterates (they do not depend on it), fewer or equal outer iterations
    with the auxiliary space, and parity with the oracle's dense constrained pencil where that is affordable."""
    nb = 6
    res = {}
    for mode in ("aux", "cheb"):
        os.environ["BLOCH_PRECOND"] = mode
        try:
            L = bloch.BravaisLattice(name)
            eq = bloch.MaxwellBlochWaveEquation(L, n, p)
        finally:
            del os.environ["BLOCH_PRECOND"]
        eps = bloch.sphere_eps(eq.element_centers())
        eq.SetMassCoef(eps)
        eq.SetKappa(np.asarray(kappa, float))
        eq.SetNumEigs(2 * nb)
        eq.SetAbsoluteTolerance(1e-8)
        eq.Setup()
        eq.Solve()
        st = eq.GetSolverStats()
        assert st["converged_bands"] == nb
        res[mode] = (eq.band_eigenvalues().copy(), st["iterations"])
        if mode == "aux":
            ops, _ = oracle_on_product_maps(eq, name, n, p, eps)
            ops.set_kappa(np.asarray(kappa, float))
    lam_a, it_a = res["aux"]
    lam_c, it_c = res["cheb"]
    scale = max(1.0, np.abs(lam_c).max())
    assert np.abs(lam_a - lam_c).max() < 1e-8 * scale
    assert it_a <= it_c + 2, (it_a, it_c)
    if ops.sp_.n_nd <= 7000:
        assert np.abs(lam_a - ops.eig_dense(nb)).max() < 1e-6 * scale


def test_batched_kpoints_with_auxiliary_space(bloch):
    """k-point batch incl. a Gamma point: the batched solve returns what single solves return."""
    L = bloch.BravaisLattice("FCC")
    ks = np.array([[0.0, 0.0, 0.0], [1.1, 0.3, 0.2], [2.5, 0.0, 1.0]])
    eq = bloch.MaxwellBlochWaveEquation(L, 4, 2)
    eps = bloch.sphere_eps(eq.element_centers())
    eq.SetMassCoef(eps)
    eq.SetNumEigs(12)
    eq.SetAbsoluteTolerance(1e-8)
    lam_b, st = eq.SolveBatch(ks)
    assert all(s["converged_bands"] == 6 for s in st)
    for i, k in enumerate(ks):
        e1 = bloch.MaxwellBlochWaveEquation(L, 4, 2)
        e1.SetMassCoef(eps)
        e1.SetNumEigs(12)
        e1.SetAbsoluteTolerance(1e-8)
        e1.SetKappa(k)
        e1.Setup()
        e1.Solve()
        assert np.abs(e1.band_eigenvalues() - lam_b[i]).max() < 1e-7 * max(1.0, np.abs(lam_b[i]).max())


@pytest.mark.parametrize("name,n,p", [("CUB", 2, 1), ("FCC", 4, 2), ("BCC", 2, 3), ("HEX", 2, 2), ("CUB", 4, 3), ("FCC", 2, 3)])
def test_multigrid_transfers_three_implementations_agree(bloch, name, n, p):
    """Nested-mesh prolongation / restriction of the H1 multigrid: explicit CSR matrices, sum-factorised parent
    kernels and element-wise kernels are the same linear maps (1e-13), restriction is the transpose of the
    prolongation, constants are reproduced (n = 2 puts periodic images of a dof inside one parent element)."""
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eq.SetKappa(np.array([0.3, 0.2, 0.1]))
    eq.Setup()
    rng = np.random.default_rng(3)
    nf, nc = eq.N_h1, eq.mg_coarse_size()
    xc = rng.uniform(-1, 1, (3, 2 * nc))
    xf = rng.uniform(-1, 1, (3, 2 * nf))
    pro = [eq.debug_mg_transfer(v, 0, xc) for v in range(3)]
    res = [eq.debug_mg_transfer(v, 1, xf) for v in range(3)]
    for v in (1, 2):
        assert rel_err(pro[v], pro[0]) < 1e-13
        assert rel_err(res[v], res[0]) < 1e-13
    # <P xc, xf> = <xc, P^T xf> (real weights: re and im parts separately)
    assert abs(np.sum(pro[1] * xf) - np.sum(xc * res[1])) < 1e-11 * abs(np.sum(pro[1] * xf))
    one = np.zeros((1, 2 * nc))
    one[0, :nc] = 1.0
    up = eq.debug_mg_transfer(1, 0, one)
    assert np.abs(up[0, :nf] - 1.0).max() < 1e-13 and np.abs(up[0, nf:]).max() == 0.0
