"""Reduced-basis k-sweep (MaxwellDispersion::buildRawBasis / approxEigenfrequencies /
traverseBrillouinZone, meta-material/meta_material_solver.cpp:3132-3410) against the oracle doing the
same Rayleigh-Ritz on the same raw basis, plus the properties the construction guarantees."""
import numpy as np
import pytest
import scipy.linalg as sla

from helpers import oracle_on_product_maps

pytestmark = pytest.mark.gpu


def _collect(eq, ks, nb):
    eq.SetNumEigs(2 * nb)
    eq.SetAbsoluteTolerance(1e-9)
    eq.ReducedBasisClear()
    raw, full = [], []
    for k in ks:
        eq.SetKappa(k)
        eq.Setup()
        eq.Solve()
        full.append(eq.band_eigenvalues())
        eq.ReducedBasisAppend()
        for i in range(nb):
            er, ei = eq.GetEigenvectorE(i)
            raw.append(er + 1j * ei)
    return np.array(raw), np.array(full)


def test_rb_matches_oracle_rayleigh_ritz(bloch):
    name, n, p, nb = "FCC", 2, 2, 4
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    eps = bloch.sphere_eps(eq.element_centers(), 0.3, 9.0, 1.0)
    eq.SetMassCoef(eps)
    X, G = (L.GetSymmetryPoint(L.GetSymmetryPointIndex(s)) for s in ("X", "Gamma"))
    ks = [X, 0.5 * X, 0.25 * X + np.array([0.0, 0.3, 0.1])]
    raw, full = _collect(eq, ks, nb)
    assert eq.ReducedBasisSize() == len(ks) * nb
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps)
    for kq in (0.75 * X, 0.6 * X + np.array([0.1, 0.05, 0.0])):
        got = eq.ApproxEigenvalues(kq, nb)
        ops.set_kappa(kq)
        P = ops.apply_projector(raw)                              # [K, N]
        A, M = ops.A_c(), ops.M_c()
        GA, GM = P.conj() @ (A @ P.T), P.conj() @ (M @ P.T)
        GA, GM = 0.5 * (GA + GA.conj().T), 0.5 * (GM + GM.conj().T)
        # same whitening as the product (drop directions below 1e-10 of the largest M-norm)
        d = 1.0 / np.sqrt(np.real(np.diag(GM)))
        w, V = np.linalg.eigh(d[:, None] * GM * d[None, :])
        keep = w > 1e-10 * w[-1]
        T = (d[:, None] * V[:, keep]) / np.sqrt(w[keep])
        ref = sla.eigh(T.conj().T @ GA @ T, eigvals_only=True)[:nb]
        assert np.allclose(got, ref, rtol=1e-6, atol=1e-8), (got, ref)
        # Rayleigh-Ritz in a subspace of the constrained space: upper bounds of the true bands
        true = ops.eig_dense(nb)
        assert (got >= true - 1e-7).all()
        assert np.abs(got - true).max() < 0.05 * true.max()


def test_rb_exact_at_basis_points_and_traverse(bloch):
    L = bloch.BravaisLattice("CUB")
    eq = bloch.MaxwellBlochWaveEquation(L, 4, 1)
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers(), 0.3, 6.0, 1.0))
    nb = 4
    md = bloch.MaxwellDispersion(eq, L, nb, samp_pow=2, mid_pts=True, tol=1e-9)
    segs = md.traverseBrillouinZone()
    n_sp = len(md.sp_eigs)
    assert md.eq.ReducedBasisSize() == n_sp * nb
    assert len(segs) == L.GetNumberPaths()
    # at a k-point whose eigenvectors are in the basis the approximation is the full solve
    e0, e1 = L.GetPathSegmentEndPointIndices(0, 0)
    lab = L.GetSymmetryPointLabel(e1)
    om = md.approxEigenfrequencies(L.GetSymmetryPoint(e1))
    assert np.allclose(om, md.sp_eigs[lab], atol=2e-5), (om, md.sp_eigs[lab])
    # every sample is filled, omega >= 0, and quarter points stay close to full solves
    for p in range(L.GetNumberPaths()):
        for s in range(L.GetNumberPathSegments(p)):
            row = segs[p][s]
            assert len(row) == 5 and all(r is not None and len(r) == nb and (r >= 0).all() for r in row)
    k0, k1 = L.GetSymmetryPoint(e0), L.GetSymmetryPoint(e1)
    kq = 0.75 * k0 + 0.25 * k1
    full = np.sqrt(np.abs(eq.GetEigenvalues(2 * nb, kq)[0::2]))
    assert np.abs(segs[0][0][1] - full).max() < 0.03 * full.max()
