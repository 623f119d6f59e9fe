"""Sweep history (csrc/solver.cu): along a straight walk the eigenvectors of the last-but-one k-point join the first
Rayleigh-Ritz as the P block.  The k-points stay independent eigenproblems (maxwell_dispersion.cpp:475-531): same
eigenvalues with and without the history, fewer outer iterations with it."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _walk(bloch, kappas, history, batch=None):
    os.environ["BLOCH_HISTORY"] = "1" if history else "0"
    try:
        L = bloch.BravaisLattice("FCC")
        eq = bloch.MaxwellBlochWaveEquation(L, 4, 2)
        eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
        eq.SetNumEigs(12)
        eq.SetAbsoluteTolerance(1e-8)
        lams, its = [], []
        for k in kappas:
            if batch is None:
                eq.SetKappa(k)
                eq.Setup()
                eq.Solve()
                st = eq.GetSolverStats()
                assert st["converged_bands"] == 6
                lams.append(eq.band_eigenvalues().copy())
                its.append(st["iterations"])
            else:
                lam, st = eq.SolveBatch(k)
                assert all(s["converged_bands"] == 6 for s in st)
                lams.append(lam.copy())
                its.append(max(s["iterations"] for s in st))
        return np.array(lams), its
    finally:
        del os.environ["BLOCH_HISTORY"]


def test_history_along_a_line_same_bands_fewer_iterations(bloch):
    L = bloch.BravaisLattice("FCC")
    X = L.GetSymmetryPoint(L.GetSymmetryPointIndex("X"))
    ks = [(0.30 + 0.01 * i) * X + np.array([0.05, 0.02, 0.0]) for i in range(6)]
    lam1, it1 = _walk(bloch, ks, True)
    lam0, it0 = _walk(bloch, ks, False)
    assert np.abs(lam1 - lam0).max() < 1e-8 * np.abs(lam0).max()
    assert it1[:2] == it0[:2]                      # no history before the third point of the walk
    assert sum(it1[2:]) < sum(it0[2:]), (it1, it0)
    # coarse sampling (12 % of Gamma-X per step): the history is not used at all
    ks = [(0.20 + 0.12 * i) * X + np.array([0.05, 0.02, 0.0]) for i in range(4)]
    lam1, it1 = _walk(bloch, ks, True)
    lam0, it0 = _walk(bloch, ks, False)
    assert it1 == it0 and np.abs(lam1 - lam0).max() < 1e-8 * np.abs(lam0).max()


def test_history_is_not_used_when_the_walk_turns_back(bloch):
    """kappa_a -> kappa_b -> kappa_a: the history holds the solution of kappa_a itself; the direction test must keep
    it out (identical iteration counts with the switch on and off)."""
    ka, kb = np.array([1.0, 0.3, 0.2]), np.array([1.1, 0.33, 0.2])
    lam1, it1 = _walk(bloch, [ka, kb, ka, kb], True)
    lam0, it0 = _walk(bloch, [ka, kb, ka, kb], False)
    assert it1 == it0
    assert np.abs(lam1 - lam0).max() < 1e-8 * np.abs(lam0).max()


def test_history_in_a_batch_with_mixed_walks(bloch):
    """three slots: a straight walk from Gamma, a straight walk elsewhere, a slot that jumps around"""
    L = bloch.BravaisLattice("FCC")
    X = L.GetSymmetryPoint(L.GetSymmetryPointIndex("X"))
    rounds = []
    jumps = [np.array([2.0, 1.0, 0.5]), np.array([0.4, 2.2, 0.1]), np.array([1.1, 0.2, 1.9]), np.array([0.3, 0.3, 2.5])]
    for r in range(4):
        rounds.append(np.array([0.02 * r * X, (0.5 + 0.015 * r) * X + np.array([0.0, 0.3, 0.1]), jumps[r]]))
    lam1, it1 = _walk(bloch, rounds, True, batch=3)
    lam0, it0 = _walk(bloch, rounds, False, batch=3)
    assert np.abs(lam1 - lam0).max() < 1e-7 * np.abs(lam0).max()
