"""Multilevel warm start (MaxwellBlochWaveSolver::GetEigenfrequencies, meta-material/
meta_material_solver.cpp:2731-2881): the ND refinement operator against the oracle's literal nodal
interpolation, and the level loop against direct fine solves."""
import numpy as np
import pytest

from oracle.bloch_oracle import Lattice, Mesh, RefElem

pytestmark = pytest.mark.gpu


def _oracle_prolong(name, n, p, eq_c, eq_f, xc):
    """fine dof = t . J_f^T E_c(x_dof): evaluate the coarse field at every fine dof node."""
    lat = Lattice(name)
    mc, mf = Mesh(lat, n), Mesh(lat, 2 * n)
    ref = RefElem(p)
    gc, sc = eq_c.dofmap("nd")
    gf, sf = eq_f.dofmap("nd")
    eye = np.eye(3)
    out = np.zeros(int(gf.max()) + 1, complex)
    cnt = np.zeros(len(out))
    for e in range(mf.ne):
        Jf = mf.J[mf.cls[e]]
        xn = mf.x0[e] + ref.nd_nodes @ Jf.T                         # physical dof nodes
        ctr = mf.x0[e] + Jf @ np.full(3, 0.5)
        par = None
        for ec in range(mc.ne):                                      # parent = coarse element holding the centre
            Jc = mc.J[mc.cls[ec]]
            xi = np.linalg.solve(Jc, ctr - mc.x0[ec])
            if (xi > -1e-9).all() and (xi < 1 + 1e-9).all():
                par = ec
                break
        assert par is not None
        Jc = mc.J[mc.cls[par]]
        xi = np.clip(np.linalg.solve(Jc, (xn - mc.x0[par]).T).T, 0.0, 1.0)
        val, _ = ref.nd_shapes(xi)                                   # [L_f, L_c, 3] reference values
        Ehat = np.einsum("qaj,a->qj", val, sc[par] * xc[gc[par]])
        Ephys = Ehat @ np.linalg.inv(Jc)                             # J^-T Ehat (row form)
        dof = np.einsum("qi,qi->q", eye[ref.nd_comp] @ Jf.T, Ephys)  # t . J_f^T E
        np.add.at(out, gf[e], sf[e] * dof)
        np.add.at(cnt, gf[e], 1.0)
    return out / cnt


@pytest.mark.parametrize("name,n,p", [("CUB", 3, 1), ("CUB", 2, 2), ("FCC", 1, 2), ("BCC", 1, 1), ("CUB", 1, 3)])
def test_nd_prolongation_matches_nodal_interpolation(bloch, name, n, p):
    L = bloch.BravaisLattice(name)
    eq_c, eq_f = bloch.MaxwellBlochWaveEquation(L, n, p), bloch.MaxwellBlochWaveEquation(L, 2 * n, p)
    k = np.array([0.5, 0.3, -0.2])
    nb = 2
    for eq in (eq_c, eq_f):
        eq.SetNumEigs(2 * nb)
        eq.SetAbsoluteTolerance(1e-8)
        eq.SetKappa(k)
        eq.Setup()
    eq_c.Solve()
    eq_c.ProlongEigenvectorsTo(eq_f)
    for i in range(nb):
        er, ei = eq_c.GetEigenvectorE(i)
        fr, fi = eq_f.GetEigenvectorE(i)                            # the installed starting block
        ref = _oracle_prolong(name, n, p, eq_c, eq_f, er + 1j * ei)
        assert np.abs(fr + 1j * fi - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    # the interpolant of a coarse field is the same function: M-norms agree (eps = 1 on both levels)
    xc = np.concatenate(eq_c.GetEigenvectorE(0))
    xf = np.concatenate(eq_f.GetEigenvectorE(0))
    assert abs(xc @ eq_c.MultM(xc) - xf @ eq_f.MultM(xf)) < 1e-10 * abs(xc @ eq_c.MultM(xc))


def test_prolong_rejects_unrelated_meshes(bloch):
    L = bloch.BravaisLattice("CUB")
    a, b = bloch.MaxwellBlochWaveEquation(L, 2, 1), bloch.MaxwellBlochWaveEquation(L, 3, 1)
    a.SetNumEigs(4)
    a.SetKappa([0.3, 0.1, 0.2])
    a.Solve()
    with pytest.raises(bloch.BlochError):
        a.ProlongEigenvectorsTo(b)


def test_multilevel_matches_direct_fine_solve_and_saves_iterations(bloch):
    L = bloch.BravaisLattice("FCC")
    eps_fn = lambda c: bloch.sphere_eps(c, 0.3, 8.0, 1.0)
    k = L.GetSymmetryPoint(L.GetSymmetryPointIndex("X")) * 0.6
    ml = bloch.MaxwellBlochWaveSolver(L, 2, 2, 6, eps_fn=eps_fn, max_lvl=3, tol=1e-12, solver_tol=1e-7)
    ml.SetKappa(k)
    om = ml.GetEigenfrequencies()
    assert len(ml.level_eigs) == 3 and ml.fine_level == 2
    direct = bloch.MaxwellBlochWaveEquation(L, 8, 2)
    direct.SetMassCoef(eps_fn(direct.element_centers()))
    direct.SetAbsoluteTolerance(1e-7)
    lam = direct.GetEigenvalues(12, k)[0::2]
    assert np.allclose(om ** 2, lam, rtol=1e-6, atol=1e-7)
    cold = direct.GetSolverStats()["iterations"]
    assert ml.level_iters[-1] < cold, (ml.level_iters, cold)
    # the level loop stops early once successive levels agree to tol
    ml2 = bloch.MaxwellBlochWaveSolver(L, 2, 2, 6, eps_fn=eps_fn, max_lvl=4, tol=1e3)
    ml2.SetKappa(k)
    ml2.GetEigenfrequencies()
    assert len(ml2.level_eigs) == 2
