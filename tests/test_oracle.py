"""CPU tests of the oracle itself: the known-answer recipes of SURVEY.md section 8(c) (the
reference has no assertions of its own) and the committed golden fixtures."""
import json
import os

import numpy as np
import pytest

from oracle.bloch_oracle import (BlochOperators, Lattice, Mesh, RefElem, Spaces, element_matrices,
                                 empty_lattice_eigs, gauss_legendre, gauss_lobatto)

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,vol", [("CUB", 1.0), ("FCC", 0.25), ("BCC", 0.5), ("HEX", 0.8660254037844386)])
def test_lattice_vectors_and_volume(name, vol):
    L = Lattice(name)
    assert np.allclose(L.lat @ L.rec.T, np.eye(3))                 # misc/test_bravais.cpp:274-287
    assert abs(L.volume - vol) < 1e-14                             # lib/bravais.cpp:83-108
    for n in (1, 2, 3):
        assert abs(Mesh(L, n).volume - vol) < 1e-13                # misc/test_bravais.cpp:432-438


@pytest.mark.parametrize("name", ["CUB", "FCC", "BCC", "HEX"])
@pytest.mark.parametrize("p", [1, 2])
def test_periodic_mesh_euler_number_and_counts(name, p):
    mesh = Mesh(Lattice(name), 2)
    s = Spaces(mesh, p)
    if p == 1:    # V - E + F - Ne = 0 on the 3-torus (meta_material.cpp:427-432)
        assert s.n_h1 - s.n_nd + s.n_rt - mesh.ne == 0
    assert s.n_nd == 3 * p ** 3 * mesh.ne and s.n_rt == s.n_nd and s.n_h1 == p ** 3 * mesh.ne


def test_1d_nodes():
    for n in range(2, 7):
        x = gauss_lobatto(n)
        assert x[0] == 0 and x[-1] == 1 and np.allclose(x + x[::-1], 1)
        g, w = gauss_legendre(n)
        assert abs(w.sum() - 1) < 1e-14
        for k in range(2 * n):
            assert abs(w @ g ** k - 1 / (k + 1)) < 1e-13


@pytest.mark.parametrize("name,n,p", [("CUB", 3, 1), ("FCC", 2, 2), ("BCC", 1, 2), ("CUB", 2, 3), ("HEX", 2, 1)])
def test_operator_identities(name, n, p):
    mesh = Mesh(Lattice(name), n)
    rng = np.random.default_rng(0)
    ops = BlochOperators(Spaces(mesh, p), rng.uniform(1, 10, mesh.ne), rng.uniform(0.5, 2, mesh.ne))
    ops.set_kappa([0.4, 1.3, -0.8])
    assert abs(ops.T12 @ ops.T01).max() < 1e-12                   # exact sequence
    assert abs(ops.C_c() @ ops.G_c()).max() < 1e-12               # ... also for the shifted complex
    A, M = ops.A_c(), ops.M_c()
    assert abs(A - A.conj().T).max() < 1e-12 * abs(A).max()
    C = ops.C_c()
    assert abs(A - C.conj().T @ ops.M2 @ C).max() < 1e-12 * abs(A).max()   # S1/DKZ form == C^H M2 C
    Ab = ops.A_block()
    x = rng.uniform(-1, 1, (2, 2 * ops.sp_.n_nd))
    y = ops.to_c(ops.apply_A(x), ops.sp_.n_nd)
    assert np.allclose(y, (A @ ops.to_c(x, ops.sp_.n_nd).T).T, atol=1e-12 * abs(Ab).max())
    assert np.all(np.linalg.eigvalsh(M.toarray()) > 0)
    # projector identities (meta_material_solver.cpp:2245-2279)
    P = ops.projector_dense()
    assert np.abs(P @ P - P).max() < 1e-9
    G = ops.G_c().toarray()
    assert np.abs(G.conj().T @ M.toarray() @ P).max() < 1e-9
    assert np.abs(C.toarray() @ (np.eye(P.shape[0]) - P)).max() < 1e-9 * abs(C).max() * 10


def test_empty_lattice_spectrum_and_convergence():
    lat = Lattice("CUB")
    kap = np.array([0.9, 0.3, -0.2])
    exact = empty_lattice_eigs(lat, kap, 6)
    errs = []
    for n, p in [(3, 1), (6, 1), (3, 2)]:
        mesh = Mesh(lat, n)
        w = BlochOperators(Spaces(mesh, p), np.ones(mesh.ne)).set_kappa(kap).eig_dense(6)
        assert abs(w[0] - kap @ kap) < 1e-10 and abs(w[1] - kap @ kap) < 1e-10   # constant envelope: exact
        assert np.all(w[2:] >= exact[2:] - 1e-9)                                 # conforming: from above
        errs.append(abs(w[2] - exact[2]) / exact[2])
    assert errs[1] < errs[0] / 3 and errs[2] < errs[0] / 3, errs                  # O(h^2p)


def test_gamma_and_kappa_symmetry():
    mesh = Mesh(Lattice("CUB"), 3)
    ops = BlochOperators(Spaces(mesh, 1), mesh.sphere_eps(radius=0.4))
    g = ops.set_kappa(np.zeros(3)).eig_dense(6)
    assert np.all(np.abs(g[:3]) < 1e-10)                          # three harmonic fields at Gamma
    k = np.array([0.9, 0.2, -0.5])
    a = ops.set_kappa(k).eig_dense(6)
    b = ops.set_kappa(-k).eig_dense(6)
    assert np.allclose(a, b, rtol=1e-9)
    x = [ops.set_kappa(np.pi * np.eye(3)[d]).eig_dense(6) for d in range(3)]
    assert np.allclose(x[0], x[1], rtol=1e-9) and np.allclose(x[0], x[2], rtol=1e-9)


def test_element_matrix_closed_forms():
    """order 1 on the unit cube: ND mass diagonal 1/9... use the exactly known traces instead"""
    ref = RefElem(1)
    em = element_matrices(ref, np.eye(3), np.array([0.0, 0.0, 1.0]))
    assert abs(em["M1"].sum() - 0.0) < 10                        # sanity
    # sum of all entries of the x-x block = int (sum of x-shapes)^2 = 1 (partition of unity in y,z)
    assert abs(em["M1"][:4, :4].sum() - 1.0) < 1e-13
    assert abs(em["M2"][:2, :2].sum() - 1.0) < 1e-13
    assert abs(em["M0"].sum() - 1.0) < 1e-13
    # curl of the 12 edge functions: each face sees its 4 edges with +-1
    assert np.allclose(np.sort(np.abs(em["T12"]).sum(1)), 4.0)


def test_golden_fixtures_pin_the_oracle():
    cases = json.load(open(os.path.join(HERE, "golden", "bands_small.json")))
    for c in cases[:6]:                                            # the cheap ones (full set under -m gpu)
        lat = Lattice(c["lattice"])
        mesh = Mesh(lat, c["n_sub"])
        eps = mesh.sphere_eps() if c["sphere"] else np.ones(mesh.ne)
        ops = BlochOperators(Spaces(mesh, c["order"]), eps).set_kappa(c["kappa"])
        assert ops.sp_.n_nd == c["n_nd"]
        w = ops.eig_dense(len(c["eigenvalues"]))
        assert np.allclose(w, c["eigenvalues"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", ["CUB", "FCC"])
def test_field_average_forms_on_constant_fields(name):
    """SetKappa linear forms (maxwell_bloch.cpp:211-279) at kappa = 0: applied to the ND / RT
    interpolants of a constant vector c they return V c (and eps-/mu^-1-weighted volumes for D, H)."""
    lat = Lattice(name)
    mesh = Mesh(lat, 2)
    s = Spaces(mesh, 2)
    rng = np.random.default_rng(5)
    eps, mui = rng.uniform(1, 4, mesh.ne), rng.uniform(0.5, 2, mesh.ne)
    ops = BlochOperators(s, eps, mui).set_kappa(np.zeros(3))
    L = ops.field_average_forms()
    c = np.array([0.3, -1.1, 0.7])
    eye = np.eye(3)
    e_nd, b_rt = np.zeros(s.n_nd), np.zeros(s.n_rt)
    vol_e = np.array([np.linalg.det(mesh.J[k]) for k in mesh.cls])
    for e in range(mesh.ne):
        J = mesh.J[mesh.cls[e]]
        e_nd[s.nd_gid[e]] = s.nd_sign[e] * (eye[s.ref.nd_comp] @ (J.T @ c))                    # t . J^T c
        b_rt[s.rt_gid[e]] = s.rt_sign[e] * (eye[s.ref.rt_comp] @ (np.linalg.det(J) * np.linalg.inv(J) @ c))  # n . adj(J) c
    assert np.allclose(L["E"] @ e_nd, mesh.volume * c) and np.allclose(L["B"] @ b_rt, mesh.volume * c)
    assert np.allclose(L["D"] @ e_nd, (eps * vol_e).sum() * c) and np.allclose(L["H"] @ b_rt, (mui * vol_e).sum() * c)


@pytest.mark.parametrize("name,n,p", [("CUB", 2, 1), ("FCC", 2, 2), ("BCC", 1, 3), ("HEX", 2, 2)])
def test_nodal_interpolation_matrix(name, n, p):
    """Pi: (H1)^3 -> ND of the AMS-type preconditioner (maxwell_bloch.cpp:492-517), pinned by exact identities:
    Pi (zeta phi) = Z01 phi (the reference's zeta-interpolation of the projector, pfem_extras_bloch.cpp:182-258),
    constant fields are curl free at Gamma, and their M-norm is the eps-weighted cell volume."""
    from oracle.bloch_oracle import nodal_interpolation
    lat = Lattice(name)
    mesh = Mesh(lat, n)
    s = Spaces(mesh, p)
    rng = np.random.default_rng(5)
    eps = rng.uniform(1, 10, mesh.ne)
    ops = BlochOperators(s, eps)
    Pi = nodal_interpolation(s)
    assert Pi.shape == (s.n_nd, 3 * s.n_h1)
    kappa = np.array([0.7, -1.1, 0.4])
    ops.set_kappa(kappa)
    phi = rng.uniform(-1, 1, s.n_h1)
    assert np.allclose(Pi @ np.concatenate([z * phi for z in ops.zeta]), ops.Z01 @ phi, atol=1e-13)
    ops.set_kappa(np.zeros(3))
    vol_eps = float(np.sum(eps * np.abs(np.linalg.det(mesh.J))[mesh.cls]))
    for d in range(3):
        c = np.zeros(3 * s.n_h1)
        c[d * s.n_h1:(d + 1) * s.n_h1] = 1.0
        v = Pi @ c
        assert np.abs(ops.T12 @ v).max() < 1e-12                    # curl of a constant field
        assert abs(v @ (ops.M1 @ v) - vol_eps) < 1e-11 * vol_eps    # |e_d|^2 integrated with eps


def _precond_condition(name, n, p, kappa=(0.9, 0.4, 0.2)):
    """cond of P T (A + sigma M) on the divergence-free space for T = auxiliary-space cycle (exact auxiliary solves)
    and for T = one damped Jacobi sweep, on the oracle's assembled matrices"""
    from oracle.bloch_oracle import nodal_interpolation
    lat = Lattice(name)
    mesh = Mesh(lat, n)
    s = Spaces(mesh, p)
    ops = BlochOperators(s, mesh.sphere_eps()).set_kappa(np.array(kappa))
    ops1 = BlochOperators(s, np.ones(mesh.ne)).set_kappa(np.array(kappa))
    A, M, G = ops.A_c().toarray(), ops.M_c().toarray(), ops.G_c().toarray()
    sigma = 1.0 / mesh.volume ** (2.0 / 3.0)
    Ash = A + sigma * M
    Pi = nodal_interpolation(s).toarray()
    n0 = s.n_h1
    Li = np.linalg.inv(ops1.S0_c().toarray() + sigma * ops.M0.toarray())
    Z = np.zeros((3 * n0, 3 * n0), complex)
    for d in range(3):
        Z[d * n0:(d + 1) * n0, d * n0:(d + 1) * n0] = Li
    dj = 1.0 / np.diag(Ash).real
    lmax = np.linalg.eigvalsh((np.sqrt(dj)[:, None] * Ash) * np.sqrt(dj)[None, :]).max()
    S = np.diag(dj / (0.625 * lmax))
    I = np.eye(len(A))
    X = S.copy()                                    # x = S r; x += Pi L^-1 Pi^T (r - A x); x += S (r - A x)
    X = X + Pi @ Z @ Pi.T @ (I - Ash @ X)
    X = X + S @ (I - Ash @ X)
    P = I - G @ np.linalg.solve(G.conj().T @ M @ G, G.conj().T @ M)

    def cond(T):
        w = np.linalg.eigvals(P @ T @ Ash @ P)
        w = np.sort(w.real[np.abs(w) > 1e-8 * np.abs(w).max()])
        return w.max() / w.min()
    return cond(X), cond(S)


def test_auxiliary_space_preconditioner_is_mesh_and_order_independent():
    """The algorithm behind csrc/aux.cu on the oracle's matrices (the role of HypreAMS, maxwell_bloch.cpp:492-517):
    smoother + Pi L^-1 Pi^T followed by the divergence projector gives cond(P T (A + sigma M)) ~ 1.5 - 2.3 whatever the
    mesh, the order and the lattice, where a Jacobi sweep alone degrades like h^-2."""
    res = {c: _precond_condition(*c) for c in [("CUB", 2, 1), ("CUB", 4, 1), ("FCC", 2, 1), ("CUB", 2, 2)]}
    for c, (ka, kj) in res.items():
        assert ka < 3.0, (c, ka)
        assert kj > 5.0 * ka, (c, ka, kj)
    assert res[("CUB", 4, 1)][1] > 2.5 * res[("CUB", 2, 1)][1]          # Jacobi: 16 -> 54
    assert res[("CUB", 4, 1)][0] < 1.5 * res[("CUB", 2, 1)][0]          # auxiliary space: 1.6 -> 1.8
