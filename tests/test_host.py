"""CPU tests of the host side: the C-ABI library loads and exports every symbol of
include/bloch_b200.h, the lattice / k-path API matches the reference tables, the element -> dof
maps agree with the oracle's independent geometric identification, and compute entry points
fail loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.bloch_oracle import Lattice, Mesh, Spaces

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(bloch):
    hdr = open(os.path.join(ROOT, "include", "bloch_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(bloch_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert len(names) > 40
    lib = C.CDLL(bloch.lib_path)
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n
    from mfem_bravais_b200 import capi
    assert names == set(capi.SIGNATURES), names ^ set(capi.SIGNATURES)
    assert bloch.lib().bloch_version() >= 100


@pytest.mark.parametrize("name", ["CUB", "FCC", "BCC", "HEX"])
def test_lattice_api_matches_reference_tables(bloch, name):
    L = bloch.BravaisLattice(name)
    o = Lattice(name)
    assert L.GetLatticeTypeLabel() == name
    assert np.allclose(L.GetLatticeVectors(), o.lat) and np.allclose(L.GetReciprocalLatticeVectors(), o.rec)
    assert abs(L.GetUnitCellVolume() - o.volume) < 1e-14
    assert L.GetNumberSymmetryPoints() == len(o.sp)
    for i in range(L.GetNumberSymmetryPoints()):
        lab = L.GetSymmetryPointLabel(i)
        assert np.allclose(L.GetSymmetryPoint(i), o.kappa(lab))   # 2 pi * sp (lib/bravais.cpp:201-206)
        assert L.GetSymmetryPointIndex(lab) == i
    assert L.GetSymmetryPointIndex("nope") == -1
    assert L.GetNumberPaths() == len(o.paths)
    for p, path in enumerate(o.paths):
        assert L.GetNumberPathSegments(p) == len(path) - 1
        for s in range(len(path) - 1):
            e0, e1 = L.GetPathSegmentEndPointIndices(p, s)
            assert L.GetSymmetryPointLabel(e0) == path[s] and L.GetSymmetryPointLabel(e1) == path[s + 1]
            mid = 0.5 * (o.kappa(path[s]) + o.kappa(path[s + 1]))
            assert np.allclose(L.GetIntermediatePoint(p, s), mid)  # lib/bravais.cpp:59-74
    # translation vectors are lattice vectors: integer in reciprocal coordinates
    t = L.GetTranslationVectors() @ o.rec.T
    assert np.allclose(t, np.round(t)) and len(L.GetFaceRadii()) == len(t)


def test_reference_labels_and_paths(bloch):
    L = bloch.BravaisLattice("FCC")       # lib/bravais.cpp:2420-2460
    assert [L.GetSymmetryPointLabel(i) for i in range(6)] == ["Gamma", "X", "W", "K", "L", "U"]
    assert L.GetIntermediatePointLabel(0, 0) == "Delta" and L.GetIntermediatePointLabel(1, 0) == "T"
    assert np.allclose(L.GetSymmetryPoint(1), 2 * np.pi * np.array([0.0, 1.0, 0.0]))
    H = bloch.BravaisLattice("HEX")       # lib/bravais.cpp:6091-6124
    assert [H.GetSymmetryPointLabel(i) for i in range(6)] == ["Gamma", "A", "H", "K", "L", "M"]
    assert H.GetNumberPaths() == 3 and H.GetNumberPathSegments(0) == 7 and H.GetIntermediatePointLabel(2, 0) == "HK"
    assert abs(H.GetUnitCellVolume() - np.sqrt(3) / 2) < 1e-14
    B = bloch.BravaisLattice("BCC")       # lib/bravais.cpp:2708-2736
    assert [B.GetSymmetryPointLabel(i) for i in range(4)] == ["Gamma", "H", "N", "P"]
    ks = bloch.k_path(L, ["Gamma", "X", "W", "L", "Gamma"], 8)
    assert ks.shape == (32, 3) and np.allclose(ks[7], L.GetSymmetryPoint(1)) and np.allclose(ks[-1], 0)
    mapped, ipt = L.MapToPrimitiveCell([0.45, 0.45, 0.05])
    assert np.linalg.norm(ipt) <= np.linalg.norm([0.45, 0.45, 0.05]) + 1e-14


def _equivalent(ga, sa, gb, sb):
    n = int(ga.max()) + 1
    if int(gb.max()) + 1 != n:
        return False
    f = np.full(n, -1)
    sg = np.zeros(n)
    for x, y, z in zip(ga.ravel(), gb.ravel(), (sa * sb).ravel()):
        if f[x] < 0:
            f[x], sg[x] = y, z
        elif f[x] != y or sg[x] != z:
            return False
    return len(np.unique(f)) == n


@pytest.mark.parametrize("name,n,p", [("CUB", 1, 1), ("CUB", 2, 2), ("CUB", 3, 3), ("FCC", 1, 2), ("FCC", 2, 3),
                                      ("FCC", 3, 1), ("BCC", 1, 3), ("BCC", 2, 2), ("BCC", 3, 1), ("HEX", 1, 2),
                                      ("HEX", 2, 1), ("HEX", 2, 3)])
def test_dofmaps_agree_with_geometric_identification(bloch, name, n, p):
    """entity-based numbering (product) vs node-position hashing (oracle): same identification
    up to renumbering and a per-dof orientation sign, including the degenerate n = 1 meshes"""
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p, device=-2)        # topology-only handle
    mesh = Mesh(Lattice(name), n)
    sp = Spaces(mesh, p)
    x0, cls, J = eq.element_geometry()
    assert np.allclose(x0, mesh.x0) and np.allclose(J, mesh.J) and (cls == mesh.cls).all()
    assert np.allclose(eq.element_centers(), mesh.centers)
    v, e, f, vol = eq.mesh_counts()
    assert v - e + f - eq.n_elem == 0 and abs(vol - mesh.volume) < 1e-13
    assert (eq.N, eq.N_rt, eq.N_h1) == (sp.n_nd, sp.n_rt, sp.n_h1)
    for space, og, osg in [("h1", sp.h1_gid, sp.h1_sign), ("nd", sp.nd_gid, sp.nd_sign), ("rt", sp.rt_gid, sp.rt_sign)]:
        g, s = eq.dofmap(space)
        assert _equivalent(g, np.abs(s) if space == "h1" else s, og, osg), space


def test_no_cpu_fallback(bloch):
    L = bloch.BravaisLattice("CUB")
    eq = bloch.MaxwellBlochWaveEquation(L, 2, 1, device=-2)
    x = np.zeros(2 * eq.N)
    for fn in (eq.MultA, eq.MultM, eq.MultProjector):
        with pytest.raises(bloch.BlochError):
            fn(x)
    with pytest.raises(bloch.BlochError):
        eq.Solve()
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(bloch.BlochError):
            bloch.MaxwellBlochWaveEquation(L, 2, 1)                 # real handle needs a GPU


def test_argument_errors(bloch):
    L = bloch.BravaisLattice("FCC")
    with pytest.raises(bloch.BlochError):
        bloch.MaxwellBlochWaveEquation(L, 2, 7, device=-2)         # unsupported order
    with pytest.raises(bloch.BlochError):
        bloch.MaxwellBlochWaveEquation(L, 0, 1, device=-2)
    with pytest.raises(bloch.BlochError):
        bloch.BravaisLattice(3)                                     # 2-D lattice not on this path
    eq = bloch.MaxwellBlochWaveEquation(L, 1, 1, device=-2)
    with pytest.raises(bloch.BlochError):
        eq.SetMassCoef(np.zeros(eq.n_elem))                         # eps must be positive
    with pytest.raises(bloch.BlochError):
        eq.SetNumEigs(1000)


def test_sphere_coefficient_and_omega_format(bloch):
    from mfem_bravais_b200.dispersion import omega_of_lambda
    c = np.array([[0.1, 0.1, 0.1], [0.2, 0.2, 0.0], [0.25, 0.0, 0.0], [0.3, 0.0, 0.0]])
    assert bloch.sphere_eps(c).tolist() == [10.0, 1.0, 10.0, 1.0]   # maxwell_dispersion.cpp:1464-1467
    assert omega_of_lambda([4.0, 0.0, -1e-9, -1.0]).tolist() == [2.0, 0.0, 0.0, -1.0]  # :1072-1083


@pytest.mark.parametrize("n,m", [(12, 4), (48, 16), (63, 21)])
def test_dense_rayleigh_ritz_solver_against_scipy(n, m):
    """csrc/dense.hpp (Cholesky + Householder/QL, lowest-m back-transformation) on Hermitian pencils with
    degenerate pairs like the Bloch spectra; and the pivoted-Cholesky values-only variant on a nearly dependent
    basis.  Host code only."""
    import ctypes as C
    import scipy.linalg as sla
    from mfem_bravais_b200 import capi
    lib = capi.lib()
    rng = np.random.default_rng(n)
    D = 3 * n
    lam_true = np.repeat(np.sort(rng.uniform(1, 50, (D + 1) // 2)), 2)[:D]    # every eigenvalue twice
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D)))
    A = (Q * lam_true) @ Q.conj().T
    X = rng.standard_normal((D, n)) + 1j * rng.standard_normal((D, n))
    GA, GM = X.conj().T @ A @ X, X.conj().T @ X
    GA, GM = 0.5 * (GA + GA.conj().T), 0.5 * (GM + GM.conj().T)
    ref_w, ref_v = sla.eigh(GA, GM)
    pack = lambda Z: np.ascontiguousarray(np.stack([Z.real, Z.imag], axis=-1))
    lam, Cr = np.zeros(m), np.zeros((n, m, 2))
    ga, gm = pack(GA), pack(GM)
    capi.check(lib.bloch_debug_hegv(n, m, capi.dptr(ga), capi.dptr(gm), capi.dptr(lam), capi.dptr(Cr), 0))
    assert np.allclose(lam, ref_w[:m], rtol=1e-11, atol=1e-11)
    Cm = Cr[..., 0] + 1j * Cr[..., 1]
    assert np.abs(GA @ Cm - (GM @ Cm) * lam).max() < 1e-9 * np.abs(GA).max()          # eigen-residual
    assert np.abs(Cm.conj().T @ GM @ Cm - np.eye(m)).max() < 1e-10                    # GM-orthonormal
    # nearly dependent basis: duplicate columns up to 1e-9 noise
    X2 = np.concatenate([X, X[:, : n // 2] + 1e-9 * rng.standard_normal((D, n // 2))], axis=1)
    G2A, G2M = X2.conj().T @ A @ X2, X2.conj().T @ X2
    G2A, G2M = 0.5 * (G2A + G2A.conj().T), 0.5 * (G2M + G2M.conj().T)
    lam2 = np.zeros(m)
    n2 = X2.shape[1]
    g2a, g2m = pack(G2A), pack(G2M)
    capi.check(lib.bloch_debug_hegv(n2, m, capi.dptr(g2a), capi.dptr(g2m), capi.dptr(lam2), None, 1))
    assert np.allclose(lam2, ref_w[:m], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name,n", [("CUB", 2), ("FCC", 2), ("BCC", 1), ("HEX", 2)])
def test_mfem_mesh_export(bloch, name, n, tmp_path):
    """write_mfem_mesh: the refined Wigner-Seitz cell as a non-periodic `MFEM mesh v1.0` file + translation vectors,
    the input of the reference's MakePeriodicMesh pipeline (lib/bravais.cpp:343-355, 9548-9832)."""
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, 1, device=-2)
    eps = bloch.sphere_eps(eq.element_centers())
    path = str(tmp_path / "cell.mesh")
    nv, ne, nb = bloch.write_mfem_mesh(eq, path, lattice=L, eps=eps)
    V, H, attr, B = bloch.read_mfem_mesh(path)
    assert (len(V), len(H), len(B)) == (nv, ne, nb) and ne == eq.n_elem
    # positively oriented MFEM hexes filling the cell
    vol = 0.0
    for h in H:
        P = V[h]
        d = np.linalg.det(np.stack([P[1] - P[0], P[3] - P[0], P[4] - P[0]]))
        assert d > 0
        vol += d
    assert abs(vol - L.GetUnitCellVolume()) < 1e-13
    # attributes reproduce the element-wise coefficients through the .coef table
    table = {int(a): (float(e), float(m)) for a, e, m in (ln.split() for ln in open(path + ".coef"))}
    assert [table[a][0] for a in attr] == list(eps) and len(table) == len(set(eps))
    # Euler characteristic of the solid cell = 1, of its boundary surface = 2
    edges = {tuple(sorted((h[a], h[b]))) for h in H for a, b in
             [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]}
    faces = {tuple(sorted(h[list(f)])) for h in H for f in [(3, 2, 1, 0), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7)]}
    assert len(V) - len(edges) + len(faces) - len(H) == 1
    bv = set(B.ravel())
    be = {tuple(sorted((q[a], q[(a + 1) % 4]))) for q in B for a in range(4)}
    assert len(bv) - len(be) + len(B) == 2
    # the translation vectors pair up the boundary vertices exactly like the product's periodic H1 numbering:
    # vertices u, v are identified iff they differ by an integer combination of translation vectors
    T = np.loadtxt(path + ".trans").reshape(-1, 3)
    assert np.allclose(T, L.GetTranslationVectors())
    gid, _ = eq.dofmap("h1")
    x0, cls, J = eq.element_geometry()
    ref = np.array([[i, j, k] for k in (0, 1) for j in (0, 1) for i in (0, 1)], float)
    key = {tuple(np.round(v, 9) + 0.0): i for i, v in enumerate(V)}
    g_of_v = {}
    for e in range(ne):
        P = x0[e] + ref @ J[cls[e]].T
        for l in range(8):
            v = key[tuple(np.round(P[l], 9) + 0.0)]
            assert g_of_v.setdefault(v, int(gid[e, l])) == int(gid[e, l])
    rec = L.GetReciprocalLatticeVectors()
    for u in bv:
        for t in np.concatenate([T, -T]):
            k = tuple(np.round(V[u] + t, 9) + 0.0)
            if k in key:
                assert g_of_v[key[k]] == g_of_v[u]
    frac = V @ rec.T
    classes = {}
    for v in range(len(V)):
        f = frac[v] - np.floor(frac[v] + 1e-9)
        classes.setdefault(tuple(np.round(f, 6) % 1.0), set()).add(g_of_v[v])
    assert all(len(s) == 1 for s in classes.values()) and len(classes) == eq.N_h1


@pytest.mark.parametrize("name,n,p", [("CUB", 2, 1), ("FCC", 2, 2), ("BCC", 1, 3), ("HEX", 2, 1)])
def test_plane_wave_initial_vectors(bloch, name, n, p):
    """CreateInitialVectors (maxwell_dispersion.cpp:735-1060) restated on the host: ND nodal interpolation against the
    oracle's literal dof functionals, and the physics on the oracle's assembled empty-lattice pencil - each wave is
    divergence-free to discretisation accuracy when E0 is orthogonal to the true wave vector, and its Rayleigh
    quotient is the empty-lattice band |kappa + 2 pi G|^2.  No GPU (topology-only handle)."""
    from helpers import oracle_on_product_maps
    from oracle.bloch_oracle import RefElem
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p, device=-2)
    kappa = 0.5 * L.GetSymmetryPoint(1) + np.array([0.11, -0.07, 0.05])
    # (1) interpolation = t_k . J^T v(x_k) with the oracle's node table
    G = np.array([1.0, -1.0, 0.0]) @ L.GetReciprocalLatticeVectors()
    e0 = np.array([0.3, 0.5, -0.2])
    f = lambda X: np.exp(2j * np.pi * (X @ G))[:, None] * e0[None, :]
    v = bloch.nd_interpolate(eq, f)
    vc = v[:eq.N] + 1j * v[eq.N:]
    ref = RefElem(p)
    x0, cls, J = eq.element_geometry()
    gid, sign = eq.dofmap("nd")
    for e in range(0, eq.n_elem, max(1, eq.n_elem // 7)):
        X = x0[e] + ref.nd_nodes @ J[cls[e]].T
        t = J[cls[e]][:, ref.nd_comp].T
        loc = np.einsum("ki,ki->k", t, f(X))
        assert np.allclose(vc[gid[e]], sign[e] * loc, atol=1e-13)
    # (2) the block of the reference recipe
    W = bloch.plane_wave_initial_vectors(eq, L, kappa, literal=False)
    tab = {"FCC": 9, "BCC": 13}.get(name, 7)
    assert W.shape == (2 * tab, 2 * eq.N)                   # generic kappa: two polarisations per shift
    assert bloch.plane_wave_initial_vectors(eq, L, np.zeros(3)).shape[0] == 2 * tab + 1     # k = 0: three
    assert bloch.plane_wave_initial_vectors(eq, L, kappa, count=5).shape[0] == 5
    ops, _ = oracle_on_product_maps(eq, name, n, p, np.ones(eq.n_elem))
    ops.set_kappa(kappa)
    A, M, Gc = ops.A_c(), ops.M_c(), ops.G_c()
    Wc = W[:4, :eq.N] + 1j * W[:4, eq.N:]                   # the four lowest waves
    b = L.GetReciprocalLatticeVectors()
    lows = sorted(np.linalg.norm(kappa + 2 * np.pi * np.array(nn, float) @ b) ** 2
                  for nn in ([(0, 0, 0)] + [tuple(int(s) for s in row) for row in np.vstack([np.eye(3), -np.eye(3)])]
                             if tab == 7 else [(0, 0, 0)]))
    h = 1.0 / (n * p)
    for x in Wc[:2]:
        rq = np.vdot(x, A @ x).real / np.vdot(x, M @ x).real
        assert abs(rq - lows[0]) < 0.6 * h ** (2 * p) * max(1.0, lows[0]) + 1e-9     # lowest band |kappa|^2 (exact pair)
        div = Gc.conj().T @ (M @ x)
        assert np.linalg.norm(div) < 1e-8 * np.linalg.norm(M @ x)                     # exactly divergence-free


@pytest.mark.parametrize("name,n,p", [("FCC", 2, 2), ("CUB", 2, 1), ("BCC", 1, 3)])
def test_cpp_wrapper_mesh_and_initial_vectors_match_python(bloch, name, n, p, tmp_path):
    """include/maxwell_bloch_b200.hpp (the C++ host side): WriteMesh and CreateInitialVectors produce what the Python
    mirror produces (same files; same plane-wave block up to the free choice of the two polarisations per shift).
    Compiles a small program against the header with a topology-only handle - no GPU."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "maxwell_bloch_b200.hpp"
#include <cstdlib>
using namespace bloch_b200;
int main(int argc, char **argv) {
  BravaisLattice L(std::atoi(argv[1]));
  MaxwellBlochWaveEquation eq(L, std::atoi(argv[2]), std::atoi(argv[3]), BLOCH_DEVICE_NONE);
  std::vector<double> c, eps(eq.GetNE());
  eq.GetElementCenters(c);
  for (int64_t e = 0; e < eq.GetNE(); e++)
    eps[e] = std::sqrt(c[3 * e] * c[3 * e] + c[3 * e + 1] * c[3 * e + 1] + c[3 * e + 2] * c[3 * e + 2]) <= 0.25 ? 10.0 : 1.0;
  eq.WriteMesh(argv[4], L, eps, {});
  std::vector<double> kappa = {0.7, -0.4, 1.1}, vecs;
  int nv = 0;
  eq.CreateInitialVectors(L, kappa, vecs, nv, 0, false);
  FILE *f = std::fopen(argv[5], "wb");
  std::fwrite(vecs.data(), sizeof(double), vecs.size(), f);
  std::fclose(f);
  std::printf("%d\n", nv);
  return 0;
}
''')
    exe = str(tmp_path / "t")
    libdir = os.path.join(root, "mfem-bravais_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), str(src), "-L" + libdir,
                           "-lbloch_b200", "-Wl,-rpath," + libdir, "-o", exe])
    from mfem_bravais_b200 import capi
    cm, cv = str(tmp_path / "cpp.mesh"), str(tmp_path / "cpp.vecs")
    nv = int(subprocess.check_output([exe, str(capi.LATTICE_TYPES[name]), str(n), str(p), cm, cv]).split()[0])
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p, device=-2)
    pm = str(tmp_path / "py.mesh")
    bloch.write_mfem_mesh(eq, pm, lattice=L, eps=bloch.sphere_eps(eq.element_centers()))
    Vc, Hc, ac, Bc = bloch.read_mfem_mesh(cm)
    Vp, Hp, ap, Bp = bloch.read_mfem_mesh(pm)
    assert np.array_equal(Hc, Hp) and np.array_equal(ac, ap) and np.allclose(Vc, Vp, atol=1e-15)
    assert {tuple(sorted(q)) for q in Bc} == {tuple(sorted(q)) for q in Bp}
    assert open(cm + ".coef").read() == open(pm + ".coef").read()
    assert np.allclose(np.loadtxt(cm + ".trans"), np.loadtxt(pm + ".trans"))
    Wp = bloch.plane_wave_initial_vectors(eq, L, [0.7, -0.4, 1.1], literal=False)
    Wc = np.fromfile(cv).reshape(nv, 2 * eq.N)
    assert Wc.shape == Wp.shape
    zp = (Wp[:, :eq.N] + 1j * Wp[:, eq.N:]).T
    zc = (Wc[:, :eq.N] + 1j * Wc[:, eq.N:]).T
    # same subspace (on the coarsest meshes the +-G waves alias at the nodes, so the blocks may be rank deficient)
    for a, b in ((zp, zc), (zc, zp)):
        coef = np.linalg.lstsq(a, b, rcond=None)[0]
        assert np.linalg.norm(a @ coef - b) < 1e-9 * np.linalg.norm(b)


def test_scalar_wrapper_cpp_and_python_agree_on_the_bloch_vector(bloch, tmp_path):
    """ScalarFloquetWaveEquation mirrors (C++ header and Python): kappa = beta pi / 180 * zeta(azimuth, inclination)
    as in misc/scalar3d.cpp:665-669, 733; sizes from the topology-only handle; Solve() without a device fails loudly."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "s.cpp"
    src.write_text(r'''
#include "maxwell_bloch_b200.hpp"
#include <cstdlib>
using namespace bloch_b200;
int main(int argc, char **argv) {
  BravaisLattice L(std::atoi(argv[1]));
  ScalarFloquetWaveEquation eq(L, 2, 3, BLOCH_DEVICE_NONE);
  std::vector<double> z, k;
  eq.GetZeta(z);                                   // class defaults: azimuth 0, inclination 90 -> (0, 0, 1)
  std::printf("%.17g %.17g %.17g\n", z[0], z[1], z[2]);
  eq.SetBeta(75.0); eq.SetAzimuth(30.0); eq.SetInclination(20.0);
  eq.GetKappa(k);
  std::printf("%.17g %.17g %.17g\n", k[0], k[1], k[2]);
  std::printf("%lld %lld\n", (long long)eq.GetH1TrueVSize(), (long long)eq.GetNE());
  try { eq.Setup(); eq.Solve(); std::printf("solved\n"); } catch (const std::exception &ex) { std::printf("error\n"); }
  return 0;
}
''')
    exe = str(tmp_path / "s")
    libdir = os.path.join(root, "mfem-bravais_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), str(src), "-L" + libdir,
                           "-lbloch_b200", "-Wl,-rpath," + libdir, "-o", exe])
    from mfem_bravais_b200 import capi
    out = subprocess.check_output([exe, str(capi.LATTICE_TYPES["BCC"])]).decode().split("\n")
    z = np.array(out[0].split(), float)
    k = np.array(out[1].split(), float)
    nh1, ne = (int(x) for x in out[2].split())
    assert np.allclose(z, [0.0, 0.0, 1.0], atol=1e-15)
    d = np.pi / 180.0
    zeta = np.array([np.cos(20 * d) * np.cos(30 * d), np.cos(20 * d) * np.sin(30 * d), np.sin(20 * d)])
    assert np.allclose(k, 75.0 * d * zeta, rtol=1e-14)
    assert out[3] == "error"                          # no CPU path
    L = bloch.BravaisLattice("BCC")
    eq = bloch.ScalarFloquetWaveEquation(L, 2, 3, device=-2)
    assert (eq.N, eq.n_elem) == (nh1, ne)
    eq.SetBeta(75.0)
    eq.SetAzimuth(30.0)
    eq.SetInclination(20.0)
    assert np.allclose(eq._eq._kappa_set, k, rtol=1e-14)


@pytest.mark.parametrize("name,n,p", [("CUB", 2, 1), ("CUB", 1, 2), ("FCC", 2, 2), ("BCC", 1, 3), ("HEX", 2, 2), ("FCC", 1, 3)])
def test_nodal_interpolation_of_the_product_equals_the_oracle(bloch, name, n, p):
    """Pi: (H1)^3 -> ND of the auxiliary-space preconditioner (csrc/aux.cu, host code) on a topology-only handle against
    the oracle's assembled interpolation on the product's dof maps (incl. one-element periodic meshes)."""
    from helpers import oracle_on_product_maps
    from oracle.bloch_oracle import nodal_interpolation
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p, device=-2)
    ops, _ = oracle_on_product_maps(eq, name, n, p, np.ones(eq.n_elem))
    Pi = nodal_interpolation(ops.sp_)
    P2 = eq.pi_matrix()
    assert P2.shape == Pi.shape == (eq.N, 3 * eq.N_h1)
    assert abs(Pi - P2).max() < 1e-14
    assert P2.nnz <= eq.N * 3 * (p + 1)
