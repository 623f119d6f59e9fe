"""CPU oracle for the Bloch-periodic Maxwell eigen path of mlstowell/mfem-bravais.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mfem-bravais_b200/,
include/) may import this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
or as the timed CPU baseline.

PARITY UNPINNED for the finite-element operators: the reference has no golden
vectors, no assertions and cannot be built here (MFEM dev branch + hypre + MPI are
absent).  The LATTICE layer (class Lattice below: vectors,
symmetry points, k-paths, coarse Wigner-Seitz hex cells, periodic identification) IS
pinned: the reference's own lib/bravais.cpp is compiled unmodified against an
interface stand-in for MFEM (oracle/ref_shim, oracle/Makefile) and its output
(tests/golden/ref_bravais.json) is compared table by table in tests/test_ref_bravais.py.
For everything else this file is a *restatement* of the reference algorithm,
assembled-matrix style, with every third-party (MFEM) convention that is not visible in the
reference tree written down explicitly:

  * spaces: H1_p, ND_p, RT_{p-1} (= MFEM RT_FECollection(p-1),
    misc/pfem_extras_bloch.cpp:49-54) on affine hexahedra; closed 1-D basis =
    Lagrange on p+1 Gauss-Lobatto points, open 1-D basis = Lagrange on p
    Gauss-Legendre points (MFEM defaults BasisType::GaussLobatto / GaussLegendre);
  * ND dof functional k:  t_k . J^T v(x_k)      (MFEM Project_ND)
    RT dof functional k:  n_k . adj(J) v(x_k)   (MFEM Project_RT)
    ND physical shape  =  J^-T w_hat            (covariant Piola)
    RT physical shape  =  J w_hat / det J       (contravariant Piola);
  * T12 = discrete curl, T01 = discrete gradient (pfem_extras_bloch.cpp:128-146),
    Z12 = nodal RT interpolation of zeta x (ND shape) (pfem_extras_bloch.cpp:221-258),
    Z01 = nodal ND interpolation of zeta * (H1 shape) (pfem_extras_bloch.cpp:182-219);
  * M1 = ND mass(eps), M2 = RT mass(1/mu) with one constant per element
    (maxwell/maxwell_dispersion.cpp:398-420);
  * S1 = C^T M2 C + beta^2 Z^T M2 Z,  DKZ = C^T M2 Z - Z^T M2 C,
    A = [[S1, +beta DKZ], [-beta DKZ, S1]],  M = diag(M1, M1)
    (maxwell/maxwell_bloch.cpp:398-469), i.e. A_c = S1 - i beta DKZ acting on
    E = Er + i Ei stored [Er; Ei];
  * projector P = I - G (G^T M G)^-1 G^T M with G = [[T01, beta Z01], [-beta Z01, T01]]
    (maxwell/maxwell_bloch.cpp:2086-2096, 2280-2290), complex form G_c = T01 - i beta Z01.

Global DOF identification is done *geometrically* (node position modulo the
lattice, direction for the sign) - deliberately a different algorithm from the
product's entity-based numbering, so that agreement of the two is evidence.
"""
import math
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree

TWO_PI = 2.0 * math.pi

# --------------------------------------------------------------------------
# 1-D nodes and Lagrange bases on [0,1]
# --------------------------------------------------------------------------

def gauss_legendre(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto(n):
    """n >= 2 Gauss-Lobatto points on [0,1] (end points + roots of P'_{n-1})."""
    if n == 2:
        return np.array([0.0, 1.0])
    c = np.zeros(n)
    c[-1] = 1.0
    r = np.polynomial.legendre.Legendre(c).deriv().roots()
    x = np.concatenate([[-1.0], np.sort(r.real), [1.0]])
    x = 0.5 * (x - x[::-1])          # symmetrise
    return 0.5 * (x + 1.0)


def lagrange(nodes, x):
    """V[q,i] = l_i(x_q) and dV[q,i] = l_i'(x_q) for the Lagrange basis on nodes."""
    nodes = np.asarray(nodes, float)
    x = np.atleast_1d(np.asarray(x, float))
    n = len(nodes)
    V = np.ones((len(x), n))
    dV = np.zeros((len(x), n))
    for i in range(n):
        den = 1.0
        for j in range(n):
            if j != i:
                den *= nodes[i] - nodes[j]
        for j in range(n):
            if j != i:
                V[:, i] *= x - nodes[j]
        for k in range(n):
            if k == i:
                continue
            t = np.ones(len(x))
            for j in range(n):
                if j != i and j != k:
                    t *= x - nodes[j]
            dV[:, i] += t
        V[:, i] /= den
        dV[:, i] /= den
    return V, dV


# --------------------------------------------------------------------------
# Lattices (restated from lib/bravais.cpp; a = 1 unless given)
# --------------------------------------------------------------------------

class Lattice:
    """lat/rec vectors (a_i . b_j = delta_ij, no 2 pi), symmetry points, paths and
    the coarse Wigner-Seitz hex mesh.  kappa at a symmetry point = 2 pi * sp
    (lib/bravais.cpp:201-206)."""

    def __init__(self, name, a=1.0):
        self.name = name
        self.a = a
        if name == "CUB":        # lib/bravais.cpp:1881-1953, 2275-2313
            self.lat = a * np.eye(3)
            self.rec = np.eye(3) / a
            b = self.rec
            self.sp = {"Gamma": 0 * b[0], "X": 0.5 * b[1], "M": 0.5 * b[0] + 0.5 * b[1],
                       "R": 0.5 * (b[0] + b[1] + b[2])}
            self.paths = [["Gamma", "X", "M", "Gamma", "R", "X"], ["M", "R"]]
            h = 0.5 * a
            self.ws_vert = np.array([[-h, -h, -h], [h, -h, -h], [h, h, -h], [-h, h, -h],
                                     [-h, -h, h], [h, -h, h], [h, h, h], [-h, h, h]])
            self.ws_hex = np.array([[0, 1, 2, 3, 4, 5, 6, 7]])
        elif name == "FCC":      # lib/bravais.cpp:2361-2460, 2494-2526
            self.lat = a * np.array([[0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
            self.rec = np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1.]]) / a
            b = self.rec
            self.sp = {"Gamma": 0 * b[0], "X": 0.5 * b[0] + 0.5 * b[2],
                       "W": 0.5 * b[0] + 0.25 * b[1] + 0.75 * b[2],
                       "K": 0.375 * b[0] + 0.375 * b[1] + 0.75 * b[2],
                       "L": 0.5 * (b[0] + b[1] + b[2]),
                       "U": 0.625 * b[0] + 0.25 * b[1] + 0.625 * b[2]}
            self.paths = [["Gamma", "X", "W", "K", "Gamma", "L", "U", "W", "L", "K"], ["U", "X"]]
            v = np.zeros((15, 3))
            for d in range(3):
                v[1 + 2 * d, d] = -0.5 * a
                v[2 + 2 * d, d] = 0.5 * a
            for i in range(2):
                for j in range(2):
                    for k in range(2):
                        v[7 + 4 * i + 2 * j + k] = [(0.5 * i - 0.25) * a, (0.5 * j - 0.25) * a,
                                                    (0.5 * k - 0.25) * a]
            self.ws_vert = v
            self.ws_hex = np.array([[0, 9, 5, 11, 8, 1, 7, 3], [0, 11, 5, 9, 14, 2, 13, 4],
                                    [0, 8, 6, 14, 9, 1, 10, 4], [0, 14, 6, 8, 11, 2, 12, 3]])
        elif name == "BCC":      # lib/bravais.cpp:2655-2736, 2804-2893
            self.lat = a * np.array([[-.5, .5, .5], [.5, -.5, .5], [.5, .5, -.5]])
            self.rec = np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0.]]) / a
            b = self.rec
            self.sp = {"Gamma": 0 * b[0], "H": 0.5 * b[0] - 0.5 * b[1] + 0.5 * b[2],
                       "N": 0.5 * b[2], "P": 0.25 * (b[0] + b[1] + b[2])}
            self.paths = [["Gamma", "H", "N", "Gamma", "P", "H"], ["P", "N"]]
            q, h = 0.25, 0.5
            sq = []
            # 6 square faces x = -h, +h, y = -h, +h, z = -h, +h (4 vertices each), in the
            # reference's vertex order
            sq += [[-h, q, 0], [-h, 0, q], [-h, -q, 0], [-h, 0, -q]]
            sq += [[h, q, 0], [h, 0, q], [h, -q, 0], [h, 0, -q]]
            sq += [[q, -h, 0], [0, -h, q], [-q, -h, 0], [0, -h, -q]]
            sq += [[q, h, 0], [0, h, q], [-q, h, 0], [0, h, -q]]
            sq += [[q, 0, -h], [0, q, -h], [-q, 0, -h], [0, -q, -h]]
            sq += [[q, 0, h], [0, q, h], [-q, 0, h], [0, -q, h]]
            for i in (-q, q):
                for j in (-q, q):
                    for k in (-q, q):
                        sq.append([i, j, k])
            sq += [[-q, 0, 0], [q, 0, 0], [0, -q, 0], [0, q, 0], [0, 0, -q], [0, 0, q]]
            self.ws_vert = a * np.array(sq)
            self.ws_hex = np.array([
                [0, 1, 2, 3, 26, 32, 24, 18], [26, 32, 24, 18, 15, 35, 36, 17],
                [15, 35, 36, 17, 30, 33, 28, 16], [30, 33, 28, 16, 4, 5, 6, 7],
                [9, 8, 11, 10, 23, 29, 34, 25], [23, 29, 34, 25, 22, 37, 32, 1],
                [22, 37, 32, 1, 21, 31, 35, 27], [21, 31, 35, 27, 13, 12, 15, 14],
                [16, 17, 18, 19, 28, 36, 24, 11], [28, 36, 24, 11, 33, 35, 32, 34],
                [33, 35, 32, 34, 5, 31, 37, 29], [5, 31, 37, 29, 20, 21, 22, 23],
                [24, 11, 34, 32, 2, 10, 25, 1], [11, 28, 33, 34, 8, 6, 5, 29],
                [30, 15, 35, 33, 4, 12, 31, 5], [15, 26, 32, 35, 14, 0, 1, 27]])
        elif name == "HEX":      # lib/bravais.cpp:6023-6124, 6-hex layout :6167-6199 (a = c)
            c = a
            s3 = math.sqrt(3.0)
            self.lat = np.array([[0.5 * a, -math.sqrt(0.75) * a, 0], [0.5 * a, math.sqrt(0.75) * a, 0], [0, 0, c]])
            self.rec = np.array([[1 / a, -1 / (s3 * a), 0], [1 / a, 1 / (s3 * a), 0], [0, 0, 1 / c]])
            b = self.rec
            self.sp = {"Gamma": 0 * b[0], "A": 0.5 * b[2], "H": (b[0] + b[1]) / 3 + 0.5 * b[2],
                       "K": (b[0] + b[1]) / 3, "L": 0.5 * b[0] + 0.5 * b[2], "M": 0.5 * b[0]}
            self.paths = [["Gamma", "M", "K", "Gamma", "A", "L", "H", "A"], ["L", "M"], ["K", "H"]]
            h = a / s3
            ring = [[0, -h], [0.5 * a, -0.5 * h], [0.5 * a, 0.5 * h], [0, h], [-0.5 * a, 0.5 * h], [-0.5 * a, -0.5 * h]]
            v = []
            for layer in range(3):
                z = 0.5 * c * (layer - 1)
                v.append([0, 0, z])
                v += [[x, y, z] for x, y in ring]
            self.ws_vert = np.array(v, float)
            hexes = []
            for i in range(2):
                for j in range(3):
                    o = 7 * i
                    q = [0 + o, 2 * j + 1 + o, 2 * j + 2 + o, ((2 * j + 2) % 6) + 1 + o]
                    hexes.append(q + [k + 7 for k in q])
            self.ws_hex = np.array(hexes)
        else:
            raise ValueError("oracle lattice %r not restated" % name)
        self.volume = abs(np.linalg.det(self.lat))

    def kappa(self, label):
        return TWO_PI * self.sp[label]

    def kpath(self, labels, npts):
        """npts points per segment: kappa0 + (i/npts)(kappa1-kappa0), i = 1..npts
        (the start point of each segment is the end point of the previous one)."""
        ks = []
        for s in range(len(labels) - 1):
            k0, k1 = self.kappa(labels[s]), self.kappa(labels[s + 1])
            for i in range(1, npts + 1):
                ks.append(k0 + (i / npts) * (k1 - k0))
        return np.array(ks)


# --------------------------------------------------------------------------
# Mesh: affine hexes (x0, J), n^3 uniform subdivision of each coarse hex
# --------------------------------------------------------------------------

class Mesh:
    def __init__(self, lattice, n):
        self.lattice = lattice
        self.n = n
        x0s, Js, cls = [], [], []
        for c, hexv in enumerate(lattice.ws_hex):
            V = lattice.ws_vert[hexv]
            J = np.stack([V[1] - V[0], V[3] - V[0], V[4] - V[0]], axis=1)   # MFEM vertex order
            # is_affine assert: remaining vertices follow the affine map
            ref = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0],
                            [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1.]])
            assert np.allclose(V[0] + ref @ J.T, V, atol=1e-12), "coarse hex %d not affine" % c
            if np.linalg.det(J) < 0:        # MFEM fixes inverted elements by reordering
                J = J[:, [1, 0, 2]]
            Jn = J / n
            # element order: coarse cell, then k, j, i (i fastest)
            kk, jj, ii = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
            ijk = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(float)
            x0s.append(V[0] + ijk @ Jn.T)
            Js.append(Jn)
            cls.append(np.full(n ** 3, c))
        self.x0 = np.concatenate(x0s)
        self.cls = np.concatenate(cls)
        self.J = np.array(Js)                    # per class
        self.ne = len(self.x0)
        self.centers = self.x0 + np.einsum("cij,j->ci", self.J, [0.5, 0.5, 0.5])[self.cls]
        self.volume = float(np.sum(np.abs(np.linalg.det(self.J))[self.cls]))

    def sphere_eps(self, radius=0.25, eps_in=10.0, eps_out=1.0):
        """mass_coef case 2 sampled at element centres (maxwell_dispersion.cpp:1464-1467)."""
        r = np.linalg.norm(self.centers, axis=1)
        return np.where(r <= radius, eps_in, eps_out)


# --------------------------------------------------------------------------
# Reference element: node positions, directions and shapes
# Local ordering (shared with the product so dof maps can be exchanged):
#   ND : [x-comp (p, p+1, p+1)] [y-comp (p+1, p, p+1)] [z-comp (p+1, p+1, p)], i fastest
#   RT : [x-comp (p+1, p, p)]   [y-comp (p, p+1, p)]   [z-comp (p, p, p+1)]
#   H1 : (p+1)^3, i fastest
# --------------------------------------------------------------------------

class RefElem:
    def __init__(self, p):
        self.p = p
        self.g, self.wg = gauss_legendre(p)
        self.l = gauss_lobatto(p + 1)
        g, l = self.g, self.l

        def grid(nx, ny, nz):
            kk, jj, ii = np.meshgrid(nz, ny, nx, indexing="ij")
            return np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)

        self.nd_nodes = np.concatenate([grid(g, l, l), grid(l, g, l), grid(l, l, g)])
        self.nd_comp = np.repeat([0, 1, 2], p * (p + 1) ** 2)
        self.rt_nodes = np.concatenate([grid(l, g, g), grid(g, l, g), grid(g, g, l)])
        self.rt_comp = np.repeat([0, 1, 2], p * p * (p + 1))
        self.h1_nodes = grid(l, l, l)
        self.n_nd, self.n_rt, self.n_h1 = len(self.nd_nodes), len(self.rt_nodes), len(self.h1_nodes)

    def _tp(self, fx, fy, fz):
        """tensor product of 1-D tables f[q, i] -> [q, i + nx*(j + ny*k)]"""
        return np.einsum("qi,qj,qk->qkji", fx, fy, fz).reshape(fx.shape[0], -1)

    def nd_shapes(self, X):
        """value [Q, n_nd, 3] and reference curl [Q, n_nd, 3] at points X[Q,3]."""
        p = self.p
        O = [lagrange(self.g, X[:, d]) for d in range(3)]
        C = [lagrange(self.l, X[:, d]) for d in range(3)]
        Q = X.shape[0]
        val = np.zeros((Q, self.n_nd, 3))
        curl = np.zeros((Q, self.n_nd, 3))
        nb = p * (p + 1) ** 2
        # x-comp u = o(x) c(y) c(z): curl = (0, du/dz, -du/dy)
        val[:, 0:nb, 0] = self._tp(O[0][0], C[1][0], C[2][0])
        curl[:, 0:nb, 1] = self._tp(O[0][0], C[1][0], C[2][1])
        curl[:, 0:nb, 2] = -self._tp(O[0][0], C[1][1], C[2][0])
        # y-comp u = c(x) o(y) c(z): curl = (-du/dz, 0, du/dx)
        val[:, nb:2 * nb, 1] = self._tp(C[0][0], O[1][0], C[2][0])
        curl[:, nb:2 * nb, 0] = -self._tp(C[0][0], O[1][0], C[2][1])
        curl[:, nb:2 * nb, 2] = self._tp(C[0][1], O[1][0], C[2][0])
        # z-comp u = c(x) c(y) o(z): curl = (du/dy, -du/dx, 0)
        val[:, 2 * nb:, 2] = self._tp(C[0][0], C[1][0], O[2][0])
        curl[:, 2 * nb:, 0] = self._tp(C[0][0], C[1][1], O[2][0])
        curl[:, 2 * nb:, 1] = -self._tp(C[0][1], C[1][0], O[2][0])
        return val, curl

    def rt_shapes(self, X):
        p = self.p
        O = [lagrange(self.g, X[:, d])[0] for d in range(3)]
        C = [lagrange(self.l, X[:, d])[0] for d in range(3)]
        val = np.zeros((X.shape[0], self.n_rt, 3))
        nb = p * p * (p + 1)
        val[:, 0:nb, 0] = self._tp(C[0], O[1], O[2])
        val[:, nb:2 * nb, 1] = self._tp(O[0], C[1], O[2])
        val[:, 2 * nb:, 2] = self._tp(O[0], O[1], C[2])
        return val

    def h1_shapes(self, X):
        C = [lagrange(self.l, X[:, d]) for d in range(3)]
        val = self._tp(C[0][0], C[1][0], C[2][0])
        grad = np.stack([self._tp(C[0][1], C[1][0], C[2][0]),
                         self._tp(C[0][0], C[1][1], C[2][0]),
                         self._tp(C[0][0], C[1][0], C[2][1])], axis=2)
        return val, grad

    def quadrature(self, nq=None):
        nq = nq or self.p + 2
        x, w = gauss_legendre(nq)
        kk, jj, ii = np.meshgrid(np.arange(nq), np.arange(nq), np.arange(nq), indexing="ij")
        X = np.stack([x[ii.ravel()], x[jj.ravel()], x[kk.ravel()]], axis=1)
        W = w[ii.ravel()] * w[jj.ravel()] * w[kk.ravel()]
        return X, W


def element_matrices(ref, J, zeta):
    """Element matrices of one affine class with Jacobian J (unit coefficients).
    Every formula is the literal MFEM definition (Piola maps + dof functionals)."""
    detJ = np.linalg.det(J)
    assert detJ > 0
    Jinv = np.linalg.inv(J)
    adjJ = detJ * Jinv
    X, W = ref.quadrature()
    nd, _ = ref.nd_shapes(X)
    rt = ref.rt_shapes(X)
    nd_phys = np.einsum("ji,qaj->qai", Jinv, nd)            # J^-T w
    rt_phys = np.einsum("ij,qaj->qai", J, rt) / detJ        # J w / det
    M1 = np.einsum("q,qai,qbi->ab", W * detJ, nd_phys, nd_phys)
    M2 = np.einsum("q,qai,qbi->ab", W * detJ, rt_phys, rt_phys)
    h1, _ = ref.h1_shapes(X)
    M0 = np.einsum("q,qa,qb->ab", W * detJ, h1, h1)
    # interpolation matrices: dof functional of (operator applied to shape l)
    eye = np.eye(3)
    ndv, ndc = ref.nd_shapes(ref.rt_nodes)                  # at RT nodes
    curl_phys = np.einsum("ij,qaj->qai", J, ndc) / detJ
    ndv_phys = np.einsum("ji,qaj->qai", Jinv, ndv)
    nk = eye[ref.rt_comp]                                   # [n_rt, 3]
    T12 = np.einsum("ki,ij,kaj->ka", nk, adjJ, curl_phys)
    Z12 = None
    Z01 = None
    h1v, h1g = ref.h1_shapes(ref.nd_nodes)                  # at ND nodes
    grad_phys = np.einsum("ji,qaj->qai", Jinv, h1g)
    tk = eye[ref.nd_comp]
    T01 = np.einsum("ki,ji,kaj->ka", tk, J, grad_phys)      # t . J^T v
    if zeta is not None:
        zx = np.cross(np.broadcast_to(zeta, ndv_phys.shape), ndv_phys)
        Z12 = np.einsum("ki,ij,kaj->ka", nk, adjJ, zx)
        zphi = h1v[:, :, None] * np.asarray(zeta)[None, None, :]
        Z01 = np.einsum("ki,ji,kaj->ka", tk, J, zphi)
    return dict(M0=M0, M1=M1, M2=M2, T12=T12, Z12=Z12, T01=T01, Z01=Z01)


# --------------------------------------------------------------------------
# Geometric global DOF identification
# --------------------------------------------------------------------------

def _identify(mesh, nodes_ref, dirs_phys_per_class, tol=1e-7):
    """nodes_ref[L,3] reference nodes; dirs_phys_per_class[c][L,3] physical direction
    of each local dof (None for scalar).  Returns gid[ne,L], sign[ne,L], nglobal."""
    L = len(nodes_ref)
    ne = mesh.ne
    pts = mesh.x0[:, None, :] + np.einsum("eij,lj->eli", mesh.J[mesh.cls], nodes_ref)
    frac = pts.reshape(-1, 3) @ mesh.lattice.rec.T
    frac -= np.floor(frac)
    frac[frac >= 1.0] = 0.0
    tree = cKDTree(frac, boxsize=1.0)
    pairs = tree.query_pairs(tol, output_type="ndarray")
    m = ne * L
    g = sp.coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(m, m))
    ncomp, lab = connected_components(g, directed=False)
    # renumber by first occurrence
    _, first = np.unique(lab, return_index=True)
    order = np.argsort(first)
    remap = np.empty(ncomp, dtype=np.int64)
    remap[order] = np.arange(ncomp)
    gid = remap[lab].reshape(ne, L)
    sign = np.ones((ne, L))
    if dirs_phys_per_class is not None:
        d = np.array(dirs_phys_per_class)[mesh.cls].reshape(-1, 3)       # [m,3]
        d = d / np.linalg.norm(d, axis=1, keepdims=True)
        rep = np.zeros((ncomp, 3))
        flat = gid.ravel()
        # representative = first member's direction, made lexicographically positive
        firstmember = np.full(ncomp, -1)
        firstmember[flat[::-1]] = np.arange(m)[::-1]
        rep = d[firstmember].copy()
        for i in range(ncomp):
            v = rep[i]
            for c in range(3):
                if abs(v[c]) > 1e-9:
                    if v[c] < 0:
                        rep[i] = -v
                    break
        dots = np.einsum("mi,mi->m", d, rep[flat])
        assert np.all(np.abs(np.abs(dots) - 1.0) < 1e-8), "dof directions not parallel"
        sign = np.sign(dots).reshape(ne, L)
    return gid, sign, ncomp


class Spaces:
    """H1_p, ND_p, RT_{p-1} on the periodic mesh."""

    def __init__(self, mesh, p, dofmaps=None):
        self.mesh, self.p = mesh, p
        self.ref = ref = RefElem(p)
        eye = np.eye(3)
        if dofmaps is None:
            nd_dirs = [np.einsum("ij,lj->li", J, eye[ref.nd_comp]) for J in mesh.J]        # J t
            # physical normal (area vector) of RT dof: adj(J)^T n = det * J^-T n
            rt_dirs = [np.einsum("ji,lj->li", np.linalg.inv(J), eye[ref.rt_comp]) for J in mesh.J]
            self.nd_gid, self.nd_sign, self.n_nd = _identify(mesh, ref.nd_nodes, nd_dirs)
            self.rt_gid, self.rt_sign, self.n_rt = _identify(mesh, ref.rt_nodes, rt_dirs)
            self.h1_gid, _, self.n_h1 = _identify(mesh, ref.h1_nodes, None)
        else:   # maps exported by the product (signed 1-based style: gid, sign arrays)
            self.nd_gid, self.nd_sign = dofmaps["nd_gid"], dofmaps["nd_sign"]
            self.h1_gid = dofmaps["h1_gid"]
            self.n_nd = int(self.nd_gid.max()) + 1
            self.n_h1 = int(self.h1_gid.max()) + 1
            if "rt_gid" in dofmaps:
                self.rt_gid, self.rt_sign = dofmaps["rt_gid"], dofmaps["rt_sign"]
                self.n_rt = int(self.rt_gid.max()) + 1
            else:
                rt_dirs = [np.einsum("ji,lj->li", np.linalg.inv(J), eye[ref.rt_comp]) for J in mesh.J]
                self.rt_gid, self.rt_sign, self.n_rt = _identify(mesh, ref.rt_nodes, rt_dirs)
        self.h1_sign = np.ones_like(self.h1_gid, dtype=float)


def _assemble_sum(gr, sr, gc, sc, nrow, ncol, elmats, cls, coef):
    """sum_e coef_e * S_r elmat[cls_e] S_c  into a CSR matrix"""
    ne, Lr = gr.shape
    Lc = gc.shape[1]
    vals = elmats[cls] * coef[:, None, None] * sr[:, :, None] * sc[:, None, :]
    rows = np.broadcast_to(gr[:, :, None], (ne, Lr, Lc))
    cols = np.broadcast_to(gc[:, None, :], (ne, Lr, Lc))
    A = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(nrow, ncol)).tocsr()
    A.eliminate_zeros()
    return A


def _assemble_set(gr, sr, gc, sc, nrow, ncol, elmats, cls):
    """interpolation-type operator: entries are *set* (not added); duplicates coming
    from neighbouring elements must agree - checked."""
    ne, Lr = gr.shape
    Lc = gc.shape[1]
    vals = elmats[cls] * sr[:, :, None] * sc[:, None, :]
    rows = np.broadcast_to(gr[:, :, None], (ne, Lr, Lc)).ravel()
    cols = np.broadcast_to(gc[:, None, :], (ne, Lr, Lc)).ravel()
    v = vals.ravel()
    # A local row owns the full global row: average over the local-row occurrences.
    rowcount = np.bincount(gr.ravel(), minlength=nrow).astype(float)
    S = sp.coo_matrix((v, (rows, cols)), shape=(nrow, ncol)).tocsr()
    A = (sp.diags(1.0 / rowcount) @ S).tocsr()
    # consistency (conformity) check: restricted to any element the global operator
    # reproduces the element matrix:  S_r (A u)[gr_e] == elmat (S_c u[gc_e])
    u = np.random.default_rng(7).uniform(-1, 1, ncol)
    lhs = sr * (A @ u)[gr]
    rhs = np.einsum("ekl,el->ek", elmats[cls], sc * u[gc])
    err = np.abs(lhs - rhs).max() / max(1.0, np.abs(rhs).max())
    assert err < 1e-10, "interpolation operator not conforming across elements: %g" % err
    A.data[np.abs(A.data) < 1e-14] = 0.0
    A.eliminate_zeros()
    return A


def nodal_interpolation(spaces):
    """Pi: (H1_p)^3 -> ND_p, the nodal interpolation of continuous vector fields - the 'Pi' matrix HypreAMS builds
    for the preconditioner the reference creates at maxwell/maxwell_bloch.cpp:492-517 (MFEM: DiscreteLinearOperator
    with an IdentityInterpolator).  ND dof functional: v -> (J t) . v(x_node).  Columns: component d of H1 node g
    at d * n_h1 + g.  With a constant vector zeta, Pi [zeta_x phi; zeta_y phi; zeta_z phi] is the reference's
    Z01 phi (pfem_extras_bloch.cpp:182-258) - the identity the CPU test pins this matrix with."""
    ref, mesh = spaces.ref, spaces.mesh
    h1v, _ = ref.h1_shapes(ref.nd_nodes)            # [n_nd_loc, n_h1_loc]
    em = []
    for J in mesh.J:
        E = np.zeros((ref.n_nd, 3 * ref.n_h1))
        for d in range(3):
            E[:, d * ref.n_h1:(d + 1) * ref.n_h1] = h1v * J[d, ref.nd_comp][:, None]
        em.append(E)
    gc = np.concatenate([spaces.h1_gid + d * spaces.n_h1 for d in range(3)], axis=1)
    return _assemble_set(spaces.nd_gid, spaces.nd_sign, gc, np.ones_like(gc, dtype=float), spaces.n_nd,
                         3 * spaces.n_h1, np.array(em), mesh.cls)


class BlochOperators:
    """Assembled operators of MaxwellBlochWaveEquation::Setup for one (mesh, p, eps, muinv)
    and, per kappa, the beta/zeta dependent ones."""

    def __init__(self, spaces, eps, muinv=None):
        self.sp_ = s = spaces
        mesh = s.mesh
        self.eps = np.asarray(eps, float)
        self.muinv = np.ones(mesh.ne) if muinv is None else np.asarray(muinv, float)
        em = [element_matrices(s.ref, J, None) for J in mesh.J]
        one = np.ones(mesh.ne)
        st = lambda key: np.array([e[key] for e in em])
        self.M1 = _assemble_sum(s.nd_gid, s.nd_sign, s.nd_gid, s.nd_sign, s.n_nd, s.n_nd,
                                st("M1"), mesh.cls, self.eps)
        self.M2 = _assemble_sum(s.rt_gid, s.rt_sign, s.rt_gid, s.rt_sign, s.n_rt, s.n_rt,
                                st("M2"), mesh.cls, self.muinv)
        self.M0 = _assemble_sum(s.h1_gid, s.h1_sign, s.h1_gid, s.h1_sign, s.n_h1, s.n_h1,
                                st("M0"), mesh.cls, one)
        self.T12 = _assemble_set(s.rt_gid, s.rt_sign, s.nd_gid, s.nd_sign, s.n_rt, s.n_nd,
                                 st("T12"), mesh.cls)
        self.T01 = _assemble_set(s.nd_gid, s.nd_sign, s.h1_gid, s.h1_sign, s.n_nd, s.n_h1,
                                 st("T01"), mesh.cls)
        self.CMC = (self.T12.T @ self.M2 @ self.T12).tocsr()
        self.kappa = None

    def set_kappa(self, kappa):
        """SetKappa + the beta/zeta part of Setup (maxwell_bloch.cpp:200-210, 375-425) and of
        the projector Setup (maxwell_bloch.cpp:2064-2128)."""
        s, mesh = self.sp_, self.sp_.mesh
        kappa = np.asarray(kappa, float)
        self.kappa = kappa
        self.beta = beta = float(np.linalg.norm(kappa))
        if beta > 0.0:
            self.zeta = zeta = kappa / beta
            em = [element_matrices(s.ref, J, zeta) for J in mesh.J]
            st = lambda key: np.array([e[key] for e in em])
            self.Z12 = _assemble_set(s.rt_gid, s.rt_sign, s.nd_gid, s.nd_sign, s.n_rt, s.n_nd,
                                     st("Z12"), mesh.cls)
            self.Z01 = _assemble_set(s.nd_gid, s.nd_sign, s.h1_gid, s.h1_sign, s.n_nd, s.n_h1,
                                     st("Z01"), mesh.cls)
            T12, Z12, M2 = self.T12, self.Z12, self.M2
            ZMZ = Z12.T @ M2 @ Z12
            CMZ = T12.T @ M2 @ Z12
            ZMC = Z12.T @ M2 @ T12
            self.DKZ = (CMZ - ZMC).tocsr()
            self.S1 = (self.CMC + beta * beta * ZMZ).tocsr()
            T01, Z01, M1 = self.T01, self.Z01, self.M1
            GMG = T01.T @ M1 @ T01
            self.DKZ0 = (Z01.T @ M1 @ T01 - T01.T @ M1 @ Z01).tocsr()
            self.A0 = (GMG + beta * beta * (Z01.T @ M1 @ Z01)).tocsr()
        else:
            self.zeta = np.zeros(3)
            self.S1 = self.CMC
            self.DKZ = None
            self.A0 = (self.T01.T @ self.M1 @ self.T01).tocsr()
            self.DKZ0 = None
        return self

    # ---- real 2x2 block form, exactly as the reference stores it ----
    def A_block(self):
        if self.DKZ is None:
            return sp.bmat([[self.S1, None], [None, self.S1]]).tocsr()
        b = self.beta
        return sp.bmat([[self.S1, b * self.DKZ], [-b * self.DKZ, self.S1]]).tocsr()

    def M_block(self):
        return sp.bmat([[self.M1, None], [None, self.M1]]).tocsr()

    # ---- complex forms ----
    def A_c(self):
        if self.DKZ is None:
            return self.S1.astype(complex)
        return (self.S1 - 1j * self.beta * self.DKZ).tocsr()

    def M_c(self):
        return self.M1.astype(complex)

    def C_c(self):
        if self.DKZ is None:
            return self.T12.astype(complex)
        return (self.T12 - 1j * self.beta * self.Z12).tocsr()

    def G_c(self):
        if self.DKZ0 is None:
            return self.T01.astype(complex)
        return (self.T01 - 1j * self.beta * self.Z01).tocsr()

    def S0_c(self):
        """G^T M G in complex form: [[A0, -beta DKZ0], [beta DKZ0, A0]] (maxwell_bloch.cpp:2130-2144)"""
        if self.DKZ0 is None:
            return self.A0.astype(complex)
        return (self.A0 + 1j * self.beta * self.DKZ0).tocsr()

    # ---- applications on [re; im] stored vectors (the C-ABI layout) ----
    @staticmethod
    def to_c(x, n):
        return x[..., :n] + 1j * x[..., n:]

    @staticmethod
    def to_ri(z):
        return np.concatenate([z.real, z.imag], axis=-1)

    def apply_A(self, x):
        return (self.A_block() @ x.T).T

    def apply_M(self, x):
        return (self.M_block() @ x.T).T

    def projector_dense(self):
        G = self.G_c().toarray()
        M = self.M_c().toarray()
        S0 = G.conj().T @ M @ G
        return np.eye(G.shape[0]) - G @ np.linalg.solve(S0, G.conj().T @ M)

    def apply_projector(self, xc):
        """y = x - G S0^-1 G^H M x, complex vectors xc[nvec, N] (maxwell_bloch.cpp:2280-2290)."""
        import scipy.sparse.linalg as spla
        G, M = self.G_c(), self.M_c()
        S0 = (G.conj().T @ M @ G).tocsc()
        lu = spla.splu(S0)
        rhs = G.conj().T @ (M @ xc.T)
        return (xc.T - G @ lu.solve(np.ascontiguousarray(rhs))).T

    # ---- field averages ----
    def field_average_forms(self):
        """The 24 linear forms of SetKappa (maxwell/maxwell_bloch.cpp:211-279), assembled literally:
        L[w] = int coef(x) {cos, sin}(kappa.x) e_i . w(x) dx with coef in {1, eps} on ND and
        {1, mu^-1} on RT; quadrature = MFEM's default for VectorFEDomainLFIntegrator
        (2 * element order -> p + 1 Gauss points per direction).  Returned as complex vectors
        cos + i sin: dict name -> [3, n]."""
        s, mesh = self.sp_, self.sp_.mesh
        ref = s.ref
        X, W = ref.quadrature(ref.p + 1)
        nd, _ = ref.nd_shapes(X)
        rt = ref.rt_shapes(X)
        out = {k: np.zeros((3, n), complex) for k, n in
               (("E", s.n_nd), ("D", s.n_nd), ("B", s.n_rt), ("H", s.n_rt))}
        for e in range(mesh.ne):
            J = mesh.J[mesh.cls[e]]
            det = np.linalg.det(J)
            xq = mesh.x0[e] + X @ J.T
            ph = np.exp(1j * (xq @ self.kappa)) * W * det                      # [Q]
            nd_phys = np.einsum("ji,qaj->qai", np.linalg.inv(J), nd)          # J^-T w
            rt_phys = np.einsum("ij,qaj->qai", J, rt) / det                   # J w / det
            le = np.einsum("q,qai->ia", ph, nd_phys) * s.nd_sign[e][None, :]
            lb = np.einsum("q,qai->ia", ph, rt_phys) * s.rt_sign[e][None, :]
            for i in range(3):
                np.add.at(out["E"][i], s.nd_gid[e], le[i])
                np.add.at(out["D"][i], s.nd_gid[e], self.eps[e] * le[i])
                np.add.at(out["B"][i], s.rt_gid[e], lb[i])
                np.add.at(out["H"][i], s.rt_gid[e], self.muinv[e] * lb[i])
        return out

    def field_averages(self, Ec, lam):
        """GetFieldAverages (maxwell/maxwell_bloch.cpp:1550-1632) of the mode Ec = Er + i Ei with
        eigenvalue lam.  (Br, Bi) follow GetEigenvectorB (:1432-1457): Bi = Re(C E)/sqrt|lam|,
        Br = -Im(C E)/sqrt|lam|.  Er_avg = cos.Er - sin.Ei, Ei_avg = sin.Er + cos.Ei etc., i.e. the
        plain (non-conjugated) product of (cos + i sin) with (Fr + i Fi)."""
        L = self.field_average_forms()
        CE = self.C_c() @ Ec
        sc = 1.0 / np.sqrt(abs(lam)) if abs(lam) > 0 else 1.0
        Bc = sc * (-CE.imag + 1j * CE.real)
        return {"E": L["E"] @ Ec, "D": L["D"] @ Ec, "B": L["B"] @ Bc, "H": L["H"] @ Bc}

    # ---- eigen-solves ----
    def eig_dense(self, nev):
        """Lowest nev eigenvalues of the pencil (A_c, M_c) restricted to {x : G_c^H M x = 0}
        (kappa != 0) or to the M-orthogonal complement of range(T01) (kappa == 0), dense."""
        import scipy.linalg as sla
        A, M, G = self.A_c().toarray(), self.M_c().toarray(), self.G_c().toarray()
        Cn = G.conj().T @ M                                  # constraint rows
        # null space of Cn
        _, sv, Vh = np.linalg.svd(Cn, full_matrices=True)
        rank = int(np.sum(sv > 1e-10 * sv[0]))
        Q = Vh[rank:].conj().T
        Ar, Mr = Q.conj().T @ A @ Q, Q.conj().T @ M @ Q
        Ar = 0.5 * (Ar + Ar.conj().T)
        Mr = 0.5 * (Mr + Mr.conj().T)
        w = sla.eigh(Ar, Mr, eigvals_only=True)
        return w[:nev]

    def eig_shift_invert(self, nev, sigma, extra=12):
        """ARPACK shift-invert around sigma with a sparse LU; returns the nev lowest
        eigenvalues above zero_tol found near sigma.  Needs lambda_nev < 2 sigma so that the
        gradient null space (lambda = 0) is farther from sigma than the wanted bands."""
        import scipy.sparse.linalg as spla
        A, M = self.A_c().tocsc(), self.M_c().tocsc()
        k = nev + extra
        w = spla.eigsh(A, k=k, M=M, sigma=sigma, which="LM", return_eigenvectors=False, tol=1e-12)
        w = np.sort(w.real)
        return w


class ScalarOperators:
    """Scalar H1 Bloch Helmholtz variant, ScalarFloquetWaveEquation::Setup (misc/scalar3d.cpp:662-818):
    S0 = G^T M1(k) G + b^2 Z^T M1(k) Z,  DKZ = G^T M1 Z - Z^T M1 G,  A = [[S0, b DKZ], [-b DKZ, S0]],
    M = diag(M0(m), M0(m)), with b = beta * pi / 180 (beta in degrees) and kappa = b * zeta."""

    def __init__(self, spaces, k_coef, m_coef):
        s, mesh = spaces, spaces.mesh
        self.sp_ = s
        em = [element_matrices(s.ref, J, None) for J in mesh.J]
        st = lambda key: np.array([e[key] for e in em])
        self.M1 = _assemble_sum(s.nd_gid, s.nd_sign, s.nd_gid, s.nd_sign, s.n_nd, s.n_nd, st("M1"), mesh.cls,
                                np.asarray(k_coef, float))
        self.M0 = _assemble_sum(s.h1_gid, s.h1_sign, s.h1_gid, s.h1_sign, s.n_h1, s.n_h1, st("M0"), mesh.cls,
                                np.asarray(m_coef, float))
        self.T01 = _assemble_set(s.nd_gid, s.nd_sign, s.h1_gid, s.h1_sign, s.n_nd, s.n_h1, st("T01"), mesh.cls)

    def set_kappa(self, kappa):
        s, mesh = self.sp_, self.sp_.mesh
        kappa = np.asarray(kappa, float)
        self.b = b = float(np.linalg.norm(kappa))
        G, M1 = self.T01, self.M1
        GMG = G.T @ M1 @ G
        if b > 0:
            zeta = kappa / b
            em = [element_matrices(s.ref, J, zeta) for J in mesh.J]
            Z = _assemble_set(s.nd_gid, s.nd_sign, s.h1_gid, s.h1_sign, s.n_nd, s.n_h1,
                              np.array([e["Z01"] for e in em]), mesh.cls)
            self.S0 = (GMG + b * b * (Z.T @ M1 @ Z)).tocsr()
            self.DKZ = (G.T @ M1 @ Z - Z.T @ M1 @ G).tocsr()
        else:
            self.S0, self.DKZ = GMG.tocsr(), None
        return self

    def A_c(self):
        return self.S0.astype(complex) if self.DKZ is None else (self.S0 - 1j * self.b * self.DKZ).tocsr()

    def A_block(self):
        if self.DKZ is None:
            return sp.bmat([[self.S0, None], [None, self.S0]]).tocsr()
        return sp.bmat([[self.S0, self.b * self.DKZ], [-self.b * self.DKZ, self.S0]]).tocsr()

    def M_block(self):
        return sp.bmat([[self.M0, None], [None, self.M0]]).tocsr()

    def eig_dense(self, nev):
        import scipy.linalg as sla
        return sla.eigh(self.A_c().toarray(), self.M0.toarray().astype(complex), eigvals_only=True)[:nev]


def empty_lattice_eigs(lattice, kappa, nev, nmax=3):
    """Exact spectrum for eps = mu = 1: |kappa + 2 pi sum n_i b_i|^2, multiplicity 2
    (the plane-wave ansatz behind CreateInitialVectors, maxwell_dispersion.cpp:869-922)."""
    r = range(-nmax, nmax + 1)
    vals = []
    for i in r:
        for j in r:
            for k in r:
                q = kappa + TWO_PI * (i * lattice.rec[0] + j * lattice.rec[1] + k * lattice.rec[2])
                vals += [q @ q, q @ q]
    return np.sort(np.array(vals))[:nev]
