"""maxwell_dispersion-style sweep logic (maxwell/maxwell_dispersion.cpp:475-648, 1062-1087,
1449-1556): k-path walk with symmetry-point cache, piecewise-constant coefficients sampled at
element centres, omega = sqrt(lambda) formatting."""
import os

import numpy as np


def sphere_eps(centers, radius=0.25, eps_in=10.0, eps_out=1.0):
    """mass_coef case 2 (maxwell_dispersion.cpp:1464-1467) at element centres."""
    r = np.linalg.norm(centers, axis=1)
    return np.where(r <= radius, eps_in, eps_out)


def lattice_coefficient(lattice, centers, frac=0.5, val0=0.0, val1=1.0):
    """bravais::LatticeCoefficient (lib/bravais.cpp:9834-9862) at element centres: rods along the lattice's
    translation vectors, radius frac * (inscribed face radius); val1 inside any rod, val0 elsewhere."""
    x = np.asarray(centers, float)
    out = np.full(len(x), float(val0))
    T, R = lattice.GetTranslationVectors(), lattice.GetFaceRadii()
    for t, r in zip(T, R):
        a = t / np.linalg.norm(t)
        perp = x - np.outer(x @ a, a)
        out[np.linalg.norm(perp, axis=1) < frac * r] = val1
    return out


def k_path(lattice, labels, npts):
    """kappa points along the path through the named symmetry points, `npts` per segment:
    kappa0 + i/npts (kappa1 - kappa0), i = 1..npts (segment start points are the previous
    segment's end points, i.e. already-solved symmetry points, maxwell_dispersion.cpp:506)."""
    ks = []
    for s in range(len(labels) - 1):
        k0 = lattice.GetSymmetryPoint(lattice.GetSymmetryPointIndex(labels[s]))
        k1 = lattice.GetSymmetryPoint(lattice.GetSymmetryPointIndex(labels[s + 1]))
        for i in range(1, npts + 1):
            ks.append(k0 + (i / npts) * (k1 - k0))
    return np.array(ks)


def omega_of_lambda(lam):
    """disp.dat convention (maxwell_dispersion.cpp:1072-1083)"""
    lam = np.asarray(lam, float)
    out = np.full(lam.shape, -1.0)
    out[lam > 0] = np.sqrt(lam[lam > 0])
    out[(lam <= 0) & (lam > -1e-6)] = 0.0
    return out


def dispersion_sweep(eq, kappas, n_bands, tol=1e-6, max_iter=2000):
    """Solves the listed k-points one after another on one handle; returns (lambda[nk, n_bands], stats)."""
    eq.SetNumEigs(2 * n_bands)
    eq.SetAbsoluteTolerance(tol, max_iter)
    out, stats = [], []
    for k in kappas:
        eq.SetKappa(k)
        eq.Setup()
        eq.Solve()
        out.append(eq.band_eigenvalues())
        stats.append(eq.GetSolverStats())
    return np.array(out), stats


def shard_kpoints(n_points, world_size, rank):
    """Contiguous chunk [lo, hi) of the k-point list owned by `rank`:
    contiguity keeps consecutive k-points on one GPU so eigenvectors can warm-start the next
    solve.  The first n_points % world_size ranks get one extra point."""
    base, extra = divmod(n_points, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_sweep(solve_fn, kappas, n_bands, dist=None):
    """Solves every k-point exactly once across the ranks of `dist` (a torch.distributed-like
    module, None = single process): rank r solves its contiguous chunk with `solve_fn(kappa) ->
    array of n_bands eigenvalues`; the per-rank blocks are gathered (no collective on the solve
    path, one all_gather of n_bands doubles per k-point at the end) and returned in path order on
    every rank as an array [len(kappas), n_bands]."""
    import numpy as _np
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = shard_kpoints(len(kappas), world, rank)
    mine = _np.zeros((hi - lo, n_bands))
    for i in range(lo, hi):
        mine[i - lo] = _np.asarray(solve_fn(kappas[i]))[:n_bands]
    if dist is None:
        return mine
    parts = [None] * world
    dist.all_gather_object(parts, (lo, mine))
    out = _np.zeros((len(kappas), n_bands))
    for plo, block in parts:
        out[plo:plo + len(block)] = block
    return out


def slot_chunks(n_points, n_slots):
    """Contiguous chunks of the index range [0, n_points) for the slots of a batched sweep (slot = one k-point
    position of one handle's batch): chunk s = [b[s], b[s+1]); sizes differ by at most one, empty chunks last."""
    bounds = [n_points * s // n_slots for s in range(n_slots + 1)]
    return [list(range(bounds[s], bounds[s + 1])) for s in range(n_slots)]


def choose_batch(n_points, n_handles, max_batch):
    """Batch size for a sweep of n_points k-points on n_handles handles: the smallest batch that keeps the number of
    rounds of the largest allowed batch, so that the last round is not mostly padding (31 points on 2 handles, at
    most 10 per batch: 2 rounds either way - a batch of 8 pads 1 solve, a batch of 10 pads 9)."""
    if n_points <= 0:
        return 1
    slots_max = max(1, n_handles * max_batch)
    rounds = -(-n_points // slots_max)
    return max(1, min(max_batch, -(-n_points // (n_handles * rounds))))


def batched_sweep(eqs, kappas, n_bands, batch, tol=1e-6, max_iter=2000, per_solve=None):
    """Solves every k-point of `kappas` exactly once on len(eqs) handles x `batch` slots per handle
    (bloch_set_kappa_batch).  The list is cut into len(eqs) * batch contiguous chunks; a handle walks its `batch`
    chunks in lock step, one batched Solve per round, so that consecutive k-points of a chunk warm-start each other
    (the reference's k-loop, maxwell_dispersion.cpp:475-531, taken len(eqs) * batch points at a time).  Handles run
    on their own host threads / streams.  A slot whose chunk is exhausted repeats its last k-point so that the batch
    size - and with it every captured graph and buffer - stays fixed; those repeats are counted in `wasted`.
    per_solve(eq): optional hook called before every batched solve (the bench's end-to-end leg re-uploads the
    coefficient field there).  Returns dict: lam[n, n_bands], iterations[n], converged[n], wasted, rounds."""
    import threading
    kappas = np.asarray(kappas, float).reshape(-1, 3)
    n = len(kappas)
    T = len(eqs)
    chunks = slot_chunks(n, T * batch)
    lam = np.zeros((n, n_bands))
    its = np.zeros(n, int)
    conv = np.zeros(n, int)
    info = {"wasted": 0, "rounds": 0}
    errors = []

    def worker(t):
        try:
            eq = eqs[t]
            eq.SetNumEigs(2 * n_bands)
            eq.SetAbsoluteTolerance(tol, max_iter)
            mine = [c for c in chunks[t * batch:(t + 1) * batch] if len(c) > 0]
            if not mine:
                return
            for r in range(max(len(c) for c in mine)):
                idx = [c[r] if r < len(c) else c[-1] for c in mine]
                if per_solve is not None:
                    per_solve(eq)
                lm, st = eq.SolveBatch(kappas[idx])
                for j, (c, i) in enumerate(zip(mine, idx)):
                    if r < len(c):
                        lam[i], its[i], conv[i] = lm[j], st[j]["iterations"], st[j]["converged_bands"]
                    else:
                        info["wasted"] += 1
                info["rounds"] = max(info["rounds"], r + 1)
        except BaseException as e:      # re-raised by the caller: a failed solve must fail the sweep
            errors.append(e)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errors:
        raise errors[0]
    return {"lam": lam, "iterations": its, "converged": conv, **info}


def dispersion_path(lattice, np_per_segment):
    """The k-points maxwell_dispersion visits, in output order (maxwell_dispersion.cpp:475-531, 596-648): for every
    path and segment of the lattice the points i = 0..np, kappa = ((np+1-i) kappa0 + i kappa1) / (np+1), plus the
    closing symmetry point of every path; a blank line separates paths in disp.dat.  Symmetry points are solved
    once and cached by label (:506, 604-614).  Returns (rows, unique_kappas): rows = list of
    (label, unique_index, path_index), unique_kappas[unique_index] = kappa to solve."""
    npt = int(np_per_segment)
    rows, uniq, by_label = [], [], {}

    def add(label, kappa, p, cache):
        if cache and label in by_label:
            u = by_label[label]
        else:
            u = len(uniq)
            uniq.append(np.array(kappa, float))
            if cache:
                by_label[label] = u
        rows.append((label, u, p))

    for p in range(lattice.GetNumberPaths()):
        ns = lattice.GetNumberPathSegments(p)
        for sgm in range(ns):
            e0, e1 = lattice.GetPathSegmentEndPointIndices(p, sgm)
            k0, k1 = lattice.GetSymmetryPoint(e0), lattice.GetSymmetryPoint(e1)
            for i in range(npt + 1):
                kap = ((npt + 1 - i) * k0 + i * k1) / (npt + 1)
                label = "-"
                if i == 0:
                    label = lattice.GetSymmetryPointLabel(e0)
                elif npt % 2 == 1 and i == (npt + 1) // 2:
                    label = lattice.GetIntermediatePointLabel(p, sgm)
                add(label, kap, p, cache=(i == 0))
            if sgm + 1 == ns:
                add(lattice.GetSymmetryPointLabel(e1), k1, p, cache=True)
    return rows, np.array(uniq)


def sharded_dispersion_sweep(eqs, lattice, np_per_segment, n_bands, batch, world=1, rank=0, tol=1e-6):
    """This rank's share of the maxwell_dispersion sweep: the UNIQUE k-points of
    dispersion_path() are split into `world` contiguous chunks (shard_kpoints), rank `rank` solves its chunk with
    batched_sweep on its handles - no collective on the solve path.  Returns (rows, unique_kappas, lo, result) with
    result = batched_sweep's dict for the k-points [lo, lo + len(result["lam"])); the caller gathers the per-rank
    blocks (n_bands doubles per k-point) and writes disp.dat with write_dispersion_data."""
    rows, uk = dispersion_path(lattice, np_per_segment)
    lo, hi = shard_kpoints(len(uk), world, rank)
    res = batched_sweep(eqs, uk[lo:hi], n_bands, batch, tol) if hi > lo else \
        {"lam": np.zeros((0, n_bands)), "iterations": np.zeros(0, int), "converged": np.zeros(0, int), "wasted": 0, "rounds": 0}
    return rows, uk, lo, res


def write_dispersion_data(path, rows, lam_unique):
    """disp.dat exactly as WriteDispersionData prints it (maxwell_dispersion.cpp:1062-1087): counter, label, then
    omega = sqrt(lambda) (0 for -1e-6 < lambda <= 0, -1 otherwise) of every REAL mode - each complex band twice,
    like the reference's block form - tab separated, a blank line after every path."""
    with open(path, "w") as f:
        last_p = None
        for c, (label, u, p) in enumerate(rows):
            if last_p is not None and p != last_p:
                f.write("\n")
            last_p = p
            om = np.repeat(omega_of_lambda(lam_unique[u]), 2)
            f.write("%d\t%s" % (c, label) + "".join("\t%.10g" % v for v in om) + "\n")
        f.write("\n")


def write_hypre_ij(path, mat, rank=0):
    """Writes a real scipy sparse matrix the way HypreParMatrix::Print does (hypre_ParCSRMatrixPrintIJ:
    file `<path>.<rank as %05d>`, header `ilower iupper jlower jupper`, then `i j value` per entry, %.14e)."""
    m = mat.tocsr()
    m.sort_indices()
    fn = "%s.%05d" % (path, rank)
    with open(fn, "w") as f:
        f.write("%d %d %d %d\n" % (0, m.shape[0] - 1, 0, m.shape[1] - 1))
        for i in range(m.shape[0]):
            for q in range(m.indptr[i], m.indptr[i + 1]):
                f.write("%d %d %.14e\n" % (i, m.indices[q], m.data[q]))
    return fn


def read_hypre_ij(fn):
    import scipy.sparse as sp
    with open(fn) as f:
        i0, i1, j0, j1 = (int(t) for t in f.readline().split())
        rows, cols, vals = [], [], []
        for line in f:
            a, b, c = line.split()
            rows.append(int(a)); cols.append(int(b)); vals.append(float(c))
    return sp.csr_matrix((vals, (rows, cols)), shape=(i1 - i0 + 1, j1 - j0 + 1))


def write_matrices(eq, prefix, label=""):
    """The -wm dump of maxwell_dispersion.cpp:553-590: Ar<label>.mat, Ai<label>.mat, M<label>.mat in
    hypre IJ text format (Ai = Im(A), i.e. the reference's (1,0) block times its block coefficient)."""
    A, M = eq.AssembleMatrix("A"), eq.AssembleMatrix("M")
    return [write_hypre_ij("%s/Ar%s.mat" % (prefix, label), A.real),
            write_hypre_ij("%s/Ai%s.mat" % (prefix, label), A.imag),
            write_hypre_ij("%s/M%s.mat" % (prefix, label), M)]


# ---- plane-wave initial vectors (CreateInitialVectors, maxwell/maxwell_dispersion.cpp:735-1060) ----
_MODE_TABLES = {      # integer shifts n of the reciprocal lattice tried by the reference, per lattice (:806-856)
    "default": [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)],
    "FCC": [(0, 0, 0)] + [(a, b, c) for a in (1, -1) for b in (1, -1) for c in (1, -1)],
    "BCC": [(0, 0, 0)] + [(0, b, c) for b in (1, -1) for c in (1, -1)] + [(a, 0, c) for a in (1, -1) for c in (1, -1)]
           + [(a, b, 0) for a in (1, -1) for b in (1, -1)],
}


def _gauss_legendre01(p):
    x, _ = np.polynomial.legendre.leggauss(p)
    return 0.5 * (x + 1.0)


def _gauss_lobatto01(n):
    if n == 2:
        return np.array([0.0, 1.0])
    inner = np.polynomial.legendre.Legendre.basis(n - 1).deriv().roots()
    return 0.5 * (np.concatenate([[-1.0], np.sort(inner.real), [1.0]]) + 1.0)


def nd_interpolate(eq, field):
    """Nodal interpolation into the Nedelec space of `eq` (what ParGridFunction::ProjectCoefficient does for a smooth
    vector field): dof k = t_k . J^T v(x_k) = (edge vector J e_c) . v(x_k) at the ND node x_k (open direction c on the
    Gauss-Legendre points, closed directions on the Gauss-Lobatto points).  `field(x[n,3]) -> complex v[n,3]`.
    Returns the vector in the boundary layout [re(N); im(N)]."""
    p = eq.order
    g, l = _gauss_legendre01(p), _gauss_lobatto01(p + 1)
    x0, cls, J = eq.element_geometry()
    gid, sign = eq.dofmap("nd")

    def grid(ax, ay, az):
        kk, jj, ii = np.meshgrid(az, ay, ax, indexing="ij")          # i fastest
        return np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)

    nodes = [grid(g, l, l), grid(l, g, l), grid(l, l, g)]
    out = np.zeros(eq.N, complex)
    nb = p * (p + 1) ** 2
    for c in range(3):
        X = x0[:, None, :] + np.einsum("eij,kj->eki", J[cls], nodes[c])      # [ne, nb, 3]
        v = field(X.reshape(-1, 3)).reshape(len(x0), nb, 3)
        t = J[cls][:, :, c]                                                    # physical edge vector of direction c
        loc = np.einsum("ei,eki->ek", t, v)
        sl = slice(c * nb, (c + 1) * nb)
        out[gid[:, sl].ravel()] = (sign[:, sl] * loc).ravel()                 # copies agree (tangential continuity)
    return np.concatenate([out.real, out.imag])


def _lagrange(nodes, x):
    """values of the Lagrange basis on `nodes` at the points x: [len(x), len(nodes)]"""
    nodes, x = np.asarray(nodes, float), np.asarray(x, float)
    out = np.ones((len(x), len(nodes)))
    for j in range(len(nodes)):
        for k in range(len(nodes)):
            if k != j:
                out[:, j] *= (x - nodes[k]) / (nodes[j] - nodes[k])
    return out


def evaluate_fields(eq, e_reim=None, b_reim=None, ref_points=None):
    """Point values of a Nedelec field E (dofs in the boundary layout [re(N); im(N)]) and / or a Raviart-Thomas
    field B ([re(N_rt); im(N_rt)]) at the reference points `ref_points` [npts, 3] of EVERY element (default: the 8
    hex corners in MFEM vertex order) - what a GridFunction evaluation does: covariant Piola map J^-T for ND,
    contravariant J / det for RT, MFEM's tensor bases (open = Lagrange on Gauss-Legendre, closed = Lagrange on
    Gauss-Lobatto points).  Returns (x[ne, npts, 3], E[ne, npts, 3] complex or None, B[...] complex or None)."""
    p = eq.order
    g, l = _gauss_legendre01(p), _gauss_lobatto01(p + 1)
    pts = _HEX_REF if ref_points is None else np.asarray(ref_points, float)
    x0, cls, J = eq.element_geometry()
    Je = J[cls]
    X = x0[:, None, :] + np.einsum("eij,kj->eki", Je, pts)
    O = [_lagrange(g, pts[:, d]) for d in range(3)]        # open basis per direction  [npts, p]
    Cc = [_lagrange(l, pts[:, d]) for d in range(3)]       # closed basis per direction [npts, p + 1]

    def shapes(kind):
        """[3 components][npts, n_loc_per_component], natural local order (i fastest)"""
        out = []
        for c in range(3):
            f = [(O[d] if (d == c) == (kind == "nd") else Cc[d]) for d in range(3)]
            out.append(np.einsum("qi,qj,qk->qkji", f[0], f[1], f[2]).reshape(len(pts), -1))
        return out

    E = B = None
    if e_reim is not None:
        e_reim = np.asarray(e_reim, float)
        ec = e_reim[:eq.N] + 1j * e_reim[eq.N:]
        gid, sign = eq.dofmap("nd")
        loc = ec[gid] * sign                                # [ne, L_nd]
        nb = p * (p + 1) ** 2
        ref = np.stack([loc[:, c * nb:(c + 1) * nb] @ sh.T for c, sh in enumerate(shapes("nd"))], axis=-1)   # [ne, npts, 3]
        Jinv = np.linalg.inv(Je)
        E = np.einsum("eji,eqj->eqi", Jinv, ref)            # J^-T v
    if b_reim is not None:
        b_reim = np.asarray(b_reim, float)
        bc = b_reim[:eq.N_rt] + 1j * b_reim[eq.N_rt:]
        gid, sign = eq.dofmap("rt")
        loc = bc[gid] * sign
        nb = p * p * (p + 1)
        ref = np.stack([loc[:, c * nb:(c + 1) * nb] @ sh.T for c, sh in enumerate(shapes("rt"))], axis=-1)
        det = np.linalg.det(Je)
        B = np.einsum("eij,eqj->eqi", Je, ref) / det[:, None, None]
    return X, E, B


def write_vtk_fields(eq, path, fields, cell_data=None):
    """Legacy-VTK unstructured grid of the refined Wigner-Seitz cell with point data on per-element corner copies
    (fields may be discontinuous across faces in their normal / tangential parts, like the FE fields themselves):
    `fields` = {name: real array [ne, 8, 3]}, cell_data = {name: array [ne]}.  Stand-in for the reference's VisIt
    data collection (maxwell_bloch.cpp:1730-1822), readable by VisIt / ParaView."""
    x0, cls, J = eq.element_geometry()
    X = x0[:, None, :] + np.einsum("eij,kj->eki", J[cls], _HEX_REF)
    ne = len(x0)
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 3.0\nmfem-bravais_b200 Bloch fields\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        f.write("POINTS %d double\n" % (8 * ne))
        np.savetxt(f, X.reshape(-1, 3), fmt="%.12g")
        f.write("CELLS %d %d\n" % (ne, 9 * ne))
        np.savetxt(f, np.column_stack([np.full(ne, 8), np.arange(8 * ne).reshape(ne, 8)]), fmt="%d")
        f.write("CELL_TYPES %d\n" % ne)
        np.savetxt(f, np.full(ne, 12), fmt="%d")
        if cell_data:
            f.write("CELL_DATA %d\n" % ne)
            for name, v in cell_data.items():
                f.write("SCALARS %s double 1\nLOOKUP_TABLE default\n" % name)
                np.savetxt(f, np.asarray(v, float), fmt="%.12g")
        f.write("POINT_DATA %d\n" % (8 * ne))
        for name, v in fields.items():
            f.write("VECTORS %s double\n" % name)
            np.savetxt(f, np.asarray(v, float).reshape(-1, 3), fmt="%.12g")


def read_vtk_fields(path):
    """minimal reader of write_vtk_fields' output (tests): returns (points[n,3], {name: array[n,3]})"""
    lines = open(path).read().split("\n")
    i, pts, fields = 0, None, {}
    while i < len(lines):
        tok = lines[i].split()
        if tok and tok[0] == "POINTS":
            n = int(tok[1])
            pts = np.array([[float(v) for v in lines[i + 1 + k].split()] for k in range(n)])
            i += n
        elif tok and tok[0] == "VECTORS":
            fields[tok[1]] = np.array([[float(v) for v in lines[i + 1 + k].split()] for k in range(len(pts))])
            i += len(pts)
        i += 1
    return pts, fields


def plane_wave_initial_vectors(eq, lattice, kappa, count=None, literal=True):
    """The reference's initial block (CreateInitialVectors): for every shift n of the lattice's mode table the
    plane waves E0 exp(i 2 pi sum_j n_j b_j . x) - the periodic envelopes of exp(i (kappa + 2 pi G) . x) - with E0 the
    two unit vectors orthogonal to k (three Cartesian ones when |k| < 1e-2), nodally interpolated into ND.  The
    reference builds the real pairs (E, iE) (:1039-1057); in this complex-native solver iE is not an independent
    vector, so one complex vector per (n, E0) is returned, ordered by |kappa + 2 pi G| (lowest empty-lattice
    frequencies first), at most `count`.  literal=True keeps the reference's k = kappa + sum_j n_j b_j (its reciprocal
    vectors carry no 2 pi, `sic` in SURVEY.md section 8(a) row a16); literal=False uses the physical wave vector."""
    kappa = np.asarray(kappa, float)
    b = lattice.GetReciprocalLatticeVectors()
    table = _MODE_TABLES.get(lattice.GetLatticeTypeLabel(), _MODE_TABLES["default"])
    waves = []
    for n in table:
        G = np.asarray(n, float) @ b
        k = kappa + (G if literal else 2.0 * np.pi * G)
        if np.linalg.norm(k) < 1e-2:
            pol = list(np.eye(3))
        else:
            w, V = np.linalg.eigh(np.eye(3) - np.outer(k, k) / (k @ k))       # eigenvalues 0, 1, 1
            pol = [V[:, 1], V[:, 2]]
        for e0 in pol:
            waves.append((np.linalg.norm(kappa + 2.0 * np.pi * G), G, e0))
    waves.sort(key=lambda t: t[0])
    if count is not None:
        waves = waves[:count]
    vecs = [nd_interpolate(eq, lambda X, G=G, e0=e0: np.exp(2j * np.pi * (X @ G))[:, None] * e0[None, :])
            for _, G, e0 in waves]
    return np.array(vecs)


_HEX_REF = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
_HEX_FACES = [(3, 2, 1, 0), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7)]


def write_mfem_mesh(eq, path, lattice=None, eps=None, muinv=None):
    """Interchange: the refined Wigner-Seitz cell of `eq` as a NON-periodic MFEM mesh file
    (`MFEM mesh v1.0`, hexahedra in MFEM vertex order, boundary = faces of a single element).  Together with the
    lattice's translation vectors (written to `<path>.trans`, one per line) this is exactly the input of the
    reference's pipeline - Wigner-Seitz mesh, refine, `MakePeriodicMesh(mesh, trans_vecs)` (lib/bravais.cpp:343-355,
    9548-9832; in current MFEM: `Mesh::MakePeriodic(mesh, mesh.CreatePeriodicVertexMapping(trans))`) - so a real
    MFEM/hypre run elsewhere works on the very same elements, in the same order.  Element attributes enumerate the
    distinct (eps, 1/mu) pairs (element-wise constants), listed in `<path>.coef` as `attribute eps muinv`, so that
    `PWConstantCoefficient`s reproduce the coefficients exactly.  Returns (n_vertices, n_elements, n_boundary)."""
    x0, cls, J = eq.element_geometry()
    ne = len(x0)
    eps = np.ones(ne) if eps is None else np.asarray(eps, float)
    muinv = np.ones(ne) if muinv is None else np.asarray(muinv, float)
    pairs = {}
    attr = np.zeros(ne, dtype=int)
    for e in range(ne):
        attr[e] = pairs.setdefault((float(eps[e]), float(muinv[e])), len(pairs) + 1)
    vid, verts, hexes = {}, [], np.zeros((ne, 8), dtype=int)
    for e in range(ne):
        P = x0[e] + _HEX_REF @ J[cls[e]].T
        for l in range(8):
            key = tuple(np.round(P[l], 10) + 0.0)
            if key not in vid:
                vid[key] = len(verts)
                verts.append(P[l])
            hexes[e, l] = vid[key]
    faces = {}
    for e in range(ne):
        for f in _HEX_FACES:
            fv = tuple(int(hexes[e, k]) for k in f)
            faces.setdefault(tuple(sorted(fv)), []).append(fv)
    bdr = [v[0] for v in faces.values() if len(v) == 1]
    with open(path, "w") as f:
        f.write("MFEM mesh v1.0\n\n#\n# Wigner-Seitz cell written by mfem_bravais_b200 (non-periodic; translation vectors in %s.trans)\n"
                "# MFEM geometry types: SQUARE = 3, CUBE = 5\n#\n\ndimension\n3\n\nelements\n%d\n" % (os.path.basename(path), ne))
        for e in range(ne):
            f.write("%d 5 %s\n" % (attr[e], " ".join(str(int(v)) for v in hexes[e])))
        f.write("\nboundary\n%d\n" % len(bdr))
        for fv in bdr:
            f.write("1 3 %d %d %d %d\n" % fv)
        f.write("\nvertices\n%d\n3\n" % len(verts))
        for v in verts:
            f.write("%.17g %.17g %.17g\n" % tuple(v + 0.0))
    with open(path + ".coef", "w") as f:
        for (e_, m_), a_ in sorted(pairs.items(), key=lambda kv: kv[1]):
            f.write("%d %.17g %.17g\n" % (a_, e_, m_))
    if lattice is not None:
        with open(path + ".trans", "w") as f:
            for t in lattice.GetTranslationVectors():
                f.write("%.17g %.17g %.17g\n" % tuple(t))
    return len(verts), ne, len(bdr)


def read_mfem_mesh(path):
    """Minimal reader of the files written by write_mfem_mesh (tests): (vertices[nv,3], hexes[ne,8], attr[ne], bdr[nb,4])."""
    toks = [ln.strip() for ln in open(path) if ln.strip() and not ln.startswith("#")]
    assert toks[0] == "MFEM mesh v1.0" and toks[toks.index("dimension") + 1] == "3"
    i = toks.index("elements")
    ne = int(toks[i + 1])
    el = np.array([[int(t) for t in toks[i + 2 + e].split()] for e in range(ne)])
    assert (el[:, 1] == 5).all()
    i = toks.index("boundary")
    nb = int(toks[i + 1])
    bd = np.array([[int(t) for t in toks[i + 2 + b].split()] for b in range(nb)])
    i = toks.index("vertices")
    nv = int(toks[i + 1])
    assert toks[i + 2] == "3"
    V = np.array([[float(t) for t in toks[i + 3 + v].split()] for v in range(nv)])
    return V, el[:, 2:], el[:, 0], bd[:, 2:]


def homogenization_sweep(eq, kappa0, num_beta, n_bands, a=1.0, num_a_per_lambda=10.0, tol=1e-6):
    """Small-kappa sweep of maxwell_homogenization.cpp:562-593: for i = 1 .. num_beta-1,
    kappa = kappa0 * 2 pi (i / (num_beta-1)) / (a * num_a_per_lambda); solve, then the field averages
    of every band (GetFieldAverages).  Returns a list of dicts {kappa, lambda, averages[band]} - the
    inputs of the reference's effective-medium fit (CalcCoefs, not part of this path)."""
    eq.SetNumEigs(2 * n_bands)
    eq.SetAbsoluteTolerance(tol)
    out = []
    for i in range(1, num_beta):
        frac = i / (num_beta - 1) if num_beta > 1 else 1.0
        kappa = np.asarray(kappa0, float) * (2.0 * np.pi * frac / (a * num_a_per_lambda))
        eq.SetKappa(kappa)
        eq.Setup()
        eq.Solve()
        out.append({"kappa": kappa, "lambda": eq.band_eigenvalues(),
                    "averages": [eq.GetFieldAverages(b) for b in range(n_bands)]})
    return out


class MaxwellBlochWaveSolver:
    """Multilevel warm start: MaxwellBlochWaveSolver::GetEigenfrequencies
    (meta-material/meta_material_solver.cpp:2731-2881).  Level l is the cell with n0 * 2**l
    subdivisions; the coarse level is solved first, its eigenvectors are interpolated to the next
    level and used as initial vectors there, until sum_i |lambda_i(fine) - lambda_i(coarse)| <= tol
    or max_lvl levels were used.  `eps_fn(centers) -> eps per element` re-samples the coefficient on
    every level, like the reference's Coefficient objects."""

    def __init__(self, lattice, n0, order, n_bands, eps_fn=None, muinv_fn=None, max_lvl=3, tol=1e-3,
                 solver_tol=1e-6, equation_cls=None):
        from .equation import MaxwellBlochWaveEquation
        self._cls = equation_cls or MaxwellBlochWaveEquation
        self.lat, self.n0, self.order, self.nb = lattice, int(n0), int(order), int(n_bands)
        self.eps_fn, self.muinv_fn = eps_fn, muinv_fn
        self.max_lvl, self.tol, self.solver_tol = int(max_lvl), float(tol), float(solver_tol)
        self.kappa = np.zeros(3)
        self.levels = []
        self.level_eigs, self.level_iters = [], []

    def _level(self, lvl):
        while len(self.levels) <= lvl:
            eq = self._cls(self.lat, self.n0 * 2 ** len(self.levels), self.order)
            if self.eps_fn is not None:
                eq.SetMassCoef(self.eps_fn(eq.element_centers()))
            if self.muinv_fn is not None:
                eq.SetStiffnessCoef(self.muinv_fn(eq.element_centers()))
            eq.SetNumEigs(2 * self.nb)
            eq.SetAbsoluteTolerance(self.solver_tol)
            self.levels.append(eq)
        return self.levels[lvl]

    def SetKappa(self, kappa):
        self.kappa = np.asarray(kappa, float)

    def GetEigenfrequencies(self):
        eq = self._level(0)
        eq.SetKappa(self.kappa)
        eq.Setup()
        eq.Solve()
        fine = eq.band_eigenvalues()
        self.level_eigs, self.level_iters = [fine], [eq.GetSolverStats()["iterations"]]
        lvl, err = 1, 2.0 * self.tol
        while lvl < self.max_lvl and err > self.tol:
            coarse = fine
            nxt = self._level(lvl)
            nxt.SetKappa(self.kappa)
            nxt.Setup()
            self.levels[lvl - 1].ProlongEigenvectorsTo(nxt)
            nxt.Solve()
            fine = nxt.band_eigenvalues()
            self.level_eigs.append(fine)
            self.level_iters.append(nxt.GetSolverStats()["iterations"])
            err = float(np.abs(fine - coarse).sum())
            lvl += 1
        self.fine_level = lvl - 1
        return np.sqrt(np.abs(fine))

    def ReturnFineEigenvector(self, i):
        return self.levels[self.fine_level].GetEigenvectorE(i)


class MaxwellDispersion:
    """Reduced-basis band-structure sweep: the reference's MaxwellDispersion
    (meta-material/meta_material_solver.cpp:3132-3410).

    buildRawBasis(): full eigen-solves at every symmetry point of the lattice's paths (and at the
    labelled intermediate points when mid_pts), eigenvectors kept on the device as the raw basis.
    approxEigenfrequencies(kappa): Rayleigh-Ritz in the span of that basis projected with the
    kappa's divergence projector; omega = sqrt|lambda|.
    traverseBrillouinZone(): 2**samp_pow intervals per path segment; end (and mid) points take
    the full-solve values, the rest the approximation.  Result: seg_eigs[p][s][i] = omega array."""

    def __init__(self, eq, lattice, n_bands, samp_pow=2, mid_pts=True, tol=1e-6):
        self.eq, self.lat, self.nb = eq, lattice, int(n_bands)
        self.samp_pow, self.mid_pts = int(samp_pow), bool(mid_pts)
        self.sp_eigs = {}
        self.seg_eigs = []
        eq.SetNumEigs(2 * self.nb)
        eq.SetAbsoluteTolerance(tol)

    def _full(self, label, kappa):
        if label in self.sp_eigs:
            return
        self.eq.SetKappa(kappa)
        self.eq.Setup()
        self.eq.Solve()
        self.sp_eigs[label] = np.sqrt(np.abs(self.eq.band_eigenvalues()))
        self.eq.ReducedBasisAppend()

    def buildRawBasis(self):
        lat = self.lat
        self.eq.ReducedBasisClear()
        self.sp_eigs = {}
        for p in range(lat.GetNumberPaths()):
            for s in range(lat.GetNumberPathSegments(p)):
                e0, e1 = lat.GetPathSegmentEndPointIndices(p, s)
                self._full(lat.GetSymmetryPointLabel(e0), lat.GetSymmetryPoint(e0))
                if self.mid_pts:
                    self._full(lat.GetIntermediatePointLabel(p, s), lat.GetIntermediatePoint(p, s))
                self._full(lat.GetSymmetryPointLabel(e1), lat.GetSymmetryPoint(e1))
        return self.eq.ReducedBasisSize()

    def approxEigenfrequencies(self, kappa):
        return np.sqrt(np.abs(self.eq.ApproxEigenvalues(kappa, self.nb)))

    def traverseBrillouinZone(self):
        lat = self.lat
        self.buildRawBasis()
        ni = 2 ** self.samp_pow
        self.seg_eigs = []
        for p in range(lat.GetNumberPaths()):
            segs = []
            for s in range(lat.GetNumberPathSegments(p)):
                e0, e1 = lat.GetPathSegmentEndPointIndices(p, s)
                k0, k1 = lat.GetSymmetryPoint(e0), lat.GetSymmetryPoint(e1)
                row = [None] * (ni + 1)
                row[0] = self.sp_eigs[lat.GetSymmetryPointLabel(e0)]
                row[ni] = self.sp_eigs[lat.GetSymmetryPointLabel(e1)]
                for i in range(1, ni):
                    if ni == 2 * i and self.mid_pts:
                        row[i] = self.sp_eigs[lat.GetIntermediatePointLabel(p, s)]
                    else:
                        # (on segments touching Gamma the reference interpolates beta along a fixed
                        # zeta, :3355-3374 - the same points as this linear interpolation)
                        row[i] = self.approxEigenfrequencies(((ni - i) * k0 + i * k1) / ni)
                segs.append(row)
            self.seg_eigs.append(segs)
        return self.seg_eigs
