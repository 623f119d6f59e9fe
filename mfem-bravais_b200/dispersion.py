"""maxwell_dispersion-style sweep logic (maxwell/maxwell_dispersion.cpp:475-648, 1062-1087,
1449-1556): k-path walk with symmetry-point cache, piecewise-constant coefficients sampled at
element centres, omega = sqrt(lambda) formatting."""
import numpy as np


def sphere_eps(centers, radius=0.25, eps_in=10.0, eps_out=1.0):
    """mass_coef case 2 (maxwell_dispersion.cpp:1464-1467) at element centres."""
    r = np.linalg.norm(centers, axis=1)
    return np.where(r <= radius, eps_in, eps_out)


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
