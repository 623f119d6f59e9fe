"""TEST INFRASTRUCTURE ONLY: CPU baseline / reference arm of bench.py ("port" of the reference algorithm).

What the reference does per k-point (maxwell/maxwell_bloch.cpp:337-620, 809-825): assemble the
zeta-dependent interpolation matrices, form S1/DKZ (4 sparse triple products) and the projector's
A0/DKZ0 (4 more), then run a projected, preconditioned block LOBPCG whose every operator
application is a CSR SpMV.  This file times exactly that on the host cores with the oracle's
assembled matrices: FULL, CONVERGED solves (no extrapolation).

Stated differences from the reference (MFEM/hypre cannot be built here):
  * hypre's AMS is replaced by the degree-24 Chebyshev-Jacobi polynomial in D^-1 (A + sigma M) that the
    CUDA path uses as well (eigenvalues do not depend on the preconditioner);
  * the projector's inner solve is a block Jacobi-PCG to 1e-2 * tol (the reference: MINRES to 1e-13 per
    vector, which is slower); the GPU path preconditions the same solve with a multigrid V-cycle;
  * the complex Hermitian form is iterated (m complex vectors instead of the reference's 2m real ones);
  * consecutive k-points are warm-started from the previous eigenvectors like the GPU sweep (the reference
    starts every k-point cold, maxwell_bloch.cpp:553).
All heavy loops are threaded: CSR x block products, the fused Chebyshev update and the whole PCG loop in C with
OpenMP (oracle/csr_spmm.c), dense Gram / rotation products in OpenBLAS.  The thread count is set explicitly
(torchrun exports OMP_NUM_THREADS=1) and reported as `cores`.
"""
import ctypes as C
import os
import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from .bloch_oracle import BlochOperators, Lattice, Mesh, Spaces

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_threads = None
_blas_limit = None


def _spmm_lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "_ref", "libcsr_spmm.so")
        if os.path.exists(p):
            _lib = C.CDLL(p)
            _lib.csr_num_threads.restype = C.c_int
            _lib.pcg_jacobi_z.restype = C.c_int
        else:
            _lib = False
    return _lib


def set_threads(n=None):
    """use n host threads (default: all cores) in the C kernels and in OpenBLAS, whatever OMP_NUM_THREADS says"""
    global _threads, _blas_limit
    n = int(n or os.cpu_count() or 1)
    L = _spmm_lib()
    if L:
        L.csr_set_threads(C.c_int(n))
    try:
        import threadpoolctl
        _blas_limit = threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    _threads = n
    return n


def threads():
    if _threads is None:
        set_threads()
    L = _spmm_lib()
    return L.csr_num_threads() if L else 1


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Csr:
    def __init__(self, A):
        A = A.tocsr().astype(complex)
        A.sort_indices()
        self.A = A
        self.n = A.shape[0]
        self.indptr = np.ascontiguousarray(A.indptr, np.int32)
        self.indices = np.ascontiguousarray(A.indices, np.int32)
        self.data = np.ascontiguousarray(A.data, complex)

    def __call__(self, X, out=None):
        L = _spmm_lib()
        X = np.ascontiguousarray(X, complex)
        if not L or X.shape[1] > 64:
            return self.A @ X
        Y = np.empty((self.n, X.shape[1]), complex) if out is None else out
        L.csr_spmm_z(C.c_int64(self.n), C.c_int(X.shape[1]), _p(self.indptr), _p(self.indices), _p(self.data), _p(X), _p(Y))
        return Y


def lobpcg_cpu(ops, nb, tol, max_iter=2000, sigma_scale=1.0, cheb_degree=24, cheb_ratio=300.0,
               proj_tol=None, seed=1, verbose=False, timing=None, X0=None, precond_factory=None):
    """Projected, preconditioned complex block LOBPCG on the assembled operators of `ops` (one kappa).
    X0: starting block (warm start from the previous k-point) or None (seeded random).
    Returns (eigenvalues[nb], iterations, counts, X) with X the final block (for the next warm start)."""
    L = _spmm_lib()
    mesh = ops.sp_.mesh
    Ac, Mc = ops.A_c(), ops.M_c()
    sigma = sigma_scale / mesh.volume ** (2.0 / 3.0)
    Ash = (Ac + sigma * Mc).tocsr()
    A, M, G, SH = Csr(Ac), Csr(Mc), Csr(ops.G_c()), Csr(Ash)
    GH = Csr(ops.G_c().conj().T)
    S0 = Csr(ops.S0_c())
    N = A.n
    mb = min(21, nb + max(6, nb // 4))
    if proj_tol is None:
        proj_tol = 1e-2 * tol if tol > 0 else 1e-8
    dA = Ash.diagonal().real
    jac = np.ascontiguousarray(1.0 / dA)
    jac0 = np.ascontiguousarray(1.0 / ops.S0_c().diagonal().real)
    counts = {"A": 0, "S0": 0}

    # rigorous bound of lambda_max(D^-1 (A + sigma M)): Gershgorin row sums of the scaled matrix
    # (the GPU path uses the element-local spectra for the same purpose)
    dsq = 1.0 / np.sqrt(dA)
    Asc = sp.diags(dsq) @ Ash @ sp.diags(dsq)
    lmax = float(abs(Asc).sum(axis=1).max())
    lmin = lmax / cheb_ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    s1 = theta / delta

    def precond(R):
        r = np.ascontiguousarray(R, complex).copy()
        d = np.empty_like(r)
        x = np.empty_like(r)
        q = np.empty_like(r)
        m = r.shape[1]
        if L:
            L.cheb_first_z(C.c_int64(N), C.c_int(m), _p(jac), _p(r), _p(d), _p(x), C.c_double(1.0 / theta))
        else:
            d[:] = jac[:, None] * r / theta
            x[:] = d
        rho = 1.0 / s1
        for _ in range(1, cheb_degree):
            SH(d, out=q) if L else q.__setitem__(slice(None), SH(d))
            counts["A"] += m
            rho_n = 1.0 / (2.0 * s1 - rho)
            if L:
                L.cheb_step_z(C.c_int64(N), C.c_int(m), _p(jac), _p(q), _p(r), _p(d), _p(x),
                              C.c_double(rho_n * rho), C.c_double(2.0 * rho_n / delta))
            else:
                r -= q
                d[:] = rho_n * rho * d + (2.0 * rho_n / delta) * (jac[:, None] * r)
                x += d
            rho = rho_n
        return x

    if precond_factory is not None:     # experiments with other preconditioners (scratch/aux_proto.py)
        precond = precond_factory(dict(ops=ops, Ash=Ash, sigma=sigma, jac=jac, cheb=precond, counts=counts))

    def project(X, rel):
        rhs = np.ascontiguousarray(GH(M(X)))
        if ops.beta == 0.0:
            rhs -= rhs.mean(axis=0)          # S0 singular on constants at Gamma
        if not np.any(rhs):
            return X
        m = rhs.shape[1]
        if L:
            phi = np.empty_like(rhs)
            its = L.pcg_jacobi_z(C.c_int64(S0.n), C.c_int(m), _p(S0.indptr), _p(S0.indices), _p(S0.data), _p(jac0),
                                 _p(rhs), _p(phi), C.c_double(rel), C.c_int(3000))
            counts["S0"] += its * m
        else:
            rr0 = np.sum(np.abs(rhs) ** 2, axis=0)
            phi = np.zeros_like(rhs)
            r = rhs
            z = jac0[:, None] * r
            p = z.copy()
            rz = np.sum((r.conj() * z).real, axis=0)
            for it in range(3000):
                q = S0(p)
                counts["S0"] += m
                pq = np.sum((p.conj() * q).real, axis=0)
                alpha = np.where(pq != 0, rz / np.where(pq != 0, pq, 1), 0)
                phi += alpha * p
                r -= alpha * q
                z = jac0[:, None] * r
                rzn = np.sum((r.conj() * z).real, axis=0)
                if np.all(np.sum(np.abs(r) ** 2, axis=0) <= rel * rel * rr0):
                    break
                beta = np.where(rz != 0, rzn / np.where(rz != 0, rz, 1), 0)
                p = z + beta * p
                rz = rzn
        return X - G(phi)

    def rr(S, AS, MS):
        GA = S.conj().T @ AS
        GM = S.conj().T @ MS
        GA = 0.5 * (GA + GA.conj().T)
        GM = 0.5 * (GM + GM.conj().T)
        w, Cv = sla.eigh(GA, GM)
        return w[:mb], Cv[:, :mb]

    if X0 is None or X0.shape != (N, mb):
        rng = np.random.default_rng(seed)
        X0 = rng.uniform(-1, 1, (N, mb)) + 1j * rng.uniform(-1, 1, (N, mb))
    X = project(np.ascontiguousarray(X0, complex), min(proj_tol, 1e-10))
    AX, MX = A(X), M(X)
    lam, Cv = rr(X, AX, MX)
    X, AX, MX = X @ Cv, AX @ Cv, MX @ Cv
    P = AP = MP = None
    it = 0
    t_loop = time.time()
    if timing is not None:
        timing["t_init"] = t_loop - timing["t_start"]
    converged = False
    for it in range(max_iter):
        R = AX - MX * lam
        rn = np.linalg.norm(R[:, :nb], axis=0)
        if verbose:
            print("[cpu lobpcg] it %d maxres %.3e" % (it, rn.max()))
        if rn.max() <= tol:
            converged = True
            break
        W = project(precond(R), proj_tol)
        AW, MW = A(W), M(W)
        counts["A"] += mb
        blocks = [X, W] + ([P] if P is not None else [])
        Ab = [AX, AW] + ([AP] if P is not None else [])
        Mb = [MX, MW] + ([MP] if P is not None else [])
        S, AS, MS = np.hstack(blocks), np.hstack(Ab), np.hstack(Mb)
        try:
            lam, Cv = rr(S, AS, MS)
        except np.linalg.LinAlgError:
            S, AS, MS = np.hstack(blocks[:2]), np.hstack(Ab[:2]), np.hstack(Mb[:2])
            lam, Cv = rr(S, AS, MS)
        Cp = Cv.copy()
        Cp[:mb] = 0
        P, AP, MP = S @ Cp, AS @ Cp, MS @ Cp
        X, AX, MX = S @ Cv, AS @ Cv, MS @ Cv
    if timing is not None:
        timing["t_iters"] = time.time() - t_loop
        timing["converged"] = converged
    return lam[:nb], it, counts, X


def time_kpoints(lattice, nsub, order, labels, pts, bands, tol, steps, first=0, warmup=0, max_iter=2000, nthreads=None):
    """Solves warmup + steps consecutive k-points of the path (starting at index `first`, warm-starting every
    k-point from the previous one, the very first from a seeded random block) and times the last `steps`:
    per-k sparse products (the reference's Setup) + full converged LOBPCG solves."""
    ncores = set_threads(nthreads)
    lat = Lattice(lattice)
    mesh = Mesh(lat, nsub)
    t0 = time.time()
    spc = Spaces(mesh, order)
    ops = BlochOperators(spc, mesh.sphere_eps())
    t_once = time.time() - t0
    ks = lat.kpath(labels, pts)
    t_setup = t_solve = 0.0
    its, lams, conv = [], [], []
    X = None
    for s in range(warmup + steps):
        k = ks[(first + s) % len(ks)]
        t0 = time.time()
        ops.set_kappa(k)          # per-k assembly + sparse triple products (reference: Setup())
        t1 = time.time()
        timing = {"t_start": t1}
        lam, it, cnt, X = lobpcg_cpu(ops, bands, tol, max_iter=max_iter, timing=timing, X0=X)
        t2 = time.time()
        if s >= warmup:
            t_setup += t1 - t0
            t_solve += t2 - t1
            its.append(int(it))
            lams.append(np.asarray(lam).tolist())
            conv.append(bool(timing.get("converged", False)))
    total = t_setup + t_solve
    L = _spmm_lib()
    return {"value": steps / total, "unit": "k-points/s", "cores": int(L.csr_num_threads()) if L else 1, "kind": "port",
            "sample": "%d consecutive k-point(s) (path index %d on, %d untimed before them) of %s order %d n_sub=%d "
                      "(N=%d), %d bands, tol %g: full converged solves, warm-started from the previous k-point "
                      "(the first one of the run from a random block); per-k sparse products %.1f s + projected "
                      "LOBPCG %.1f s in total (one-time assembly %.1f s not counted)"
                      % (steps, first + warmup, warmup, lattice, order, nsub, spc.n_nd, bands, tol, t_setup, t_solve, t_once),
            "steps": steps, "warmup": warmup, "iterations": its, "all_converged": all(conv),
            "seconds_per_step": total / steps, "eigenvalues_last": lams[-1] if lams else None}
