"""TEST INFRASTRUCTURE ONLY: CPU baseline for bench.py ("port" of the reference algorithm).

What the reference does per k-point (maxwell/maxwell_bloch.cpp:337-620, 809-825): assemble the
zeta-dependent interpolation matrices, form S1/DKZ (4 sparse triple products) and the projector's
A0/DKZ0 (4 more), then run a projected, preconditioned block LOBPCG whose every operator
application is a CSR SpMV.  This file times exactly that on the host cores with the oracle's
assembled matrices.  hypre's AMS is not available, so the preconditioner is the same Chebyshev /
Jacobi polynomial the CUDA path uses and the projector's inner solve is Jacobi-PCG - i.e. the
SAME iteration as the GPU path, on assembled CSR matrices with a threaded SpMM (csr_spmm.c).
"""
import ctypes as C
import os
import time

import numpy as np
import scipy.linalg as sla

from .bloch_oracle import BlochOperators, Lattice, Mesh, Spaces

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _spmm_lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "_ref", "libcsr_spmm.so")
        if os.path.exists(p):
            _lib = C.CDLL(p)
            _lib.csr_num_threads.restype = C.c_int
        else:
            _lib = False
    return _lib


class Csr:
    def __init__(self, A):
        A = A.tocsr().astype(complex)
        A.sort_indices()
        self.A = A
        self.n = A.shape[0]
        self.indptr = np.ascontiguousarray(A.indptr, np.int32)
        self.indices = np.ascontiguousarray(A.indices, np.int32)
        self.data = np.ascontiguousarray(A.data, complex)

    def __call__(self, X):
        L = _spmm_lib()
        X = np.ascontiguousarray(X, complex)
        if not L:
            return self.A @ X
        Y = np.empty((self.n, X.shape[1]), complex)
        L.csr_spmm_z(C.c_int64(self.n), C.c_int(X.shape[1]), self.indptr.ctypes.data_as(C.c_void_p),
                     self.indices.ctypes.data_as(C.c_void_p), self.data.ctypes.data_as(C.c_void_p),
                     X.ctypes.data_as(C.c_void_p), Y.ctypes.data_as(C.c_void_p))
        return Y


def threads():
    L = _spmm_lib()
    return L.csr_num_threads() if L else 1


def lobpcg_cpu(ops, nb, tol, max_iter=2000, sigma_scale=1.0, cheb_degree=24, cheb_ratio=300.0,
               proj_tol=1e-9, seed=1, verbose=False, timing=None):
    mesh = ops.sp_.mesh
    A, M, G = Csr(ops.A_c()), Csr(ops.M_c()), Csr(ops.G_c())
    GH = Csr(ops.G_c().conj().T)
    S0 = Csr(ops.S0_c())
    N = A.n
    mb = min(32, nb + max(4, nb // 4))
    sigma = sigma_scale / mesh.volume ** (2.0 / 3.0) + ops.beta ** 2
    dA = ops.A_c().diagonal().real + sigma * ops.M_c().diagonal().real
    jac = (1.0 / dA)[:, None]
    jac0 = (1.0 / ops.S0_c().diagonal().real)[:, None]
    counts = {"A": 0, "S0": 0}

    def shifted(X):
        counts["A"] += X.shape[1]
        return A(X) + sigma * M(X)

    rng = np.random.default_rng(seed)
    # rigorous bound of lambda_max(D^-1 (A + sigma M)): Gershgorin row sums of the scaled matrix
    # (the GPU path uses the element-local spectra for the same purpose)
    Ash = (ops.A_c() + sigma * ops.M_c()).tocsr()
    dsq = 1.0 / np.sqrt(dA)
    import scipy.sparse as _sp
    Asc = _sp.diags(dsq) @ Ash @ _sp.diags(dsq)
    lmax = float(abs(Asc).sum(axis=1).max())
    lmin = lmax / cheb_ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    s1 = theta / delta

    def precond(R):
        r = R.copy()
        d = jac * r / theta
        x = d.copy()
        rho = 1.0 / s1
        for _ in range(1, cheb_degree):
            r -= shifted(d)
            rho_n = 1.0 / (2.0 * s1 - rho)
            d = rho_n * rho * d + (2.0 * rho_n / delta) * (jac * r)
            x += d
            rho = rho_n
        return x

    def project(X, rel):
        rhs = GH(M(X))
        rr0 = np.sum(np.abs(rhs) ** 2, axis=0)
        if rr0.max() == 0:
            return X
        phi = np.zeros_like(rhs)
        r = rhs.copy()
        z = jac0 * r
        p = z.copy()
        rz = np.sum((r.conj() * z).real, axis=0)
        for it in range(3000):
            q = S0(p)
            counts["S0"] += p.shape[1]
            pq = np.sum((p.conj() * q).real, axis=0)
            alpha = np.where(pq != 0, rz / np.where(pq != 0, pq, 1), 0)
            phi += alpha * p
            r -= alpha * q
            z = jac0 * r
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

    X = project(rng.uniform(-1, 1, (N, mb)) + 1j * rng.uniform(-1, 1, (N, mb)), 1e-10)
    AX, MX = A(X), M(X)
    lam, Cv = rr(X, AX, MX)
    X, AX, MX = X @ Cv, AX @ Cv, MX @ Cv
    P = AP = MP = None
    it = 0
    t_loop = time.time()
    if timing is not None:
        timing["t_init"] = t_loop - timing["t_start"]
    for it in range(max_iter):
        R = AX - MX * lam
        rn = np.linalg.norm(R[:, :nb], axis=0)
        if verbose:
            print("[cpu lobpcg] it %d maxres %.3e" % (it, rn.max()))
        if rn.max() <= tol:
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
    return lam[:nb], it, counts


def time_kpoints(lattice, nsub, order, labels, pts, bands, tol, steps, first=0, sample_iters=0,
                 assumed_iterations=20):
    """sample_iters == 0: full solves.  sample_iters > 0: each step runs the per-k setup, the initial
    projection + Rayleigh-Ritz and `sample_iters` LOBPCG iterations at FULL size and extrapolates
    linearly to `assumed_iterations` (the count the same algorithm needs; stated in `sample`)."""
    lat = Lattice(lattice)
    mesh = Mesh(lat, nsub)
    t0 = time.time()
    sp = Spaces(mesh, order)
    ops = BlochOperators(sp, mesh.sphere_eps())
    t_once = time.time() - t0
    ks = lat.kpath(labels, pts)
    t_setup = t_solve = 0.0
    its, lams = [], []
    for s in range(steps):
        k = ks[(first + s) % len(ks)]
        t0 = time.time()
        ops.set_kappa(k)          # per-k assembly + sparse triple products (reference: Setup())
        t1 = time.time()
        timing = {"t_start": t1}
        if sample_iters > 0:
            lam, it, cnt = lobpcg_cpu(ops, bands, 0.0, max_iter=sample_iters, timing=timing)
            t_s = timing["t_init"] + timing["t_iters"] / sample_iters * assumed_iterations
            it = assumed_iterations
        else:
            lam, it, cnt = lobpcg_cpu(ops, bands, tol, timing=timing)
            t_s = time.time() - t1
        t_setup += t1 - t0
        t_solve += t_s
        its.append(it)
        lams.append(lam.tolist())
    total = t_setup + t_solve
    how = ("full solves" if sample_iters == 0 else
           "setup + initial projection/Rayleigh-Ritz + %d LOBPCG iterations timed at full size, extrapolated "
           "linearly to %d iterations" % (sample_iters, assumed_iterations))
    return {"value": steps / total, "unit": "k-points/s", "cores": threads(), "kind": "port",
            "sample": "%d k-point(s) of %s order %d n_sub=%d (N=%d), %d bands, tol %g; %s; per-k sparse products "
                      "%.1f s + projected LOBPCG %.1f s per sample set (one-time assembly %.1f s not counted)"
                      % (steps, lattice, order, nsub, sp.n_nd, bands, tol, how, t_setup, t_solve, t_once),
            "steps": steps, "iterations": its}
