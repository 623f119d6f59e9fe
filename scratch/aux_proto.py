"""Prototype of the auxiliary-space (Hiptmair-Xu / AMS-like) preconditioner on the oracle's assembled matrices:
T = smoother + Pi L^-1 Pi^H,  Pi: (H1_p)^3 -> ND_p nodal interpolation, L = Bloch Laplacian(mu^-1) [+ sigma*mass].
Counts LOBPCG iterations against the Chebyshev-Jacobi polynomial.  CPU only; projections by sparse LU."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl, scipy.linalg as sla
sys.path.insert(0, ".")
from oracle.bloch_oracle import Lattice, Mesh, Spaces, BlochOperators, _assemble_set
from oracle import cpu_solver as cs


def build_pi(spc):
    ref, mesh = spc.ref, spc.mesh
    h1v, _ = ref.h1_shapes(ref.nd_nodes)            # [n_nd_loc, n_h1_loc]
    comp = ref.nd_comp
    em = []
    for J in mesh.J:
        E = np.zeros((ref.n_nd, 3 * ref.n_h1))
        for d in range(3):
            E[:, d * ref.n_h1:(d + 1) * ref.n_h1] = h1v * J[d, comp][:, None]     # (J e_c) . e_d * phi_a(X_i)
        em.append(E)
    gc = np.concatenate([spc.h1_gid + d * spc.n_h1 for d in range(3)], axis=1)
    sc = np.ones_like(gc, dtype=float)
    return _assemble_set(spc.nd_gid, spc.nd_sign, gc, sc, spc.n_nd, 3 * spc.n_h1, np.array(em), mesh.cls)


def lobpcg(A, M, project, precond, N, nb, tol, max_iter=300, seed=1):
    mb = nb + max(6, nb // 4)
    rng = np.random.default_rng(seed)
    X = project(rng.uniform(-1, 1, (N, mb)) + 1j * rng.uniform(-1, 1, (N, mb)))
    def rr(S, AS, MS):
        GA = S.conj().T @ AS; GM = S.conj().T @ MS
        w, C = sla.eigh(0.5 * (GA + GA.conj().T), 0.5 * (GM + GM.conj().T))
        return w[:mb], C[:, :mb]
    AX, MX = A @ X, M @ X
    lam, C = rr(X, AX, MX); X, AX, MX = X @ C, AX @ C, MX @ C
    P = None
    for it in range(max_iter):
        R = AX - MX * lam
        rn = np.linalg.norm(R[:, :nb], axis=0)
        if rn.max() <= tol:
            return lam[:nb], it
        W = project(precond(R))
        AW, MW = A @ W, M @ W
        bl = [X, W] + ([P] if P is not None else []); Ab = [AX, AW] + ([AP] if P is not None else []); Mb = [MX, MW] + ([MP] if P is not None else [])
        S, AS, MS = np.hstack(bl), np.hstack(Ab), np.hstack(Mb)
        try:
            lam, C = rr(S, AS, MS)
        except np.linalg.LinAlgError:
            S, AS, MS = np.hstack(bl[:2]), np.hstack(Ab[:2]), np.hstack(Mb[:2]); lam, C = rr(S, AS, MS)
        Cp = C.copy(); Cp[:mb] = 0
        P, AP, MP = S @ Cp, AS @ Cp, MS @ Cp
        X, AX, MX = S @ C, AS @ C, MS @ C
    return lam[:nb], max_iter


def run(lat, n, p, modes, nb=10, kappa=(2.1, 0.9, 0.4)):
    L = Lattice(lat); mesh = Mesh(L, n); spc = Spaces(mesh, p)
    eps = mesh.sphere_eps()
    ops = BlochOperators(spc, eps); ops.set_kappa(np.array(kappa))
    class W:
        def __init__(self, A): self.c = cs.Csr(A.tocsr().astype(complex))
        def __matmul__(self, X): return self.c(np.ascontiguousarray(X, complex))
    A0, M0c, G0 = ops.A_c().tocsr(), ops.M_c().tocsr(), ops.G_c().tocsr()
    A, M, G = W(A0), W(M0c), W(G0)
    GH = W(G0.conj().T.tocsr())
    luS0 = spl.splu(ops.S0_c().tocsc())
    project = lambda X: X - G @ luS0.solve(np.ascontiguousarray(GH @ (M @ X)))
    sigma = 1.0 / mesh.volume ** (2.0 / 3.0)
    Ash0 = (A0 + sigma * M0c).tocsr()
    dA = Ash0.diagonal().real; jac = 1.0 / dA
    dsq = 1.0 / np.sqrt(dA)
    lmax = float(abs(sp.diags(dsq) @ Ash0 @ sp.diags(dsq)).sum(axis=1).max())
    Ash = W(Ash0)
    N = spc.n_nd
    Pi0 = build_pi(spc).tocsr(); Pi = W(Pi0); PiH = W(Pi0.T.tocsr())
    ops1 = BlochOperators(spc, np.ones(mesh.ne)); ops1.set_kappa(np.array(kappa))
    n0 = spc.n_h1
    luL = spl.splu(ops1.S0_c().tocsc())
    luLs = spl.splu((ops1.S0_c() + sigma * ops.M0).tocsc())

    def cheb(R, degree=24, ratio=300.0, lo=None):
        lmin = lmax / ratio if lo is None else lo
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin); s1 = theta / delta
        r = R.copy(); d = jac[:, None] * r / theta; x = d.copy(); rho = 1.0 / s1
        for _ in range(1, degree):
            r -= Ash @ d
            rho_n = 1.0 / (2.0 * s1 - rho)
            d = rho_n * rho * d + (2.0 * rho_n / delta) * (jac[:, None] * r)
            x += d; rho = rho_n
        return x

    def aux(R, lu):
        r3 = PiH @ R
        z = np.empty_like(r3)
        for d in range(3):
            z[d * n0:(d + 1) * n0] = lu.solve(np.ascontiguousarray(r3[d * n0:(d + 1) * n0]))
        return Pi @ z

    def mult(R, lu, smooth):
        x = smooth(R); r = R - Ash @ x
        x = x + aux(r, lu); r = R - Ash @ x
        return x + smooth(r)

    table = {
        "cheb": lambda R: cheb(R),
        "aux_j": lambda R: mult(R, luL, lambda r: (4.0 / (3.0 * lmax)) * jac[:, None] * r),
        "aux_js": lambda R: mult(R, luLs, lambda r: (4.0 / (3.0 * lmax)) * jac[:, None] * r),
        "aux_c2": lambda R: mult(R, luL, lambda r: cheb(r, 2, 4.0)),
        "aux_c3": lambda R: mult(R, luL, lambda r: cheb(r, 3, 6.0)),
        "aux_add": lambda R: (4.0 / (3.0 * lmax)) * jac[:, None] * R + aux(R, luL),
    }
    print("setup done, lmax %.3f" % lmax, flush=True)
    for mode in modes:
        t = time.time()
        lam, it = lobpcg(A, M, project, table[mode], N, nb, 1e-6)
        print("%s n=%d p=%d N=%d %-8s %3d iterations  %.1f s  lam[:3]=%s" % (lat, n, p, N, mode, it, time.time() - t, np.round(lam[:3], 6)), flush=True)


if __name__ == "__main__":
    run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4:])
