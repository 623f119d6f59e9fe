"""Assembled operators (the reference's -wm dump, maxwell_dispersion.cpp:553-590) against the oracle's
sparse matrices, entry by entry on the product's dof numbering."""
import numpy as np
import pytest

from helpers import oracle_on_product_maps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n,p", [("CUB", 2, 1), ("FCC", 1, 2), ("BCC", 1, 2), ("HEX", 1, 1), ("CUB", 1, 3)])
def test_assembled_matrices_match_oracle(bloch, name, n, p, tmp_path):
    L = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(L, n, p)
    rng = np.random.default_rng(2)
    eps, mui = rng.uniform(1, 5, eq.n_elem), rng.uniform(0.5, 2, eq.n_elem)
    eq.SetMassCoef(eps)
    eq.SetStiffnessCoef(mui)
    k = np.array([0.6, -0.2, 0.9])
    eq.SetKappa(k)
    eq.Setup()
    ops, _ = oracle_on_product_maps(eq, name, n, p, eps, mui)
    ops.set_kappa(k)
    A, M = eq.AssembleMatrix("A"), eq.AssembleMatrix("M")
    Ao, Mo = ops.A_c(), ops.M_c()
    assert abs(A - Ao).max() < 1e-11 * abs(Ao).max()
    assert abs(M - Mo).max() < 1e-12 * abs(Mo).max()
    assert abs(A - A.getH()).max() < 1e-12 * abs(Ao).max()
    # consistency with the matrix-free apply
    x = rng.standard_normal(2 * eq.N)
    y = eq.MultA(x)
    xc = x[: eq.N] + 1j * x[eq.N:]
    assert np.abs((A @ xc) - (y[: eq.N] + 1j * y[eq.N:])).max() < 1e-11 * np.abs(y).max()
    # hypre IJ round trip of the dump
    files = bloch.write_matrices(eq, str(tmp_path), "-t")
    assert [f.split("/")[-1] for f in files] == ["Ar-t.mat.00000", "Ai-t.mat.00000", "M-t.mat.00000"]
    Ar = bloch.read_hypre_ij(files[0])
    assert abs(Ar - A.real).max() < 1e-13 * abs(Ao).max()
