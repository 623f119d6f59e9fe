"""CPU tests of the host logic added for the k-point batch and the sharded sweep: slot chunking, the batched sweep
scheduler (driven with a stand-in equation object, no GPU), the enumeration of the maxwell_dispersion path with its
symmetry-point cache (maxwell_dispersion.cpp:475-531, 596-648), the disp.dat writer, field evaluation / VTK export,
and the post-processing helpers of the reference interface (DetermineBasis, IdentifyDegeneracies)."""
import os

import numpy as np
import pytest

from oracle.bloch_oracle import Lattice, empty_lattice_eigs


class FakeEq:
    """stand-in with the batch interface of MaxwellBlochWaveEquation: eigenvalues = exact empty-lattice spectrum"""

    def __init__(self, name):
        self.olat = Lattice(name)
        self.batches = []
        self.nb = 0

    def SetNumEigs(self, nev):
        self.nb = nev // 2

    def SetAbsoluteTolerance(self, tol, max_iter=2000):
        pass

    def SolveBatch(self, kappas):
        ks = np.asarray(kappas, float).reshape(-1, 3)
        self.batches.append(len(ks))
        lam = np.array([empty_lattice_eigs(self.olat, k, self.nb) for k in ks])
        return lam, [{"iterations": 7, "converged_bands": self.nb} for _ in ks]


def test_choose_batch(bloch):
    assert bloch.choose_batch(31, 2, 10) == 8          # 2 rounds, 1 padded solve instead of 9
    assert bloch.choose_batch(249, 2, 10) == 10        # 13 rounds of 20 slots
    assert bloch.choose_batch(20, 2, 10) == 10 and bloch.choose_batch(5, 2, 10) == 3 and bloch.choose_batch(0, 2, 10) == 1
    for n in range(1, 80):
        for T in (1, 2, 3):
            B = bloch.choose_batch(n, T, 10)
            assert 1 <= B <= 10 and -(-n // (T * B)) == -(-n // (T * 10))     # never more rounds than the largest batch


def test_slot_chunks(bloch):
    for n in (0, 1, 5, 20, 33):
        for s in (1, 3, 8, 10, 16):
            ch = bloch.slot_chunks(n, s)
            assert len(ch) == s and sum(ch, []) == list(range(n))
            sizes = [len(c) for c in ch]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n,T,B", [(20, 2, 5), (20, 2, 10), (13, 2, 4), (3, 2, 4), (32, 1, 8)])
def test_batched_sweep_scheduler(bloch, n, T, B):
    lat = bloch.BravaisLattice("FCC")
    ks = bloch.k_path(lat, ["Gamma", "X", "W", "L", "Gamma"], 8)[:n]
    eqs = [FakeEq("FCC") for _ in range(T)]
    res = bloch.batched_sweep(eqs, ks, 6, B)
    ref = np.array([empty_lattice_eigs(Lattice("FCC"), k, 6) for k in ks])
    assert np.allclose(res["lam"], ref)
    assert (res["converged"] == 6).all() and (res["iterations"] == 7).all()
    solved = sum(sum(e.batches) for e in eqs)
    assert solved == n + res["wasted"]                           # every k-point once, repeats only as padding
    for e in eqs:
        assert len(set(e.batches)) <= 1                           # the batch size of a handle never changes
    if n % (T * B) == 0:
        assert res["wasted"] == 0 and res["rounds"] == n // (T * B)


def test_batched_sweep_propagates_failures(bloch):
    class Broken(FakeEq):
        def SolveBatch(self, kappas):
            raise bloch.BlochError("bloch_solve failed (-4)")
    lat = bloch.BravaisLattice("FCC")
    ks = bloch.k_path(lat, ["Gamma", "X"], 4)
    with pytest.raises(bloch.BlochError):
        bloch.batched_sweep([FakeEq("FCC"), Broken("FCC")], ks, 4, 2)


@pytest.mark.parametrize("name,npt", [("FCC", 1), ("CUB", 3), ("HEX", 27)])
def test_dispersion_path_and_sharding(bloch, name, npt, tmp_path):
    lat = bloch.BravaisLattice(name)
    rows, uk = bloch.dispersion_path(lat, npt)
    nseg = sum(lat.GetNumberPathSegments(p) for p in range(lat.GetNumberPaths()))
    assert len(rows) == nseg * (npt + 1) + lat.GetNumberPaths()
    # symmetry points are solved once: a label that occurs twice maps to the same unique k-point
    seen = {}
    for label, u, p in rows:
        if label != "-":
            sp = lat.GetSymmetryPointIndex(label)
            if sp >= 0:
                assert seen.setdefault(label, u) == u
                assert np.allclose(uk[u], lat.GetSymmetryPoint(sp))
    assert len(uk) == len(set(u for _, u, _ in rows)) and max(u for _, u, _ in rows) == len(uk) - 1
    if name == "HEX":
        assert len(rows) == 255                                    # configs[3]: "256 k-points"
    # sharded over 1, 2, 3 ranks (emulated): identical disp.dat
    outs = []
    for world in (1, 2, 3):
        lam = np.zeros((len(uk), 4))
        for rank in range(world):
            r, u2, lo, res = bloch.sharded_dispersion_sweep([FakeEq(name)], lat, npt, 4, 3, world, rank)
            assert r == rows and np.array_equal(u2, uk)
            lam[lo:lo + len(res["lam"])] = res["lam"]
        path = tmp_path / ("disp_%d.dat" % world)
        bloch.write_dispersion_data(str(path), rows, lam)
        outs.append(open(path).read())
    assert outs[0] == outs[1] == outs[2]
    lines = [l for l in outs[0].split("\n") if l.strip()]
    assert len(lines) == len(rows)
    first = lines[0].split("\t")
    assert first[0] == "0" and first[1] == rows[0][0] and len(first) == 2 + 8      # counter, label, 2 x 4 real modes
    assert outs[0].count("\n\n") == lat.GetNumberPaths()                           # blank line after every path


@pytest.mark.parametrize("name,n,p", [("FCC", 2, 1), ("BCC", 1, 2), ("HEX", 2, 3)])
def test_field_evaluation_and_vtk_export(bloch, name, n, p, tmp_path):
    lat = bloch.BravaisLattice(name)
    eq = bloch.MaxwellBlochWaveEquation(lat, n, p, device=-2)
    a = np.array([0.3, -1.2, 0.7])
    const = lambda X: np.broadcast_to(a * (1.0 + 0.5j), X.shape)
    dofs = bloch.nd_interpolate(eq, const)
    X, E, _ = bloch.evaluate_fields(eq, dofs)
    assert E.shape == (eq.n_elem, 8, 3)
    assert np.abs(E - a * (1.0 + 0.5j)).max() < 1e-13            # constants are in the Nedelec space: exact
    # a lattice-periodic plane wave: interpolation error drops under refinement
    G = lat.GetReciprocalLatticeVectors()[0]
    pw = lambda Y: np.exp(2j * np.pi * (Y @ G))[:, None] * np.array([0.0, 1.0, 0.5])[None, :]
    errs = []
    for nn in (n, 2 * n):
        e2 = bloch.MaxwellBlochWaveEquation(lat, nn, p, device=-2)
        Xc, Ec, _ = bloch.evaluate_fields(e2, bloch.nd_interpolate(e2, pw), ref_points=np.array([[0.5, 0.5, 0.5]]))
        errs.append(np.abs(Ec - pw(Xc.reshape(-1, 3)).reshape(Ec.shape)).max())
    assert errs[1] < 0.6 * errs[0]
    path = tmp_path / "f.vtk"
    bloch.write_vtk_fields(eq, str(path), {"E_r": E.real, "E_i": E.imag}, {"epsilon": np.arange(eq.n_elem, dtype=float)})
    pts, fields = bloch.read_vtk_fields(str(path))
    assert pts.shape == (8 * eq.n_elem, 3) and np.allclose(pts, X.reshape(-1, 3))
    assert np.allclose(fields["E_r"], E.real.reshape(-1, 3)) and np.allclose(fields["E_i"], E.imag.reshape(-1, 3))
    txt = open(path).read()
    assert "CELL_TYPES %d" % eq.n_elem in txt and "SCALARS epsilon double 1" in txt


def test_determine_basis_and_degeneracies(bloch):
    lat = bloch.BravaisLattice("CUB")
    eq = bloch.MaxwellBlochWaveEquation(lat, 2, 1, device=-2)
    eq.SetKappa([0.3, -0.2, 0.9])
    e = eq.DetermineBasis([1.0, 0.0, 0.0])
    M = np.array(e)
    assert np.allclose(M @ M.T, np.eye(3), atol=1e-14) and np.linalg.det(M) > 0.999
    assert np.allclose(e[2], np.array([0.3, -0.2, 0.9]) / np.linalg.norm([0.3, -0.2, 0.9]))
    eq.SetKappa([0.0, 0.0, 0.0])
    assert np.allclose(np.array(eq.DetermineBasis([1.0, 2.0, 3.0])), np.eye(3))
    # IdentifyDegeneracies on a stubbed spectrum (maxwell_bloch.cpp:1493-1548)
    eq.GetEigenvalues = lambda: np.array([1e-9, 2e-9, 1.0, 1.0 + 1e-6, 2.0, 2.5, 2.5 + 1e-7])
    groups = eq.IdentifyDegeneracies(1e-4, 1e-4)
    assert groups == [{0, 1}, {2, 3}, {4}, {5, 6}]
    assert eq.ComputeHomogenizedCoefs() is None


def test_cpu_reference_arm_solver_small():
    """the CPU arm of bench.py (oracle/cpu_solver.py): full converged solves, checked against the oracle's dense
    constrained pencil on a mesh small enough for it"""
    from oracle import cpu_solver
    from oracle.bloch_oracle import BlochOperators, Mesh, Spaces
    res = cpu_solver.time_kpoints("FCC", 2, 2, ["Gamma", "X", "W", "L", "Gamma"], 8, 6, 1e-7, steps=2, first=2, warmup=1, nthreads=2)
    assert res["all_converged"] and res["steps"] == 2 and res["warmup"] == 1 and res["cores"] == 2
    olat = Lattice("FCC")
    mesh = Mesh(olat, 2)
    ops = BlochOperators(Spaces(mesh, 2), mesh.sphere_eps())
    ops.set_kappa(olat.kpath(["Gamma", "X", "W", "L", "Gamma"], 8)[4])
    assert np.allclose(res["eigenvalues_last"], ops.eig_dense(6), rtol=1e-7)
