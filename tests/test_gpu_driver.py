"""maxwell_dispersion-style sweeps: the Python sweep helpers and the C++ driver binary (built on
include/maxwell_bloch_b200.hpp) must agree with direct solves; disp.dat follows the reference's
format (maxwell/maxwell_dispersion.cpp:1062-1087)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dispersion_sweep_and_sharded_sweep_agree(bloch):
    L = bloch.BravaisLattice("HEX")
    eq = bloch.MaxwellBlochWaveEquation(L, 2, 2)
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
    ks = bloch.k_path(L, ["Gamma", "M", "K", "Gamma"], 2)
    lam, stats = bloch.dispersion_sweep(eq, ks, 4, tol=1e-8)
    assert lam.shape == (6, 4) and all(s["converged_bands"] == 4 for s in stats)
    assert np.all(np.diff(lam, axis=1) >= -1e-9)                        # ascending per k-point
    assert np.all(np.abs(lam[-1, :3]) < 1e-7)                           # back at Gamma: 3 zero modes

    def solve(k):
        eq.SetKappa(k)
        eq.Setup()
        eq.Solve()
        return eq.band_eigenvalues()

    eq.SetNumEigs(8)
    again = bloch.sharded_sweep(solve, ks, 4, None)
    assert np.allclose(again, lam, rtol=1e-6, atol=1e-7)


def test_cpp_driver_writes_reference_style_disp_dat(bloch, tmp_path):
    exe = os.path.join(ROOT, "mfem-bravais_b200", "lib", "maxwell_dispersion_b200")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built")
    r = subprocess.run([exe, "-bl", "1", "-o", "1", "-sr", "0", "-pr", "2", "-p", "2", "-np", "1", "-nb", "4",
                        "-out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [l.split("\t") for l in open(tmp_path / "disp.dat") if l.strip()]
    # CUB paths: Gamma-X-M-Gamma-R-X (5 segments) and M-R (1): per segment np+1 points + the closing point
    assert rows[0][1] == "Gamma" and rows[1][1] == "Delta" and rows[2][1] == "X"
    assert all(len(row) == 2 + 8 for row in rows)                        # nev = 2 * nb real modes
    w = np.array([[float(x) for x in row[2:]] for row in rows])
    assert np.allclose(w[:, 0::2], w[:, 1::2])                           # every band twice
    # same k-point solved through the Python mirror: X = pi * e_y
    L = bloch.BravaisLattice("CUB")
    eq = bloch.MaxwellBlochWaveEquation(L, 4, 1)
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
    lam = eq.GetEigenvalues(8, L.GetSymmetryPoint(1))
    assert np.allclose(w[2], np.sqrt(np.maximum(lam, 0)), rtol=1e-5, atol=1e-6)
    assert (tmp_path / "stats_0.out").read_text().startswith("Timings:")


def test_cpp_driver_matrix_dump(bloch, tmp_path):
    """-wm: Ar / Ai / M at the labelled k-points in hypre IJ format (maxwell_dispersion.cpp:553-590)"""
    exe = os.path.join(ROOT, "mfem-bravais_b200", "lib", "maxwell_dispersion_b200")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built")
    r = subprocess.run([exe, "-bl", "1", "-o", "1", "-sr", "0", "-pr", "1", "-p", "2", "-np", "0", "-nb", "2",
                        "-wm", "-out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    L = bloch.BravaisLattice("CUB")
    eq = bloch.MaxwellBlochWaveEquation(L, 2, 1)
    eq.SetMassCoef(bloch.sphere_eps(eq.element_centers()))
    eq.SetKappa(L.GetSymmetryPoint(L.GetSymmetryPointIndex("X")))
    eq.Setup()
    A, M = eq.AssembleMatrix("A"), eq.AssembleMatrix("M")
    Ar = bloch.read_hypre_ij(str(tmp_path / "ArX.mat.00000"))
    Ai = bloch.read_hypre_ij(str(tmp_path / "AiX.mat.00000"))
    Mr = bloch.read_hypre_ij(str(tmp_path / "MX.mat.00000"))
    assert abs(Ar - A.real).max() < 1e-12 and abs(Ai - A.imag).max() < 1e-12 and abs(Mr - M).max() < 1e-12
    assert abs(Ai + Ai.T).max() < 1e-12 and abs(Ai).max() > 0          # the beta DKZ block is antisymmetric


def test_cpp_driver_plane_wave_init_and_mesh_export(bloch, tmp_path):
    """-iv: the reference's CreateInitialVectors block per k-point (maxwell_dispersion.cpp:531, 735-1060) gives the
    same disp.dat as the built-in starting guess; -wmesh writes the cell for MFEM cross-checks."""
    exe = os.path.join(ROOT, "mfem-bravais_b200", "lib", "maxwell_dispersion_b200")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built")
    outs = []
    for extra in ([], ["-iv", "-wmesh"]):
        d = tmp_path / ("b" if extra else "a")
        d.mkdir()
        r = subprocess.run([exe, "-bl", "1", "-o", "1", "-sr", "0", "-pr", "2", "-p", "2", "-np", "0", "-nb", "4",
                            "-out", str(d)] + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        rows = [l.split("\t") for l in open(d / "disp.dat") if l.strip()]
        outs.append(np.array([[float(x) for x in row[2:]] for row in rows]))
    assert outs[0].shape == outs[1].shape and np.allclose(outs[0], outs[1], rtol=1e-5, atol=2e-5)
    V, H, attr, B = bloch.read_mfem_mesh(str(tmp_path / "b" / "ws-cell.mesh"))
    assert len(H) == 64 and set(attr) == {1, 2} and len(B) == 6 * 16


def test_cpp_driver_batched_walk_equals_sequential(bloch, tmp_path):
    """-kb N: the driver collects the unique k-points of the walk and solves them N at a time (bloch_set_kappa_batch);
    disp.dat must be the one the reference-style sequential walk writes."""
    exe = os.path.join(ROOT, "mfem-bravais_b200", "lib", "maxwell_dispersion_b200")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built")
    common = ["-bl", "2", "-o", "2", "-sr", "0", "-pr", "1", "-np", "1", "-nb", "4"]
    seq, bat = tmp_path / "seq", tmp_path / "bat"
    seq.mkdir(); bat.mkdir()
    for d, extra in ((seq, []), (bat, ["-kb", "5"])):
        r = subprocess.run([exe] + common + extra + ["-out", str(d)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
    ta, tb = (seq / "disp.dat").read_text(), (bat / "disp.dat").read_text()
    la, lb = [l.split("\t") for l in ta.split("\n")], [l.split("\t") for l in tb.split("\n")]
    assert len(la) == len(lb)
    for ra, rb in zip(la, lb):
        assert ra[:2] == rb[:2]                                           # counters, labels, blank lines between paths
        if len(ra) > 2:
            assert np.allclose([float(x) for x in ra[2:]], [float(x) for x in rb[2:]], rtol=1e-5, atol=1e-5)


def test_cpp_scalar3d_driver_matches_the_python_mirror(bloch):
    """scalar3d_b200 (C++ ScalarFloquetWaveEquation mirror + the reference's flags and coefficients,
    misc/scalar3d.cpp:280-557) prints the eigenvalues the Python mirror computes for the same problem."""
    exe = os.path.join(ROOT, "mfem-bravais_b200", "lib", "scalar3d_b200")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built")
    r = subprocess.run([exe, "-bl", "3", "-o", "2", "-sr", "0", "-pr", "2", "-nev", "8", "-b", "40", "-az", "20", "-inc", "30"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("Eigenvalues:")][0]
    ev = np.array(line.split()[1:], float)
    assert ev.shape == (8,) and np.allclose(ev[0::2], ev[1::2])
    L = bloch.BravaisLattice("BCC")
    eq = bloch.ScalarFloquetWaveEquation(L, 4, 2)
    r2 = np.linalg.norm(eq.element_centers(), axis=1)
    eq.SetMassCoef(np.where(r2 <= 0.5, 10.0, 1.0))
    eq.SetStiffnessCoef(np.where(r2 <= 0.5, 5.0, 0.1))
    eq.SetBeta(40.0)
    eq.SetAzimuth(20.0)
    eq.SetInclination(30.0)
    eq.SetNumEigs(8)
    eq.SetAbsoluteTolerance(1e-6)
    eq.Setup()
    eq.Solve()
    assert np.allclose(ev, eq.GetEigenvalues(), rtol=1e-6, atol=1e-8), (ev, eq.GetEigenvalues())
