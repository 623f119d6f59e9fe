"""Field output of the eigenmodes (WriteVisitFields, maxwell/maxwell_bloch.cpp:1730-1822; GetEigenvector,
:1371-1458) and the sharded dispersion sweep on the device.  Known answer: in the empty lattice (eps = mu = 1) the
lowest band at kappa is the constant envelope E0 perpendicular to kappa with lambda = |kappa|^2 exactly, and
B (GetEigenvectorB convention: i (C - i kappa x) E / sqrt(lambda)) has the constant envelope kappa x E0 / |kappa|."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_write_visit_fields_empty_lattice(bloch, tmp_path):
    lat = bloch.BravaisLattice("FCC")
    eq = bloch.MaxwellBlochWaveEquation(lat, 2, 2)
    eq.SetMassCoef(np.ones(eq.n_elem))
    kappa = 0.3 * lat.GetSymmetryPoint(lat.GetSymmetryPointIndex("L"))
    eq.SetAbsoluteTolerance(1e-10)
    lam = eq.GetEigenvalues(8, kappa)
    assert abs(lam[0] - kappa @ kappa) < 1e-9 * (kappa @ kappa)
    names = eq.WriteVisitFields(str(tmp_path), "Maxwell-Dispersion-test", eps=np.ones(eq.n_elem))
    assert len(names) == 8 and os.path.exists(tmp_path / "Maxwell-Dispersion-test.visit")
    times = [l.split() for l in open(tmp_path / "Maxwell-Dispersion-test.times")]
    assert abs(float(times[0][2]) - np.linalg.norm(kappa)) < 1e-8          # time = omega of the mode
    pts, f = bloch.read_vtk_fields(str(tmp_path / names[0]))
    E = f["E_r"] + 1j * f["E_i"]
    B = f["B_r"] + 1j * f["B_i"]
    # the lowest pair is degenerate: any combination of the two constant polarisations, still constant and transverse
    assert np.abs(E - E[0]).max() < 1e-7 and np.abs(E @ kappa).max() < 1e-7
    assert np.abs(B - B[0]).max() < 1e-7
    # GetEigenvectorB returns Bi = Re(C E)/omega, Br = -Im(C E)/omega, i.e. B = i (C - i kappa x) E / omega = kappa x E / omega
    expect = np.cross(kappa, E[0]) / np.linalg.norm(kappa)
    assert np.abs(B[0] - expect).max() < 1e-7
    # second real mode of the band = i times the first (the pairing of the reference's real block form)
    _, f2 = bloch.read_vtk_fields(str(tmp_path / names[1]))
    assert np.allclose(f2["E_r"], -f["E_i"], atol=1e-12) and np.allclose(f2["E_i"], f["E_r"], atol=1e-12)
    er, ei, br, bi = eq.GetEigenvector(0)
    assert er.shape == (eq.N,) and br.shape == (eq.N_rt,)


def test_cpp_driver_writes_visit_fields(bloch, tmp_path):
    exe = os.path.join(ROOT, "mfem-bravais_b200", "lib", "maxwell_dispersion_b200")
    r = subprocess.run([exe, "-bl", "1", "-o", "1", "-sr", "1", "-pr", "1", "-np", "0", "-nb", "3", "-visit", "-out", str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    files = sorted(p for p in os.listdir(tmp_path) if p.endswith(".vtk"))
    assert any(p.startswith("Maxwell-Dispersion-Gamma_") for p in files) and len(files) >= 6
    # the C++ writer and the Python reader agree on the format; fields are finite and not identically zero
    pts, f = bloch.read_vtk_fields(str(tmp_path / [p for p in files if "-X_" in p][0]))
    assert pts.shape[1] == 3 and set(f) == {"E_r", "E_i", "B_r", "B_i"}
    assert np.isfinite(f["E_r"]).all() and np.abs(f["E_r"]).max() + np.abs(f["E_i"]).max() > 1e-6


def test_sharded_dispersion_sweep_equals_single_rank(bloch, tmp_path):
    """configs[3] in small: the HEX sweep split over 2 and 3 emulated ranks gives the single-rank disp.dat"""
    lat = bloch.BravaisLattice("HEX")
    eqs = [bloch.MaxwellBlochWaveEquation(lat, 2, 2) for _ in range(2)]
    eps = bloch.sphere_eps(eqs[0].element_centers())
    for eq in eqs:
        eq.SetMassCoef(eps)
    outs = []
    for world in (1, 2, 3):
        lam = None
        for rank in range(world):
            rows, uk, lo, res = bloch.sharded_dispersion_sweep(eqs, lat, 1, 6, 3, world, rank, tol=1e-8)
            assert (res["converged"] == 6).all()
            if lam is None:
                lam = np.zeros((len(uk), 6))
            lam[lo:lo + len(res["lam"])] = res["lam"]
        outs.append(lam)
        bloch.write_dispersion_data(str(tmp_path / ("disp%d.dat" % world)), rows, lam)
    assert np.allclose(outs[0], outs[1], rtol=1e-6, atol=1e-7) and np.allclose(outs[0], outs[2], rtol=1e-6, atol=1e-7)
    a = np.loadtxt(tmp_path / "disp1.dat", usecols=range(2, 14))
    b = np.loadtxt(tmp_path / "disp3.dat", usecols=range(2, 14))
    assert a.shape == (len(rows), 12) and np.allclose(a, b, rtol=1e-5, atol=1e-6)
