"""B200-native Bloch Maxwell eigen path (drop-in for mfem-bravais' MaxwellBlochWaveEquation).

The compute path lives in csrc/ (CUDA, sm_100a) behind the C ABI of include/bloch_b200.h;
this package is the thin host-side mirror used by the tests, bench.py and the Python driver.
There is no CPU fallback: creating an equation without the built library or without a CUDA
device raises.
"""
from .capi import lib, lib_path, BlochError, LATTICE_TYPES  # noqa: F401
from .equation import BravaisLattice, MaxwellBlochWaveEquation, ScalarFloquetWaveEquation  # noqa: F401
from .dispersion import choose_batch, sharded_dispersion_sweep, batched_sweep, slot_chunks, dispersion_path, write_dispersion_data, k_path, sphere_eps, lattice_coefficient, dispersion_sweep, shard_kpoints, sharded_sweep, MaxwellDispersion, MaxwellBlochWaveSolver, homogenization_sweep, write_hypre_ij, read_hypre_ij, write_matrices, write_mfem_mesh, read_mfem_mesh, plane_wave_initial_vectors, nd_interpolate, evaluate_fields, write_vtk_fields, read_vtk_fields  # noqa: F401
