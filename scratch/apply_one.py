"""One configuration of the ND apply for ncu captures: python scratch/apply_one.py LATTICE P N NVEC [reps]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfem_bravais_b200 as m

name, p, n, nv = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 8
lat = m.BravaisLattice(name)
eq = m.MaxwellBlochWaveEquation(lat, n, p)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
eq.SetKappa(0.5 * lat.GetSymmetryPoint(1)); eq.Setup()
x = torch.rand(eq.N * nv * 2, device="cuda", dtype=torch.float64) * 2 - 1
y = torch.empty_like(x)
for _ in range(reps):
    eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
torch.cuda.synchronize()
print("ok", eq.N, nv)
