"""one operator-apply workload for ncu: BCC order 3 n_sub 12 (2.24M complex DOF), 10 vectors"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import mfem_bravais_b200 as m
name, p, n, nv = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lat = m.BravaisLattice(name); eq = m.MaxwellBlochWaveEquation(lat, n, p)
eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
eq.SetKappa(0.5 * lat.GetSymmetryPoint(1)); eq.Setup()
x = torch.rand(eq.N * nv * 2, device="cuda", dtype=torch.float64) * 2 - 1
y = torch.empty_like(x)
for _ in range(6):
    eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
torch.cuda.synchronize()
print("done", eq.N)
