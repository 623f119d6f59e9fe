"""A/B timing of the ND operator apply (env BLOCH_ND_ITEM / BLOCH_ND_ITEM_THREADS select the kernel)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfem_bravais_b200 as m

st = torch.cuda.Stream()
flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
cases = [("CUB", 1, 48), ("FCC", 2, 8), ("FCC", 2, 16), ("HEX", 2, 16), ("BCC", 3, 8), ("BCC", 3, 12)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] + str(c[1]) in sys.argv[1:]] or cases
for name, p, n in cases:
    lat = m.BravaisLattice(name)
    eq = m.MaxwellBlochWaveEquation(lat, n, p)
    eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
    eq.set_stream(st.cuda_stream)
    eq.SetKappa(0.5 * lat.GetSymmetryPoint(1)); eq.Setup()
    for nv in (1, 10, 16, 30):
        x = torch.rand(eq.N * nv * 2, device="cuda", dtype=torch.float64) * 2 - 1
        y = torch.empty_like(x)
        ts = []
        with torch.cuda.stream(st):
            for _ in range(5):
                eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
            for _ in range(15):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
                e1.record(st)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
        st.synchronize()
        t = float(np.median(ts))
        print(json.dumps({"case": f"{name} p{p} n{n}", "N": eq.N, "nv": nv, "us": round(t * 1e6, 1),
                          "gdofs": round(eq.N * nv / t / 1e9, 2), "hbm_frac": round(32 * eq.N * nv / t / 6544.7e9, 4),
                          "item": os.environ.get("BLOCH_ND_ITEM", "1"), "comp": os.environ.get("BLOCH_ND_COMP", "4"), "ov": os.environ.get("BLOCH_ND_ITEM_OVERLAY", "0")}),
              flush=True)
        del x, y
    del eq
