#!/usr/bin/env python
"""bench.py - headline benchmark of the Bloch Maxwell eigen path (BASELINE.json metric:
"Bloch curl-curl apply GDOF/s; k-points/sec (10 bands, tol 1e-6)").

A *step* is one k-point eigen-solve (SetKappa + Setup + Solve, 10 complex bands, abs. residual
tol 1e-6) of configs[1]: FCC lattice, dielectric sphere (eps 10 inside r <= 0.25), ND order 2,
the 32-point path Gamma-X-W-L-Gamma.  `value` = k-points/sec over all ranks.  The apply kernel
(Y = A X, the curl-curl operator) is timed separately with CUDA events and reported in
`roofline` (algorithmic 32 B per complex DOF per vector) and `apply_gdofs`.

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference ...                     # CPU restatement (oracle port)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PATH_LABELS = ["Gamma", "X", "W", "L", "Gamma"]
ALG_BYTES_PER_DOF = 32.0   # read x (16 B) + write y (16 B) per complex DOF per vector


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", default="FCC")
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--n-sub", type=int, default=int(os.environ.get("BLOCH_BENCH_NSUB", "8")))
    ap.add_argument("--bands", type=int, default=10)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--pts-per-segment", type=int, default=8)
    ap.add_argument("--streams", type=int, default=int(os.environ.get("BLOCH_BENCH_STREAMS", "4")),
                    help="concurrent k-point solves per GPU (independent handles on separate streams)")
    ap.add_argument("--chunk", type=int, default=0,
                    help="k-points per work unit of the stream pool (0 = steps / (2 * streams))")
    ap.add_argument("--apply-vectors", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--apply-study", action="store_true",
                    help="operator-apply roofline study (configs[2]: BCC order 3, ~2.2M complex DOF, 1/4/10/30 RHS)")
    ap.add_argument("--cpu-sample-nsub", type=int, default=0, help="0 = same mesh as the workload")
    ap.add_argument("--cpu-sample-iters", type=int, default=1,
                    help="LOBPCG iterations timed per CPU sample (0 = full solves)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md).  Sampled through NVML from a
    Python thread (every 100 ms): polling with a `nvidia-smi -lms` child process was measured to slow the timed
    region by up to 40 % on some hosts (its queries serialise with kernel launches in the driver); nvidia-smi is
    only the fallback when the NVML binding is missing."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, "/tmp/bloch_clocks_%d_%d.csv" % (os.getpid(), index)
        self.thread, self.stop_flag, self.sm, self.mx, self.reasons = None, threading.Event(), [], [], set()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if self.index < len(ids) else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def _poll(self, nv, h):
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)),
                       "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "source": "nvidia-smi"}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def k_points(m, lat, args):
    return m.k_path(lat, PATH_LABELS, args.pts_per_segment)


# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import mfem_bravais_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lat = m.BravaisLattice(args.lattice)
    ks = k_points(m, lat, args)
    nk = len(ks)
    cores = os.cpu_count() or 1
    T = max(1, min(args.streams, max(1, cores // max(1, world))))   # solver threads per rank <= cores per rank
    if T * world * 2 > cores:
        os.environ["BLOCH_BLOCKING_SYNC"] = "1"                    # sleep, do not spin, when threads are scarce
    eqs = [m.MaxwellBlochWaveEquation(lat, args.n_sub, args.order, device=local) for _ in range(T)]
    eps = m.sphere_eps(eqs[0].element_centers())
    for eq in eqs:
        eq.SetMassCoef(eps)
        eq.SetNumEigs(2 * args.bands)
        eq.SetAbsoluteTolerance(args.tol, 2000)
    N = eqs[0].N

    iters_seen = []

    def solve_range(eq, idxs, out, e2e):
        for i in idxs:
            if e2e:
                eq.SetMassCoef(eps)          # host -> device copy of this step's coefficient field
            eq.SetKappa(ks[i % nk])
            eq.Setup()
            eq.Solve()
            out[i] = eq.band_eigenvalues() if e2e else None
            iters_seen.append(eq.GetSolverStats()["iterations"])

    def sweep(first, count, e2e):
        """solves k-points first..first+count-1 (mod path length) of this rank on T streams"""
        out = {}
        idxs = [first + j for j in range(count)]
        if T == 1:
            solve_range(eqs[0], idxs, out, e2e)
        else:
            # contiguous chunks (neighbouring k-points warm-start each other) handed out from a shared
            # queue, so a stream that drew easy k-points takes another chunk instead of idling
            csz = args.chunk if args.chunk > 0 else max(1, len(idxs) // (2 * T))
            chunks = [idxs[i:i + csz] for i in range(0, len(idxs), csz)]
            lock = threading.Lock()

            def worker(eq):
                while True:
                    with lock:
                        if not chunks:
                            return
                        mine = chunks.pop(0)
                    solve_range(eq, mine, out, e2e)

            th = [threading.Thread(target=worker, args=(eqs[t],)) for t in range(T)]
            [t.start() for t in th]
            [t.join() for t in th]
        return out

    base = rank * args.steps          # weak scaling: every rank gets its own `steps` k-points
    sweep(base, max(args.warmup, T), False)

    def timed(e2e):
        l0 = sum(eq.GetSolverStats()["kernel_launches"] for eq in eqs)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        out = sweep(base, args.steps, e2e)
        torch.cuda.synchronize()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        l1 = sum(eq.GetSolverStats()["kernel_launches"] for eq in eqs)
        return ms, l1 - l0, out

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, _ = timed(False)
    clocks = sampler.stop()
    ms_e2e, _, out = timed(True)
    its = [eq.GetSolverStats()["iterations"] for eq in eqs]
    its_all = iters_seen if iters_seen else its

    # ---- apply kernel: CUDA events on the handle's stream (= torch's current stream), L2 flushed ----
    eq = eqs[0]
    nv = args.apply_vectors
    st = torch.cuda.Stream()            # a real (non-null) stream shared by the events and the handle
    eq.set_stream(st.cuda_stream)
    eq.SetKappa(ks[3]); eq.Setup()
    g = torch.Generator(device="cuda"); g.manual_seed(12345)
    x = (torch.rand(N * nv * 2, device="cuda", dtype=torch.float64, generator=g) * 2 - 1)
    y = torch.empty_like(x)
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
    torch.cuda.synchronize()
    times = []
    with torch.cuda.stream(st):
        for _ in range(5):
            eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
        for _ in range(20):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
            e1.record(st)
            e1.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
    st.synchronize()
    eq.set_stream(0)
    t_apply = float(np.mean(times))
    hbm, how = peaks()
    achieved = ALG_BYTES_PER_DOF * N * nv / t_apply / 1e9
    traffic, shares = None, {}
    try:   # per-launch DRAM traffic of k_nd_apply from the committed ncu --set full capture (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
        key = "%s_p%d_n%d" % (args.lattice.lower(), args.order, args.n_sub)
        if key in tr and tr[key]["vectors"] == nv:
            traffic = tr[key]["dram_bytes_per_launch"]
        shares = json.load(open(os.path.join(ROOT, "profiles", "kernel_shares_r1.json")))
    except Exception:
        pass
    nd_kernel = ("k_nd_item<%d> (lane pair per item)" if args.order <= 2 else "k_nd_comp<%d> (six lanes per item)") % args.order
    nd_share = sum(v["share"] for k, v in shares.items() if k.startswith(("k_nd_item", "k_nd_comp", "k_nd_apply")))
    roof = {"bound": "hbm", "kernel": nd_kernel + ", y = A x, memset of y included",
            "achieved": achieved, "peak": hbm, "peak_source": how, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": traffic, "traffic_source": "profiles/ncu_nd_apply_r1.md" if traffic else None,
            "share_of_step_ncu": nd_share if shares else None,
            "dominant_kernel_of_step": max(shares.items(), key=lambda kv: kv[1]["share"])[0] if shares else None,
            "dominant_kernel_share": max((v["share"] for v in shares.values()), default=None),
            "launch_us_mean": t_apply * 1e6, "launch_us_best": float(np.min(times)) * 1e6,
            "dofs": N, "vectors": nv, "alg_bytes_per_dof_vector": ALG_BYTES_PER_DOF}
    # the same kernel in its throughput regime (the bench mesh is in the launch-latency regime): one mesh
    # refinement up, full 16-column solver block, measured live the same way
    try:
        del x, y
        big = m.MaxwellBlochWaveEquation(lat, 2 * args.n_sub, args.order, device=local)
        big.SetMassCoef(m.sphere_eps(big.element_centers()))
        big.set_stream(st.cuda_stream)
        big.SetKappa(ks[3]); big.Setup()
        nvb = 16
        xb = torch.rand(big.N * nvb * 2, device="cuda", dtype=torch.float64, generator=g) * 2 - 1
        yb = torch.empty_like(xb)
        tb = []
        with torch.cuda.stream(st):
            for _ in range(5):
                big.apply_A_device(xb.data_ptr(), yb.data_ptr(), nvb)
            for _ in range(20):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                big.apply_A_device(xb.data_ptr(), yb.data_ptr(), nvb)
                e1.record(st)
                e1.synchronize()
                tb.append(e0.elapsed_time(e1) * 1e-3)
        st.synchronize()
        tbm = float(np.mean(tb))
        Q = args.order + 1
        cfma = 12 * args.order * Q ** 3 + 12 * args.order ** 2 * Q * Q + 9 * args.order ** 2 * Q
        flops = 4.0 * cfma * big.n_elem * nvb
        fp64 = big.fp64_peak_tflops()
        roof["at_scale"] = {"workload": "%s order %d n_sub=%d, N=%d, %d vectors" % (args.lattice, args.order, 2 * args.n_sub, big.N, nvb),
                            "achieved": ALG_BYTES_PER_DOF * big.N * nvb / tbm / 1e9, "unit": "GB/s",
                            "frac": ALG_BYTES_PER_DOF * big.N * nvb / tbm / 1e9 / hbm, "gdofs": big.N * nvb / tbm / 1e9,
                            "launch_us_mean": tbm * 1e6, "fp64_tflops": flops / tbm / 1e12,
                            "fp64_peak_tflops_measured": fp64, "fp64_frac": flops / tbm / 1e12 / fp64}
        del big, xb, yb
    except Exception as ex:   # informational only
        roof["at_scale"] = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    total_steps = args.steps * world
    line = {
        "metric": "k-points/sec (10 bands, tol 1e-6)", "value": total_steps / (ms * 1e-3), "unit": "k-points/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, T), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s dielectric-sphere band structure along Gamma-X-W-L-Gamma, ND order %d, "
                               "n_sub=%d (N=%d complex DOF), %d k-points path, %d bands, tol %g"
                               % (args.lattice, args.order, args.n_sub, N, nk, args.bands, args.tol),
                   "k_points_per_rank": args.steps, "concurrent_solves_per_gpu": T,
                   "l2": "solver working set (basis [N][3m] x3 + temporaries) exceeds L2 only for n_sub>=16; "
                         "apply micro-benchmark flushes L2 (256 MiB write) between launches"},
        "e2e": {"value": total_steps / (ms_e2e * 1e-3), "unit": "k-points/s",
                "h2d_bytes_per_step": int(eps.nbytes + 24), "d2h_bytes_per_step": int(8 * args.bands)},
        "gpu_launches": int(launches), "lobpcg_iterations_mean": float(np.mean(its_all)),
        "apply_gdofs": N * nv / t_apply / 1e9,
        "roofline": roof, "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(args, steps=1, assumed_iterations=int(round(np.mean(its_all))))
        except Exception as e:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"error": repr(e)}
    _emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
def cpu_baseline(args, steps, assumed_iterations=20):
    """CPU restatement of the reference algorithm (oracle port): assembled CSR operators, per
    k-point RAP-style products, projected block LOBPCG.  Bounded sample of the same workload."""
    from oracle import cpu_solver
    nsub = args.cpu_sample_nsub or args.n_sub
    t0 = time.time()
    res = cpu_solver.time_kpoints(args.lattice, nsub, args.order, PATH_LABELS, args.pts_per_segment,
                                  args.bands, args.tol, steps, first=3, sample_iters=args.cpu_sample_iters,
                                  assumed_iterations=assumed_iterations)
    res["wall_s"] = time.time() - t0
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_baseline(args, steps=max(1, min(args.steps, 2)))
    line = {"impl": "reference", "metric": "k-points/sec (10 bands, tol 1e-6)", "value": res["value"],
            "unit": "k-points/s", "n_gpus": 0, "steps": res["steps"], "warmup": 0,
            "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": res["sample"]}, "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "k-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


def apply_study(args):
    """Y = A X throughput (GDOF/s) and fractions of the HBM (32 B/DOF/vector) and fp64 rooflines."""
    import torch
    import mfem_bravais_b200 as m
    hbm, how = peaks()
    out = []
    st = torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
    fp64 = None
    for name, p, n in [("CUB", 1, 48), ("FCC", 2, 16), ("BCC", 3, 12)]:
        lat = m.BravaisLattice(name)
        eq = m.MaxwellBlochWaveEquation(lat, n, p)
        eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
        eq.set_stream(st.cuda_stream)
        eq.SetKappa(0.5 * lat.GetSymmetryPoint(1)); eq.Setup()
        if fp64 is None:
            fp64 = eq.fp64_peak_tflops()
        Q = p + 1
        cfma = 12 * p * Q ** 3 + 12 * p * p * Q * Q + 9 * p * p * Q      # complex-by-real FMAs per element-vector
        flops_per_dof = 4.0 * cfma * eq.n_elem / eq.N
        for nv in (1, 4, 10, 30):
            x = torch.rand(eq.N * nv * 2, device="cuda", dtype=torch.float64) * 2 - 1
            y = torch.empty_like(x)
            ts = []
            with torch.cuda.stream(st):
                for _ in range(5):
                    eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
                for _ in range(20):
                    flush.fill_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
                    e1.record(st)
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e-3)
            st.synchronize()
            t = float(np.median(ts))
            gd = eq.N * nv / t / 1e9
            out.append({"lattice": name, "order": p, "n_sub": n, "N": eq.N, "vectors": nv, "median_us": t * 1e6,
                        "best_us": float(np.min(ts)) * 1e6, "gdofs": gd,
                        "hbm_frac": ALG_BYTES_PER_DOF * gd / hbm, "flops_per_dof": flops_per_dof,
                        "tflops": gd * flops_per_dof / 1e3, "fp64_frac": gd * flops_per_dof / 1e3 / fp64})
            del x, y
        del eq
    _emit(json.dumps({"apply_study": out, "hbm_peak_gbs": hbm, "hbm_peak_source": how, "fp64_peak_tflops": fp64,
                      "note": "timing = cudaMemset of y + k_nd_apply, CUDA events, L2 flushed between launches"}))


def _emit(line):
    """The JSON line goes to the process's ORIGINAL stdout; everything else that writes to fd 1 while the
    benchmark runs (e.g. NCCL's version banner at communicator creation) is sent to stderr."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.apply_study:
        apply_study(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
