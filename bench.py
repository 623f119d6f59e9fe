#!/usr/bin/env python
"""bench.py - headline benchmark of the Bloch Maxwell eigen path (BASELINE.json metric:
"Bloch curl-curl apply GDOF/s; k-points/sec (10 bands, tol 1e-6)").

A *step* is one k-point eigen-solve (SetKappa + Setup + Solve, 10 complex bands, abs. residual tol 1e-6) of
configs[1]: FCC lattice, dielectric sphere (eps 10 inside r <= 0.25), ND order 2, the 32-point path
Gamma-X-W-L-Gamma.  `value` = k-points/sec over all ranks at n_sub = 8 (N = 49 152 complex DOF, the size the CPU
arm can solve completely); the same sweep one refinement up (n_sub = 16, N = 393 216) is reported in `n_sub16`.
Every timed solve is checked (all bands converged; eigenvalues against the committed oracle fixture
tests/golden/bands_baseline.json) - an unvalidated run exits non-zero.
`roofline` is the operator-apply kernel Y = A X on configs[2] (BCC order 3, 2.24 M complex DOF, 10 right-hand
sides; 1/4/16/30 in `study`), timed live with CUDA events, L2 flushed: the "apply GDOF/s" half of the metric.

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference algorithm (oracle port)
  python bench.py --sweep hex --gpus N                     # configs[3]: HEX 256-point sweep sharded over N ranks
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import numpy as np  # noqa: E402

PATH_LABELS = ["Gamma", "X", "W", "L", "Gamma"]
ALG_BYTES_PER_DOF = 32.0   # read x (16 B) + write y (16 B) per complex DOF per vector
GOLDEN = os.path.join(ROOT, "tests", "golden", "bands_baseline.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", default="FCC")
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--n-sub", type=int, default=int(os.environ.get("BLOCH_BENCH_NSUB", "8")))
    ap.add_argument("--bands", type=int, default=10)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--pts-per-segment", type=int, default=8)
    ap.add_argument("--streams", type=int, default=int(os.environ.get("BLOCH_BENCH_STREAMS", "4")),
                    help="concurrent handles per GPU (independent streams / host threads)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("BLOCH_BENCH_BATCH", "5")),
                    help="k-points iterated together inside one handle (bloch_set_kappa_batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the config-3 apply study (quick runs)")
    ap.add_argument("--no-n16", action="store_true", help="skip the n_sub = 16 leg")
    ap.add_argument("--apply-study", action="store_true",
                    help="operator-apply roofline study only (CUB p1 / FCC p2 / BCC p3, 1/4/10/16/30 RHS)")
    ap.add_argument("--sweep", default="", choices=["", "hex"],
                    help="hex: configs[3], the 256-point HEX dispersion sweep sharded over the ranks (strong scaling)")
    ap.add_argument("--sweep-np", type=int, default=27, help="points per path segment of --sweep (27 -> 255 rows)")
    ap.add_argument("--out", default="", help="--sweep: directory for disp.dat")
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
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md), sampled through NVML from a
    Python thread every 100 ms (a polling nvidia-smi child process slowed the timed region on some hosts)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
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
        except Exception:
            self.thread = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)),
                       "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        return out


def workload_config(args, N, nk):
    """the `config` object, identical for both arms"""
    return {"workload": "%s dielectric-sphere band structure along Gamma-X-W-L-Gamma, ND order %d, n_sub=%d "
                        "(N=%d complex DOF), %d-point path, %d bands, tol %g; secondary legs of the b200 arm: the same "
                        "sweep at n_sub=%d, operator apply on BCC order 3 n_sub=12"
                        % (args.lattice, args.order, args.n_sub, N, nk, args.bands, args.tol, 2 * args.n_sub),
            "lattice": args.lattice, "order": args.order, "n_sub": args.n_sub, "bands": args.bands, "tol": args.tol,
            "k_points_per_rank": args.steps, "k_points": "path indices 0..%d on every rank" % (args.steps - 1),
            "l2": "solver working set (basis [N][3 x batch x 16 columns] x3 + temporaries, ~1 GB) exceeds the 126 MB "
                  "L2; the apply micro-benchmarks flush L2 (256 MiB write) between launches"}


def golden_bands(lattice, order, n_sub):
    """{path index: eigenvalues} of the committed oracle fixture for this mesh (tests/golden/bands_baseline.json)"""
    out = {}
    try:
        for rec in json.load(open(GOLDEN)):
            if rec["lattice"] == lattice and rec["order"] == order and rec["n_sub"] == n_sub and rec.get("path_index") is not None:
                out[int(rec["path_index"])] = np.array(rec["eigenvalues"])
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------------------
def time_apply(m, torch, eq, nv, st, flush, reps=20):
    """CUDA-event timing of Y = A X (memset of y + apply kernel) on the handle's stream, L2 flushed between launches"""
    x = torch.rand(eq.N * nv * 2, device="cuda", dtype=torch.float64) * 2 - 1
    y = torch.empty_like(x)
    ts = []
    with torch.cuda.stream(st):
        for _ in range(5):
            eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            eq.apply_A_device(x.data_ptr(), y.data_ptr(), nv)
            e1.record(st)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
    st.synchronize()
    del x, y
    return float(np.mean(ts)), float(np.median(ts)), float(np.min(ts))


def apply_entry(eq, name, p, n, nv, t_mean, t_med, t_best, hbm, fp64):
    Q = p + 1
    cfma = 12 * p * Q ** 3 + 12 * p * p * Q * Q + 9 * p * p * Q      # complex-by-real FMAs per element-vector
    flops_per_dof = 4.0 * cfma * eq.n_elem / eq.N
    gd = eq.N * nv / t_mean / 1e9
    return {"lattice": name, "order": p, "n_sub": n, "N": eq.N, "vectors": nv, "launch_us_mean": t_mean * 1e6,
            "launch_us_median": t_med * 1e6, "launch_us_best": t_best * 1e6, "gdofs": gd,
            "achieved_gbs": ALG_BYTES_PER_DOF * gd, "hbm_frac": ALG_BYTES_PER_DOF * gd / hbm,
            "flops_per_dof": flops_per_dof, "fp64_tflops": gd * flops_per_dof / 1e3,
            "fp64_frac": gd * flops_per_dof / 1e3 / fp64 if fp64 else None}


def apply_roofline(m, torch, local, cases, vectors):
    """operator-apply study: list of entries (one per case x vector count)"""
    hbm, how = peaks()
    st = torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024 // 8, device="cuda", dtype=torch.float64)
    out, fp64 = [], None
    for name, p, n in cases:
        lat = m.BravaisLattice(name)
        eq = m.MaxwellBlochWaveEquation(lat, n, p, device=local)
        eq.SetMassCoef(m.sphere_eps(eq.element_centers()))
        eq.set_stream(st.cuda_stream)
        eq.SetKappa(0.5 * lat.GetSymmetryPoint(1))
        eq.Setup()
        if fp64 is None:
            fp64 = eq.fp64_peak_tflops()
        for nv in vectors:
            out.append(apply_entry(eq, name, p, n, nv, *time_apply(m, torch, eq, nv, st, flush), hbm, fp64))
        eq.set_stream(0)
        del eq
    return out, hbm, how, fp64


def traffic_for(kernel_key):
    """per-launch DRAM bytes of the roofline kernel from the committed `ncu --set full` capture of this round"""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
        return tr[kernel_key]["dram_bytes_per_launch"], tr[kernel_key].get("source")
    except Exception:
        return None, None


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
    ks = m.k_path(lat, PATH_LABELS, args.pts_per_segment)
    nk = len(ks)
    cores = os.cpu_count() or 1
    # solver threads per rank: they sleep while their stream works (blocking waits), 4 of them lose nothing on 2 cores
    # (profiles/tb_scan3_r2.log), so the cap is two threads per core of this rank's share; capped handles get a
    # larger batch so that the number of k-points in flight stays the same
    slots = max(1, args.streams) * max(1, args.batch)
    T = max(1, min(args.streams, max(1, 2 * cores // max(1, world))))
    B = max(1, min(16, -(-slots // T)))
    while T * B > args.steps and B > 1:
        B -= 1
    # prefer a batch size that fills every round (no padded repeats): the largest B' in [B/2, B] with steps % (T B') == 0
    for cand in range(B, max(1, B // 2) - 1, -1):
        if args.steps % (T * cand) == 0:
            B = cand
            break
    # host side: a handle's thread sleeps while it waits for its stream (a batched iteration takes ~20 ms, the wake-up
    # latency of a blocking wait is noise) instead of spinning on a core that the Rayleigh-Ritz threads of the other
    # handles / ranks need; the per-handle Rayleigh-Ritz pool is sized to the cores this rank can count on
    os.environ.setdefault("BLOCH_BLOCKING_SYNC", "1")
    os.environ.setdefault("BLOCH_RR_THREADS", str(max(1, min(4, cores // max(1, world * T)))))
    eqs = [m.MaxwellBlochWaveEquation(lat, args.n_sub, args.order, device=local) for _ in range(T)]
    eps = m.sphere_eps(eqs[0].element_centers())
    for eq in eqs:
        eq.SetMassCoef(eps)
    N = eqs[0].N
    # weak scaling: every rank solves the SAME `steps` k-points (path indices 0 .. steps-1), i.e. identical work per
    # GPU, so that the max-over-ranks time measures the machine and not which stretch of the path (the Gamma point,
    # degenerate symmetry points) a rank happened to draw; the sweep of ONE path sharded over the ranks is --sweep
    base = 0

    def kappa_at(pos):
        """position on the closed path in units of path points (fractional positions interpolate)"""
        i0 = int(np.floor(pos)) % nk
        f = pos - np.floor(pos)
        return ks[i0] if f == 0 else (1 - f) * ks[i0] + f * ks[(i0 + 1) % nk]

    timed_idx = [base + j for j in range(args.steps)]
    timed_kappas = np.array([kappa_at(i) for i in timed_idx])
    # warm-up: every slot solves the HALF-STEP predecessor of its first timed k-point - not a member of the timed
    # set, and as close to it as the previous path point is in the steady state of a long sweep; no timed solve
    # ever starts from its own eigenvectors
    chunks = m.slot_chunks(args.steps, T * B)
    warm_kappas = np.array([kappa_at(timed_idx[c[0]] - 0.5) for c in chunks if len(c) > 0])
    n_warm = len(warm_kappas)

    def upload_eps(eq):
        eq.SetMassCoef(eps)          # host -> device copy of this step's coefficient field (end-to-end leg)

    def warm():     # n_warm == T * B points on T * B slots: slot s gets warm_kappas[s], like chunk s of the timed sweep
        return m.batched_sweep(eqs, warm_kappas, args.bands, B, args.tol)

    def launches():
        return sum(eq.GetSolverStats()["kernel_launches"] for eq in eqs)

    def timed(e2e):
        warm()                         # state before the timed region: slots hold their half-step predecessors
        l0 = launches()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        res = m.batched_sweep(eqs, timed_kappas, args.bands, B, args.tol, per_solve=upload_eps if e2e else None)
        torch.cuda.synchronize()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms, launches() - l0, res

    warm()                             # first touch: kernel attributes, graphs, work space
    sampler = ClockSampler(local)
    sampler.start()
    ms, n_launch, res = timed(False)
    clocks = sampler.stop()
    ms_e2e, _, res_e2e = timed(True)

    # ---- validation of what was timed ----
    def validate(r, what):
        bad = np.nonzero(r["converged"] != args.bands)[0]
        if len(bad):
            raise SystemExit("bench.py: %s: k-point(s) %s did not converge all %d bands" % (what, bad.tolist(), args.bands))
        gold = golden_bands(args.lattice, args.order, args.n_sub)
        checked, worst = [], 0.0
        for j, i in enumerate(timed_idx):
            if (i % nk) in gold:
                err = float(np.max(np.abs(r["lam"][j] - gold[i % nk]) / np.abs(gold[i % nk])))
                checked.append(int(i % nk))
                worst = max(worst, err)
        if checked and worst > 1e-6:
            raise SystemExit("bench.py: %s: eigenvalues deviate from the oracle fixture by %.2e (> 1e-6)" % (what, worst))
        return {"all_bands_converged": True, "golden_path_indices": checked, "max_rel_err_vs_oracle": worst if checked else None,
                "tolerance": 1e-6, "fixture": "tests/golden/bands_baseline.json"}

    val = validate(res, "timed sweep")
    validate(res_e2e, "end-to-end sweep")
    if np.max(np.abs(res["lam"] - res_e2e["lam"]) / np.maximum(np.abs(res["lam"]), 1e-3)) > 1e-6:
        raise SystemExit("bench.py: timed and end-to-end sweeps disagree")
    validated = bool(val["golden_path_indices"]) or rank != 0

    # ---- live phase shares of one batched solve (CUDA events inside the handle, outside the timed region) ----
    shares = None
    try:
        eq = eqs[0]
        eq.SetProfile(True)
        b = min(B, args.steps)
        eq.SolveBatch(warm_kappas[:b] if n_warm >= b else timed_kappas[:b])
        eq.SolveBatch(timed_kappas[:b])
        pr = eq.GetProfile()
        eq.SetProfile(False)
        tot = pr["solve"]
        nd = pr["nd_apply_outside_precond"] + pr["nd_apply_in_precond"]
        groups = {"nd_operator_apply": nd, "projector_multigrid": pr["projector"], "lifted_operator_extra": pr["lift_extra"],
                  "chebyshev_vector_updates": pr["precond"] - pr["nd_apply_in_precond"], "gram_and_rotation": pr["gram_rotation"],
                  "host_rayleigh_ritz": pr["host_rr"]}
        shares = {"solve_ms": tot, "batch": b, "share": {k: v / tot for k, v in groups.items()},
                  "dominant": max(groups.items(), key=lambda kv: kv[1])[0],
                  "how": "CUDA events around the phases of one batched Solve() (bloch_set_profile), host part by wall clock"}
    except Exception as ex:   # informational
        shares = {"error": repr(ex)}

    # ---- secondary leg: the same sweep one refinement up ----
    n16 = None
    if not args.no_n16:
        try:
            big = m.MaxwellBlochWaveEquation(lat, 2 * args.n_sub, args.order, device=local)
            big.SetMassCoef(m.sphere_eps(big.element_centers()))
            b16, k16 = 4, 8
            c16 = m.slot_chunks(k16, b16)
            m.batched_sweep([big], np.array([kappa_at(base + c[0] - 0.5) for c in c16]), args.bands, b16, args.tol)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r16 = m.batched_sweep([big], timed_kappas[:k16], args.bands, b16, args.tol)
            torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            ms16 = e0.elapsed_time(e1)
            n16 = {"n_sub": 2 * args.n_sub, "N": big.N, "k_points": k16, "batch": b16, "value": k16 / (ms16 * 1e-3),
                   "unit": "k-points/s (this rank)", "ms_per_step": ms16 / k16,
                   "all_bands_converged": bool(np.all(r16["converged"] == args.bands)),
                   "lobpcg_iterations_mean": float(np.mean(r16["iterations"])),
                   "max_rel_dev_from_n_sub_%d" % args.n_sub: float(np.max(np.abs(r16["lam"] - res["lam"][:k16]) / np.abs(res["lam"][:k16])))}
            del big
        except Exception as ex:
            n16 = {"error": repr(ex)}

    # ---- roofline: operator apply on configs[2] (BCC order 3, n_sub 12), FCC order 2 as secondary ----
    roof, roof_fcc = None, None
    if not args.no_roofline and rank == 0:
        try:
            study, hbm, how, fp64 = apply_roofline(m, torch, local, [("BCC", 3, 12)], [10, 1, 4, 16, 30])
            head = study[0]
            traffic, tsrc = traffic_for("bcc_p3_n12_v10")
            nd_share = shares["share"]["nd_operator_apply"] if shares and "share" in shares else None
            roof = {"bound": "hbm", "kernel": "k_nd_comp<3> (six lanes per item), y = A x on BCC order 3 n_sub=12 (N=%d), "
                                              "%d right-hand sides, clearing of y included" % (head["N"], head["vectors"]),
                    "achieved": head["achieved_gbs"], "peak": hbm, "peak_source": how, "unit": "GB/s", "frac": head["hbm_frac"],
                    "traffic": traffic, "traffic_source": tsrc, "gdofs": head["gdofs"],
                    "fp64_tflops": head["fp64_tflops"], "fp64_peak_tflops_measured": fp64, "fp64_frac": head["fp64_frac"],
                    "launch_us_mean": head["launch_us_mean"], "launch_us_best": head["launch_us_best"], "dofs": head["N"],
                    "vectors": head["vectors"], "alg_bytes_per_dof_vector": ALG_BYTES_PER_DOF,
                    "study": [{k: e[k] for k in ("vectors", "gdofs", "hbm_frac", "fp64_frac", "launch_us_mean")} for e in study],
                    "share_of_headline_step": nd_share, "phase_shares_of_headline_step": shares}
            sf, _, _, _ = apply_roofline(m, torch, local, [(args.lattice, args.order, args.n_sub), (args.lattice, args.order, 2 * args.n_sub)],
                                         [16, 16 * B])
            roof_fcc = [{k: e[k] for k in ("lattice", "order", "n_sub", "N", "vectors", "gdofs", "hbm_frac", "fp64_frac", "launch_us_mean")} for e in sf]
        except Exception as ex:
            roof = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    total_steps = args.steps * world
    line = {
        "metric": "k-points/sec (10 bands, tol 1e-6)", "value": total_steps / (ms * 1e-3), "unit": "k-points/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, n_warm), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, N, nk),
        "impl_config": {"handles_per_gpu": T, "k_points_batched_per_handle": B, "slots": T * B,
                        "padded_solves": int(res["wasted"]), "rounds": int(res["rounds"]),
                        "preconditioner": ("chebyshev-jacobi degree 24" if os.environ.get("BLOCH_PRECOND") == "cheb" else
                                           "auxiliary space: Chebyshev(2)-Jacobi smoother + Pi (H1)^3 multigrid V-cycles"),
                        "warmup_solves": "each of the %d slots solves the half-step predecessor of its first timed "
                                         "k-point (not in the timed set) before every timed region" % n_warm},
        "e2e": {"value": total_steps / (ms_e2e * 1e-3), "unit": "k-points/s",
                "h2d_bytes_per_step": int(eps.nbytes + 24), "d2h_bytes_per_step": int(8 * args.bands)},
        "gpu_launches": int(n_launch), "lobpcg_iterations_mean": float(np.mean(res["iterations"])),
        "validated": validated, "validation": val,
        "n_sub16": n16, "roofline": roof, "roofline_fcc": roof_fcc, "clocks": clocks,
    }
    if roof and "gdofs" in roof:
        line["apply_gdofs"] = roof["gdofs"]
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(args, steps=1, warmup=1)
        except Exception as e:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"error": repr(e)}
    _emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
def cpu_baseline(args, steps, warmup):
    """CPU restatement of the reference algorithm (oracle port): assembled CSR operators, per k-point RAP-style
    products, projected block LOBPCG; FULL converged solves on all host cores (oracle/cpu_solver.py)."""
    from oracle import cpu_solver
    t0 = time.time()
    nk = 4 * args.pts_per_segment
    # untimed k-points first, so that the timed ones are path indices 0 .. steps-1 like rank 0 of the GPU arm
    res = cpu_solver.time_kpoints(args.lattice, args.n_sub, args.order, PATH_LABELS, args.pts_per_segment,
                                  args.bands, args.tol, steps, first=(-warmup) % nk, warmup=warmup, nthreads=os.cpu_count())
    res["wall_s"] = time.time() - t0
    gold = golden_bands(args.lattice, args.order, args.n_sub)
    last = (steps - 1) % nk
    if last in gold and res.get("eigenvalues_last"):
        res["max_rel_err_vs_fixture"] = float(np.max(np.abs(np.array(res["eigenvalues_last"]) - gold[last]) / np.abs(gold[last])))
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mfem_bravais_b200 as m   # lattice / sizes only (topology handle, no GPU)
    lat = m.BravaisLattice(args.lattice)
    nk = len(m.k_path(lat, PATH_LABELS, args.pts_per_segment))
    topo = m.MaxwellBlochWaveEquation(lat, args.n_sub, args.order, device=-2)
    res = cpu_baseline(args, steps=args.steps, warmup=args.warmup)
    if not res["all_converged"]:
        raise SystemExit("bench.py --impl reference: a CPU solve did not converge")
    line = {"impl": "reference", "metric": "k-points/sec (10 bands, tol 1e-6)", "value": res["value"],
            "unit": "k-points/s", "n_gpus": 0, "steps": res["steps"], "warmup": res["warmup"],
            "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, topo.N, nk), "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "k-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------
def hex_sweep(args):
    """configs[3]: the HEX dispersion sweep (Gamma-M-K-Gamma-A-L-H-A, L-M, K-H; 255 rows, symmetry points solved
    once) sharded over the ranks in contiguous chunks, no collective on the solve path; results gathered, disp.dat
    written by rank 0.  Strong scaling: total work fixed.  Reports the per-rank times (tail imbalance)."""
    import torch
    import mfem_bravais_b200 as m
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lat = m.BravaisLattice("HEX")
    rows, uk = m.dispersion_path(lat, args.sweep_np)
    lo, hi = m.shard_kpoints(len(uk), world, rank)
    T = max(1, args.streams)
    B = m.choose_batch(hi - lo, T, max(1, args.batch))      # no mostly-padded last round on this rank's chunk
    eqs = [m.MaxwellBlochWaveEquation(lat, 8, 2, device=local) for _ in range(T)]
    eps = m.sphere_eps(eqs[0].element_centers())
    for eq in eqs:
        eq.SetMassCoef(eps)
    m.batched_sweep(eqs, uk[lo:lo + T * B] * 1.0 + 0.01, args.bands, B, args.tol)    # first touch (attributes, graphs)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows, uk, lo, res = m.sharded_dispersion_sweep(eqs, lat, args.sweep_np, args.bands, B, world, rank, args.tol)
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    my_ms = e0.elapsed_time(e1)
    parts = [(lo, res["lam"], my_ms, int(np.sum(res["iterations"])), bool(np.all(res["converged"] == args.bands)))]
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, parts[0])
        parts = gathered
    if rank == 0:
        lam = np.zeros((len(uk), args.bands))
        for plo, block, _, _, _ in parts:
            lam[plo:plo + len(block)] = block
        times = [p[2] for p in parts]
        outdir = args.out or os.path.join(ROOT, "gpurun_out")
        os.makedirs(outdir, exist_ok=True)
        path = os.path.join(outdir, "disp_hex_n%d.dat" % world)
        m.write_dispersion_data(path, rows, lam)
        np.save(os.path.join(outdir, "disp_hex_n%d.npy" % world), lam)
        line = {"metric": "HEX 256-point dispersion sweep, wall time (configs[3])", "value": max(times) * 1e-3, "unit": "s",
                "higher_is_better": False, "scaling": "strong", "n_gpus": world, "rows": len(rows), "unique_k_points": len(uk),
                "k_points_per_s": len(uk) / (max(times) * 1e-3), "per_rank_s": [t * 1e-3 for t in times],
                "tail_imbalance_max_over_mean": max(times) / float(np.mean(times)),
                "per_rank_lobpcg_iterations": [p[3] for p in parts], "all_bands_converged": all(p[4] for p in parts),
                "config": {"workload": "HEX a=c=1 dielectric sphere, ND order 2, n_sub=8 (N=%d), %d rows / %d unique k-points, "
                                       "%d bands, tol %g" % (eqs[0].N, len(rows), len(uk), args.bands, args.tol),
                           "handles_per_gpu": T, "k_points_batched_per_handle": B}, "disp_dat": os.path.relpath(path, ROOT)}
        _emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def apply_study(args):
    """Y = A X throughput (GDOF/s) and fractions of the HBM (32 B/DOF/vector) and fp64 rooflines."""
    import torch
    import mfem_bravais_b200 as m
    out, hbm, how, fp64 = apply_roofline(m, torch, 0, [("CUB", 1, 48), ("FCC", 2, 16), ("BCC", 3, 12)], [1, 4, 10, 16, 30])
    _emit(json.dumps({"apply_study": out, "hbm_peak_gbs": hbm, "hbm_peak_source": how, "fp64_peak_tflops": fp64,
                      "note": "timing = clearing of y + apply kernel, CUDA events, L2 flushed between launches"}))


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
    elif a.sweep == "hex":
        hex_sweep(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
