"""Builds the tracked round-2 summaries under profiles/ from gpurun_out/ (scratch)."""
import collections, csv, json, os, re, shutil, subprocess
os.makedirs("profiles", exist_ok=True)
G = "gpurun_out/"
# ---- launch list of the batched bench ----
lines = [l for l in open(G + "launches_r2.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("bloch_b200::", "").replace("<unnamed>::", "")
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    key = (name, row["Grid Size"].replace(" ", ""))
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu launch list, round 2\n",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 4000 --csv python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16`",
       "(FCC order 2, n_sub 8, N = 49152, ONE handle iterating 10 k-points together = 160-column block vectors; window inside the timed sweep;",
       "per-launch times are cold-cache and serialised: compare SHARES.  final round-2 code: auxiliary-space preconditioner, sum-factorised transfers.  The window's launches = %.1f ms of kernel time.)\n" % (tot / 1e3),
       "| kernel | grid | launches | total us | avg us | share |\n|---|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %s | %d | %.1f | %.2f | %.1f %% |" % (k[0][:70], k[1], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
byname = collections.defaultdict(float)
for k, v in agg.items():
    byname[re.sub(r"<.*", "", k[0])] += v[1]
out.append("\nBy kernel family: " + ", ".join("`%s` %.1f %%" % (k, 100 * v / tot) for k, v in sorted(byname.items(), key=lambda kv: -kv[1])[:12]))
open("profiles/launches_r2.md", "w").write("\n".join(out) + "\n")
json.dump({"%s %s" % k: {"launches": v[0], "total_us": v[1], "share": v[1] / tot} for k, v in agg.items()},
          open("profiles/kernel_shares_r2.json", "w"), indent=1)

# ---- ncu --set full captures of the apply kernels: traffic + table ----
def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(r.splitlines()))
    return dict(zip(rows[0], zip(rows[2], rows[1])))
def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
traffic, md = {}, ["# ncu --set full captures of the ND operator-apply kernels, round 2\n",
                   "`ncu --set full --clock-control none --import-source on -k regex:k_nd_(item|comp) -s 6 -c 1 python scratch/apply_one.py <lattice> <p> <n_sub> <vectors>`",
                   "(kernel only; the clearing of y that precedes it - one cudaMemset2D, 16 B per dof and vector - is a separate node)\n"]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
for tag, desc, N, nv in [("bcc_p3_n12_v10", "k_nd_comp<3>: BCC order 3 n_sub 12 (configs[2], the bench line's roofline kernel), N = 2239488, 10 vectors", 2239488, 10),
                         ("fcc_p2_n8_v160", "k_nd_item<2>: FCC order 2 n_sub 8, N = 49152, 160 vectors (the batched solver block: 10 k-points x 16 columns)", 49152, 160),
                         ("fcc_p2_n16_v16", "k_nd_item<2>: FCC order 2 n_sub 16, N = 393216, 16 vectors", 393216, 16)]:
    rep = G + "prof_nd_%s_r2.ncu-rep" % tag
    if not os.path.exists(rep):
        continue
    m = raw(rep)
    md.append("## %s\n\n| metric | value | unit |\n|---|---|---|" % desc)
    for k in want:
        if k in m:
            md.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
    tr = tobytes(*m["dram__bytes_read.sum"]) + tobytes(*m["dram__bytes_write.sum"])
    alg = 32.0 * N * nv
    traffic[tag] = {"dram_bytes_per_launch": tr, "algorithmic_bytes_per_launch": alg, "ratio": tr / alg, "N": N, "vectors": nv,
                    "source": "profiles/ncu_nd_apply_r2.md (ncu --set full, kernel only; + 16 B x N x vectors written by the clearing of y)"}
    md.append("\nDRAM traffic %.1f MB per launch vs algorithmic %.1f MB (32 B x N x %d) -> ratio %.2f (round 1: RED for every dof)\n" % (tr / 1e6, alg / 1e6, nv, tr / alg))
open("profiles/ncu_nd_apply_r2.md", "w").write("\n".join(md) + "\n")
json.dump(traffic, open("profiles/traffic_r2.json", "w"), indent=1)
for f in ["bench_r2.json", "apply_study_r2.json", "bench_reference_r2.json", "configs_r2.json", "hex_sweep_1gpu_r2.json", "bench_2gpu_r2.json",
          "bench_4gpu_r2.json", "bench_8gpu_r2.json", "hex_n1.json", "hex_n2.json", "hex_n4.json", "hex_n8.json", "tb_scan.log", "ab_env_2_5.log"]:
    if os.path.exists(G + f):
        shutil.copy(G + f, "profiles/" + (f if "r2" in f else f.replace(".json", "_r2.json").replace(".log", "_r2.log")))
# ---- ncu --set full of the kernels added with the auxiliary-space preconditioner ----
want2 = want + ["launch__grid_size", "launch__block_size", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
md = ["# ncu --set full captures: sum-factorised multigrid transfers and the order-3 lane-pair S0 kernel, round 2\n",
      "`NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> -c <n> python scratch/solve_profile.py BCC 8 3`",
      "(third, warm solve of BCC order 3 n_sub 8, N = 663552, 16 columns; the FCC capture: `scratch/batch_profile.py 8 10`, 160 columns).",
      "Of the captured launches the one with the largest grid (the fine level) is tabulated.\n"]
import glob
def raw_all(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(r.splitlines()))
    return [dict(zip(rows[0], zip(row, rows[1]))) for row in rows[2:]]
for rep in sorted(glob.glob(G + "prof_k_h1_*_r2.ncu-rep")):
    ms = raw_all(rep)
    if not ms:
        continue
    m = max(ms, key=lambda d: float(d["launch__grid_size"][0].replace(",", "")))
    md.append("## %s  (`%s`)\n\n| metric | value | unit |\n|---|---|---|" % (m["Kernel Name"][0][:80].replace("void ", "").replace("<unnamed>::", ""), os.path.basename(rep)))
    for k in want2:
        if k in m:
            md.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
    md.append("")
open("profiles/ncu_mg_aux_r2.md", "w").write("\n".join(md) + "\n")
for src, dst in [("aux_cmp_final.json", "aux_cmp_r2.json"), ("small_kappa_r2.log", "small_kappa_r2.log"), ("ab_env_aux_r2.log", "ab_env_aux_r2.log"),
                 ("ab_env_mgdeg_r2.log", "ab_env_mgdeg_r2.log")]:
    if os.path.exists(G + src):
        shutil.copy(G + src, "profiles/" + dst)
print(open("profiles/launches_r2.md").read()[:1800])
print(json.dumps(traffic, indent=1))
