"""Builds the tracked summaries under profiles/ from gpurun_out/ (scratch)."""
import csv, collections, json, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs("profiles", exist_ok=True)
out = []
# ---- launch list ----
lines = [l for l in open("gpurun_out/launches_%s.csv" % R) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("bloch_b200::", "").replace("<unnamed>::", "")
    try:
        v = float(row["Metric Value"])
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out.append("# ncu launch list, round %s\n" % R)
out.append("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline`")
out.append("(FCC order 2, n_sub 8, N = 49152, 16-column block, window inside the warm-started lifted solves; per-launch times are cold-cache and serialised: compare SHARES)\n")
out.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (k[:70], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
out.append("\ntotal %.1f us over %d launches\n" % (tot, sum(v[0] for v in agg.values())))
open("profiles/launches_%s.md" % R, "w").write("\n".join(out))
json.dump({k: {"launches": v[0], "total_us": v[1], "share": v[1] / tot} for k, v in agg.items()},
          open("profiles/kernel_shares_%s.json" % R, "w"), indent=1)

# ---- full captures ----
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
md = ["# ncu --set full captures of the ND operator apply kernels, round %s\n" % R,
      "`ncu --set full --clock-control none --import-source on -k regex:k_nd_(item|comp) -s 6 -c 1 python scratch/apply_one.py <lattice> <p> <n_sub> <vectors>`\n",
      "k_nd_item = lane pair per item (orders 1-2, nd_item.cu), k_nd_comp = six lanes per item (order 3, nd_comp.cu); the superseded "
      "cooperative-tile kernel k_nd_apply (kernels.cu) measured 26.3 us / 865 us on the first / third workload below.\n"]
traffic = {}
for tag, desc, N, nvec in [("fcc_p2_n8", "k_nd_item: FCC order 2 n_sub 8 (bench workload), N = 49152, 10 vectors", 49152, 10),
                           ("fcc_p2_n16", "k_nd_item: FCC order 2 n_sub 16, N = 393216, 16 vectors (solver block)", 393216, 16),
                           ("bcc_p3_n12", "k_nd_comp: BCC order 3 n_sub 12 (roofline study), N = 2239488, 10 vectors", 2239488, 10)]:
    rep = "gpurun_out/prof_nd_%s_%s.ncu-rep" % (tag, R)
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    md.append("## %s\n\n| metric | value | unit |\n|---|---|---|" % desc)
    vals = {}
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            md.append("| %s | %s | %s |" % (k, r[i], units[i]))
            vals[k] = (r[i], units[i])
    def tobytes(k):
        v, u = vals[k]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tr = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    alg = 32.0 * N * nvec
    traffic[tag] = {"dram_bytes_per_launch": tr, "algorithmic_bytes_per_launch": alg, "ratio": tr / alg, "N": N, "vectors": nvec}
    md.append("\nDRAM traffic %.1f MB per launch vs algorithmic %.1f MB (32 B x N x %d) -> ratio %.2f\n" % (tr / 1e6, alg / 1e6, nvec, tr / alg))
open("profiles/ncu_nd_apply_%s.md" % R, "w").write("\n".join(md))
json.dump(traffic, open("profiles/traffic_%s.json" % R, "w"), indent=1)
for f in ["bench_%s.json" % R, "apply_study_%s.json" % R]:
    if os.path.exists("gpurun_out/" + f):
        open("profiles/" + f, "w").write(open("gpurun_out/" + f).read())
print(open("profiles/launches_%s.md" % R).read()[:2500])
print(json.dumps(traffic, indent=1))
