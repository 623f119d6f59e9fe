"""Summarises an `ncu --set full` report with many kernels: one row per (kernel, grid) with the metrics that decide
which roof a kernel sits under.  usage: ncu_table.py report.ncu-rep [title] > profiles/xyz.md"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = {"us": "gpu__time_duration.sum", "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "regs": "launch__registers_per_thread", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "long_sb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
     "math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
     "lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
     "barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "short_sb": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "atom": "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed"}
def num(r, key, scale_bytes=False):
    i = col.get(M[key])
    if i is None or r[i] == "":
        return float("nan")
    v = float(r[i].replace(",", ""))
    u = units[i]
    if scale_bytes:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if key == "us":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    return v
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("bloch_b200::", "").replace("<unnamed>::", "")
    key = (name, r[col["Grid Size"]], r[col["Block Size"]])
    agg.setdefault(key, []).append(r)
print("# %s\n" % title)
print("Per (kernel, grid): mean over the captured launches.  fp64 / issue / warps: % of peak while active; DRAM / L2 / L2-atomic: % of peak over the kernel;")
print("stalls: warps stalled per issued instruction (the two largest are named).\n")
print("| kernel | grid x block | n | us | fp64 pipe % | issue % | DRAM % | L2 % | L2 atomic % | warps % | regs | DRAM MB (rd+wr) | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (name, grid, block), rs in agg.items():
    mean = lambda k, sb=False: sum(num(r, k, sb) for r in rs) / len(rs)
    stalls = sorted(((mean(k), k) for k in ("long_sb", "mio", "math", "lg", "barrier", "short_sb", "wait")), reverse=True)[:2]
    print("| `%s` | %s x %s | %d | %.1f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %d | %.1f | %s |" % (
        name[:60], grid.replace(" ", ""), block.replace(" ", ""), len(rs), mean("us"), mean("fp64"), mean("issue"), mean("dram"), mean("l2"),
        mean("atom"), mean("warps"), int(mean("regs")), (mean("rd", True) + mean("wr", True)) / 1e6,
        ", ".join("%s %.1f" % (k, v) for v, k in stalls)))
