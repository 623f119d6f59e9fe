"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> profiles/launches_<tag>.md + kernel_shares_<tag>.json.
usage: python scratch/launch_shares.py gpurun_out/launches.csv r1 "<command line that produced it>" """
import collections, csv, json, re, sys

path, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*$", "", name).replace("bloch_b200::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((int|bool)\)", "", name)
    t = float(r[iv].replace(",", ""))
    t = t / 1000.0 if r[iu] in ("ns", "nsecond") else (t * 1000.0 if r[iu] in ("ms", "msecond") else t)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values()); n = sum(a[0] for a in agg.values())
items = sorted(agg.items(), key=lambda kv: -kv[1][1])
with open("profiles/launches_%s.md" % tag, "w") as f:
    f.write("# ncu launch list, round %s\n\nCommand: `%s`\n(per-launch times are cold-cache and serialised: compare SHARES)\n\n" % (tag, cmd))
    f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
    for k, (c, t) in items:
        f.write("| `%s` | %d | %.1f | %.2f | %.1f %% |\n" % (k, c, t, t / c, 100 * t / tot))
    f.write("\ntotal %.1f us over %d launches\n" % (tot, n))
json.dump({k: {"launches": c, "total_us": t, "share": t / tot} for k, (c, t) in items},
          open("profiles/kernel_shares_%s.json" % tag, "w"), indent=1)
print("total_us", tot, "launches", n)
for k, (c, t) in items[:12]:
    print("%-40s %5d %9.1f %6.2f %5.1f%%" % (k, c, t, t / c, 100 * t / tot))
