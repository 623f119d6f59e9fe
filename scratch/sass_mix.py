"""Static SASS opcode mix and resource usage of the operator-apply kernels in libbloch_b200.so -> profiles/sass_mix_r1.md
(cuobjdump only; no GPU needed)."""
import collections, re, subprocess, sys

so = "mfem-bravais_b200/lib/libbloch_b200.so"
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
mix, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        mix[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        mix[cur][m.group(2)] += 1
rows = []
for fn, c in mix.items():
    if "k_nd_item" not in fn and "k_nd_comp" not in fn and "k_nd_apply" not in fn:
        continue
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()
    cut = dem.find(">(")
    name = (dem[:cut + 1] if cut > 0 else dem).replace("void ", "").replace("bloch_b200::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((int|bool)\)", "", name)
    tot = sum(c.values())
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    mem = c["LDS"] + c["STS"] + c["LDG"] + c["REDG"] + c["RED"] + c["LDGSTS"] + c["ATOMG"]
    rows.append((name, usage.get(fn, (0, 0, 0, 0)), tot, fp, c["LDS"], c["STS"], c["LDG"], c["REDG"] + c["RED"], c["SHFL"], c["LDL"] + c["STL"], c["BAR"]))
rows.sort()
with open("profiles/sass_mix_r1.md", "w") as f:
    f.write("# Static SASS mix of the operator-apply kernels (cuobjdump -sass / --dump-resource-usage of libbloch_b200.so)\n\n"
            "Counts are static instructions of the whole kernel (the lane-pair kernels are straight-line code per tile, so\n"
            "static = dynamic per tile up to the prologue; `k_nd_comp<3>` and `k_nd_apply` contain rolled loops).\n"
            "Template arguments: `k_nd_item<P, HAS_A, HAS_M, threads, gather variant>`, `k_nd_comp<P, HAS_A, HAS_M, threads, items per warp>`, `k_nd_apply<P, warps>`.\n\n"
            "| kernel | regs | stack B | instr | fp64 (DFMA+DMUL+DADD) | fp64 share | LDS | STS | LDG | RED | SHFL | LDL+STL | BAR |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for name, u, tot, fp, lds, sts, ldg, red, shfl, ll, bar in rows:
        f.write("| `%s` | %d | %d | %d | %d | %.0f %% | %d | %d | %d | %d | %d | %d | %d |\n" % (name, u[0], u[1], tot, fp, 100.0 * fp / max(tot, 1), lds, sts, ldg, red, shfl, ll, bar))
print(open("profiles/sass_mix_r1.md").read()[:3000])
