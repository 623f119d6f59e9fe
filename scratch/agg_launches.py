import csv, collections, re, sys
fn = sys.argv[1]
with open(fn) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('bloch_b200::', '').replace('<unnamed>::', '').replace('void ', '')
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    k = (name[:40], row['Grid Size'])
    agg[k][0] += 1; agg[k][1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-42s %-14s n=%4d avg %7.2f us  share %5.1f%%" % (k[0], k[1], a[0], a[1] / a[0], 100 * a[1] / tot))
print("total us", tot)
