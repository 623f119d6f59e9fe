python bench.py --apply-study 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for r in d['apply_study']:
    if r['vectors']>=10: print(r['lattice'],r['order'],r['N'],r['vectors'],'%.1f us %.1f GDOF/s hbm %.3f fp64 %.3f'%(r['median_us'],r['gdofs'],r['hbm_frac'],r['fp64_frac']))
"
