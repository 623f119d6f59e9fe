for b in 0 1; do BLOCH_BLOCKING_SYNC=$b python bench.py --no-cpu-baseline --n-sub 8 --steps 32 --streams 4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('blocking $b: value %.2f e2e %.2f its %.2f'%(d['value'],d['e2e']['value'],d['lobpcg_iterations_mean']))
    elif 'rror' in l: print(l.strip()[:200])
"; done
nproc
