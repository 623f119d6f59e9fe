python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for t in 1 2 4 6 8; do python bench.py --no-cpu-baseline --n-sub 8 --steps 32 --streams $t 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('streams $t: value %.2f e2e %.2f its %.2f'%(d['value'],d['e2e']['value'],d['lobpcg_iterations_mean']))
    elif 'rror' in l: print(l.strip()[:200])
"; done
