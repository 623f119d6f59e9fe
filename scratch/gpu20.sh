for g in 4 6 8; do BLOCH_GUARD=$g python bench.py --no-cpu-baseline --n-sub 8 --steps 32 --streams 1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('guard $g value',d['value'],'its',d['lobpcg_iterations_mean'])
    elif 'rror' in l: print(l.strip()[:200])
"; done
