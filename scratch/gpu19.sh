for n in 8 16; do python bench.py --no-cpu-baseline --n-sub $n --steps 32 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n_sub $n value',d['value'],'e2e',d['e2e']['value'],'its',d['lobpcg_iterations_mean'],'launches',d['gpu_launches'],'apply',d['apply_gdofs'])
"; done
