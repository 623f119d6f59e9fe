python scratch/dbg_golden.py 2>&1 | cut -c1-150
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --no-cpu-baseline --steps 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'e2e',d['e2e']['value'],'its',d['lobpcg_iterations_mean'],'launches',d['gpu_launches'])
"
