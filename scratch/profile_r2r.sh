set -x
for v in 2 1 2 1; do BLOCH_MG_SMOOTH_DEGREE=$v python bench.py --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deg $v', round(d['value'],2), round(d['e2e']['value'],2), d['lobpcg_iterations_mean'], 'n16', round(d['n_sub16']['value'],2), d['n_sub16']['lobpcg_iterations_mean'])"; done
BLOCH_MG_SMOOTH_DEGREE=1 python scratch/configs.py deg1 > gpurun_out/configs_deg1.log 2>&1; python -c "
import json
for r in json.load(open('gpurun_out/configs_deg1.json')): print(r['config'], 'solve %.2f s' % r['solve_s'], r['iterations'], r['converged'])"
