timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_apply_fullsize.py tests/test_gpu_parity_baseline.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -4
for v in 0 1; do echo "PARTIAL_CLEAR=$v"; BLOCH_ND_PARTIAL_CLEAR=$v timeout 300 python bench.py --apply-study 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['apply_study']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('lattice','order','n_sub','N','nvec','vectors','gdofs','ms','frac_hbm','hbm_frac','us')})
"; done
