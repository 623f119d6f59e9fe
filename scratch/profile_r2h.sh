set -x
timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_apply_fullsize.py tests/test_gpu_parity_baseline.py tests/test_gpu_batch.py tests/test_gpu_fields.py -x -q > gpurun_out/t_r2h.log 2>&1; tail -4 gpurun_out/t_r2h.log
BLOCH_ND_FRESH_Y=0 python bench.py --apply-study > gpurun_out/apply_study_fresh0.json 2>/dev/null
BLOCH_ND_FRESH_Y=1 python bench.py --apply-study > gpurun_out/apply_study_fresh1.json 2>/dev/null
python - <<'PY'
import json
a=json.load(open('gpurun_out/apply_study_fresh0.json'))['apply_study']; b=json.load(open('gpurun_out/apply_study_fresh1.json'))['apply_study']
for x,y in zip(a,b): print(x['lattice'],x['order'],x['n_sub'],x['vectors'],'RED-only %.1f GDOF/s  interior stores %.1f GDOF/s (%.3f of HBM roof, %.3f fp64)'%(x['gdofs'],y['gdofs'],y['hbm_frac'],y['fp64_frac']))
PY
python bench.py --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; cut -c1-300 gpurun_out/bench_r2h.json
