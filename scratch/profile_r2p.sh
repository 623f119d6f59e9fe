set -x
timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_apply_fullsize.py tests/test_gpu_parity_baseline.py tests/test_gpu_batch.py -x -q > gpurun_out/t_r2p.log 2>&1; tail -4 gpurun_out/t_r2p.log
show() { python - "$1" <<'PY'
import json, sys
for e in json.load(open(sys.argv[1]))['apply_study']:
    if e['vectors'] in (10, 16, 30): print(e['lattice'], e['order'], e['vectors'], '%.1f GDOF/s  hbm %.3f fp64 %.3f' % (e['gdofs'], e['hbm_frac'], e['fp64_frac']))
PY
}
python bench.py --apply-study > gpurun_out/apply_study_pair2.json 2>/dev/null; show gpurun_out/apply_study_pair2.json
BLOCH_ND_COMP_WARPS=16 python bench.py --apply-study > gpurun_out/apply_study_w16.json 2>/dev/null; show gpurun_out/apply_study_w16.json | grep BCC
BLOCH_ND_COMP_ITEMS=4 python bench.py --apply-study > gpurun_out/apply_study_i4.json 2>/dev/null; show gpurun_out/apply_study_i4.json | grep BCC
BLOCH_ND_COMP_ITEMS=4 BLOCH_ND_COMP_WARPS=16 python bench.py --apply-study > gpurun_out/apply_study_i4w16.json 2>/dev/null; show gpurun_out/apply_study_i4w16.json | grep BCC
python bench.py --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | cut -c1-160
