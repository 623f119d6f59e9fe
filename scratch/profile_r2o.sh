set -x
timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_apply_fullsize.py tests/test_gpu_parity_baseline.py -x -q > gpurun_out/t_r2o.log 2>&1; tail -4 gpurun_out/t_r2o.log
python bench.py --apply-study > gpurun_out/apply_study_pair.json 2>/dev/null
python - <<'PY'
import json
for e in json.load(open('gpurun_out/apply_study_pair.json'))['apply_study']:
    if e['order'] == 3: print(e['lattice'], e['order'], e['vectors'], '%.1f GDOF/s  hbm %.3f fp64 %.3f' % (e['gdofs'], e['hbm_frac'], e['fp64_frac']))
PY
ncu --set full --clock-control none --import-source on -k regex:k_nd_comp -s 6 -c 1 -f -o gpurun_out/prof_nd_bcc_p3_n12_v10_r2 python scratch/apply_one.py BCC 3 12 10 > gpurun_out/ncu_a.log 2>&1
python scratch/regress.py maxwell 2>&1 | grep -E "its; ms|setup"
