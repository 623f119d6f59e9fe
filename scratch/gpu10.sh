set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --apply-study > gpurun_out/apply_study_r1.json 2> gpurun_out/apply_study.err; tail -3 gpurun_out/apply_study.err; python -c "
import json
d=json.load(open('gpurun_out/apply_study_r1.json'))
print('fp64 peak',d['fp64_peak_tflops'])
for r in d['apply_study']: print(r['lattice'],r['order'],r['N'],r['vectors'],'%.1f us %.1f GDOF/s hbm %.3f fp64 %.3f'%(r['median_us'],r['gdofs'],r['hbm_frac'],r['fp64_frac']))
"
./mfem-bravais_b200/lib/maxwell_dispersion_b200 -bl 2 -o 1 -pr 2 -np 1 -nb 6 -out gpurun_out && head -5 gpurun_out/disp.dat && cat gpurun_out/stats_0.out
