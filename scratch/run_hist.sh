timeout 900 python -m pytest tests/test_gpu_eigen.py tests/test_gpu_batch.py tests/test_gpu_reduced_basis.py tests/test_gpu_multilevel.py tests/test_gpu_driver.py -x -q -m gpu 2>&1 | tail -4
for v in 0 1; do echo "BLOCH_HISTORY=$v"; BLOCH_HISTORY=$v timeout 300 python bench.py --sweep hex 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['k_points_per_s'], d['per_rank_lobpcg_iterations'], d['all_bands_converged'])"; cp gpurun_out/disp_hex_n1.npy gpurun_out/disp_hex_hist$v.npy; done
python -c "
import numpy as np
a=np.load('gpurun_out/disp_hex_hist0.npy'); b=np.load('gpurun_out/disp_hex_hist1.npy'); print('max rel diff', np.max(np.abs(a-b)/np.maximum(1.0,np.abs(a))))"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], d['lobpcg_iterations_mean'], d['validated'], (d.get('n_sub16') or {}).get('value'))"
