set -x
timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_batch.py tests/test_gpu_eigen.py tests/test_gpu_api.py -x -q > gpurun_out/t_r2n.log 2>&1; tail -4 gpurun_out/t_r2n.log
for v in 1 0 1 0; do BLOCH_MG_SM2=$v python bench.py --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | cut -c1-160; done
BLOCH_MG_SM2=1 python scratch/regress.py maxwell 2>&1 | grep -E "its; ms|setup"
BLOCH_MG_SM2=0 python scratch/regress.py maxwell 2>&1 | grep -E "its; ms|setup"
