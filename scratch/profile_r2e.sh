set -x
python scratch/tb_scan.py > gpurun_out/tb_scan.log 2>&1; cat gpurun_out/tb_scan.log
BLOCH_MG_SMOOTH_DEGREE=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2e_deg1.log 2>&1; tail -4 gpurun_out/t_r2e_deg1.log
