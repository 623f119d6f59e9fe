set -x
timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_batch.py tests/test_gpu_eigen.py -x -q > gpurun_out/t_r2k.log 2>&1; tail -4 gpurun_out/t_r2k.log
for v in 0 1 2; do BLOCH_H1_PAIR=$v python bench.py --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | cut -c1-160; done
BLOCH_H1_PAIR=1 python scratch/batch_profile.py 8 10 | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gram2_basis|k_rr_update|k_h1_s0_pair" --launch-skip 30 --launch-count 10 -f -o gpurun_out/prof_gram_r2 python scratch/batch_profile.py 8 10 > gpurun_out/ncu_gram_r2.log 2>&1
tail -1 gpurun_out/ncu_gram_r2.log | cut -c1-200
