# round 2, call B: GPU tests incl. the k-point batch, batch-size / stream timing
set -x
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q > gpurun_out/t_r2b_batch.log 2>&1; tail -15 gpurun_out/t_r2b_batch.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2b.log 2>&1; tail -5 gpurun_out/t_r2b.log
timeout 900 python scratch/batch_time.py 8 32 > gpurun_out/batch_time_n8.log 2>&1; cat gpurun_out/batch_time_n8.log | tail -12
