set -x
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_eigen.py -x -q > gpurun_out/t_r2d.log 2>&1; tail -5 gpurun_out/t_r2d.log
python bench.py --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 1500 gpurun_out/bench_r2d.json; tail -5 gpurun_out/bench_r2d.err
python scratch/ab_env.py 2 5 20 > gpurun_out/ab_env_2_5.log 2>&1; cat gpurun_out/ab_env_2_5.log
