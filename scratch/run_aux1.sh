set -x
timeout 900 python -m pytest tests/test_gpu_aux.py -x -q > gpurun_out/aux_tests.log 2>&1; tail -15 gpurun_out/aux_tests.log
timeout 600 python scratch/aux_cmp.py a fcc8 bcc4 bcc8 fcc16 cub16 hex8 > gpurun_out/aux_cmp_a.log 2>&1; cat gpurun_out/aux_cmp_a.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err; tail -c 1500 gpurun_out/bench_aux.json; tail -5 gpurun_out/bench_aux.err
BLOCH_PRECOND=cheb timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/bench_cheb.json 2> gpurun_out/bench_cheb.err; tail -c 600 gpurun_out/bench_cheb.json
