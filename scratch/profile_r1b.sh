set -x
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 300 gpurun_out/bench_r1.json
python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 3000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scratch/configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log | cut -c1-200
