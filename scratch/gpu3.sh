set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py > gpurun_out/bench_r1_first.json 2> gpurun_out/bench_r1_first.err; tail -c 3000 gpurun_out/bench_r1_first.json; tail -5 gpurun_out/bench_r1_first.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
