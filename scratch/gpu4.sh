set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 1800 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
