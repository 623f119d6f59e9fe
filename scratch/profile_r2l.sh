set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2l.log 2>&1; tail -4 gpurun_out/t_r2l.log
python bench.py --no-cpu-baseline --no-roofline > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; cut -c1-200 gpurun_out/bench_r2l.json
python bench.py --no-cpu-baseline --no-roofline --no-n16 --steps 32 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['impl_config'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gram2_basis|k_rr_update" --launch-skip 6 --launch-count 6 -f -o gpurun_out/prof_gram_r2 python scratch/batch_profile.py 8 10 > gpurun_out/ncu_gram_r2.log 2>&1
tail -1 gpurun_out/ncu_gram_r2.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()"
