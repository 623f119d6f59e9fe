set -x
# 1. plain bench (default workload)
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.json
# 2. launch list of a short bench (after the same command ran clean)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# 3. full capture of the roofline kernel at the bench size (FCC p2 n8, 10 vectors) and at the study size
python scratch/apply_once.py FCC 2 8 10 > gpurun_out/plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_nd_apply -s 3 -c 1 -o gpurun_out/prof_nd_apply_fcc_p2_n8_r1 python scratch/apply_once.py FCC 2 8 10 > gpurun_out/ncu_a.log 2>&1
python scratch/apply_once.py BCC 3 12 10 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_nd_apply -s 3 -c 1 -o gpurun_out/prof_nd_apply_bcc_p3_n12_r1 python scratch/apply_once.py BCC 3 12 10 > gpurun_out/ncu_b.log 2>&1
python bench.py --apply-study > gpurun_out/apply_study_r1.json 2>/dev/null
ls -la gpurun_out | tail -15
