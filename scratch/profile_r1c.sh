# regenerates the measured artefacts behind profiles/ (run under gpurun; summarise with scratch/summarise_profiles.py r1)
set -x
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 300 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_r1.json 2>> gpurun_out/bench_r1.err
python bench.py --apply-study > gpurun_out/apply_study_r1.json 2> gpurun_out/apply_study.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 4 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nd_item -s 6 -c 1 -f -o gpurun_out/prof_nd_fcc_p2_n8_r1 python scratch/apply_one.py FCC 2 8 10 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nd_item -s 6 -c 1 -f -o gpurun_out/prof_nd_fcc_p2_n16_r1 python scratch/apply_one.py FCC 2 16 16 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nd_comp -s 6 -c 1 -f -o gpurun_out/prof_nd_bcc_p3_n12_r1 python scratch/apply_one.py BCC 3 12 10 > gpurun_out/ncu_c.log 2>&1
python scratch/configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log | cut -c1-200
