# regenerates the measured artefacts behind profiles/*_r2.* (run under gpurun; summarise with scratch/summarise_r2.py)
set -x
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 600 gpurun_out/bench_r2.json
python bench.py --apply-study > gpurun_out/apply_study_r2.json 2> gpurun_out/apply_study_r2.err
python bench.py --sweep hex > gpurun_out/hex_sweep_1gpu_r2.json 2> gpurun_out/hex_sweep_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 8000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/ncu_launch_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nd_comp -s 6 -c 1 -f -o gpurun_out/prof_nd_bcc_p3_n12_v10_r2 python scratch/apply_one.py BCC 3 12 10 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nd_item -s 6 -c 1 -f -o gpurun_out/prof_nd_fcc_p2_n8_v160_r2 python scratch/apply_one.py FCC 2 8 160 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nd_item -s 6 -c 1 -f -o gpurun_out/prof_nd_fcc_p2_n16_v16_r2 python scratch/apply_one.py FCC 2 16 16 > gpurun_out/ncu_c.log 2>&1
python scratch/configs.py r2 > gpurun_out/configs_r2.log 2>&1; tail -3 gpurun_out/configs_r2.log | cut -c1-200
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/bench_reference_r2.json 2> gpurun_out/bench_reference_r2.err; cut -c1-200 gpurun_out/bench_reference_r2.json
ls -la gpurun_out | head -30
