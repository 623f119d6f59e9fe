# regenerates the measured artefacts of the final round-2 code (auxiliary-space preconditioner, sum-factorised transfers,
# order-3 lane-pair S0 kernel); summarise with scratch/summarise_r2.py
set -x
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 400 gpurun_out/bench_r2.json
python bench.py --sweep hex > gpurun_out/hex_sweep_1gpu_r2.json 2> gpurun_out/hex_sweep_r2.err; tail -c 300 gpurun_out/hex_sweep_1gpu_r2.json
python scratch/configs.py r2 > gpurun_out/configs_r2.log 2>&1; tail -6 gpurun_out/configs_r2.log | cut -c1-250
AUX_CMP_MODES=aux,cheb python scratch/aux_cmp.py final cub16 fcc8 fcc16 bcc8 bcc12 hex8 > gpurun_out/aux_cmp_final.log 2>&1; cut -c1-200 gpurun_out/aux_cmp_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/ncu_launch_r2.log 2>&1
python scratch/agg_launches.py gpurun_out/launches_r2.csv 12
for k in k_h1_restrict_sf k_h1_prolong_sf k_h1_s0_pair; do
  NCU_RANGE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -s 12 -c 4 -f -o gpurun_out/prof_${k}_bcc_p3_n8_r2 python scratch/solve_profile.py BCC 8 3 > gpurun_out/ncu_$k.log 2>&1
done
NCU_RANGE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_h1_restrict_sf -c 1 -f -o gpurun_out/prof_k_h1_restrict_sf_fcc_p2_n8_r2 python scratch/batch_profile.py 8 10 > gpurun_out/ncu_rsf2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
