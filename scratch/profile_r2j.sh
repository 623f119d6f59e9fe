set -x
for w in maxwell scalar; do
  python scratch/regress.py $w 2>&1 | grep -E "its; ms|setup" 
  BLOCH_CHEB_THREE_TERM=0 python scratch/regress.py $w 2>&1 | grep -E "its; ms|setup"
  BLOCH_MG_CSR_TRANSFER=0 python scratch/regress.py $w 2>&1 | grep -E "its; ms|setup"
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 8000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/ncu_launch_r2.log 2>&1
tail -2 gpurun_out/ncu_launch_r2.log | cut -c1-200; wc -l gpurun_out/launches_r2.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gram2_basis|k_rr_update|k_resid_norm|k_cheb3" --launch-skip 40 --launch-count 8 -f -o gpurun_out/prof_gram_r2 python scratch/batch_profile.py 8 10 > gpurun_out/ncu_gram_r2.log 2>&1
tail -1 gpurun_out/ncu_gram_r2.log | cut -c1-200
