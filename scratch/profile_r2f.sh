set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2f.log 2>&1; tail -6 gpurun_out/t_r2f.log
python bench.py --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; cut -c1-400 gpurun_out/bench_r2f.json; tail -3 gpurun_out/bench_r2f.err
BLOCH_MG_CSR_TRANSFER=0 python bench.py --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | cut -c1-200
BLOCH_CHEB_THREE_TERM=0 python bench.py --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_h1_s0_item|k_cheb3|k_nd_item|k_csr_apply|k_gram2_basis|k_rr_update|k_h1_op|k_cheb_step|k_resid_norm|k_col_dot" \
  --launch-skip 4000 --launch-count 60 -f -o gpurun_out/prof_batch_r2 \
  python scratch/batch_profile.py 8 10 > gpurun_out/ncu_batch_r2.log 2>&1
tail -2 gpurun_out/ncu_batch_r2.log | cut -c1-200
