# round 2, call A: GPU tests after the round-1 ADVICE fixes + ncu --set full captures of the solver's non-ND kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2a.log 2>&1; tail -3 gpurun_out/t_r2a.log
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_gram|k_rr_update|k_h1_dense|k_h1_s0_item|k_h1_reduce|k_cheb_step_z|k_h1_restrict|k_h1_prolong_rep|k_pcg_a|k_resid_norm" \
  --launch-skip 1500 --launch-count 40 -f -o gpurun_out/prof_solver_r2 \
  python bench.py --steps 2 --warmup 2 --streams 1 --no-cpu-baseline > gpurun_out/ncu_solver_r2.log 2>&1
tail -2 gpurun_out/ncu_solver_r2.log | cut -c1-300
ls -la gpurun_out/prof_solver_r2.ncu-rep
