set -x
timeout 1800 python -m pytest tests/test_gpu_fields.py tests/test_gpu_matrices.py tests/test_gpu_multilevel.py tests/test_gpu_parity_baseline.py tests/test_gpu_reduced_basis.py tests/test_gpu_scalar.py -x -q > gpurun_out/t_r2g.log 2>&1; tail -6 gpurun_out/t_r2g.log
nproc
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/bench_reference_r2.json 2> gpurun_out/bench_reference_r2.err; cut -c1-600 gpurun_out/bench_reference_r2.json; tail -4 gpurun_out/bench_reference_r2.err
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_h1_s0_item|k_cheb3|k_nd_item|k_csr_apply|k_gram2_basis|k_rr_update|k_h1_op" \
  --launch-skip 3000 --launch-count 16 -f -o gpurun_out/prof_batch_r2 \
  python scratch/batch_profile.py 8 10 > gpurun_out/ncu_batch_r2.log 2>&1
tail -2 gpurun_out/ncu_batch_r2.log | cut -c1-200; ls -la gpurun_out/
