set -x
timeout 600 python -m pytest tests/test_gpu_rr_device.py tests/test_gpu_batch.py tests/test_gpu_eigen.py -x -q > gpurun_out/t_r2q.log 2>&1; tail -4 gpurun_out/t_r2q.log
for i in 1 2; do python bench.py --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | cut -c1-140; done
for i in 1 2 3; do MALLOC_CHECK_=3 MALLOC_PERTURB_=165 python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/mc_$i.json 2> gpurun_out/mc_$i.err; echo rc=$?; cut -c1-120 gpurun_out/mc_$i.json; tail -2 gpurun_out/mc_$i.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 8000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/ncu_launch_r2.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_launch_r2.log | cut -c1-200; wc -l gpurun_out/launches_r2.csv
BLOCH_RR_DEVICE=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 2000 --csv --log-file gpurun_out/launches_r2_hostrr.csv python bench.py --steps 20 --warmup 5 --streams 1 --batch 10 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/ncu_launch_r2_hostrr.log 2>&1; echo rc=$?
tail -2 gpurun_out/ncu_launch_r2_hostrr.log | cut -c1-200
