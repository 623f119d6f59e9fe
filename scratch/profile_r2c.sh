set -x
python scratch/batch_profile.py 8 8 > gpurun_out/bprof_n8_k8.log 2>&1; cat gpurun_out/bprof_n8_k8.log
python scratch/batch_profile.py 8 1 > gpurun_out/bprof_n8_k1.log 2>&1; cat gpurun_out/bprof_n8_k1.log
python scratch/batch_profile.py 16 4 > gpurun_out/bprof_n16_k4.log 2>&1; cat gpurun_out/bprof_n16_k4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 6000 --csv --log-file gpurun_out/launches_b8.csv python scratch/batch_profile.py 8 8 > gpurun_out/ncu_launch_b8.log 2>&1
tail -3 gpurun_out/ncu_launch_b8.log
