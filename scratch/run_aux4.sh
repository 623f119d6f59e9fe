python scratch/solve_profile.py BCC 8 3
NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bcc8.csv python scratch/solve_profile.py BCC 8 3 > gpurun_out/ncu_bcc8.log 2>&1
python scratch/agg_launches.py gpurun_out/launches_bcc8.csv 30
