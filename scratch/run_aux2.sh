set -x
python scratch/batch_profile.py 8 10 > gpurun_out/bp_aux.log 2>&1; cat gpurun_out/bp_aux.log
BLOCH_PRECOND=cheb python scratch/batch_profile.py 8 10 > gpurun_out/bp_cheb.log 2>&1; cat gpurun_out/bp_cheb.log
NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_aux.csv python scratch/batch_profile.py 8 10 > gpurun_out/ncu_aux.log 2>&1
python scratch/agg_launches.py gpurun_out/launches_aux.csv 40
