export BLOCH_TWO_PASS=0
python scratch/apply_once.py FCC 2 16 10 > gpurun_out/plain_apply2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_nd_apply -s 3 -c 1 -o gpurun_out/prof_nd_apply_fcc_p2_v2 python scratch/apply_once.py FCC 2 16 10 > gpurun_out/ncu_apply2.log 2>&1
tail -2 gpurun_out/ncu_apply2.log
