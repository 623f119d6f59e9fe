export BLOCH_TWO_PASS=1
for cfg in "CUB 1 48 10" "FCC 2 16 10" "BCC 3 12 10"; do
python scratch/apply_once.py $cfg > gpurun_out/plain_tp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_nd_apply|k_nd_reduce" -s 6 -c 4 --csv --log-file gpurun_out/tp.csv python scratch/apply_once.py $cfg > /dev/null 2>&1
echo "== $cfg"; grep -v "^==" gpurun_out/tp.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:40], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
" | sort | uniq | head -12
done
