echo "cores: $(nproc)"
for cfg in "0-15 4 5" "0-3 4 5" "0-1 4 5" "0-3 2 10" "0-1 2 10"; do set -- $cfg
taskset -c $1 timeout 200 python bench.py --streams $2 --batch $3 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cpus=$1 T=$2 B=$3', round(d['value'],2), round(d['e2e']['value'],2), d['lobpcg_iterations_mean'], d['validated'])"
done
for tb in "2 10" "4 5"; do set -- $tb
timeout 200 python bench.py --sweep hex --streams $1 --batch $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hex T=$1 B=$2', d['value'], d['k_points_per_s'])"
done
