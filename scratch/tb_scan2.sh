for tb in "2 10 20" "4 5 20" "3 8 24" "2 12 24" "1 16 16" "4 6 24" "2 10 20"; do set -- $tb
timeout 200 python bench.py --streams $1 --batch $2 --steps $3 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T=$1 B=$2 steps=$3', round(d['value'],2), round(d['e2e']['value'],2), d['lobpcg_iterations_mean'], d['validated'], d['impl_config']['rounds'], d['impl_config']['padded_solves'])"
done
