for cfg in "2 2" "1 2" "2 1" "1 1"; do
  set -- $cfg
  export BLOCH_AUX_SMOOTH_DEGREE=$1 BLOCH_AUX_MG_DEGREE=$2
  echo "=== ND smoother degree $1, aux multigrid smoother degree $2"
  AUX_CMP_MODES=aux timeout 300 python scratch/aux_cmp.py t$1$2 bcc8 fcc16 hex8 2>&1 | cut -c1-260
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['lobpcg_iterations_mean'], d['validated'])"
done
unset BLOCH_AUX_SMOOTH_DEGREE BLOCH_AUX_MG_DEGREE
AUX_CMP_MODES=aux timeout 300 python scratch/aux_cmp.py t12 bcc12 2>&1 | cut -c1-260
