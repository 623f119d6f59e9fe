timeout 900 python -m pytest tests/test_gpu_aux.py -x -q 2>&1 | tail -8
AUX_CMP_MODES=aux timeout 300 python scratch/aux_cmp.py sf bcc8 bcc12 2>&1 | cut -c1-260
python scratch/solve_profile.py BCC 8 3
for v in 0 1; do echo "FCC bench SF_TRANSFER=$v"; BLOCH_MG_SF_TRANSFER=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['lobpcg_iterations_mean'], d['validated'])"; done
