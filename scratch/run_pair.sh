timeout 600 python -m pytest tests/test_gpu_scalar.py tests/test_gpu_aux.py tests/test_gpu_apply.py -x -q -m gpu 2>&1 | tail -5
for v in 0 1; do echo "BLOCH_H1_PAIR=$v"; BLOCH_H1_PAIR=$v python scratch/solve_profile.py BCC 8 3 | head -4; done
BLOCH_H1_PAIR=1 AUX_CMP_MODES=aux timeout 300 python scratch/aux_cmp.py pair bcc12 2>&1 | cut -c1-260
