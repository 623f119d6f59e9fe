BLOCH_CG_DEBUG=1 python scratch/proj_time.py 2>&1 | grep -v "^\[proj_cg\]" ; BLOCH_CG_DEBUG=1 python scratch/proj_time.py 2>&1 | grep "^\[proj_cg\]" | head -2
python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k "projector" 2>&1 | tail -2
