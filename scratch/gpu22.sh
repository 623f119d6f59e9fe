python scratch/dbg_golden.py 2>&1 | cut -c1-120
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
