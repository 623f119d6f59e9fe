python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --no-cpu-baseline --n-sub 16 --steps 8 2>&1 | tail -c 1500
