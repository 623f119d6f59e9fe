timeout 800 python scratch/configs.py 2>&1 | cut -c1-400
