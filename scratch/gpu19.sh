python scratch/sweeps.py
