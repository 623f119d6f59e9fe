BLOCH_VERBOSE=1 python scratch/sweep2.py 8 | grep -v "lobpcg\] it" | grep -v "its; ms"
timeout 500 python scratch/sweep2.py 16
