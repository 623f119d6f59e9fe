for mg in 1 0; do echo "BLOCH_MG=$mg"; BLOCH_MG=$mg BLOCH_VERBOSE=1 python scratch/proj_time16.py 16 2>&1 | grep -v "lobpcg\] it " | cut -c1-250; done
BLOCH_MG=1 BLOCH_VERBOSE=1 python scratch/proj_time16.py 8 2>&1 | grep -v "lobpcg\] it " | cut -c1-250
