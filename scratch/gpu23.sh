python scratch/sanitize.py > gpurun_out/san_plain.log 2>&1 && tail -2 gpurun_out/san_plain.log && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/sanitize.py > gpurun_out/memcheck_r1.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck_r1.log | head -10; tail -3 gpurun_out/memcheck_r1.log
