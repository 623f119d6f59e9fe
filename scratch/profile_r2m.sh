set -x
timeout 900 python -m pytest tests/test_gpu_rr_device.py -x -q > gpurun_out/t_r2m.log 2>&1; tail -15 gpurun_out/t_r2m.log
for v in 1 0; do BLOCH_RR_DEVICE=$v python bench.py --no-cpu-baseline --no-roofline --no-n16 2>gpurun_out/bench_rr$v.err | cut -c1-200; tail -2 gpurun_out/bench_rr$v.err; done
for v in 1 0; do BLOCH_RR_DEVICE=$v python bench.py --no-cpu-baseline --no-roofline --no-n16 --streams 1 2>/dev/null | cut -c1-160; done
BLOCH_RR_DEVICE=1 python scratch/batch_profile.py 8 10 | head -12
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2m_all.log 2>&1; tail -4 gpurun_out/t_r2m_all.log
