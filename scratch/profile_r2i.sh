# multi-GPU: weak-scaling bench line at N ranks, HEX 256-point sweep (configs[3]) strong scaling, reference arm under torchrun
set -x
NG=${NG:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 "${@:2}"; }
timeout 600 python -m pytest tests/test_gpu_driver.py -x -q > gpurun_out/t_r2i.log 2>&1; tail -3 gpurun_out/t_r2i.log
python bench.py --sweep hex --gpus 1 > gpurun_out/hex_n1.json 2> gpurun_out/hex_n1.err; cut -c1-500 gpurun_out/hex_n1.json
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    run $n --sweep hex > gpurun_out/hex_n$n.json 2> gpurun_out/hex_n$n.err; cut -c1-500 gpurun_out/hex_n$n.json
    run $n --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-n16 > gpurun_out/bench_${n}gpu_r2.json 2> gpurun_out/bench_${n}gpu_r2.err; cut -c1-260 gpurun_out/bench_${n}gpu_r2.json
  fi
done
run 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_torchrun_r2.json 2> gpurun_out/bench_reference_torchrun_r2.err; cut -c1-200 gpurun_out/bench_reference_torchrun_r2.json
python - <<'PY'
import numpy as np, glob
a = np.load('gpurun_out/disp_hex_n1.npy')
for f in sorted(glob.glob('gpurun_out/disp_hex_n[248].npy')):
    b = np.load(f); print(f, 'max rel dev from the single-GPU sweep: %.2e' % np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-3)))
PY
