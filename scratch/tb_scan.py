import json, os, subprocess, sys
for T, B, extra in [(2, 5, {}), (2, 10, {}), (4, 5, {}), (1, 10, {}), (1, 5, {}), (1, 16, {}), (3, 5, {}), (2, 5, {"BLOCH_MG_SMOOTH_DEGREE": "1"}), (2, 10, {"BLOCH_MG_SMOOTH_DEGREE": "1"}), (4, 5, {"BLOCH_MG_SMOOTH_DEGREE": "1"})]:
    env = dict(os.environ); env.update(extra)
    p = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-roofline", "--no-n16", "--streams", str(T), "--batch", str(B)],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print("T %d B %2d %-36s value %6.2f e2e %6.2f its %.1f padded %d rounds %d" % (T, B, extra, d["value"], d["e2e"]["value"],
              d["lobpcg_iterations_mean"], d["impl_config"]["padded_solves"], d["impl_config"]["rounds"]), flush=True)
    except Exception as e:
        print(T, B, "FAILED", p.stderr[-300:], flush=True)
