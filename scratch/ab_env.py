"""A/B of solver environment switches on the bench workload: runs bench.py (headline legs only) per setting.
usage: ab_env.py [streams] [batch] [steps]"""
import json, os, subprocess, sys
T, B, K = (sys.argv[1:4] + ["2", "5", "20"])[:3] if len(sys.argv) >= 4 else ("2", "5", "20")
SETTINGS = [
    {},
    {"BLOCH_MG_SMOOTH_DEGREE": "1"},
    {"BLOCH_LIFT_PROJ_TOL": "0.25"},
    {"BLOCH_LIFT_X_TOL": "0.25"},
    {"BLOCH_LIFT_PROJ_TOL": "0.25", "BLOCH_LIFT_X_TOL": "0.25"},
    {"BLOCH_GUARD": "4"},
    {"BLOCH_LIFT_TAU": "4"},
    {"BLOCH_LIFT_TAU": "16"},
    {"BLOCH_AUX_SMOOTH_RATIO": "8"},
    {"BLOCH_AUX_SMOOTH_DEGREE": "3"},
    {"BLOCH_SIGMA_SCALE": "0.25"},
    {},
]
if os.environ.get("AB_SETTINGS"):
    SETTINGS = json.loads(os.environ["AB_SETTINGS"])
for st in SETTINGS:
    env = dict(os.environ); env.update(st)
    p = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-roofline", "--no-n16", "--streams", T, "--batch", B,
                        "--steps", K], env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        sh = d["roofline"] or {}
        print("%-70s value %6.2f e2e %6.2f its %.1f launches %d valid %s" % (st, d["value"], d["e2e"]["value"], d["lobpcg_iterations_mean"],
              d["gpu_launches"], d["validated"]), flush=True)
    except Exception as e:
        print(st, "FAILED", p.stderr[-400:], flush=True)
