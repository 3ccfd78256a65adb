"""Runs bench.py once per plan configuration (GT_TILE_LEAVES / GT_SEG_POSITIONS / GT_ROWS_PER_CTA) and prints the
kernel times, to pick the defaults.  Usage: python tools/sweep.py "T,Q,R" "T,Q,R" ... [-- extra bench args]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
extra = []
if "--" in args:
    k = args.index("--")
    args, extra = args[:k], args[k + 1:]
for spec in args:
    T, Q, R = spec.split(",")
    env = dict(os.environ, GT_TILE_LEAVES=T, GT_SEG_POSITIONS=Q, GT_ROWS_PER_CTA=R)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-sampler", "--e2e-steps", "2"] + extra,
                         env=env, capture_output=True, text=True)
    try:
        line = json.loads(out.stdout.strip().splitlines()[-1])
        print(spec, "value %.0f dist/s  ms/step %.4f  kernel_ms %s  roofline %.3f" % (
            line["value"], line["ms_per_step"], json.dumps(line["kernel_ms"]), line["roofline"]["frac"]), flush=True)
    except Exception:
        print(spec, "FAILED", out.stdout[-300:], out.stderr[-1500:], flush=True)
