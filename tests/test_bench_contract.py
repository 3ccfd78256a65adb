"""bench.py's reference arm runs without a GPU: it must print ONE JSON line with the contract's keys, must not load the
product package, and must use every host thread whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--vocab", "3000", "--batch", "16", "--steps", "3", "--warmup", "1", env={"OMP_NUM_THREADS": "1"})
    assert d["impl"] == "reference" and d["unit"] == "distributions/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["product_package_loaded"] is False
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"]
    want = min(16, len(os.sched_getaffinity(0)))  # OpenMP over the 16 rows of a step, OMP_NUM_THREADS ignored
    assert cb["cores"] == want, cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("batch_weight_sum + batch_weight_max") and d["config"]["vocab"] == 3000


def test_reference_arm_other_ranks_and_workloads():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--vocab", "3000"], capture_output=True,
                         text=True, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""  # ranks other than 0 exit without work
    d = run_bench("--impl", "reference", "--workload", "smc4096")
    assert d["impl"] == "reference" and "unavailable" in d
