"""The reference arm of bench.py runs on the CPU (no GPU needed): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pair-scores/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "BASELINE.json configs[1]" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    # every host core is used even when the launcher exports OMP_NUM_THREADS=1 (torch.distributed.run does)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) == d["cpu_baseline"]["threads"]
    assert d["cpu_baseline"]["extrapolated"] is True and "extrapolated" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode != 0 and "no CPU path" in (res.stderr + res.stdout)
