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


def test_parity_block_and_topk_rule_on_cpu():
    """bench.parity_block is the in-run checker of every bench line: on CPU, fed with the oracle's own logits (plus a
    perturbation) it must report zero (resp. the perturbation), all documented keys, and its top-k rule must agree with
    tests/helpers.topk_agree."""
    import torch
    import bench
    from oracle.protnote_oracle import EncoderCfg, ScorerCfg, protnote_forward
    from tests.helpers import topk_agree
    torch.manual_seed(0)
    model = bench.base_config_model("strict")
    onehots, lengths, labels = bench.synthetic_inputs(3, 48, 40, pinned=False)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref = protnote_forward(sd, onehots, lengths, labels, EncoderCfg(), ScorerCfg())
    blk = bench.parity_block(model, ref.clone(), onehots, lengths, labels, 1, world=2, proteins=(0, 1, 2))
    for key in ("checker", "proteins", "pairs_checked", "max_abs_err", "tol", "within_tol", "logit_std",
                "top10_identical_where_decided", "shard_boundary_columns", "shard_boundary_max_abs_err"):
        assert key in blk, key
    assert blk["max_abs_err"] == 0.0 and blk["within_tol"] and blk["top10_identical_where_decided"]
    assert blk["pairs_checked"] == 3 * 40 and blk["shard_boundary_columns"] == [0, 19, 20, 39]
    off = ref.clone()
    off[1, 20] += 3e-4
    blk = bench.parity_block(model, off, onehots, lengths, labels, 1, world=2, proteins=(0, 1, 2))
    assert not blk["within_tol"] and abs(blk["max_abs_err"] - 3e-4) < 1e-6 and abs(blk["shard_boundary_max_abs_err"] - 3e-4) < 1e-6
    # the top-k rule of the bench and of the tests is the same function of (reference, candidate)
    g = torch.Generator().manual_seed(1)
    for _ in range(20):
        r = torch.randn(4, 30, generator=g)
        c = r + torch.randn(4, 30, generator=g) * 2e-4
        assert bench.topk_decided_agree(r, c, 10, 1e-4)[0] == topk_agree(r, c, 10, 1e-4)
