"""Prints the interesting parts of a bench.py JSON line (file may contain other stdout lines, e.g. NCCL's banner)."""
import json
import sys

for path in sys.argv[1:]:
    line = [ln for ln in open(path).read().splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    print(f"== {path}: N={d['n_gpus']} value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.1f}  e2e {d['e2e']['value']:.4g}")
    r, e = d.get("roofline", {}), d.get("roofline_encoder", {})
    print(f"   roofline frac {r.get('frac')}  executed {r.get('executed_frac')}  encoder frac {e.get('frac')}  clocks {d.get('clocks')}")
    p = d.get("parity") or {}
    print(f"   parity: max_abs_err {p.get('max_abs_err')} within {p.get('within_tol')} top10 {p.get('top10_identical_where_decided')} "
          f"std {p.get('logit_std')} absmax {p.get('logit_absmax')} ranks_agree {(p.get('ranks_agree') or {}).get('identical')}")
    c = d.get("configs") or {}
    if c:
        ec = c.get("ec") or {}
        print(f"   EC: {ec.get('value')} pairs/s, frac {ec.get('frac_of_peak')}, parity {(ec.get('parity') or {}).get('max_abs_err')} err={ec.get('error')}")
        for pt in (c.get("sweep") or {}).get("points", []):
            print(f"   sweep T={pt['seq_len']:5d} L={pt['label_rows']:6d}: {pt['value']:.4g} pairs/s  frac {pt['frac_of_peak']:.3f}  scorer GEMM {pt['scorer_gemm_tflops']:.0f} TF")
        t = c.get("train") or {}
        print(f"   train parity: {json.dumps(t.get('parity'))[:400]}")
        for m in ("fast", "strict"):
            print(f"   train {m}: {json.dumps(t.get(m))[:500]}")
        print(f"   configs seconds {c.get('seconds')}  train error {t.get('error')}  memory GiB after headline/ec/before train: "
              f"{c.get('memory_allocated_gib_after_headline')} {c.get('memory_allocated_gib_after_ec')} {c.get('memory_allocated_gib_before_train')}")
    if "fast_mode" in d:
        print(f"   fast_mode: {d['fast_mode']['value']:.4g} pairs/s, scorer GEMM {d['fast_mode']['scorer_gemm_tflops']}")
    if "cpu_baseline" in d:
        print(f"   cpu_baseline: {d['cpu_baseline']['value']:.4g} kind {d['cpu_baseline']['kind']} cores {d['cpu_baseline']['cores']}")
