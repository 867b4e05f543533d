"""GPU experiment: the round-toward-zero bias of the tcgen05 fp32 accumulator as a function of the promotion period, the
compensation factor that removes its mean (engine options trunc_comp_c1 / trunc_comp_c0), and what the encoder gains from
longer promotion periods.  Product code only (pn_linear through the C ABI); the reference is torch fp64 on the same GPU.

    python tools/trunc_comp_probe.py [out.json]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from protnote_b200 import native  # noqa: E402

dev = "cuda"
g = torch.Generator().manual_seed(0)
M, N = 1024, 512
PERIODS = (32, 64, 128, 256, 512, 1024, 0)


def stats(y, ref):
    d = y.double() - ref
    bias = float((d * ref.sign()).mean() / ref.abs().mean())
    # what is left when the best single factor is taken out (least squares): the incoherent part of the error
    alpha = float((d * ref).sum() / (ref * ref).sum())
    resid = d - alpha * ref
    return bias, float(d.pow(2).mean().sqrt()), float(resid.pow(2).mean().sqrt()), alpha


def run(cases, label):
    rows = []
    for name, x, w, ref in cases:
        for kc in PERIODS:
            native.set_option("promote_k_other", kc)
            y = native.linear(x, w, None, native.PN_STRICT)
            bias, rms, resid, alpha = stats(y, ref)
            K = x.shape[1]
            rows.append(dict(case=name, K=K, period=kc if kc else K, bias=bias, rms=rms, resid=resid, alpha=alpha))
            print(f"{label:12s} {name:22s} period {kc if kc else K:5d}  rel.bias {bias:+.3e}  lsq factor {alpha:+.3e}  "
                  f"rms {rms:.3e}  rms without factor {resid:.3e}", flush=True)
    return rows


cases = []
for K, relu_in in ((576, True), (1152, True), (3072, True), (3072, False), (9900, True)):
    x = torch.randn(M, K, generator=g)
    if relu_in:
        x = x.relu()
    w = torch.randn(N, K, generator=g) / K ** 0.5
    x, w = x.to(dev), w.to(dev)
    ref = x.double() @ w.double().T
    y32 = x @ w.T
    b, r, rr, a = stats(y32, ref)
    print(f"torch fp32 (TF32 {'on' if torch.backends.cuda.matmul.allow_tf32 else 'off'}) K={K} relu_in={relu_in}: rel.bias {b:+.3e} rms {r:.3e}")
    cases.append((f"K={K},relu={int(relu_in)}", x, w, ref))

results = {}
for beta in (0, 30000, 33000, 36000):
    native.set_option("trunc_beta_ppt", beta)
    results[beta] = run(cases, f"beta={beta}ppt")
native.set_option("trunc_beta_ppt", 33000)

# encoder time vs promotion period (base_config encoder, 128 x 1024 aa)
from bench import base_config_model, synthetic_inputs  # noqa: E402
model = base_config_model("strict").to(dev)
onehots, lengths, _ = synthetic_inputs(128, 1024, 8, pinned=False)
onehots, lengths = onehots.to(dev), lengths.to(dev)
enc = []
for cta2 in (0, 1):
    native.set_option("cta2", cta2)
    for kc in (32, 64, 128, 256, 512):
        native.set_option("promote_k_encoder", kc)
        model.sequence_encoder._packed = None      # weights carry the chunk structure: re-pack
        with torch.no_grad():
            model.sequence_encoder.get_embeddings(onehots, lengths)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                model.sequence_encoder.get_embeddings(onehots, lengths)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = 128 * 1024 * 60_896_000 / (ms * 1e-3) / 1e12
        enc.append(dict(cta2=cta2, period=kc, ms=ms, algorithmic_tflops=tf))
        print(f"encoder 128 x 1024 aa, cta2={cta2}, promotion every {kc:4d}: {ms:8.2f} ms  {tf:6.1f} algorithmic TFLOP/s", flush=True)
native.set_option("cta2", -1)
native.set_option("promote_k_encoder", 32)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(dict(results={str(k): v for k, v in results.items()}, encoder=enc), f, indent=1)
