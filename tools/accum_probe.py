"""GPU experiment: how the tcgen05 fp32 accumulator's rounding shows up, and what K-chunk promotion / separate
accumulation of the hi*lo terms buys.  Uses only pn_linear (strict / fast) plus fp32 adds in torch."""
import sys
import torch
sys.path.insert(0, ".")
from protnote_b200 import native  # noqa: E402

dev = "cuda"
g = torch.Generator().manual_seed(0)
M, N = 512, 512


def report(tag, y, ref):
    d = (y.double() - ref)
    rel = d / ref.abs().clamp_min(1e-3)
    print(f"{tag:40s} max|err| {d.abs().max().item():.3e}  mean err*sign(ref) {(d*ref.sign()).mean().item():+.3e}  rms {d.pow(2).mean().sqrt().item():.3e}  mean rel {rel.abs().mean().item():.2e}")


for K, relu_in in ((3072, True), (3072, False), (9900, True)):
    x = torch.randn(M, K, generator=g)
    if relu_in:
        x = x.relu()
    w = torch.randn(N, K, generator=g) / K ** 0.5
    x, w = x.to(dev), w.to(dev)
    ref = x.double() @ w.double().T
    print(f"--- K={K} relu_in={relu_in} ref rms {ref.pow(2).mean().sqrt().item():.3f}")
    report("torch fp32 matmul (TF32 off)", (x @ w.T), ref)
    report("strict, whole K", native.linear(x, w, None, native.PN_STRICT), ref)
    for chunk in (1024, 512, 256, 128, 64):
        acc = torch.zeros(M, N, device=dev)
        for k0 in range(0, K, chunk):
            acc += native.linear(x[:, k0:k0 + chunk].contiguous(), w[:, k0:k0 + chunk].contiguous(), None, native.PN_STRICT)
        report(f"strict, promoted every {chunk}", acc, ref)
    # separate accumulation of the correction terms
    xh = x.half().float(); xl = (x - xh)
    wh = w.half().float(); wl = (w - wh)   # (unscaled split: fine for this probe, |w| ~ 0.02)
    ws = w * 4096.0
    wh = ws.half().float(); wl = ws - wh
    main = native.linear(xh, wh, None, native.PN_FAST)
    corr = native.linear(xl * 2048, wh, None, native.PN_FAST) / 2048 + native.linear(xh, wl * 2048, None, native.PN_FAST) / 2048
    report("main/corr separate accumulators", (main + corr) / 4096.0, ref)
    for chunk in (512, 128):
        acc = torch.zeros(M, N, device=dev)
        for k0 in range(0, K, chunk):
            s = slice(k0, k0 + chunk)
            acc += native.linear(xh[:, s].contiguous(), wh[:, s].contiguous(), None, native.PN_FAST)
        report(f"separate + main promoted every {chunk}", (acc + corr) / 4096.0, ref)
torch.backends.cuda.matmul.allow_tf32 = False
