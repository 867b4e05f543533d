"""GPU probe: CUDA-event time of every primitive call of one training step at bench size (no profiler attached).
Usage (GPU box): python tools/train_profile.py [labels] [precision] [sequences]"""
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, ".")
from bench import base_config_model  # noqa: E402
from protnote_b200 import train as pn_train  # noqa: E402
from protnote_b200.train_native import NativeOps  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
precision = sys.argv[2] if len(sys.argv) > 2 else "fast"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64


class Timed:
    def __init__(self, ops):
        self.ops, self.log = ops, []

    def __getattr__(self, name):
        fn = getattr(self.ops, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            rows = next((getattr(x, "rows", None) for x in a if hasattr(x, "rows")), None)
            self.log.append((name, rows, e0, e1))
            return out
        return wrapped


dev = torch.device("cuda")
model = base_config_model(precision).to(dev).train()
model.sequence_encoder.eval()
g = torch.Generator().manual_seed(0)
P_f = torch.randn(B, 1100, generator=g).to(dev)
L_f = torch.randn(L, 1024, generator=g).to(dev)
y = (torch.rand(B, L, generator=g) < 0.02).float().to(dev)
for it in range(2):
    ops = Timed(NativeOps(precision))
    t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0.record()
    logits, ctx = pn_train.forward_train(ops, None, model, P_f, L_f)
    t1.record()
    gl = (torch.sigmoid(logits) - y) / logits.numel()
    pn_train.backward_train(ops, None, ctx, gl)
    t2.record()
    torch.cuda.synchronize()
agg = OrderedDict()
for name, rows, e0, e1 in ops.log:
    big = rows is not None and rows >= B * L // 2
    key = f"{name}{' [pairs]' if big else ''}"
    d = agg.setdefault(key, [0, 0.0])
    d[0] += 1
    d[1] += e0.elapsed_time(e1)
total = sum(v[1] for v in agg.values())
print(f"B {B} L {L} {precision}: forward {t0.elapsed_time(t1):.1f} ms, backward {t1.elapsed_time(t2):.1f} ms; primitives {total:.1f} ms")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} calls {n:3d}  {ms:9.2f} ms  {100 * ms / total:5.1f}%")
