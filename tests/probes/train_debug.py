"""GPU probe: runs one training step with the CUDA primitives and with their fp64 torch statement side by side and prints,
call by call, how far each primitive's output is from the fp64 one (relative to the tensor's largest entry).
Usage (GPU box): python tests/probes/train_debug.py [case] [B] [L] [precision]"""
import sys

import torch

sys.path.insert(0, ".")
from oracle.cases import CASES  # noqa: E402
from oracle.protnote_oracle import synth_state_dict  # noqa: E402
from oracle.train_ops import TAct, TorchOps  # noqa: E402
from oracle.train_oracle import synth_targets  # noqa: E402
from protnote_b200 import train as pn_train  # noqa: E402
from protnote_b200.train_native import Act, NativeOps  # noqa: E402
from tests.helpers import build_b200_model  # noqa: E402


def to64(x):
    if isinstance(x, Act):
        v = x.hi.double()
        if x.lo is not None:
            v = v + x.lo.double()
        v = v[:, :x.cols]
        if x.sc is not None:
            v = v / x.sc[0].double()
        return v.cpu()
    if isinstance(x, TAct):
        return (x.val / x.sc).double()
    if isinstance(x, torch.Tensor):
        return x.detach().double().cpu()
    if hasattr(x, "sums"):
        return x.sums.detach().double().cpu()
    if hasattr(x, "scale") and hasattr(x, "invstd"):
        return torch.stack([x.scale, x.shift, x.mean, x.invstd]).double()
    if isinstance(x, tuple):
        return torch.cat([to64(t).reshape(-1) for t in x])
    return None


class Recorder:
    def __init__(self, ops):
        self.ops, self.log = ops, []

    def __getattr__(self, name):
        fn = getattr(self.ops, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            out = fn(*a, **k)
            v = to64(out)
            if v is not None:
                self.log.append((name, v))
            return out
        return wrapped


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "base_small"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    precision = sys.argv[4] if len(sys.argv) > 4 else "strict"
    ecfg, scfg, *_ = CASES[case]
    sd = synth_state_dict(ecfg, scfg, seed=CASES[case][6], calib_T=64)
    g = torch.Generator().manual_seed(17)
    P_f = torch.randn(B, scfg.protein_embedding_dim, generator=g)
    L_f = torch.randn(L, scfg.label_embedding_dim, generator=g)
    y = synth_targets(B, L, 17)
    logs = []
    for dev in ("cuda", "cpu"):
        model = build_b200_model(ecfg, scfg, sd, device=dev).train()
        if dev == "cpu":
            model = model.double()
            ops = Recorder(TorchOps(torch.float64))
            pf, lf, yy = P_f.double(), L_f.double(), y.double()
        else:
            ops = Recorder(NativeOps(precision))
            pf, lf, yy = P_f.cuda(), L_f.cuda(), y.cuda()
        logits, ctx = pn_train.forward_train(ops, None, model, pf, lf)
        gl = (torch.sigmoid(logits) - yy) / logits.numel()
        pn_train.backward_train(ops, None, ctx, gl.to(logits.dtype))
        logs.append(ops.log)
    for i, ((n1, a), (n2, b)) in enumerate(zip(*logs)):
        assert n1 == n2, (n1, n2)
        if a.shape != b.shape:
            print(f"{i:3d} {n1:18s} shape mismatch {tuple(a.shape)} {tuple(b.shape)}")
            continue
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        rel_l2 = float((a - b).norm() / b.norm().clamp_min(1e-300))
        print(f"{i:3d} {n1:18s} {str(tuple(b.shape)):16s} max|ref| {scale:9.3e}  max err/scale {err / max(scale, 1e-300):9.2e}  rel L2 {rel_l2:9.2e}")


if __name__ == "__main__":
    main()
