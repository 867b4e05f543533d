"""TEST INFRASTRUCTURE ONLY - generates tests/golden/train_*.pt from the UNMODIFIED reference ProtNote class in TRAIN mode.

    python -m oracle.make_golden_train          (build container: needs /root/reference)

For each case the reference module (protnote/models/ProtNote.py:9-334) is built as bin/main.py:407-446 does, strict-loaded
with the seeded synthetic state_dict, put in .train() mode and called with the cached-embedding arguments
(sequence_embeddings=, label_embeddings=; the frozen encoder is not on the gradient path, ProtNote.py:243-260).  The loss is
the reference's 'BCE' (torch.nn.BCEWithLogitsLoss, protnote/utils/losses.py:270-294); `loss.backward()` fills the gradients.
Stored: logits, loss, every parameter gradient and the updated BatchNorm running statistics, in float64.
"""
from __future__ import annotations

import os
import sys

import torch

from .cases import CASES
from .make_golden import GOLDEN_DIR, build_reference_model, weight_checksum
from .protnote_oracle import synth_state_dict
from .train_oracle import synth_targets

# name -> (model case, B, L rows, input seed)
TRAIN_CASES = {
    "train_tiny": ("tiny_concat", 6, 50, 17),
    "train_tiny_wide": ("tiny_concat", 3, 130, 18),
}


def train_inputs(case: str):
    model_case, B, L, seed = TRAIN_CASES[case]
    ecfg, scfg, *_ = CASES[model_case]
    sd = synth_state_dict(ecfg, scfg, seed=CASES[model_case][6], calib_T=64)
    g = torch.Generator().manual_seed(seed)
    P_f = torch.randn(B, scfg.protein_embedding_dim, generator=g)
    L_f = torch.randn(L, scfg.label_embedding_dim, generator=g)
    return ecfg, scfg, sd, P_f, L_f, synth_targets(B, L, seed)


def main(argv=None):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in TRAIN_CASES:
        ecfg, scfg, sd, P_f, L_f, y = train_inputs(name)
        ref = build_reference_model(ecfg, scfg, sd).double().train()
        logits, _ = ref(sequence_embeddings=P_f.double(), label_embeddings=L_f.double())
        loss = torch.nn.BCEWithLogitsLoss()(logits, y.double())
        loss.backward()
        out = {
            "case": name, "weights_checksum": weight_checksum(sd), "inputs_checksum": float(P_f.double().sum() + L_f.double().sum()),
            "logits": logits.detach().clone(), "loss": float(loss),
            "grads": {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None},
            "running": {k: b.detach().clone() for k, b in ref.named_buffers()
                        if "running_" in k and not k.startswith("sequence_encoder.")},
            "torch_version": str(torch.__version__),
        }
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(out, path)
        print(f"{name}: logits {tuple(logits.shape)}, loss {float(loss):.6f}, {len(out['grads'])} gradients -> {path} "
              f"({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    sys.exit(main())
