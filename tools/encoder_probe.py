"""GPU probe (product code only): one eval-mode pass of the base_config ProteInfer encoder, for ncu / timing.
Engine options come from PN_OPTIONS (e.g. PN_OPTIONS="promote_k_encoder=128,cta2=1").

    python tools/encoder_probe.py [sequences=128] [seq_len=1024] [reps=3]
"""
import sys

import torch

sys.path.insert(0, ".")
from bench import base_config_model, synthetic_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
model = base_config_model("strict").cuda()
onehots, lengths, _ = synthetic_inputs(B, T, 8, pinned=False)
onehots, lengths = onehots.cuda(), lengths.cuda()
with torch.no_grad():
    model.sequence_encoder.get_embeddings(onehots, lengths)
    torch.cuda.synchronize()
    if reps == 0:
        sys.exit(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model.sequence_encoder.get_embeddings(onehots, lengths)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"encoder {B} x {T} aa: {ms:.2f} ms, {B * T * 60_896_000 / (ms * 1e-3) / 1e12:.1f} algorithmic TFLOP/s")
