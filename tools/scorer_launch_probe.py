"""GPU probe (product only): one strict pass of W_p / W_l / the pair scorer at 16 proteins x 32768 label rows (one workspace
chunk = 2^19 pairs, the headline's chunk size), for an ncu launch list."""
import sys

import torch

sys.path.insert(0, ".")
from bench import base_config_model  # noqa: E402
from protnote_b200 import native  # noqa: E402

model = base_config_model("strict").cuda()
scorer = model._ensure_packed()
P_f = torch.randn(16, 1100, device="cuda")
L_f = torch.randn(32768, 1024, device="cuda")
mode = native.MODES["strict"]
_, a = scorer.project_sequences(P_f, mode)
_, c = scorer.project_labels(L_f, mode)
logits = scorer.score(a, c, mode=mode)
torch.cuda.synchronize()
print("ok", tuple(logits.shape))
