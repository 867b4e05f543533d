#!/bin/bash
# Launch list (per-kernel device time) of the scorer at B=16 x L=32768, strict, for promote_k 0 and 256.
set -x
for PK in 0 256; do
  PN_PROMOTE=$PK ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pk$PK.csv \
    python - <<PY
import os, sys, torch
sys.path.insert(0, ".")
from protnote_b200 import native
from protnote_b200.ProtNote import ProtNote
native.set_option("promote_k", int(os.environ["PN_PROMOTE"]))
torch.manual_seed(0)
model = ProtNote(protein_embedding_dim=1100, label_embedding_dim=1024, latent_dim=1024, output_mlp_hidden_dim_scale_factor=3,
                 output_mlp_num_layers=3, projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3).cuda().eval()
scorer = model._ensure_packed()
a = torch.randn(16, 3072, device="cuda"); c = torch.randn(32768, 3072, device="cuda")
for _ in range(2):
    scorer.score(a, c, mode=3)
torch.cuda.synchronize()
PY
done
