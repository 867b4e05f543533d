#!/bin/bash
# ncu --set full of the fused-generator GEMM (option fuse_features=1), scorer only, B=8 x L=32768
PN_OPTIONS=fuse_features=1 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_fused \
  python - <<PY
import sys, torch
sys.path.insert(0, ".")
from protnote_b200 import native
from protnote_b200.ProtNote import ProtNote
torch.manual_seed(0)
model = ProtNote(protein_embedding_dim=1100, label_embedding_dim=1024, latent_dim=1024, output_mlp_hidden_dim_scale_factor=3,
                 output_mlp_num_layers=3, projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3).cuda().eval()
scorer = model._ensure_packed()
a = torch.randn(8, 3072, device="cuda"); c = torch.randn(32768, 3072, device="cuda")
for _ in range(2):
    scorer.score(a, c, mode=3)
torch.cuda.synchronize()
PY
