"""GPU probe: throughput of the projection + pair-scorer kernels at the headline shape (per-kernel CUDA-event times)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from protnote_b200 import native  # noqa: E402
from protnote_b200.ProtNote import ProtNote  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["strict", "fast"]
torch.manual_seed(0)
model = ProtNote(protein_embedding_dim=1100, label_embedding_dim=1024, latent_dim=1024,
                 output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3, projection_head_num_layers=4,
                 projection_head_hidden_dim_scale_factor=3, feature_fusion="concatenation").cuda().eval()
scorer = model._ensure_packed()
P_f = torch.randn(B, 1100, device="cuda")
L_f = torch.randn(L, 1024, device="cuda")
FLOP_PAIR = 37_754_880
for mode_name in modes:
    mode = native.MODES[mode_name]
    for bk, pk, fuse in ((32, 256, 1), (32, 256, 0), (64, 256, 0)):
        native.set_option("bk", bk)
        native.set_option("promote_k_scorer", pk)
        native.set_option("fuse_features", fuse)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for it in range(2):
            ev[0].record()
            _, a = scorer.project_sequences(P_f, mode)
            ev[1].record()
            _, c = scorer.project_labels(L_f, mode)
            ev[2].record()
            logits = scorer.score(a, c, mode=mode)
            ev[3].record()
            torch.cuda.synchronize()
        t_p, t_l, t_s = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
        pairs = B * L
        print(f"{mode_name} bk={bk} promote_k={pk} fuse={fuse}: W_p {t_p:.2f} ms, W_l {t_l:.2f} ms ({L*50.33e6/t_l/1e9:.0f} TFLOP/s), "
              f"score {t_s:.2f} ms -> {pairs/t_s/1e3:.2f} M pairs/s, {pairs*FLOP_PAIR/t_s/1e9:.0f} TFLOP/s algorithmic "
              f"({(3 if mode_name=='strict' else 1)*pairs*FLOP_PAIR/t_s/1e9:.0f} executed), logits finite={torch.isfinite(logits).all().item()}")
native.set_option("bk", 0)
native.set_option("promote_k_scorer", 256)
native.set_option("fuse_features", 0)
