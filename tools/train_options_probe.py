"""What the non-default training options cost (one B200, fast and strict mode, batch 64 x 4096 label rows = one rank's share
of the 8-GPU layout, sequence embeddings given): the default configuration, OUTPUT_MLP_DROPOUT 0.1, FEATURE_FUSION
concatenation_prod, OUTPUT_MLP_BATCHNORM False.  ms per step of train_loss (focal) + backward, CUDA events.
    python tools/train_options_probe.py > gpurun_out/train_options.txt"""
import sys
import time

import torch

sys.path.insert(0, ".")
from protnote_b200 import train as pn_train  # noqa: E402
from protnote_b200.ProtNote import ProtNote  # noqa: E402

B, L = 64, 4096
dev = torch.device("cuda", 0)


def model(precision, **kw):
    torch.manual_seed(42)
    args = dict(protein_embedding_dim=1100, label_embedding_dim=1024, latent_dim=1024, label_embedding_pooling_method="mean",
                sequence_encoder=None, output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3,
                outout_mlp_add_batchnorm=True, projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3,
                feature_fusion="concatenation", precision=precision)
    args.update(kw)
    return ProtNote(**args).to(dev).train()


g = torch.Generator().manual_seed(1)
P_f, L_f = torch.randn(B, 1100, generator=g).to(dev), torch.randn(L, 1024, generator=g).to(dev)
y = (torch.rand(B, L, generator=g) < 0.02).float().to(dev)
t_start = time.time()
for precision in ("fast", "strict"):
    for name, kw in (("default", {}), ("OUTPUT_MLP_DROPOUT 0.1", dict(dropout=0.1)),
                     ("concatenation_prod", dict(feature_fusion="concatenation_prod")),
                     ("OUTPUT_MLP_BATCHNORM False", dict(outout_mlp_add_batchnorm=False))):
        if time.time() - t_start > 40:
            break
        m = model(precision, **kw)

        def step():
            for p in m.parameters():
                p.grad = None
            loss, _ = pn_train.train_loss(m, P_f, L_f, y, loss="focal", gamma=2.0, alpha=0.25)
            loss.backward()
            return loss

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        print(f"{precision:6s} {name:28s} {e0.elapsed_time(e1) / 3:8.2f} ms / step   loss {float(loss.detach()):.5f}   "
              f"peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB", flush=True)
        del m
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
