"""GPU diagnostic: where the pair scorer's logit error comes from.  The layer-1 halves a [B,H] / c [L,H] are computed in fp64
(exact) and by the library; the scorer's layers 2..3 + output dot are then run by the library on either, and everything is
compared with the fp64 oracle.  Prints, per variant, max / rms / mean-signed error and the per-protein mean-signed error."""
import sys

import torch

sys.path.insert(0, ".")
from bench import base_config_model, calibrate_model, synthetic_inputs  # noqa: E402
from oracle.protnote_oracle import EncoderCfg, ScorerCfg, batchnorm_eval, projection_head, proteinfer_embeddings, score_pairs  # noqa: E402
from protnote_b200 import native  # noqa: E402

torch.set_num_threads(16)
dev = torch.device("cuda", 0)
B, T, L = 4, 512, 4096
model = base_config_model("strict").to(dev)
calibrate_model(model, dev)
onehots, lengths, labels = synthetic_inputs(B, T, L, pinned=False)
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
ecfg, scfg = EncoderCfg(), ScorerCfg()
f64 = torch.float64
with torch.no_grad():
    emb64 = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.", f64)
    P_f = emb64.float()                                    # the scorer's input (exactly representable)
    ref64 = score_pairs(sd, P_f, labels, scfg, f64)
    P_e = projection_head(sd, "W_p", P_f.double(), scfg, f64)
    L_e = projection_head(sd, "W_l", labels.double(), scfg, f64)
    W1 = sd["output_layer.0.weight"].double()
    d = P_e.shape[1]
    g, b_, m, v = (sd[f"output_layer.1.{k}"].double() for k in ("weight", "bias", "running_mean", "running_var"))
    s = g / torch.sqrt(v + scfg.bn_eps)
    a64 = (P_e @ W1[:, :d].T) * s + (b_ - m * s)
    c64 = (L_e @ W1[:, d:].T) * s


def report(tag, logits):
    e = logits.cpu().double() - ref64
    per = " ".join(f"{float(x):+.1e}" for x in e.mean(1))
    print(f"{tag:58s} max {e.abs().max():.2e} rms {e.pow(2).mean().sqrt():.2e} mean {e.mean():+.2e} | per protein: {per}", flush=True)


scorer = model._ensure_packed()
mode = native.MODES["strict"]
for pk in (256, 128, 64):
    native.set_option("promote_k_scorer", pk)
    scorer = model._ensure_packed()
    with torch.no_grad():
        _, a = scorer.project_sequences(P_f.to(dev), mode)
        _, c = scorer.project_labels(labels.to(dev), mode)
        print(f"promote_k_scorer={pk}: a err max {(a.cpu().double()-a64).abs().max():.2e} (|a| max {a64.abs().max():.1f})   "
              f"c err max {(c.cpu().double()-c64).abs().max():.2e} (|c| max {c64.abs().max():.1f})")
        report("  library a, library c", scorer.score(a, c, mode=mode))
        report("  exact a (fp64 -> fp32), library c", scorer.score(a64.float().to(dev), c, mode=mode))
        report("  library a, exact c", scorer.score(a, c64.float().to(dev), mode=mode))
        report("  exact a, exact c  (layers 2, 3 + dot only)", scorer.score(a64.float().to(dev), c64.float().to(dev), mode=mode))
with torch.no_grad():
    report("fp32 CPU oracle", score_pairs(sd, P_f, labels, scfg, torch.float32))
