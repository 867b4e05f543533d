"""GPU diagnostic: error of the calibrated base_config bench model (bench.py) against the fp64 and fp32 CPU oracle, for a
few engine option sets.  Decomposes max |logit - oracle_fp32| into our error vs fp64 and the fp32 reference's own rounding.

    python tests/probes/bench_parity_probe.py "opt=a=1,b=2" "opt=..."      (each opt=... is one PN_OPTIONS-style set)
"""
import sys

import torch

sys.path.insert(0, ".")
from bench import base_config_model, calibrate_model, synthetic_inputs  # noqa: E402
from oracle.protnote_oracle import EncoderCfg, ScorerCfg, proteinfer_embeddings, score_pairs  # noqa: E402
from protnote_b200 import native  # noqa: E402

torch.set_num_threads(16)
dev = torch.device("cuda", 0)
B, T, L = 3, 1024, 4096
model = base_config_model("strict").to(dev)
print(calibrate_model(model, dev))
onehots, lengths, labels = synthetic_inputs(B, T, L, pinned=False)
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
ecfg, scfg = EncoderCfg(), ScorerCfg()
with torch.no_grad():
    emb64 = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.", torch.float64)
    emb32 = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.", torch.float32)
    ref64 = score_pairs(sd, emb64, labels, scfg, torch.float64)
    ref32 = score_pairs(sd, emb32, labels, scfg, torch.float32)
    ref64_e32 = score_pairs(sd, emb64.float(), labels, scfg, torch.float64)
print(f"logit std {ref64.std():.3f} absmax {ref64.abs().max():.2f};  fp32 oracle vs fp64: max {(ref32.double()-ref64).abs().max():.2e} "
      f"mean signed {(ref32.double()-ref64).mean():+.2e} rms {(ref32.double()-ref64).pow(2).mean().sqrt():.2e}; "
      f"emb32 vs emb64 max {(emb32.double()-emb64).abs().max():.2e}")
sets = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("opt=")] or [""]
for spec in sets:
    for item in filter(None, (x.strip() for x in spec.split(","))):
        name, _, value = item.partition("=")
        native.set_option(name, int(value))
    with torch.no_grad():
        emb = model.sequence_encoder.get_embeddings(onehots.to(dev), lengths.to(dev)).cpu()
        model._label_cache = None
        logits = model(sequence_onehots=onehots.to(dev), sequence_lengths=lengths.to(dev), label_embeddings=labels.to(dev))[0].cpu()
        model._label_cache = None
        lg_exact = model(sequence_embeddings=emb64.float().to(dev), label_embeddings=labels.to(dev))[0].cpu()
    d64, d32, ds = logits.double() - ref64, logits - ref32, lg_exact.double() - ref64_e32
    print(f"[{spec}] emb err max {(emb.double()-emb64).abs().max():.2e} | logits vs fp64: max {d64.abs().max():.2e} mean signed "
          f"{d64.mean():+.2e} rms {d64.pow(2).mean().sqrt():.2e} | vs fp32 oracle: max {d32.abs().max():.2e} mean|.| {d32.abs().mean():.2e} "
          f"| scorer alone (exact emb) vs fp64: max {ds.abs().max():.2e} mean signed {ds.mean():+.2e} rms {ds.pow(2).mean().sqrt():.2e}", flush=True)
