"""GPU diagnostic: where the logit error of the strict path comes from (encoder vs heads vs scorer), per golden case."""
import sys
import torch
sys.path.insert(0, ".")
from oracle.cases import CASES  # noqa: E402
from oracle import protnote_oracle as O  # noqa: E402
from tests.helpers import build_b200_model, load_case  # noqa: E402

torch.set_num_threads(8)
names = [a for a in sys.argv[1:] if "=" not in a] or list(CASES)
promote = [tuple(int(v) for v in a.split("=")[1].split(",")) for a in sys.argv[1:] if a.startswith("promote=")] or [(32, 32, 64)]
from protnote_b200 import native  # noqa: E402
# beta=<ppt>: per-K-position truncation compensation folded into the packed weights (engine option trunc_beta_ppt)
# comp=c1,c0 (ppt): uniform epilogue compensation (engine options trunc_comp_c1 / trunc_comp_c0)
for a in sys.argv[1:]:
    if a.startswith("beta="):
        native.set_option("trunc_beta_ppt", int(a.split("=")[1]))
        print(f"######## per-K-position truncation compensation beta = {a.split('=')[1]} ppt")
    if a.startswith("comp="):
        c1, c0 = (int(t) for t in a.split("=", 1)[1].split(","))
        native.set_option("trunc_comp_c1", c1)
        native.set_option("trunc_comp_c0", c0)
        print(f"######## uniform truncation compensation ({c1} * K_chunk + {c0}) ppt")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
for name in names:
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case(name)
    sd_cuda = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        lib_emb = O.proteinfer_embeddings(sd_cuda, onehots.cuda(), lengths.cuda(), ecfg, "sequence_encoder.").cpu()
        lib_logits = O.protnote_forward(sd_cuda, onehots.cuda(), lengths.cuda(), labels.cuda(), ecfg, scfg).cpu()
    print(f"== {name}: PyTorch library kernels on this GPU, fp32 (TF32 off): emb err vs golden {(lib_emb-g['embeddings']).abs().max():.2e}  logits err vs golden {(lib_logits-g['logits']).abs().max():.2e}")
for pk in promote:
  native.set_option("promote_k_encoder", pk[0]); native.set_option("promote_k_heads", pk[1]); native.set_option("promote_k_scorer", pk[2])
  print(f"######## promote_k encoder/heads/scorer = {pk}")
  for name in names:
      ecfg, scfg, sd, onehots, lengths, labels, g = load_case(name)
      model = build_b200_model(ecfg, scfg, sd)
      with torch.no_grad():
          emb = model.sequence_encoder.get_embeddings(onehots.cuda(), lengths.cuda()).cpu()
          logits, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
          logits = logits.cpu()
          emb64 = O.proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.", torch.float64)
          emb32 = O.proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.", torch.float32)
          ref64 = O.score_pairs(sd, emb64, labels, scfg, torch.float64)
          ref32 = g["logits"]
          # scorer alone: feed the exact (fp64-rounded-to-fp32) embeddings
          lg_exact_emb, _ = model(sequence_embeddings=emb64.float().cuda(), label_embeddings=labels.cuda())
          lg_exact_emb = lg_exact_emb.cpu()
          ref64_from32 = O.score_pairs(sd, emb64.float(), labels, scfg, torch.float64)
          # heads
          scorer = model._ensure_packed()
          P_e, a = scorer.project_sequences(emb64.float().cuda(), 3, want_embedding=True)
          L_e, c = scorer.project_labels(labels.cuda(), 3, want_embedding=True)
          wp = "W_p.1" if scfg.sequence_embedding_dropout > 0 else "W_p"
          P_e64 = O.projection_head(sd, wp, emb64.float().double(), scfg, torch.float64)
          L_e64 = O.projection_head(sd, "W_l", labels.double(), scfg, torch.float64)
          P_e32 = O.projection_head(sd, wp, emb64.float(), scfg, torch.float32)
      print(f"== {name}: logit std {ref32.std():.2f}")
      print(f"   emb:    ours-fp64 {(emb.double()-emb64).abs().max():.2e}   ref32-fp64 {(emb32.double()-emb64).abs().max():.2e}   |emb|max {emb64.abs().max():.2f}")
      print(f"   P_e:    ours-fp64 {(P_e.cpu().double()-P_e64).abs().max():.2e}   ref32-fp64 {(P_e32.double()-P_e64).abs().max():.2e}   |P_e|max {P_e64.abs().max():.2f}")
      print(f"   L_e:    ours-fp64 {(L_e.cpu().double()-L_e64).abs().max():.2e}   |L_e|max {L_e64.abs().max():.2f}")
      print(f"   logits (exact emb in): ours-fp64 {(lg_exact_emb.double()-ref64_from32).abs().max():.2e}")
      print(f"   logits end-to-end: ours-fp64 {(logits.double()-ref64).abs().max():.2e}   ours-ref32 {(logits-ref32).abs().max():.2e}   ref32-fp64 {(ref32.double()-ref64).abs().max():.2e}")
