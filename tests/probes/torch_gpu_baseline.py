"""GPU probe: the reference arithmetic executed by PyTorch's own CUDA kernels (cuDNN / cuBLAS) on the same B200 -
the library-kernel number SURVEY.md section 8(d) asks to beat.  The model is the oracle restatement of the reference
forward (oracle/protnote_oracle.py, test infrastructure) moved to the GPU; fp32 with TF32 off, fp32 with TF32 on, and
under torch.autocast(fp16) as ProtNoteTrainer.evaluation_step runs it (ProtNoteTrainer.py:287).
Usage (GPU box): python tests/probes/torch_gpu_baseline.py [sequences] [labels]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import base_config_model, synthetic_inputs  # noqa: E402
from oracle.protnote_oracle import EncoderCfg, ScorerCfg, proteinfer_embeddings, score_pairs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
T = 1024
dev = torch.device("cuda")
model = base_config_model("strict")
sd = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
onehots, lengths, labels = synthetic_inputs(B, T, L, pinned=False)
onehots, lengths, labels = onehots.to(dev), lengths.to(dev), labels.to(dev)
ecfg, scfg = EncoderCfg(), ScorerCfg()


def run(autocast):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
        P_f = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.")
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        # the reference materialises the joint tensor for the whole batch (ProtNote.py:112-126); 8 proteins x 32K rows
        # x 2048 floats = 2.1 GB per chunk here
        logits = score_pairs(sd, P_f.float(), labels, scfg, pair_chunk=8 * L)
        torch.cuda.synchronize()
    return t1, logits


for name, tf32, autocast in (("fp32 (TF32 off)", False, False), ("fp32 (TF32 on)", True, False), ("autocast fp16", True, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    run(autocast)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t1, logits = run(autocast)
    t2 = time.perf_counter()
    print(json.dumps({"impl": "PyTorch CUDA kernels, oracle restatement of the reference forward", "precision": name,
                      "sequences": B, "seq_len": T, "label_rows": L, "pair_scores_per_s": B * L / (t2 - t0),
                      "encoder_s": t1 - t0, "scorer_s": t2 - t1, "logit_std": float(logits.float().std())}), flush=True)
