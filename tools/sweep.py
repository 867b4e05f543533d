"""GPU probe (BASELINE.json configs[4]): throughput of the scoring path over sequence length x label rows, strict and fast.
Prints one JSON line per point: pair-scores/s of the whole forward, encoder residues/s, scorer pairs/s.
Usage (GPU box): python tools/sweep.py [sequences]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from bench import base_config_model, synthetic_inputs  # noqa: E402
from protnote_b200 import native  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
model = base_config_model("strict").to(dev)
for T in (256, 512, 1024, 2048):
    for L in (1024, 8192, 32768):
        x, lens, lab = (t.to(dev) for t in synthetic_inputs(B, T, L, pinned=False))
        for mode in ("strict", "fast"):
            model.precision = model.sequence_encoder.precision = mode
            m = native.MODES[mode]
            scorer = model._ensure_packed()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            with torch.no_grad():
                for it in range(2):
                    model._label_cache = None
                    ev[0].record()
                    P_f = model.sequence_encoder.get_embeddings(x, lens)
                    ev[1].record()
                    _, a = scorer.project_sequences(P_f, m)
                    _, c = scorer.project_labels(lab, m)
                    ev[2].record()
                    scorer.score(a, c, mode=m)
                    ev[3].record()
                    torch.cuda.synchronize()
            enc, proj, sc = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
            tot = enc + proj + sc
            print(json.dumps({"mode": mode, "sequences": B, "seq_len": T, "label_rows": L,
                              "pair_scores_per_s": B * L / tot * 1e3, "ms": tot, "encoder_ms": enc, "heads_ms": proj,
                              "scorer_ms": sc, "encoder_residues_per_s": B * T / enc * 1e3,
                              "encoder_tflops": B * T * 60.896e6 / enc / 1e9,
                              "scorer_pairs_per_s": B * L / sc * 1e3, "scorer_tflops": B * L * 37.75488e6 / sc / 1e9}),
                  flush=True)
