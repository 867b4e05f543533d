"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.pt from the UNMODIFIED reference classes.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

For each case in oracle/cases.py it builds the reference's own ``ProteInfer`` and ``ProtNote``
(protnote/models/protein_encoders.py:70-107, protnote/models/ProtNote.py:10-102), strict-loads the
seeded synthetic state_dict, runs ``ProtNote.forward`` / ``ProteInfer.get_embeddings`` in fp32 eval
mode on the seeded inputs and stores the outputs together with a float64 checksum of the weights
(so RNG drift on another machine is detected instead of silently changing the problem).
"""
from __future__ import annotations

import os
import sys

import torch

from .cases import CASES
from .protnote_oracle import synth_inputs, synth_state_dict
from .ref_import import import_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference_model(ecfg, scfg, sd):
    ProtNote, ProteInfer, _ = import_reference()
    enc = ProteInfer(num_labels=sd["sequence_encoder.output_layer.weight"].shape[0],
                     input_channels=ecfg.input_channels, output_channels=ecfg.output_channels,
                     kernel_size=ecfg.kernel_size, activation=torch.nn.ReLU,
                     dilation_base=ecfg.dilation_base, num_resnet_blocks=ecfg.num_resnet_blocks,
                     bottleneck_factor=ecfg.bottleneck_factor)
    model = ProtNote(protein_embedding_dim=scfg.protein_embedding_dim,
                     label_embedding_dim=scfg.label_embedding_dim, latent_dim=scfg.latent_dim,
                     label_embedding_pooling_method="mean", label_encoder=None, sequence_encoder=enc,
                     inference_descriptions_per_label=scfg.inference_descriptions_per_label,
                     output_mlp_hidden_dim_scale_factor=scfg.output_mlp_hidden_dim_scale_factor,
                     output_mlp_num_layers=scfg.output_mlp_num_layers,
                     outout_mlp_add_batchnorm=scfg.output_mlp_batchnorm,
                     projection_head_num_layers=scfg.projection_head_num_layers,
                     projection_head_hidden_dim_scale_factor=scfg.projection_head_hidden_dim_scale_factor,
                     label_encoder_num_trainable_layers=0, train_sequence_encoder=False,
                     sequence_embedding_dropout=scfg.sequence_embedding_dropout,
                     label_embedding_dropout=scfg.label_embedding_dropout, dropout=scfg.output_mlp_dropout,
                     feature_fusion=scfg.feature_fusion, temperature=scfg.temperature)
    model.load_state_dict(sd, strict=True)
    return model.eval()


def weight_checksum(sd) -> float:
    return float(sum(v.double().sum() for k, v in sorted(sd.items()) if v.is_floating_point()))


def main(argv=None):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    for name, (ecfg, scfg, B, T, L, ragged, wseed, iseed) in CASES.items():
        sd = synth_state_dict(ecfg, scfg, seed=wseed, calib_T=min(T, 512))
        onehots, lengths, labels = synth_inputs(B, T, L, ecfg, scfg, ragged=ragged, seed=iseed)
        if name == "tiny_long":
            lengths[1] = 1  # a length-1 protein: its mean pool is a single column
            onehots[1, :, 1:] = 0
        model = build_reference_model(ecfg, scfg, sd)
        with torch.no_grad():
            emb = model.sequence_encoder.get_embeddings(onehots, lengths)
            logits, _ = model(sequence_onehots=onehots, sequence_lengths=lengths, label_embeddings=labels)
        out = {
            "case": name, "B": B, "T": T, "L": L,
            "tokens": onehots.argmax(1).to(torch.int8), "lengths": lengths.clone(),
            "labels_checksum": float(labels.double().sum()),
            "weights_checksum": weight_checksum(sd),
            "embeddings": emb.clone(), "logits": logits.clone(),
            "torch_version": str(torch.__version__),
        }
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(out, path)
        print(f"{name}: logits {tuple(logits.shape)} std {float(logits.std()):.3f} -> {path} "
              f"({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    sys.exit(main())
