"""Shared helpers for the parity tests (the oracle is the checker, never the product)."""
from __future__ import annotations

import os

import torch

from oracle.cases import CASES
from oracle.protnote_oracle import pad_mask, synth_inputs, synth_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def weight_checksum(sd) -> float:
    return float(sum(v.double().sum() for k, v in sorted(sd.items()) if v.is_floating_point()))


def load_case(name):
    """Rebuild a golden case from its seeds and verify it against the committed fixture."""
    ecfg, scfg, B, T, L, ragged, wseed, iseed = CASES[name]
    g = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"))
    sd = synth_state_dict(ecfg, scfg, seed=wseed, calib_T=min(T, 512))
    onehots, lengths, labels = synth_inputs(B, T, L, ecfg, scfg, ragged=ragged, seed=iseed)
    # the fixture carries the exact tokens/lengths the reference saw
    lengths = g["lengths"].clone()
    tokens = g["tokens"].long()
    onehots = torch.nn.functional.one_hot(tokens, ecfg.input_channels).permute(0, 2, 1).float()
    onehots = (onehots * (~pad_mask(lengths, T))[:, None, :]).contiguous()
    assert abs(weight_checksum(sd) - g["weights_checksum"]) <= 1e-6 * max(1.0, abs(g["weights_checksum"])), \
        "synthetic weights differ from the ones the golden vectors were made with (RNG drift)"
    assert abs(float(labels.double().sum()) - g["labels_checksum"]) <= 1e-6 * max(1.0, abs(g["labels_checksum"]))
    return ecfg, scfg, sd, onehots, lengths, labels, g


def topk_agree(ref: torch.Tensor, got: torch.Tensor, k: int, tol: float) -> bool:
    """Top-k label indices must be identical at every rank the reference decides by more than `2*tol`: rank j counts
    when both of its neighbours in the reference's sorted row (j-1 and j+1) are further than 2*tol away (ties closer
    than the tolerance are not a property of the arithmetic)."""
    k = min(k, ref.shape[1])
    ri = ref.topk(k, dim=1).indices
    gi = got.topk(k, dim=1).indices
    srt = ref.sort(dim=1, descending=True).values[:, :k + 1]
    clear = (srt[:, :-1] - srt[:, 1:]) > 2 * tol          # clear[:, j]: gap between sorted ranks j and j+1
    decided = torch.ones_like(ri, dtype=torch.bool)
    decided[:, :clear.shape[1]] &= clear[:, :k]              # gap below rank j
    decided[:, 1:] &= clear[:, :k - 1]                       # gap above rank j
    return bool(((ri == gi) | ~decided).all())


def build_b200_model(ecfg, scfg, sd, device="cuda", precision="strict"):
    """protnote_b200 modules built exactly the way bin/main.py:383-446 builds the reference's, strict-loaded with `sd`."""
    from protnote_b200.ProtNote import ProtNote
    from protnote_b200.protein_encoders import ProteInfer
    enc = ProteInfer(num_labels=sd["sequence_encoder.output_layer.weight"].shape[0],
                     input_channels=ecfg.input_channels, output_channels=ecfg.output_channels,
                     kernel_size=ecfg.kernel_size, activation=torch.nn.ReLU, dilation_base=ecfg.dilation_base,
                     num_resnet_blocks=ecfg.num_resnet_blocks, bottleneck_factor=ecfg.bottleneck_factor,
                     precision=precision)
    model = ProtNote(protein_embedding_dim=scfg.protein_embedding_dim, label_embedding_dim=scfg.label_embedding_dim,
                     latent_dim=scfg.latent_dim, label_embedding_pooling_method="mean", label_encoder=None,
                     sequence_encoder=enc,
                     inference_descriptions_per_label=scfg.inference_descriptions_per_label,
                     output_mlp_hidden_dim_scale_factor=scfg.output_mlp_hidden_dim_scale_factor,
                     output_mlp_num_layers=scfg.output_mlp_num_layers,
                     outout_mlp_add_batchnorm=scfg.output_mlp_batchnorm,
                     projection_head_num_layers=scfg.projection_head_num_layers,
                     projection_head_hidden_dim_scale_factor=scfg.projection_head_hidden_dim_scale_factor,
                     label_encoder_num_trainable_layers=0, train_sequence_encoder=False,
                     sequence_embedding_dropout=scfg.sequence_embedding_dropout,
                     label_embedding_dropout=scfg.label_embedding_dropout,
                     dropout=scfg.output_mlp_dropout,
                     feature_fusion=scfg.feature_fusion, temperature=scfg.temperature, precision=precision)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()
