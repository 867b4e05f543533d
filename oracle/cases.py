"""TEST INFRASTRUCTURE ONLY - the named golden cases shared by make_golden.py and the tests."""
from __future__ import annotations

from .protnote_oracle import EncoderCfg, ScorerCfg

_TINY_E = dict(input_channels=20, output_channels=72, kernel_size=9, dilation_base=3,
               num_resnet_blocks=3, bottleneck_factor=0.5)
_TINY_S = dict(protein_embedding_dim=72, label_embedding_dim=40, latent_dim=32,
               output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3,
               projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3)

# name -> (encoder cfg, scorer cfg, B, T, L rows, ragged, weight seed, input seed)
CASES = {
    # small odd-sized model: exercises every padding path (72, 36, 40, 96 are not multiples of 64)
    "tiny_concat": (EncoderCfg(**_TINY_E), ScorerCfg(**_TINY_S), 5, 150, 24, True, 42, 1234),
    # inference ensembling over k=2 consecutive description rows (ProtNote.py:308-322)
    "tiny_k2": (EncoderCfg(**_TINY_E), ScorerCfg(**_TINY_S, inference_descriptions_per_label=2),
                3, 97, 20, True, 43, 1235),
    # [p; t; p*t] fusion (ProtNote.py:139-150)
    "tiny_prod": (EncoderCfg(**_TINY_E), ScorerCfg(**_TINY_S, feature_fusion="concatenation_prod"),
                  4, 64, 17, True, 44, 1236),
    # sequence longer than the receptive field of the widest dilation; length-1 and full-length rows
    "tiny_long": (EncoderCfg(**{**_TINY_E, "num_resnet_blocks": 5}), ScorerCfg(**_TINY_S),
                  3, 700, 9, True, 45, 1237),
    # the published architecture (base_config.yaml) at a size the CPU finishes in seconds
    "base_small": (EncoderCfg(), ScorerCfg(), 2, 300, 48, True, 46, 1238),
}
