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
    """Top-k label indices must be identical wherever the reference's ranking is decided by more
    than `2*tol` (ties closer than the tolerance are not a property of the arithmetic)."""
    k = min(k, ref.shape[1])
    rv, ri = ref.topk(k, dim=1)
    gi = got.topk(k, dim=1).indices
    srt = ref.sort(dim=1, descending=True).values
    nxt = srt[:, 1:k + 1] if ref.shape[1] > k else srt[:, 1:k]
    gaps = (srt[:, :nxt.shape[1]] - nxt).abs()
    decided = torch.ones_like(ri, dtype=torch.bool)
    decided[:, :gaps.shape[1]] &= gaps > 2 * tol
    decided[:, 1:] &= decided[:, :-1].clone() | True
    # rank j is well defined if gap(j-1,j) and gap(j,j+1) both exceed 2*tol
    ok_rank = torch.ones_like(ri, dtype=torch.bool)
    ok_rank[:, :gaps.shape[1]] &= gaps > 2 * tol
    ok_rank[:, 1:gaps.shape[1] + 1] &= (gaps > 2 * tol)[:, :ok_rank.shape[1] - 1]
    return bool(((ri == gi) | ~ok_rank).all())
