"""The travelling oracle reproduces the reference's outputs stored in tests/golden (CPU, runs everywhere)."""
import pytest
import torch

from oracle.cases import CASES
from oracle.protnote_oracle import proteinfer_embeddings, protnote_forward
from tests.helpers import load_case

# fp32 tolerance: the oracle and the reference run the same fp32 operators in a different order
# (e.g. explicit pad + conv vs padding="same"), so agreement is to rounding, not bitwise.
TOL = 2e-5


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden(name):
    torch.set_num_threads(8)
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case(name)
    with torch.no_grad():
        emb = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.")
        logits = protnote_forward(sd, onehots, lengths, labels, ecfg, scfg)
    assert emb.shape == g["embeddings"].shape and logits.shape == g["logits"].shape
    assert (emb - g["embeddings"]).abs().max() <= TOL * max(1.0, float(g["embeddings"].abs().max()))
    assert (logits - g["logits"]).abs().max() <= TOL * max(1.0, float(g["logits"].abs().max()))


def test_oracle_float64_is_close_to_float32():
    """Sizes the fp32 noise floor of the reference itself: its fp32 logits sit 1e-5..5e-5 away from an
    fp64 evaluation of the same network (logit std ~2), which is what the 1e-4 parity tolerance of the
    CUDA path has to be read against."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    hi = protnote_forward(sd, onehots, lengths, labels, ecfg, scfg, dtype=torch.float64)
    assert (hi.float() - g["logits"]).abs().max() < 1e-4


def test_padding_is_ignored():
    """Per-sequence independence: garbage in the padded columns must not change anything."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    noisy = onehots.clone()
    T = onehots.shape[-1]
    for b in range(onehots.shape[0]):
        noisy[b, :, int(lengths[b]):] = 7.0
    a = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.")
    b_ = proteinfer_embeddings(sd, noisy, lengths, ecfg, "sequence_encoder.")
    assert torch.equal(a, b_)
