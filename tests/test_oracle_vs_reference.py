"""Pins the oracle against the UNMODIFIED reference classes (only where /root/reference exists)."""
import pytest
import torch

from oracle.cases import CASES
from oracle.protnote_oracle import ScorerCfg, proteinfer_embeddings, protnote_forward, synth_inputs, synth_state_dict
from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present on this machine")


@pytest.mark.parametrize("name", ["tiny_concat", "tiny_k2", "tiny_prod"])
def test_fresh_inputs_match_reference(name):
    """New seeds (not the golden ones): oracle == reference modules to fp32 rounding."""
    from oracle.make_golden import build_reference_model
    ecfg, scfg, B, T, L, ragged, wseed, iseed = CASES[name]
    sd = synth_state_dict(ecfg, scfg, seed=wseed + 100)
    onehots, lengths, labels = synth_inputs(B + 1, T + 13, L + 2 * scfg.inference_descriptions_per_label,
                                            ecfg, scfg, ragged=True, seed=iseed + 100)
    model = build_reference_model(ecfg, scfg, sd)
    with torch.no_grad():
        ref_emb = model.sequence_encoder.get_embeddings(onehots, lengths)
        ref_logits, _ = model(sequence_onehots=onehots, sequence_lengths=lengths, label_embeddings=labels)
        emb = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.")
        logits = protnote_forward(sd, onehots, lengths, labels, ecfg, scfg)
    # fp32 vs fp32: same operators, different association order -> agreement to the fp32 noise floor
    assert (emb - ref_emb).abs().max() < 2e-6
    assert (logits - ref_logits).abs().max() < 1e-4
    # fp64 vs fp64: the algorithm itself is pinned to ~1e-12
    model = model.double()
    with torch.no_grad():
        ref64, _ = model(sequence_onehots=onehots.double(), sequence_lengths=lengths,
                         label_embeddings=labels.double())
        got64 = protnote_forward(sd, onehots, lengths, labels, ecfg, scfg, dtype=torch.float64)
    assert (got64 - ref64).abs().max() < 1e-9


def test_non_onehot_float_input_and_similarity_fusion():
    """The module accepts any float [B,Cin,T] and the 'similarity' fusion (ProtNote.py:281-284)."""
    from oracle.make_golden import build_reference_model
    ecfg, scfg, B, T, L, ragged, wseed, iseed = CASES["tiny_concat"]
    scfg = ScorerCfg(**{**scfg.__dict__, "feature_fusion": "similarity"})
    sd = synth_state_dict(ecfg, scfg, seed=7)
    onehots, lengths, labels = synth_inputs(3, 80, 11, ecfg, scfg, ragged=True, seed=8)
    x = onehots + 0.25 * torch.randn(onehots.shape, generator=torch.Generator().manual_seed(1))
    model = build_reference_model(ecfg, scfg, sd)
    with torch.no_grad():
        ref_logits, _ = model(sequence_onehots=x, sequence_lengths=lengths, label_embeddings=labels)
        logits = protnote_forward(sd, x, lengths, labels, ecfg, scfg)
    assert (logits - ref_logits).abs().max() < 1e-4
