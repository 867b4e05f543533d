"""Parity of the sm_100a path (through the C ABI) against the reference's outputs (tests/golden) and the oracle."""
import pytest
import torch

from oracle.cases import CASES
from oracle.protnote_oracle import proteinfer_embeddings, protnote_forward
from tests.helpers import build_b200_model, load_case, topk_agree

pytestmark = pytest.mark.gpu

# north_star: fp32 logits within 1e-4 of the reference PyTorch path, identical top-k label indices.
LOGIT_TOL = 1e-4


@pytest.mark.parametrize("name", list(CASES))
def test_golden_case_strict(name):
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case(name)
    model = build_b200_model(ecfg, scfg, sd)
    with torch.no_grad():
        emb = model.sequence_encoder.get_embeddings(onehots.cuda(), lengths.cuda())
        logits, extra = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(),
                              label_embeddings=labels.cuda())
    torch.cuda.synchronize()
    assert logits.dtype == torch.float32 and logits.shape == g["logits"].shape
    assert set(extra) == {"output_layer_embeddings", "joint_embeddings"}
    emb_err = (emb.cpu() - g["embeddings"]).abs().max().item()
    err = (logits.cpu() - g["logits"]).abs().max().item()
    ref64 = protnote_forward(sd, onehots, lengths, labels, ecfg, scfg, dtype=torch.float64)
    err64 = (logits.cpu().double() - ref64).abs().max().item()
    print(f"{name}: emb err {emb_err:.2e}  logits err vs reference {err:.2e}  vs fp64 oracle {err64:.2e}")
    assert emb_err <= 1e-4 * max(1.0, g["embeddings"].abs().max().item())
    assert err <= LOGIT_TOL
    assert topk_agree(g["logits"], logits.cpu(), k=10, tol=LOGIT_TOL)


def test_fast_mode_is_close():
    """fp16-operand mode is compared at the tolerance the reference's own autocast path has (~1e-2 of logit std)."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("base_small")
    model = build_b200_model(ecfg, scfg, sd, precision="fast")
    with torch.no_grad():
        logits, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(),
                          label_embeddings=labels.cuda())
    err = (logits.cpu() - g["logits"]).abs().max().item()
    print(f"fast mode: logits err {err:.2e} (logit std {g['logits'].std().item():.2f})")
    assert err <= 0.1


def test_label_cache_and_repack():
    """Second call reuses the projected labels; an in-place weight update invalidates both caches."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd)
    lab = labels.cuda()
    with torch.no_grad():
        a, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=lab)
        b, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=lab)
        assert torch.equal(a, b)
        model.output_layer[-1].bias.add_(1.0)
        c, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=lab)
    assert (c - a - 1.0).abs().max().item() < 1e-5


def test_host_inputs_and_errors():
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd)
    with torch.no_grad():
        logits, _ = model(sequence_onehots=onehots, sequence_lengths=lengths, label_embeddings=labels)  # host tensors
        assert (logits.cpu() - g["logits"]).abs().max().item() <= LOGIT_TOL
        with pytest.raises(ValueError):
            model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda())
        with pytest.raises(ValueError):
            model(label_embeddings=labels.cuda())
        emb = model.sequence_encoder.get_embeddings(onehots.cuda(), lengths.cuda())
        again, _ = model(sequence_embeddings=emb, label_embeddings=labels.cuda())
        assert torch.equal(again, logits)
