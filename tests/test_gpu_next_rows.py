"""GPU tests of the rows SURVEY.md section 8(f) ranks next to the hot path:
N1 device-side evaluation post-processing (sigmoid, per-label TP / FN / FP, top-k) against the torch ops
   ProtNoteTrainer.evaluate runs per batch (ProtNoteTrainer.py:522-537, calculate_tp_fn_fp :61-83);
N2 token-id input (1 byte per residue) against the one-hot path - bit-identical;
and the plumbing case of BASELINE.json configs[0]: 128 sequences x 256 aa x 100 label rows in batches of 8, driven the
way ProtNoteTrainer.evaluation_step drives the module (ProtNoteTrainer.py:247-292), checked against the CPU oracle."""
import pytest
import torch

from oracle.cases import CASES
from oracle.protnote_oracle import protnote_forward, synth_inputs, synth_state_dict
from tests.helpers import build_b200_model, load_case, topk_agree

pytestmark = pytest.mark.gpu


def _tp_fn_fp(probs, labels, threshold):
    preds = (probs >= threshold).float()
    return (preds * labels).sum(0), ((1 - preds) * labels).sum(0), (preds * (1 - labels)).sum(0)


@pytest.mark.parametrize("B,L,kind", [(37, 300, "int64"), (8, 32768, "float"), (1, 5, "int64")])
def test_postprocess_matches_trainer_ops(B, L, kind):
    from protnote_b200 import native
    g = torch.Generator().manual_seed(B + L)
    logits = (torch.randn(B, L, generator=g) * 3).cuda()
    labels = (torch.rand(B, L, generator=g) < 0.1)
    labels = (labels.long() if kind == "int64" else labels.float()).cuda()
    k = min(10, L)
    out = native.postprocess(logits, labels, threshold=0.5, want_probabilities=True, topk=k)
    probs = torch.sigmoid(logits)
    assert float((out["probabilities"] - probs).abs().max()) < 2e-7
    # counts are exactly those of the probabilities the kernel produced ...
    tp, fn, fp = _tp_fn_fp(out["probabilities"], labels.float(), 0.5)
    assert torch.equal(out["tp"], tp) and torch.equal(out["fn"], fn) and torch.equal(out["fp"], fp)
    # ... and differ from torch's only where a probability is within rounding of the threshold
    near = int(((probs - 0.5).abs() < 1e-6).sum())
    rtp, rfn, rfp = _tp_fn_fp(probs, labels.float(), 0.5)
    assert float((tp - rtp).abs().sum() + (fn - rfn).abs().sum() + (fp - rfp).abs().sum()) <= 2 * near
    tv, ti = logits.topk(k, dim=1)
    assert torch.equal(out["topk_values"], tv)
    assert torch.equal(out["topk_indices"].long(), ti)          # continuous random logits: no ties
    # counts accumulate across batches
    out2 = native.postprocess(logits, labels, threshold=0.5, counts=(out["tp"], out["fn"], out["fp"]))
    assert torch.equal(out2["tp"], 2 * tp)


def test_topk_ties_are_ordered_by_index():
    from protnote_b200 import native
    logits = torch.tensor([[1.0, 3.0, 3.0, -1.0, 3.0, 0.5]]).cuda()
    out = native.postprocess(logits, topk=4)
    assert out["topk_indices"].tolist() == [[1, 2, 4, 0]]
    assert out["topk_values"].tolist() == [[3.0, 3.0, 3.0, 1.0]]


@pytest.mark.parametrize("name", ["tiny_concat", "tiny_long"])
def test_token_input_is_bit_identical_to_onehot(name):
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case(name)
    model = build_b200_model(ecfg, scfg, sd)
    tokens = onehots.argmax(1)                     # padding columns are all-zero -> id 0, masked by the lengths
    with torch.no_grad():
        e1 = model.sequence_encoder.get_embeddings(onehots.cuda(), lengths.cuda())
        e2 = model.sequence_encoder.get_embeddings_from_tokens(tokens.cuda(), lengths.cuda())
        e3 = model.sequence_encoder.get_embeddings_from_tokens(tokens.to(torch.uint8), lengths)      # host tensors
        l1, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
        l2, _ = model(sequence_tokens=tokens.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
    assert torch.equal(e1, e2) and torch.equal(e1, e3)
    assert torch.equal(l1, l2)
    assert float((l2.cpu() - g["logits"]).abs().max()) < 1e-4


def test_plumbing_configuration_through_an_evaluation_loop():
    """BASELINE.json configs[0]: 128 synthetic 256-aa sequences x 100 cached label embeddings, batches of 8."""
    from protnote_b200 import native
    ecfg, scfg, *_ = CASES["tiny_concat"]
    sd = synth_state_dict(ecfg, scfg, seed=42, calib_T=128)
    N, T, L, BS = 128, 256, 100, 8
    onehots, lengths, label_embeddings = synth_inputs(N, T, L, ecfg, scfg, ragged=True, seed=77)
    multihots = (torch.rand(N, L, generator=torch.Generator().manual_seed(1)) < 0.05).long()
    model = build_b200_model(ecfg, scfg, sd)
    loss_fn = torch.nn.BCEWithLogitsLoss()
    counts = tuple(torch.zeros(L, device="cuda") for _ in range(3))
    all_logits, losses = [], []
    with torch.no_grad():
        for s in range(0, N, BS):
            # ProtNoteTrainer.evaluation_step: _to_device, autocast, model(**inputs), loss (ProtNoteTrainer.py:272-290)
            x, lens = onehots[s:s + BS].cuda(), lengths[s:s + BS].cuda()
            y, lab = multihots[s:s + BS].cuda(), label_embeddings.cuda()
            with torch.autocast("cuda"):
                logits, _ = model(sequence_onehots=x, sequence_lengths=lens, label_embeddings=lab, save_embeddings=False)
                losses.append(float(loss_fn(logits, y.float())))
            native.postprocess(logits, y, threshold=0.5, counts=counts)
            all_logits.append(logits.float().cpu())
    got = torch.cat(all_logits)
    ref = protnote_forward(sd, onehots, lengths, label_embeddings, ecfg, scfg)
    assert float((got - ref).abs().max()) < 1e-4
    assert topk_agree(ref, got, 10, 1e-4)
    ref_loss = [float(loss_fn(ref[s:s + BS], multihots[s:s + BS].float())) for s in range(0, N, BS)]
    assert max(abs(a - b) for a, b in zip(losses, ref_loss)) < 1e-5
    probs = torch.sigmoid(ref)
    decided = (probs - 0.5).abs() > 1e-4                       # pairs whose prediction does not hinge on rounding
    tp, fn, fp = _tp_fn_fp(torch.sigmoid(got), multihots.float(), 0.5)
    assert torch.equal(counts[0].cpu(), tp) and torch.equal(counts[1].cpu(), fn) and torch.equal(counts[2].cpu(), fp)
    rtp, rfn, rfp = _tp_fn_fp(probs, multihots.float(), 0.5)
    assert float((tp - rtp).abs().sum() + (fn - rfn).abs().sum() + (fp - rfp).abs().sum()) <= 2 * float((~decided).sum())


def test_label_projection_round_trip_on_disk(tmp_path):
    """N3: the projected label halves stored on disk reproduce the logits bit for bit, skip W_l, and are refused for
    other weights or other embeddings."""
    from protnote_b200 import native
    from protnote_b200._lib import ProtnoteB200Error
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd)
    path = str(tmp_path / "labels.proj.pt")
    with torch.no_grad():
        ref, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
        model.save_label_projection(path, labels)
        fresh = build_b200_model(ecfg, scfg, sd)
        lab_dev = fresh.load_label_projection(path, labels)
        P_f = fresh.sequence_encoder.get_embeddings(onehots.cuda(), lengths.cuda())
        before = native.launch_count()
        got, _ = fresh(sequence_embeddings=P_f, label_embeddings=lab_dev)
        cached_launches = native.launch_count() - before
        fresh._label_cache = None
        before = native.launch_count()
        again, _ = fresh(sequence_embeddings=P_f, label_embeddings=lab_dev)
        uncached_launches = native.launch_count() - before
    assert torch.equal(got, ref) and torch.equal(again, ref)
    assert cached_launches < uncached_launches          # W_l and the label half of layer 1 were not launched
    other = build_b200_model(ecfg, scfg, {k: (v * 1.01 if k.startswith("W_l.0") else v) for k, v in sd.items()})
    with pytest.raises(ProtnoteB200Error):
        other.load_label_projection(path, labels)
    with pytest.raises(ProtnoteB200Error):
        fresh.load_label_projection(path, labels * 1.5)
