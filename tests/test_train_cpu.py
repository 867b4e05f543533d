"""CPU tests of the training step: (1) the travelling training oracle is pinned against the reference's own ProtNote class
in train mode, (2) the primitive sequencing of protnote_b200/train.py - run here with the torch stand-in for the CUDA
primitives - reproduces that oracle's logits, gradients and running statistics."""
import copy

import pytest
import torch

from oracle.cases import CASES
from oracle.protnote_oracle import ScorerCfg, synth_state_dict
from oracle.ref_import import import_reference, reference_available
from oracle.train_ops import TorchOps
from oracle.train_oracle import synth_targets, train_step_oracle
from protnote_b200 import train as pn_train
from tests.helpers import build_b200_model


def _problem(name="tiny_concat", B=6, L=10, seed=5):
    ecfg, scfg, *_ = CASES[name]
    sd = synth_state_dict(ecfg, scfg, seed=CASES[name][6], calib_T=64)
    g = torch.Generator().manual_seed(seed)
    P_f = torch.randn(B, scfg.protein_embedding_dim, generator=g)
    L_f = torch.randn(L, scfg.label_embedding_dim, generator=g)
    return ecfg, scfg, sd, P_f, L_f, synth_targets(B, L, seed)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_train_oracle_matches_reference():
    ecfg, scfg, sd, P_f, L_f, y = _problem()
    from oracle.make_golden import build_reference_model
    ref = build_reference_model(ecfg, scfg, sd).double().train()
    logits, _ = ref(sequence_embeddings=P_f.double(), label_embeddings=L_f.double())
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.double())
    loss.backward()
    o_logits, o_loss, o_grads, o_stats = train_step_oracle(sd, P_f, L_f, y, scfg)
    assert (logits.detach() - o_logits).abs().max() < 1e-10
    assert abs(float(loss) - float(o_loss)) < 1e-12
    named = dict(ref.named_parameters())
    checked = 0
    for k, g in o_grads.items():
        assert (named[k].grad - g).abs().max() <= 1e-10 * max(1.0, float(g.abs().max())), k
        checked += 1
    assert checked >= 20
    bufs = dict(ref.named_buffers())
    for k, v in o_stats.items():
        assert (bufs[k] - v).abs().max() < 1e-10, k


def _ours_vs_oracle(scfg, ecfg, sd, P_f, L_f, y, tol):
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    ops = TorchOps(torch.float64)
    logits, ctx = pn_train.forward_train(ops, None, model, P_f.double(), L_f.double())
    o_logits, o_loss, o_grads, o_stats = train_step_oracle(sd, P_f, L_f, y, scfg)
    assert (logits - o_logits).abs().max() < tol
    # BCE-with-logits (mean reduction) gradient w.r.t. the logits
    g_logits = (torch.sigmoid(logits) - y.double()) / logits.numel()
    grads = pn_train.backward_train(ops, None, ctx, g_logits)
    named = dict(model.named_parameters())
    for k, g in o_grads.items():
        got = grads[named[k]]
        assert got.shape == g.shape, k
        assert (got - g).abs().max() <= tol * max(1.0, float(g.abs().max())), k
    bufs = dict(model.named_buffers())
    for k, v in o_stats.items():
        assert (bufs[k] - v).abs().max() < tol, k
    for k, b in bufs.items():
        if k.endswith("num_batches_tracked") and not k.startswith("sequence_encoder"):
            assert int(b) == 1, k


def test_sequencing_matches_oracle_tiny():
    ecfg, scfg, sd, P_f, L_f, y = _problem()
    _ours_vs_oracle(scfg, ecfg, sd, P_f, L_f, y, 1e-9)


def test_sequencing_matches_oracle_two_layer_mlp():
    ecfg, _, *_ = CASES["tiny_concat"]
    scfg = ScorerCfg(protein_embedding_dim=72, label_embedding_dim=40, latent_dim=32,
                     output_mlp_hidden_dim_scale_factor=2, output_mlp_num_layers=2,
                     projection_head_num_layers=2, projection_head_hidden_dim_scale_factor=2)
    sd = synth_state_dict(ecfg, scfg, seed=11, calib_T=64)
    g = torch.Generator().manual_seed(3)
    P_f, L_f = torch.randn(4, 72, generator=g), torch.randn(7, 40, generator=g)
    _ours_vs_oracle(scfg, ecfg, sd, P_f, L_f, synth_targets(4, 7, 3), 1e-9)


def test_autograd_function_delivers_parameter_gradients():
    ecfg, scfg, sd, P_f, L_f, y = _problem(B=5, L=8, seed=8)
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    before = copy.deepcopy(model.state_dict())
    logits = pn_train.train_logits(model, P_f.double(), L_f.double(), ops=TorchOps(torch.float64))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.double())
    loss.backward()
    _, o_loss, o_grads, _ = train_step_oracle(before, P_f, L_f, y, scfg)
    assert abs(float(loss) - float(o_loss)) < 1e-10
    named = dict(model.named_parameters())
    for k, g in o_grads.items():
        assert named[k].grad is not None, k
        assert (named[k].grad - g).abs().max() <= 1e-9 * max(1.0, float(g.abs().max())), k
    assert all(p.grad is None for n, p in named.items() if n.startswith("sequence_encoder."))


@pytest.mark.parametrize("name", ["train_tiny", "train_tiny_wide"])
def test_train_oracle_matches_committed_golden(name):
    """tests/golden/train_*.pt were produced by the reference's own ProtNote class in train mode (oracle/make_golden_train.py)."""
    import os
    from oracle.make_golden_train import train_inputs
    from tests.helpers import GOLDEN_DIR, weight_checksum
    g = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"))
    ecfg, scfg, sd, P_f, L_f, y = train_inputs(name)
    assert abs(weight_checksum(sd) - g["weights_checksum"]) <= 1e-6 * max(1.0, abs(g["weights_checksum"]))
    logits, loss, grads, stats = train_step_oracle(sd, P_f, L_f, y, scfg)
    assert (logits - g["logits"]).abs().max() < 1e-9
    assert abs(float(loss) - g["loss"]) < 1e-10
    assert set(grads) == set(g["grads"])
    for k, v in g["grads"].items():
        assert (grads[k] - v).abs().max() <= 1e-9 * max(1.0, float(v.abs().max())), k
    for k, v in g["running"].items():
        assert (stats[k] - v).abs().max() < 1e-9, k


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_train_mode_encoder_oracle_matches_reference():
    """The frozen encoder in .train() mode (batch-statistic BatchNorm, running statistics updated) - reference class vs oracle."""
    from oracle.make_golden import build_reference_model
    from oracle.protnote_oracle import synth_inputs
    from oracle.train_oracle import proteinfer_embeddings_train
    ecfg, scfg, *_ = CASES["tiny_concat"]
    sd = synth_state_dict(ecfg, scfg, seed=42, calib_T=64)
    onehots, lengths, _ = synth_inputs(5, 90, 4, ecfg, scfg, ragged=True, seed=31)
    ref = build_reference_model(ecfg, scfg, sd).double().train()
    with torch.no_grad():
        want = ref.sequence_encoder.get_embeddings(onehots.double(), lengths)
    got, stats = proteinfer_embeddings_train(sd, onehots, lengths, ecfg)
    assert (got - want).abs().max() < 1e-10
    bufs = dict(ref.named_buffers())
    assert len(stats) == 4 * ecfg.num_resnet_blocks
    for k, v in stats.items():
        assert (bufs[k] - v).abs().max() < 1e-10, k


def test_training_path_refuses_what_it_does_not_implement():
    """Unsupported options fail loudly (no silent fallback): dropout > 0, a fusion other than 'concatenation', an output MLP
    without BatchNorm; and the product module itself has no CPU path in training mode either."""
    from protnote_b200._lib import ProtnoteB200Error
    ecfg, scfg, sd, P_f, L_f, y = _problem()
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    ops = TorchOps(torch.float64)
    for m in model.output_layer.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.1
    with pytest.raises(NotImplementedError):
        pn_train.forward_train(ops, None, model, P_f.double(), L_f.double())
    for m in model.output_layer.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    prod_cfg = ScorerCfg(protein_embedding_dim=72, label_embedding_dim=40, latent_dim=32,
                         output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3, projection_head_num_layers=4,
                         projection_head_hidden_dim_scale_factor=3, feature_fusion="concatenation_prod")
    sd_prod = synth_state_dict(ecfg, prod_cfg, seed=3, calib_T=64)
    prod = build_b200_model(ecfg, prod_cfg, sd_prod, device="cpu").double().train()
    with pytest.raises(NotImplementedError):
        pn_train.forward_train(ops, None, prod, P_f.double(), L_f.double())
    nobn_cfg = ScorerCfg(protein_embedding_dim=72, label_embedding_dim=40, latent_dim=32,
                         output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3, projection_head_num_layers=4,
                         projection_head_hidden_dim_scale_factor=3, output_mlp_batchnorm=False)
    sd_nobn = synth_state_dict(ecfg, nobn_cfg, seed=4, calib_T=64)
    nobn = build_b200_model(ecfg, nobn_cfg, sd_nobn, device="cpu").double().train()
    with pytest.raises(NotImplementedError):
        pn_train.forward_train(ops, None, nobn, P_f.double(), L_f.double())
    # the product module: CPU tensors in training mode -> error, never a torch fallback
    cpu_model = build_b200_model(ecfg, scfg, sd, device="cpu").train()
    with pytest.raises(ProtnoteB200Error):
        cpu_model(sequence_embeddings=P_f, label_embeddings=L_f)
    with pytest.raises(ValueError):
        cpu_model(sequence_embeddings=P_f)


def test_single_row_batchnorm_raises_like_torch():
    ecfg, scfg, sd, P_f, L_f, y = _problem()
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        pn_train.forward_train(TorchOps(torch.float64), None, model, P_f[:1].double(), L_f.double())
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        pn_train.forward_train(TorchOps(torch.float64), None, model, P_f.double(), L_f[:1].double())


# ---------------------------------------------------------------------------------------------------- fused loss (N4)
LOSSES = [dict(loss="bce"), dict(loss="bce", reduction="sum"), dict(loss="bce", pos_weight="vector"),
          dict(loss="focal", gamma=2.0, alpha=0.25), dict(loss="focal", gamma=1.0, alpha=-1.0, label_smoothing=0.1),
          dict(loss="focal", gamma=3.0, alpha=0.6, reduction="sum")]


def _loss_kw(kw, L, dtype=torch.float64):
    kw = dict(kw)
    if kw.get("pos_weight") == "vector":
        kw["pos_weight"] = (torch.rand(L, generator=torch.Generator().manual_seed(2)) * 3 + 0.5).to(dtype)
    return kw


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_focal_loss_oracle_matches_reference_class(capsys):
    """oracle.train_oracle.focal_loss == the reference's FocalLoss module (protnote/utils/losses.py:171-213), values and
    gradients, over every option combination the fused kernel implements."""
    import_reference()
    from protnote.utils.losses import FocalLoss
    from oracle.train_oracle import focal_loss
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(7, 33, generator=g, dtype=torch.float64) * 4).requires_grad_(True)
    t = (torch.rand(7, 33, generator=g) < 0.2).double()
    for alpha, gamma, ls, red in ((0.25, 2.0, 0.0, "mean"), (-1.0, 1.0, 0.1, "mean"), (0.6, 3.0, 0.05, "sum"), (0.5, 0.5, 0.0, "mean")):
        ref = FocalLoss(alpha=alpha, gamma=gamma, reduction=red, label_smoothing=ls)(x, t)
        (gr,) = torch.autograd.grad(ref, x)
        ours = focal_loss(x, t, alpha, gamma, ls, red)
        (go,) = torch.autograd.grad(ours, x)
        assert abs(float(ref) - float(ours)) < 1e-13 and (gr - go).abs().max() < 1e-13


@pytest.mark.parametrize("kw", LOSSES, ids=lambda k: "-".join(f"{a}={b}" for a, b in k.items()))
def test_fused_loss_sequencing_matches_oracle(kw):
    """train.train_loss (loss + gradient seed produced by the last forward primitive) == loss(logits).backward() of the
    oracle, run with the torch stand-in for the primitives."""
    ecfg, scfg, sd, P_f, L_f, y = _problem(B=5, L=9, seed=21)
    kw = _loss_kw(kw, L_f.shape[0])
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    loss, logits = pn_train.train_loss(model, P_f.double(), L_f.double(), y.double(), ops=TorchOps(torch.float64), **kw)
    assert not logits.requires_grad
    (loss * 1.0).backward()
    o_logits, o_loss, o_grads, _ = train_step_oracle(sd, P_f, L_f, y, scfg, **kw)
    assert (logits - o_logits).abs().max() < 1e-9
    assert abs(float(loss) - float(o_loss)) <= 1e-6 * max(1.0, abs(float(o_loss)))      # the fused loss leaves as fp32
    named = dict(model.named_parameters())
    for k, g in o_grads.items():
        assert (named[k].grad - g).abs().max() <= 1e-9 * max(1.0, float(g.abs().max())), k
