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


def _variant_cfg(**kw):
    base = dict(protein_embedding_dim=72, label_embedding_dim=40, latent_dim=32, output_mlp_hidden_dim_scale_factor=3,
                output_mlp_num_layers=3, projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3)
    base.update(kw)
    return ScorerCfg(**base)


def test_training_path_refuses_what_it_does_not_implement():
    """Unsupported options fail loudly (no silent fallback): a Dropout in front of the first Linear of an MLP, the fused
    loss with FEATURE_FUSION 'similarity' (no output MLP to fuse it into); and the product module itself has no CPU path in
    training mode."""
    from protnote_b200._lib import ProtnoteB200Error
    ecfg, scfg, sd, P_f, L_f, y = _problem()
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    ops = TorchOps(torch.float64)
    model.output_layer = torch.nn.Sequential(torch.nn.Dropout(0.1), *list(model.output_layer))     # get_mlp(input_dropout=)
    with pytest.raises(NotImplementedError):
        pn_train.forward_train(ops, None, model, P_f.double(), L_f.double())
    sim_cfg = _variant_cfg(feature_fusion="similarity")
    sim = build_b200_model(ecfg, sim_cfg, synth_state_dict(ecfg, sim_cfg, seed=3, calib_T=64), device="cpu").double().train()
    with pytest.raises(NotImplementedError, match="similarity"):
        pn_train.train_loss(sim, P_f.double(), L_f.double(), y.double(), ops=ops)
    # the product module: CPU tensors in training mode -> error, never a torch fallback
    cpu_model = build_b200_model(ecfg, scfg, sd, device="cpu").train()
    with pytest.raises(ProtnoteB200Error):
        cpu_model(sequence_embeddings=P_f, label_embeddings=L_f)
    with pytest.raises(ValueError):
        cpu_model(sequence_embeddings=P_f)


VARIANTS = {"no_batchnorm": dict(output_mlp_batchnorm=False),
            "diff": dict(feature_fusion="concatenation_diff"),
            "diff_no_batchnorm": dict(feature_fusion="concatenation_diff", output_mlp_batchnorm=False),
            "two_layers_no_batchnorm": dict(output_mlp_num_layers=2, output_mlp_batchnorm=False),
            "similarity": dict(feature_fusion="similarity", temperature=0.07),
            "prod": dict(feature_fusion="concatenation_prod"),
            "prod_no_batchnorm": dict(feature_fusion="concatenation_prod", output_mlp_batchnorm=False),
            "one_layer_prod": dict(feature_fusion="concatenation_prod", output_mlp_num_layers=1),
            "one_layer": dict(output_mlp_num_layers=1),
            "one_layer_diff_no_batchnorm": dict(output_mlp_num_layers=1, output_mlp_batchnorm=False,
                                                feature_fusion="concatenation_diff")}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_sequencing_matches_oracle_variants(variant):
    """OUTPUT_MLP_BATCHNORM False (Linear+bias -> ReLU layers: the BatchNorm primitives run with a fixed affine state and
    cleared backward sums) and FEATURE_FUSION concatenation_diff (folded into the two layer-1 factors) through the same
    primitive sequence: logits, every parameter gradient (incl. the hidden biases and the third weight block) and the
    running statistics against the autograd oracle."""
    ecfg, _, *_ = CASES["tiny_concat"]
    scfg = _variant_cfg(**VARIANTS[variant])
    sd = synth_state_dict(ecfg, scfg, seed=13, calib_T=64)
    g = torch.Generator().manual_seed(17)
    P_f, L_f = torch.randn(5, 72, generator=g), torch.randn(9, 40, generator=g)
    _ours_vs_oracle(scfg, ecfg, sd, P_f, L_f, synth_targets(5, 9, 17), 1e-9)
    if not scfg.output_mlp_batchnorm:       # the hidden biases are parameters of this variant: they must have been checked
        hidden_biases = [k for k in train_step_oracle(sd, P_f, L_f, synth_targets(5, 9, 17), scfg)[2]
                         if k.startswith("output_layer.") and k.endswith(".bias")]
        assert len(hidden_biases) == scfg.output_mlp_num_layers + 1


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_train_oracle_variants_match_reference(variant):
    """The oracle's statement of those variants is the reference class's (train mode, fp64, autograd)."""
    from oracle.make_golden import build_reference_model
    ecfg, _, *_ = CASES["tiny_concat"]
    scfg = _variant_cfg(**VARIANTS[variant])
    sd = synth_state_dict(ecfg, scfg, seed=13, calib_T=64)
    g = torch.Generator().manual_seed(17)
    P_f, L_f = torch.randn(5, 72, generator=g), torch.randn(9, 40, generator=g)
    y = synth_targets(5, 9, 17)
    ref = build_reference_model(ecfg, scfg, sd).double().train()
    logits, _ = ref(sequence_embeddings=P_f.double(), label_embeddings=L_f.double())
    torch.nn.functional.binary_cross_entropy_with_logits(logits, y.double()).backward()
    o_logits, _, o_grads, o_stats = train_step_oracle(sd, P_f, L_f, y, scfg)
    assert (logits.detach() - o_logits).abs().max() < 1e-10
    named = dict(ref.named_parameters())
    trainable = [k for k, p in named.items() if not k.startswith("sequence_encoder.")]
    assert set(o_grads) == set(trainable)
    for k, gr in o_grads.items():
        assert (named[k].grad - gr).abs().max() <= 1e-10 * max(1.0, float(gr.abs().max())), k
    bufs = dict(ref.named_buffers())
    for k, v in o_stats.items():
        assert (bufs[k] - v).abs().max() < 1e-10, k


def _dropout_masks(model, P_rows, L_rows, base_seed, rank=0):
    """The multipliers of every active Dropout inside W_p / W_l / output_layer for the step seeded with `base_seed`:
    the sites of protnote_b200.train (site -> seed, p, width) and the statement of the device mask (oracle.train_ops)."""
    from oracle.train_ops import dropout_multiplier
    rows = {"p": P_rows, "l": L_rows, "o": P_rows * L_rows}
    return {(t, i): dropout_multiplier(seed, rows[t], width, p)
            for (t, i), (seed, p, width) in pn_train.dropout_sites(model, base_seed, rank).items()}


@pytest.mark.parametrize("variant", ["default", "no_batchnorm", "two_layers", "prod"])
def test_output_mlp_dropout_sequencing_matches_oracle(variant):
    """OUTPUT_MLP_DROPOUT > 0 (Dropout after every hidden ReLU and after the last Linear of W_p / W_l, after every hidden
    ReLU but the last of output_layer): the forward masks the activations, the backward the gradients, with masks that are
    a pure function of (seed, row, column) - against the autograd oracle multiplying with the same masks."""
    ecfg, _, *_ = CASES["tiny_concat"]
    kw = {"default": {}, "no_batchnorm": dict(output_mlp_batchnorm=False), "two_layers": dict(output_mlp_num_layers=2),
          "prod": dict(feature_fusion="concatenation_prod")}[variant]
    scfg = _variant_cfg(output_mlp_dropout=0.3, **kw)
    sd = synth_state_dict(ecfg, scfg, seed=13, calib_T=64)
    g = torch.Generator().manual_seed(19)
    P_f, L_f = torch.randn(6, 72, generator=g), torch.randn(10, 40, generator=g)
    y = synth_targets(6, 10, 19)
    model = build_b200_model(ecfg, scfg, sd, device="cpu").double().train()
    ops = TorchOps(torch.float64)
    logits, ctx = pn_train.forward_train(ops, None, model, P_f.double(), L_f.double(), drop_seed=987654321)
    masks = _dropout_masks(model, 6, 10, 987654321)
    n_hidden = scfg.output_mlp_num_layers
    assert set(masks) == ({("p", i) for i in range(4)} | {("l", i) for i in range(4)} | {("o", j) for j in range(n_hidden - 1)})
    o_logits, _, o_grads, o_stats = train_step_oracle(sd, P_f, L_f, y, scfg, masks=masks)
    assert (logits - o_logits).abs().max() < 1e-9
    plain, *_ = train_step_oracle(sd, P_f, L_f, y, scfg)
    assert (plain - o_logits).abs().max() > 1e-2                      # the masks really changed the step
    grads = pn_train.backward_train(ops, None, ctx, (torch.sigmoid(logits) - y.double()) / logits.numel())
    named = dict(model.named_parameters())
    for k, gr in o_grads.items():
        assert (grads[named[k]] - gr).abs().max() <= 1e-9 * max(1.0, float(gr.abs().max())), k
    bufs = dict(model.named_buffers())
    for k, v in o_stats.items():
        assert (bufs[k] - v).abs().max() < 1e-9, k
    # the base seed comes from torch's CPU generator when it is not given: same torch seed, same step
    torch.manual_seed(5)
    a, _ = pn_train.forward_train(ops, None, model, P_f.double(), L_f.double(), update_running=False)
    torch.manual_seed(5)
    b, _ = pn_train.forward_train(ops, None, model, P_f.double(), L_f.double(), update_running=False)
    c, _ = pn_train.forward_train(ops, None, model, P_f.double(), L_f.double(), update_running=False)
    assert torch.equal(a, b) and not torch.equal(a, c)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_dropout_mask_placement_is_the_reference_modules():
    """The oracle multiplies where the reference has its nn.Dropout modules: replace every Dropout of the reference
    class's W_p / W_l / output_layer (in module order) by a fixed multiplier and compare logits and gradients."""
    from oracle.make_golden import build_reference_model
    from oracle.train_ops import dropout_multiplier

    class Mul(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            return x * self.m

    ecfg, _, *_ = CASES["tiny_concat"]
    scfg = _variant_cfg(output_mlp_dropout=0.25)
    sd = synth_state_dict(ecfg, scfg, seed=13, calib_T=64)
    g = torch.Generator().manual_seed(19)
    P_f, L_f = torch.randn(5, 72, generator=g), torch.randn(7, 40, generator=g)
    y = synth_targets(5, 7, 19)
    ref = build_reference_model(ecfg, scfg, sd).double().train()
    masks, seed = {}, 100
    for tag, part, rows in (("p", ref.W_p, 5), ("l", ref.W_l, 7), ("o", ref.output_layer, 35)):
        mods, layer, width = list(part), -1, None
        for i, m in enumerate(mods):
            if isinstance(m, torch.nn.Linear):
                layer, width = layer + 1, m.out_features
            if isinstance(m, torch.nn.Dropout):
                assert m.p == 0.25
                seed += 1
                masks[(tag, layer)] = dropout_multiplier(seed, rows, width, 0.25)
                part[i] = Mul(masks[(tag, layer)])
    assert len(masks) == 4 + 4 + 2
    logits, _ = ref(sequence_embeddings=P_f.double(), label_embeddings=L_f.double())
    torch.nn.functional.binary_cross_entropy_with_logits(logits, y.double()).backward()
    o_logits, _, o_grads, _ = train_step_oracle(sd, P_f, L_f, y, scfg, masks=masks)
    assert (logits.detach() - o_logits).abs().max() < 1e-10
    named = dict(ref.named_parameters())
    for k, gr in o_grads.items():
        assert (named[k].grad - gr).abs().max() <= 1e-10 * max(1.0, float(gr.abs().max())), k


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_embedding_dropouts_follow_the_reference_rng_stream():
    """SEQUENCE_EMBEDDING_DROPOUT / LABEL_EMBEDDING_DROPOUT > 0 (ProtNote.py:83-86: W_p / W_l become Sequential(Dropout, MLP),
    keys W_p.1.* / W_l.1.*) together with the label noise: with the same torch seed the training forward draws the same
    noise and the same masks in the same order as the reference module, so logits and gradients agree to rounding."""
    ProtNoteRef, _, _ = import_reference()
    from protnote_b200.ProtNote import ProtNote
    kw = dict(protein_embedding_dim=72, label_embedding_dim=40, latent_dim=32, label_embedding_pooling_method="mean",
              sequence_encoder=None, label_encoder=None, inference_descriptions_per_label=1,
              output_mlp_hidden_dim_scale_factor=3, output_mlp_num_layers=3, outout_mlp_add_batchnorm=True,
              projection_head_num_layers=4, projection_head_hidden_dim_scale_factor=3,
              label_encoder_num_trainable_layers=0, train_sequence_encoder=False, feature_fusion="concatenation",
              sequence_embedding_dropout=0.25, label_embedding_dropout=0.4, label_embedding_noising_alpha=20.0)
    torch.manual_seed(3)
    ref = ProtNoteRef(**kw).double().train()
    ours = ProtNote(**kw)
    ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    ours = ours.double().train()
    assert pn_train.input_dropout(ours.W_p) == 0.25 and pn_train.input_dropout(ours.W_l) == 0.4
    g = torch.Generator().manual_seed(23)
    P_f, L_f = torch.randn(6, 72, generator=g).double(), torch.randn(11, 40, generator=g).double()
    counts = torch.full((11,), 7)
    y = synth_targets(6, 11, 23).double()

    torch.manual_seed(1234)
    r_logits, _ = ref(sequence_embeddings=P_f, label_embeddings=L_f, label_token_counts=counts)
    torch.nn.functional.binary_cross_entropy_with_logits(r_logits, y).backward()

    torch.manual_seed(1234)
    # what ProtNote._forward_train does ahead of the primitives (label noise, ProtNote.py:219-240) ...
    noised = L_f + (2 * torch.rand_like(L_f) - 1) * (20.0 / 40 ** 0.5)
    # ... and the primitive sequence (with the torch stand-in), which applies the two dropouts
    logits = pn_train.train_logits(ours, P_f, noised, ops=TorchOps(torch.float64))
    torch.nn.functional.binary_cross_entropy_with_logits(logits, y).backward()
    assert (logits.detach() - r_logits.detach()).abs().max() < 1e-9
    r_named, named = dict(ref.named_parameters()), dict(ours.named_parameters())
    for k, p in r_named.items():
        assert (named[k].grad - p.grad).abs().max() <= 1e-9 * max(1.0, float(p.grad.abs().max())), k
    # dropout really was active: the same inputs without it give different logits
    torch.manual_seed(1234)
    ours.W_p[0].p = ours.W_l[0].p = 0.0
    plain = pn_train.train_logits(ours, P_f, noised, ops=TorchOps(torch.float64))
    assert (plain.detach() - logits.detach()).abs().max() > 1e-3


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
