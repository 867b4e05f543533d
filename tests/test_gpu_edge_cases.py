"""Edge cases and size-independent properties of the sm_100a path (through the C ABI), checked against the oracle."""
import dataclasses

import pytest
import torch

from oracle.cases import CASES
from oracle.protnote_oracle import (EncoderCfg, ScorerCfg, proteinfer_embeddings, proteinfer_logits, protnote_forward,
                                    synth_inputs, synth_state_dict)
from tests.helpers import build_b200_model, load_case

pytestmark = pytest.mark.gpu
TOL = 1e-4

TINY_E = CASES["tiny_concat"][0]
TINY_S = CASES["tiny_concat"][1]


def run(model, onehots, lengths, labels):
    with torch.no_grad():
        out, _ = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("B,T,L", [(1, 1, 1), (1, 9, 3), (2, 129, 33), (3, 128, 1), (2, 257, 130), (1, 1500, 5)])
def test_ragged_shapes_against_oracle(B, T, L):
    """Single residue / single label / tile-boundary lengths / sequences longer than the widest dilation's reach."""
    ecfg = dataclasses.replace(TINY_E, num_resnet_blocks=5) if T > 1000 else TINY_E
    sd = synth_state_dict(ecfg, TINY_S, seed=100 + B + T + L)
    onehots, lengths, labels = synth_inputs(B, T, L, ecfg, TINY_S, ragged=True, seed=7 * T + L)
    if T > 1:
        lengths[-1] = 1                      # a one-residue protein next to a full-length one
        onehots[-1, :, 1:] = 0
    model = build_b200_model(ecfg, TINY_S, sd)
    got = run(model, onehots, lengths, labels)
    ref = protnote_forward(sd, onehots, lengths, labels, ecfg, TINY_S)
    assert got.shape == ref.shape == (B, L)
    assert (got - ref).abs().max().item() <= TOL


def test_padding_content_and_batch_position_do_not_matter():
    """set_padding_to_sentinel semantics (datasets.py:535-569): garbage beyond `length` is ignored, and a protein scored
    alone (unpadded) gets bit-identical logits to the same protein inside a padded batch."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd)
    base = run(model, onehots, lengths, labels)
    noisy = onehots.clone()
    for b in range(onehots.shape[0]):
        noisy[b, :, int(lengths[b]):] = 123.0
    assert torch.equal(run(model, noisy, lengths, labels), base)
    for b in (1, onehots.shape[0] - 1):
        n = int(lengths[b])
        solo = run(model, onehots[b:b + 1, :, :n].contiguous(), lengths[b:b + 1], labels)
        assert torch.equal(solo[0], base[b])


def test_permutation_equivariance_and_chunking_are_bitwise():
    """Pairs are independent: permuting proteins / label rows permutes the logits bit for bit, and so does forcing the
    scorer to work in small chunks (workspace-limited path)."""
    from protnote_b200 import native
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("base_small")
    model = build_b200_model(ecfg, scfg, sd)
    base = run(model, onehots, lengths, labels)
    gen = torch.Generator().manual_seed(3)
    pl = torch.randperm(labels.shape[0], generator=gen)
    pb = torch.randperm(onehots.shape[0], generator=gen)
    assert torch.equal(run(model, onehots, lengths, labels[pl]), base[:, pl])
    assert torch.equal(run(model, onehots[pb], lengths[pb], labels), base[pb])
    native.set_option("chunk_rows", 128)
    try:
        model._label_cache = None
        assert torch.equal(run(model, onehots, lengths, labels), base)
    finally:
        native.set_option("chunk_rows", 0)


@pytest.mark.parametrize("variant", ["diff", "no_batchnorm", "two_layers", "dropout_wrappers", "k3", "one_layer",
                                     "one_layer_prod_k2", "one_layer_no_batchnorm_fast", "mlp_dropout"])
def test_config_variants_against_oracle(variant):
    """Constructor options that change the arithmetic or the state_dict layout (ProtNote.py:83-102,128-138,337-378)."""
    scfg = {
        "diff": dataclasses.replace(TINY_S, feature_fusion="concatenation_diff"),
        "no_batchnorm": dataclasses.replace(TINY_S, output_mlp_batchnorm=False),
        "two_layers": dataclasses.replace(TINY_S, output_mlp_num_layers=2, projection_head_num_layers=1),
        "dropout_wrappers": dataclasses.replace(TINY_S, sequence_embedding_dropout=0.1, label_embedding_dropout=0.2),
        "k3": dataclasses.replace(TINY_S, inference_descriptions_per_label=3),
        # OUTPUT_MLP_NUM_LAYERS 1: the output neuron follows layer 1 (a per-pair dot kernel instead of the scorer GEMMs)
        "one_layer": dataclasses.replace(TINY_S, output_mlp_num_layers=1),
        "one_layer_prod_k2": dataclasses.replace(TINY_S, output_mlp_num_layers=1, feature_fusion="concatenation_prod",
                                                 inference_descriptions_per_label=2),
        "one_layer_no_batchnorm_fast": dataclasses.replace(TINY_S, output_mlp_num_layers=1, output_mlp_batchnorm=False),
        # OUTPUT_MLP_DROPOUT > 0: the Dropout modules are inactive in eval mode and do not move any state_dict index
        "mlp_dropout": dataclasses.replace(TINY_S, output_mlp_dropout=0.3),
    }[variant]
    fast = variant.endswith("_fast")
    sd = synth_state_dict(TINY_E, scfg, seed=900 + len(variant))
    L = 22 if variant == "one_layer_prod_k2" else 21
    onehots, lengths, labels = synth_inputs(4, 77, L, TINY_E, scfg, ragged=True, seed=55)
    model = build_b200_model(TINY_E, scfg, sd, precision="fast" if fast else "strict")
    got = run(model, onehots, lengths, labels)
    ref = protnote_forward(sd, onehots, lengths, labels, TINY_E, scfg)
    assert got.shape == ref.shape
    # the k-row ensemble maps a logit error e to e / (p (1 - p)) at most; the synthetic logits keep p away from 0/1
    assert (got - ref).abs().max().item() <= (0.05 * float(ref.std()) + 0.05 if fast else TOL)


def test_float_inputs_and_encoder_logits():
    """Any float [B,Cin,T] is accepted, and ProteInfer.forward = output_layer(get_embeddings)
    (protein_encoders.py:120-123, the path bin/test_proteinfer.py:303 uses)."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd)
    x = onehots + 0.25 * torch.randn(onehots.shape, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        emb = model.sequence_encoder.get_embeddings(x.cuda(), lengths.cuda()).cpu()
        enc_logits = model.sequence_encoder(x.cuda(), lengths.cuda()).cpu()
    ref_emb = proteinfer_embeddings(sd, x, lengths, ecfg, "sequence_encoder.")
    ref_logits = proteinfer_logits(sd, x, lengths, ecfg, "sequence_encoder.")
    assert (emb - ref_emb).abs().max().item() <= 2e-5
    assert (enc_logits - ref_logits).abs().max().item() <= 2e-5


def test_headline_shape_spot_check():
    """BASELINE.json's shape (1024 aa x 32768 label rows, published architecture) on a slice of the protein axis: a random
    subset of the [B, L] logits must equal the oracle's scores of exactly those (protein, label) pairs."""
    ecfg, scfg = EncoderCfg(), ScorerCfg()
    sd = synth_state_dict(ecfg, scfg, seed=46, calib_T=256)
    B, T, L = 24, 1024, 32768
    onehots, lengths, labels = synth_inputs(B, T, L, ecfg, scfg, ragged=True, seed=99)
    model = build_b200_model(ecfg, scfg, sd)
    got = run(model, onehots, lengths, labels)
    assert got.shape == (B, L) and torch.isfinite(got).all()
    gen = torch.Generator().manual_seed(5)
    rows = torch.tensor([0, 7, 23])
    cols = torch.cat([torch.tensor([0, 127, 128, L - 1]), torch.randint(0, L, (28,), generator=gen)])
    torch.set_num_threads(8)
    ref = protnote_forward(sd, onehots[rows], lengths[rows], labels[cols], ecfg, scfg)
    sub = got[rows][:, cols]
    assert (sub - ref).abs().max().item() <= TOL
    # top-k identity on the full label axis for one protein, against the oracle on the labels that matter
    top = got[0].topk(10).indices
    ref_top = protnote_forward(sd, onehots[:1], lengths[:1], labels[top], ecfg, scfg)[0]
    assert (got[0, top] - ref_top).abs().max().item() <= TOL


@pytest.mark.parametrize("k", [1, 2])
def test_similarity_fusion_against_oracle(k):
    """FEATURE_FUSION 'similarity' (ProtNote.py:281-284): cosine of the projected embeddings / temperature."""
    scfg = dataclasses.replace(TINY_S, feature_fusion="similarity", inference_descriptions_per_label=k, temperature=0.07)
    sd = synth_state_dict(TINY_E, scfg, seed=31 + k)
    onehots, lengths, labels = synth_inputs(5, 90, 26, TINY_E, scfg, ragged=True, seed=77)
    model = build_b200_model(TINY_E, scfg, sd)
    got = run(model, onehots, lengths, labels)
    ref = protnote_forward(sd, onehots, lengths, labels, TINY_E, scfg)
    assert got.shape == ref.shape == (5, 26 // k)
    assert (got - ref).abs().max().item() <= TOL


def test_save_embeddings_contract():
    """forward(save_embeddings=True) returns the joint features and the last hidden layer of every pair on the CPU
    (ProtNote.py:294-303,324-334) and the same logits as the plain call."""
    from oracle.protnote_oracle import joint_features, output_mlp, projection_head
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_prod")
    model = build_b200_model(ecfg, scfg, sd)
    with torch.no_grad():
        plain, empty = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
        logits, emb = model(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda(),
                            save_embeddings=True)
    assert empty == {"output_layer_embeddings": [], "joint_embeddings": []}
    assert torch.equal(plain, logits)
    P_f = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.")
    P_e = projection_head(sd, "W_p", P_f, scfg, torch.float32)
    L_e = projection_head(sd, "W_l", labels, scfg, torch.float32)
    joint = joint_features(P_e, L_e, scfg.feature_fusion)
    hidden = output_mlp(sd, "output_layer", joint, scfg, torch.float32, return_hidden=True)
    assert emb["joint_embeddings"].device.type == "cpu" and emb["joint_embeddings"].shape == joint.shape
    assert (emb["joint_embeddings"] - joint).abs().max().item() <= 2e-5
    assert emb["output_layer_embeddings"].shape == hidden.shape
    assert (emb["output_layer_embeddings"] - hidden).abs().max().item() <= 1e-4


def test_fused_feature_generator_matches_separate_kernel():
    """Option fuse_features=1 builds relu(a[b] + c[l]) inside the first scorer GEMM's operand producer (TMA-staged label
    tiles, generator warps writing the swizzled fp16 planes); it must give bit-identical logits to the default path."""
    from protnote_b200 import native
    ecfg, scfg = EncoderCfg(), ScorerCfg()
    sd = synth_state_dict(ecfg, scfg, seed=46, calib_T=128)
    onehots, lengths, labels = synth_inputs(3, 200, 640, ecfg, scfg, ragged=True, seed=5)
    model = build_b200_model(ecfg, scfg, sd)
    base = run(model, onehots, lengths, labels)
    native.set_option("fuse_features", 1)
    try:
        model._label_cache = None
        fused = run(model, onehots, lengths, labels)
    finally:
        native.set_option("fuse_features", 0)
    assert torch.equal(fused, base)


def test_cta_pair_kernels_match_single_cta_kernels():
    """Option cta2=1 runs every contraction as CTA pairs (tcgen05 cta_group::2, 256-row tiles, each CTA loading half of
    the weight tile, cross-CTA mbarrier protocol).  Same arithmetic per output element -> bit-identical logits."""
    from protnote_b200 import native
    for case in ("tiny_long", "base_small"):
        ecfg, scfg, sd, onehots, lengths, labels, g = load_case(case)
        model = build_b200_model(ecfg, scfg, sd)
        native.set_option("cta2", 0)
        try:
            base = run(model, onehots, lengths, labels)
            native.set_option("cta2", 1)
            model._label_cache = None
            pair = run(model, onehots, lengths, labels)
            # fast mode (where pairs are the default): same arithmetic per output element as well
            model.precision = model.sequence_encoder.precision = "fast"
            model._label_cache = None
            pair_fast = run(model, onehots, lengths, labels)
            native.set_option("cta2", 0)
            model._label_cache = None
            base_fast = run(model, onehots, lengths, labels)
        finally:
            native.set_option("cta2", -1)
        assert torch.equal(pair, base)
        assert torch.equal(pair_fast, base_fast)
        assert (pair - g["logits"]).abs().max().item() <= TOL


def test_split_accumulator_encoder_option_stays_within_tolerance():
    """Option split_corr=1: strict-mode encoder convolutions keep the hi*hi products and the lo corrections in separate
    TMEM buffers (faster, slightly less accurate; off by default)."""
    from protnote_b200 import native
    for case in ("tiny_long", "base_small"):
        ecfg, scfg, sd, onehots, lengths, labels, g = load_case(case)
        model = build_b200_model(ecfg, scfg, sd)
        native.set_option("split_corr", 1)
        try:
            got = run(model, onehots, lengths, labels)
        finally:
            native.set_option("split_corr", 0)
        assert (got - g["logits"]).abs().max().item() <= TOL
