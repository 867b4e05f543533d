"""Round-2 engine features on the GPU: accumulator-truncation compensation, the fp64 protein-side head, option-driven
re-packing, and the pair / single-CTA kernels producing the same logits."""
import pytest
import torch

from oracle.protnote_oracle import projection_head
from tests.helpers import build_b200_model, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture()
def native():
    from protnote_b200 import native as n
    yield n
    for name, value in (("trunc_beta_ppt", 33000), ("promote_k_other", 64), ("promote_k_scorer", 256), ("promote_k_heads", 32),
                        ("promote_k_encoder", 64), ("promote_k_pointwise", 64), ("cta2", 1), ("f64_protein_head", 1)):
        n.set_option(name, value)


def _bias(y, ref):
    d = y.double() - ref
    return float((d * ref.sign()).mean() / ref.abs().mean()), float(d.pow(2).mean().sqrt())


def test_truncation_compensation_removes_the_accumulator_bias(native):
    """The tcgen05 accumulator add rounds toward zero: a strict GEMM promoted every 256 K-elements comes out short by
    ~7e-7 relative; with the per-K-position compensation folded into the packed weights the bias is < 3e-8 and the
    total error is below torch's own fp32 matmul (TF32 off)."""
    g = torch.Generator().manual_seed(0)
    M, N, K = 512, 384, 3072
    x = torch.randn(M, K, generator=g).relu().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    ref = x.double() @ w.double().T
    native.set_option("promote_k_other", 256)
    native.set_option("trunc_beta_ppt", 0)
    bias0, rms0 = _bias(native.linear(x, w, None, native.PN_STRICT), ref)
    native.set_option("trunc_beta_ppt", 33000)
    bias1, rms1 = _bias(native.linear(x, w, None, native.PN_STRICT), ref)
    torch.backends.cuda.matmul.allow_tf32 = False
    _, rms_torch = _bias(x @ w.T, ref)
    print(f"relative bias {bias0:+.2e} -> {bias1:+.2e}; rms {rms0:.2e} -> {rms1:.2e}; torch fp32 rms {rms_torch:.2e}")
    assert -1.2e-6 < bias0 < -4e-7
    assert abs(bias1) < 3e-8
    assert rms1 < 0.6 * rms0 and rms1 < rms_torch


@pytest.mark.parametrize("rows", [5, 70], ids=["skinny_kernel", "tiled_kernel"])
def test_fp64_protein_side_head(native, rows):
    """a[b] (W_p + protein half of output layer 1) carries every protein's contribution to all of its logits: strict mode
    evaluates it in fp64 and only rounds the result to fp32."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("base_small")
    model = build_b200_model(ecfg, scfg, sd)
    P_f = torch.randn(rows, scfg.protein_embedding_dim, generator=torch.Generator().manual_seed(3))
    f64 = torch.float64
    P_e64 = projection_head(sd, "W_p", P_f.double(), scfg, f64)
    W1 = sd["output_layer.0.weight"].double()
    d = scfg.latent_dim
    gm, bt, mu, var = (sd[f"output_layer.1.{k}"].double() for k in ("weight", "bias", "running_mean", "running_var"))
    s = gm / torch.sqrt(var + scfg.bn_eps)
    a64 = (P_e64 @ W1[:, :d].T) * s + (bt - mu * s)
    errs = {}
    for flag in (1, 0):
        native.set_option("f64_protein_head", flag)
        scorer = model._ensure_packed()
        with torch.no_grad():
            P_e, a = scorer.project_sequences(P_f.cuda(), native.PN_STRICT, want_embedding=True)
        errs[flag] = (float((a.cpu().double() - a64).abs().max() / a64.abs().max()),
                      float((P_e.cpu().double() - P_e64).abs().max() / P_e64.abs().max()))
    print(f"relative max error of a / P_e: fp64 head {errs[1]}, tensor-core head {errs[0]}")
    assert errs[1][0] <= 1.2e-7 and errs[1][1] <= 1.2e-7        # = rounding the fp64 result to fp32
    assert errs[0][0] <= 1e-5                                    # the tensor-core head is fp32-grade


def test_engine_option_change_repacks_and_keeps_parity(native):
    """Packed weights carry the chunk structure (truncation compensation): changing a promotion period bumps the option
    epoch, the next forward re-packs, and the logits stay within the bar."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("tiny_concat")
    model = build_b200_model(ecfg, scfg, sd)
    args = dict(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
    with torch.no_grad():
        l0 = model(**args)[0].cpu()
    pack0, enc0 = model._packed, model.sequence_encoder._packed
    native.set_option("promote_k_scorer", 128)
    native.set_option("promote_k_encoder", 32)
    with torch.no_grad():
        l1 = model(**args)[0].cpu()
    assert model._packed is not pack0 and model.sequence_encoder._packed is not enc0
    assert (l0 - g["logits"]).abs().max() <= 1e-4 and (l1 - g["logits"]).abs().max() <= 1e-4
    assert not torch.equal(l0, l1)          # a different summation order really ran


def test_pair_and_single_cta_kernels_agree(native):
    """cta_group::2 pairs (default) and the single-CTA kernel issue the same MMAs in the same order per output row."""
    ecfg, scfg, sd, onehots, lengths, labels, g = load_case("base_small")
    model = build_b200_model(ecfg, scfg, sd)
    args = dict(sequence_onehots=onehots.cuda(), sequence_lengths=lengths.cuda(), label_embeddings=labels.cuda())
    out = {}
    for flag in (1, 0):
        native.set_option("cta2", flag)
        with torch.no_grad():
            out[flag] = model(**args)[0].cpu()
    assert torch.equal(out[0], out[1])
