"""The tensor-core engine (pn_linear / pn_conv1d through the C ABI) against fp64 PyTorch operators."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bk", [32, 64])
@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 520, 200), (1000, 3072, 3072), (77, 36, 72), (1, 8, 5)])
def test_linear_strict(bk, shape):
    from protnote_b200 import native
    native.set_option("bk", bk)
    try:
        M, N, K = shape
        g = torch.Generator().manual_seed(M * 7 + N)
        x = torch.randn(M, K, generator=g).cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        y = native.linear(x, w, b)
        ref = x.double() @ w.double().T + b.double()
        # fp32-grade: same error level as an fp32 SIMT GEMM (cuBLAS fp32 measures 5e-6 max at K=3072 on this data)
        assert (y.double() - ref).abs().max().item() <= 1e-5
    finally:
        native.set_option("bk", 0)


@pytest.mark.parametrize("cfg", [(2, 64, 256, 64, 1, 1), (3, 72, 150, 36, 9, 3), (2, 20, 300, 72, 9, 1),
                                 (2, 36, 700, 72, 9, 81), (1, 20, 5, 16, 9, 27)])
def test_masked_dilated_conv_strict(cfg):
    """MaskedConv1D semantics (reference protein_encoders.py:8-17): mask, Conv1d(padding='same'), mask."""
    from protnote_b200 import native
    B, cin, T, cout, taps, dil = cfg
    g = torch.Generator().manual_seed(sum(cfg))
    x = torch.randn(B, cin, T, generator=g).cuda()
    lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lengths[0] = T
    lengths = lengths.cuda()
    w = (torch.randn(cout, cin, taps, generator=g) / (cin * taps) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    y = native.conv1d_channels_last(x, lengths, w, b, dil)
    mask = torch.arange(T, device="cuda")[None, :] < lengths[:, None]
    ref = torch.nn.functional.conv1d((x * mask[:, None, :]).double(), w.double(), b.double(), padding="same", dilation=dil)
    ref = (ref * mask[:, None, :]).permute(0, 2, 1)
    assert (y.double() - ref).abs().max().item() <= 1e-5
    assert y[~mask].abs().max().item() == 0.0 if (~mask).any() else True
