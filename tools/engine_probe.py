"""GPU probe of the tensor-core engine: prints max errors of pn_linear / pn_conv1d against fp64 torch for every
(bk, mode) combination.  Diagnostic only (never asserts) - the pass/fail versions live in tests/."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from protnote_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
print("device check:", lib.pn_device_check(0), lib.pn_last_error())


def linear(x, w, bias, mode):
    M, K = x.shape
    N = w.shape[0]
    ws = torch.empty(lib.pn_linear_workspace_bytes(M, N, K), dtype=torch.uint8, device=dev)
    y = torch.full((M, N), float("nan"), device=dev)
    st = lib.pn_linear(_lib.ptr(x), M, K, K, _lib.ptr(w), N, _lib.ptr(bias), _lib.ptr(y), N, _lib.ptr(ws), ws.numel(),
                       mode, _lib.stream_ptr())
    if st:
        print("  pn_linear error:", lib.pn_last_error())
    torch.cuda.synchronize()
    return y


def conv(x, lengths, w, bias, dil, mode):
    B, cin, T = x.shape
    cout, _, taps = w.shape
    ws = torch.empty(lib.pn_conv1d_workspace_bytes(B, T, cin, cout, taps), dtype=torch.uint8, device=dev)
    y = torch.full((B, T, cout), float("nan"), device=dev)
    st = lib.pn_conv1d(_lib.ptr(x), _lib.ptr(lengths), B, cin, T, _lib.ptr(w), _lib.ptr(bias), cout, taps, dil,
                       _lib.ptr(y), _lib.ptr(ws), ws.numel(), mode, _lib.stream_ptr())
    if st:
        print("  pn_conv1d error:", lib.pn_last_error())
    torch.cuda.synchronize()
    return y


g = torch.Generator(device="cpu").manual_seed(0)
for bk in (64, 32):
    lib.pn_set_option(b"bk", bk)
    for mode, name in ((_lib.PN_STRICT, "strict"), (_lib.PN_FAST, "fast")):
        for (M, N, K) in ((128, 256, 64), (128, 256, 128), (300, 520, 200), (1000, 3072, 3072), (77, 36, 72)):
            x = torch.randn(M, K, generator=g).to(dev)
            w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
            b = torch.randn(N, generator=g).to(dev)
            try:
                y = linear(x, w, b, mode)
                ref = (x.double() @ w.double().T + b.double())
                err = (y.double() - ref).abs().max().item()
                print(f"bk={bk} {name:6s} linear M={M} N={N} K={K}: max err {err:.3e} (ref absmax {ref.abs().max().item():.2f}) nan={torch.isnan(y).sum().item()}")
            except Exception as e:  # noqa: BLE001
                print(f"bk={bk} {name} linear {M,N,K} raised {e!r}")
        for (B, cin, T, cout, taps, dil) in ((2, 64, 256, 64, 1, 1), (3, 72, 150, 36, 9, 3), (2, 20, 300, 72, 9, 1), (2, 1100, 700, 550, 9, 81)):
            x = torch.randn(B, cin, T, generator=g).to(dev)
            lengths = torch.randint(T // 2, T + 1, (B,), generator=g)
            lengths[0] = T
            lengths = lengths.to(dev)
            w = (torch.randn(cout, cin, taps, generator=g) / (cin * taps) ** 0.5).to(dev)
            b = torch.randn(cout, generator=g).to(dev)
            try:
                y = conv(x, lengths, w, b, dil, mode)
                mask = (torch.arange(T, device=dev)[None, :] < lengths[:, None])
                xm = (x * mask[:, None, :]).double()
                ref = torch.nn.functional.conv1d(xm, w.double(), b.double(), padding="same", dilation=dil)
                ref = (ref * mask[:, None, :]).permute(0, 2, 1)
                err = (y.double() - ref).abs().max().item()
                print(f"bk={bk} {name:6s} conv B={B} cin={cin} T={T} cout={cout} taps={taps} dil={dil}: max err {err:.3e} nan={torch.isnan(y).sum().item()}")
            except Exception as e:  # noqa: BLE001
                print(f"bk={bk} {name} conv raised {e!r}")

# quick throughput probe of the engine alone (pn_linear includes split+pack kernels, so time a big one twice)
lib.pn_set_option(b"bk", 0)
for mode, name in ((_lib.PN_STRICT, "strict"), (_lib.PN_FAST, "fast")):
    for bk in (32, 64):
        lib.pn_set_option(b"bk", bk)
        M, N, K = 32768, 3072, 3072
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        ws = torch.empty(lib.pn_linear_workspace_bytes(M, N, K), dtype=torch.uint8, device=dev)
        y = torch.empty(M, N, device=dev)
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.pn_linear(_lib.ptr(x), M, K, K, _lib.ptr(w), N, None, _lib.ptr(y), N, _lib.ptr(ws), ws.numel(), mode, _lib.stream_ptr())
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{name} bk={bk} linear {M}x{N}x{K} incl. split+pack: {ms:.3f} ms -> {2*M*N*K/ms/1e9:.1f} TFLOP/s algorithmic")
print("launches:", lib.pn_launch_count())
