"""Thin torch-tensor wrappers over the C ABI (include/protnote_b200.h).

PyTorch is used here for exactly three things: owning device memory (torch.empty), naming the current CUDA stream,
and handing data pointers to the library.  No arithmetic on tensor data happens in this file.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import PN_FAST, PN_STRICT, EncoderCfg, ScorerCfg, check, pointer_array, ptr, stream_ptr

MODES = {"strict": PN_STRICT, "fast": PN_FAST}

# Scratch memory is cached per (device, CUDA stream, purpose) and only ever grows.  A buffer is only ever handed to work
# queued on the stream it is keyed by, so two streams (or two models driven from two threads on their own streams) never
# share scratch, and a buffer that is replaced by a larger one is released on the stream that last used it (the caching
# allocator's stream ordering then keeps it alive until the kernels already queued on that stream have run).  Pointers
# stay stable per stream, which is what CUDA-graph capture needs.
_scratch = {}


def scratch(device: torch.device, key: str, nbytes: int) -> torch.Tensor:
    k = (device.index, torch.cuda.current_stream(device).cuda_stream, key)
    buf = _scratch.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _scratch.pop(k, None)
        with torch.cuda.device(device):
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _scratch[k] = buf
    return buf


def release_scratch():
    _scratch.clear()


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.ProtnoteB200Error(
            f"{name} is on {t.device}: protnote_b200 computes on a CUDA sm_100a device only (no CPU path)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_options_epoch = 0


def set_option(name: str, value: int):
    """Engine options are process-global (include/protnote_b200.h).  Packed weights depend on some of them (the
    truncation compensation follows the promotion period), so every change bumps an epoch that the modules fold into
    their pack keys: the next forward re-packs."""
    global _options_epoch
    check(_lib.load().pn_set_option(name.encode(), int(value)))
    _options_epoch += 1


def options_epoch() -> int:
    return _options_epoch


def _apply_env_options():
    """Experiment switches: PN_OPTIONS="cta2=1,bk=64" sets engine options when the package is imported."""
    import os
    spec = os.environ.get("PN_OPTIONS", "")
    for item in filter(None, (x.strip() for x in spec.split(","))):
        name, _, value = item.partition("=")
        set_option(name, int(value))


def launch_count() -> int:
    return int(_lib.load().pn_launch_count())


def gemm_timing(enable: bool):
    check(_lib.load().pn_gemm_timing(int(enable)))


def gemm_timing_read():
    """(summed launch ms, launches, algorithmic FLOPs) of the scorer GEMM launches since gemm_timing(True)."""
    ms, n, fl = C.c_double(), C.c_longlong(), C.c_double()
    check(_lib.load().pn_gemm_timing_read(C.byref(ms), C.byref(n), C.byref(fl)))
    return ms.value, n.value, fl.value


_apply_env_options()


class PackedEncoder:
    """Device-resident packed weights of one ProteInfer encoder + the forward call."""

    def __init__(self, input_channels: int, channels: int, bottleneck: int, kernel_size: int, dilation_base: int,
                 num_blocks: int, bn_eps: float = 1e-3):
        self.lib = _lib.load()
        self.cfg = EncoderCfg(input_channels, channels, bottleneck, kernel_size, dilation_base, num_blocks, bn_eps)
        self.packed: Optional[torch.Tensor] = None

    def pack(self, params: Sequence[torch.Tensor]):
        """params: fp32 CUDA tensors in the order include/protnote_b200.h documents for pn_encoder_pack."""
        dev = params[0].device
        for p in params:
            _require_cuda(p, "encoder parameter")
        keep = [_f32c(p.detach()) for p in params]
        nbytes = self.lib.pn_encoder_packed_bytes(C.byref(self.cfg))
        if nbytes == 0:
            check(1)
        if self.packed is None or self.packed.numel() < nbytes or self.packed.device != dev:
            self.packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.pn_encoder_pack(C.byref(self.cfg), pointer_array(keep), len(keep), ptr(self.packed),
                                           self.packed.numel(), stream_ptr()))
        del keep  # stream-ordered: the caching allocator keeps the memory alive until the kernels have run

    def forward(self, x: torch.Tensor, lengths: torch.Tensor, mode: int = PN_STRICT,
                max_workspace_bytes: int = 6 << 30) -> torch.Tensor:
        if self.packed is None:
            raise _lib.ProtnoteB200Error("encoder weights have not been packed")
        _require_cuda(x, "sequence_onehots")
        dev = self.packed.device
        x = _f32c(x)
        lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
        B, cin, T = x.shape
        if cin != self.cfg.input_channels:
            raise ValueError(f"expected {self.cfg.input_channels} input channels, got {cin}")
        out = torch.empty(B, self.cfg.channels, dtype=torch.float32, device=dev)
        if B == 0:
            return out
        need = self.lib.pn_encoder_workspace_bytes(C.byref(self.cfg), B, T)
        one = self.lib.pn_encoder_workspace_bytes(C.byref(self.cfg), 1, T) + 4096
        ws = scratch(dev, "encoder", max(min(need + 4096 * B, max_workspace_bytes), one))
        with torch.cuda.device(dev):
            check(self.lib.pn_encoder_forward(C.byref(self.cfg), ptr(self.packed), ptr(x), ptr(lengths), B, T, ptr(out),
                                              ptr(ws), ws.numel(), mode, stream_ptr()))
        return out


def _encoder_forward_tokens(self, tokens: torch.Tensor, lengths: torch.Tensor, mode: int = PN_STRICT,
                            max_workspace_bytes: int = 6 << 30) -> torch.Tensor:
    """tokens [B, T] integer ids (uint8 on the wire) -> [B, channels]; bit-identical to forward() on their one-hot."""
    if self.packed is None:
        raise _lib.ProtnoteB200Error("encoder weights have not been packed")
    _require_cuda(tokens, "sequence tokens")
    dev = self.packed.device
    if tokens.dtype != torch.uint8:
        if tokens.is_floating_point() or int(tokens.min()) < 0 or int(tokens.max()) > 255:
            raise ValueError("token ids must be integers in [0, 255]")
        tokens = tokens.to(torch.uint8)
    tokens = tokens.contiguous()
    lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
    B, T = tokens.shape
    out = torch.empty(B, self.cfg.channels, dtype=torch.float32, device=dev)
    if B == 0:
        return out
    need = self.lib.pn_encoder_workspace_bytes(C.byref(self.cfg), B, T)
    one = self.lib.pn_encoder_workspace_bytes(C.byref(self.cfg), 1, T) + 4096
    ws = scratch(dev, "encoder", max(min(need + 4096 * B, max_workspace_bytes), one))
    with torch.cuda.device(dev):
        check(self.lib.pn_encoder_forward_tokens(C.byref(self.cfg), ptr(self.packed), ptr(tokens), ptr(lengths), B, T,
                                                 ptr(out), ptr(ws), ws.numel(), mode, stream_ptr()))
    return out


PackedEncoder.forward_tokens = _encoder_forward_tokens


def _encoder_pack_raw(self, params: Sequence[torch.Tensor]):
    """Weights for the training-mode forward: nothing folded (BatchNorm uses batch statistics there)."""
    dev = params[0].device
    keep = [_f32c(p.detach()) for p in params]
    nbytes = self.lib.pn_encoder_packed_bytes(C.byref(self.cfg))
    if getattr(self, "packed_raw", None) is None or self.packed_raw.numel() < nbytes or self.packed_raw.device != dev:
        self.packed_raw = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(self.lib.pn_encoder_pack_raw(C.byref(self.cfg), pointer_array(keep), len(keep), ptr(self.packed_raw),
                                           self.packed_raw.numel(), stream_ptr()))


def _encoder_forward_train(self, x: torch.Tensor, lengths: torch.Tensor, bn_params: Sequence[torch.Tensor],
                           momentum: float, update_running: bool = True, mode: int = PN_STRICT, group=None,
                           total_sequences: Optional[int] = None) -> torch.Tensor:
    """Batch-statistic BatchNorm forward of the encoder (pn_encoder_forward_train); bn_params: per block
    bn1.{weight, bias, running_mean, running_var}, bn2.{...} - the running statistics are updated in place.
    With `total_sequences` (and an initialised torch.distributed group) `x` holds THIS RANK's sequences of a batch of
    `total_sequences`: the per-channel sums of every BatchNorm are all-reduced between the statistics pass and the
    normalisation pass (pn_encoder_forward_train_sharded), so the result is the unsharded batch's."""
    _require_cuda(x, "sequence_onehots")
    dev = self.packed_raw.device
    x = _f32c(x)
    lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
    B, cin, T = x.shape
    if cin != self.cfg.input_channels:
        raise ValueError(f"expected {self.cfg.input_channels} input channels, got {cin}")
    for p in bn_params:
        if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
            raise _lib.ProtnoteB200Error("BatchNorm parameters / buffers must be contiguous fp32 CUDA tensors")
    out = torch.empty(B, self.cfg.channels, dtype=torch.float32, device=dev)
    if B == 0:
        if total_sequences is not None and int(total_sequences) > 0:
            # a rank without sequences (more ranks than sequences) still owes the other ranks its share - zero - of every
            # BatchNorm-sum all-reduce, in the order pn_encoder_forward_train_sharded issues them
            # (and it applies the same running-statistic update: pn_t_bn_finalize on the reduced sums)
            import torch.distributed as dist
            count = float(total_sequences) * T
            state = torch.empty(4 * self.cfg.channels, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                for i in range(self.cfg.num_blocks):
                    for j, cols in enumerate((self.cfg.channels, self.cfg.bottleneck)):
                        g, b, rm, rv = bn_params[8 * i + 4 * j: 8 * i + 4 * j + 4]
                        stats = torch.zeros(2 * cols, dtype=torch.float64, device=dev)
                        dist.all_reduce(stats, group=group)
                        check(self.lib.pn_t_bn_finalize(ptr(stats), C.c_double(count), None, C.c_double(0.0), ptr(g), ptr(b),
                                                        C.c_float(self.cfg.bn_eps), C.c_float(momentum),
                                                        ptr(rm) if update_running else None,
                                                        ptr(rv) if update_running else None, cols, ptr(state), stream_ptr()))
        return out
    ws = scratch(dev, "encoder_train", self.lib.pn_encoder_train_workspace_bytes(C.byref(self.cfg), B, T))
    # sharded whenever other ranks exist (one of them may hold nothing, so `total == B` on this rank proves nothing)
    import torch.distributed as _dist
    multi = _dist.is_available() and _dist.is_initialized() and _dist.get_world_size(group) > 1
    if total_sequences is None or (int(total_sequences) == B and not multi):
        with torch.cuda.device(dev):
            check(self.lib.pn_encoder_forward_train(C.byref(self.cfg), ptr(self.packed_raw), ptr(x), ptr(lengths), B, T,
                                                    pointer_array(bn_params), len(bn_params), C.c_float(momentum),
                                                    int(update_running), ptr(out), ptr(ws), ws.numel(), mode, stream_ptr()))
        return out
    import torch.distributed as dist
    stats = torch.empty(2 * ((self.cfg.channels + 63) // 64 * 64), dtype=torch.float64, device=dev)
    failure = []

    def reduce(_ptr, count, _user, _stream):
        try:
            dist.all_reduce(stats[:count], group=group)      # stream-ordered after the sums, before the normalisation
            return 0
        except Exception as exc:  # noqa: BLE001 - reported through the C return code
            failure.append(exc)
            return 1

    cb = _lib.REDUCE_FN(reduce)
    with torch.cuda.device(dev):
        rc = self.lib.pn_encoder_forward_train_sharded(C.byref(self.cfg), ptr(self.packed_raw), ptr(x), ptr(lengths), B, T,
                                                       pointer_array(bn_params), len(bn_params), C.c_float(momentum),
                                                       int(update_running), ptr(out), ptr(ws), ws.numel(), mode,
                                                       C.c_double(float(total_sequences) * T), ptr(stats), cb, None,
                                                       stream_ptr())
    if failure:
        raise failure[0]
    check(rc)
    return out


PackedEncoder.pack_raw = _encoder_pack_raw
PackedEncoder.forward_train = _encoder_forward_train


def postprocess(logits: torch.Tensor, labels: Optional[torch.Tensor] = None, threshold: float = 0.5,
                want_probabilities: bool = False, topk: int = 0, counts=None):
    """Device-side evaluation post-processing of a [B, L] logit batch (ProtNoteTrainer.py:522-537, :61-83).
    Returns dict(probabilities, tp, fn, fp, topk_values, topk_indices); `counts` = (tp, fn, fp) fp32 [L] tensors to
    accumulate into across batches (created zeroed when omitted)."""
    lib = _lib.load()
    _require_cuda(logits, "logits")
    if logits.dtype != torch.float32 or logits.stride(1) != 1:
        logits = logits.float().contiguous()
    B, L = logits.shape
    dev = logits.device
    kind = 0
    tp = fn = fp = None
    if labels is not None:
        _require_cuda(labels, "labels")
        if labels.dtype == torch.int64:
            kind = 1
        else:
            kind, labels = 2, labels.float()
        if labels.stride(1) != 1:
            labels = labels.contiguous()
        tp, fn, fp = counts if counts is not None else (torch.zeros(L, dtype=torch.float32, device=dev) for _ in range(3))
    probs = torch.empty(B, L, dtype=torch.float32, device=dev) if want_probabilities else None
    tv = torch.empty(B, topk, dtype=torch.float32, device=dev) if topk else None
    ti = torch.empty(B, topk, dtype=torch.int32, device=dev) if topk else None
    if B > 0:
        with torch.cuda.device(dev):
            check(lib.pn_postprocess(ptr(logits), B, L, logits.stride(0), ptr(labels), kind,
                                     labels.stride(0) if labels is not None else 0, C.c_float(threshold), ptr(probs),
                                     L, ptr(tp), ptr(fn), ptr(fp), int(topk), ptr(tv), ptr(ti), stream_ptr()))
    return {"probabilities": probs, "tp": tp, "fn": fn, "fp": fp, "topk_values": tv, "topk_indices": ti}


class PackedScorer:
    """Device-resident packed weights of W_p, W_l and the output MLP + the three calls of the scorer."""

    def __init__(self, protein_dim: int, label_dim: int, latent_dim: int, proj_hidden: int, proj_layers: int,
                 out_hidden: int, out_layers: int, out_batchnorm: bool, fusion: str, descriptions_per_label: int,
                 bn_eps: float = 1e-5):
        self.lib = _lib.load()
        if fusion not in _lib.FUSIONS:
            raise ValueError(f"feature fusion '{fusion}' is not handled by the fused pair scorer")
        self.cfg = ScorerCfg(protein_dim, label_dim, latent_dim, proj_hidden, proj_layers, out_hidden, out_layers,
                             int(bool(out_batchnorm)), _lib.FUSIONS[fusion], descriptions_per_label, bn_eps)
        self.packed: Optional[torch.Tensor] = None

    def num_params(self) -> int:
        n = self.lib.pn_scorer_num_params(C.byref(self.cfg))
        if n < 0:
            check(1)
        return n

    def pack(self, params: Sequence[torch.Tensor]):
        dev = params[0].device
        for p in params:
            _require_cuda(p, "scorer parameter")
        keep = [_f32c(p.detach()) for p in params]
        nbytes = self.lib.pn_scorer_packed_bytes(C.byref(self.cfg))
        if nbytes == 0:
            check(1)
        if self.packed is None or self.packed.numel() < nbytes or self.packed.device != dev:
            self.packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.pn_scorer_pack(C.byref(self.cfg), pointer_array(keep), len(keep), ptr(self.packed),
                                          self.packed.numel(), stream_ptr()))
        del keep

    def _project(self, fn, x: torch.Tensor, want_embedding: bool, mode: int):
        dev = self.packed.device
        x = _f32c(x)
        n = x.shape[0]
        similarity = self.cfg.fusion == _lib.FUSIONS["similarity"]
        emb = torch.empty(n, self.cfg.latent_dim, dtype=torch.float32, device=dev) if (want_embedding or similarity) else None
        half = None if similarity else torch.empty(n, self.cfg.out_hidden, dtype=torch.float32, device=dev)
        if n == 0:
            return emb, half
        rows = min(n, 1 << 15)
        ws = scratch(dev, "project", self.lib.pn_project_workspace_bytes(C.byref(self.cfg), rows))
        with torch.cuda.device(dev):
            check(fn(C.byref(self.cfg), ptr(self.packed), ptr(x), n, ptr(emb), ptr(half), ptr(ws), ws.numel(), mode,
                     stream_ptr()))
        return emb, half

    def project_sequences(self, P_f: torch.Tensor, mode: int = PN_STRICT, want_embedding: bool = False):
        """P_f [n, protein_dim] -> (P_e [n, latent] or None, a [n, out_hidden])"""
        _require_cuda(P_f, "sequence embeddings")
        return self._project(self.lib.pn_project_sequences, P_f, want_embedding, mode)

    def project_labels(self, L_f: torch.Tensor, mode: int = PN_STRICT, want_embedding: bool = False):
        """L_f [n, label_dim] -> (L_e [n, latent] or None, c [n, out_hidden])"""
        _require_cuda(L_f, "label embeddings")
        return self._project(self.lib.pn_project_labels, L_f, want_embedding, mode)

    def score(self, a: torch.Tensor, c: torch.Tensor, P_e: Optional[torch.Tensor] = None,
              L_e: Optional[torch.Tensor] = None, mode: int = PN_STRICT, out: Optional[torch.Tensor] = None,
              max_workspace_bytes: int = 8 << 30, hidden_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """a [B, H], c [L, H] -> logits [B, L / k] fp32; hidden_out [B*L, H] optionally receives the last hidden layer"""
        dev = self.packed.device
        B, L = a.shape[0], c.shape[0]
        k = self.cfg.descriptions_per_label
        if L % k != 0:
            raise ValueError(f"{L} label rows is not a multiple of inference_descriptions_per_label={k}")
        if out is None:
            out = torch.empty(B, L // k, dtype=torch.float32, device=dev)
        if B == 0 or L == 0:
            return out
        need = self.lib.pn_scorer_workspace_bytes(C.byref(self.cfg), B, L)
        floor = self.lib.pn_scorer_min_workspace_bytes(C.byref(self.cfg))
        ws = scratch(dev, "scorer", max(min(need, max_workspace_bytes), floor))
        with torch.cuda.device(dev):
            check(self.lib.pn_score_pairs_ex(C.byref(self.cfg), ptr(self.packed), ptr(a), ptr(c), ptr(P_e), ptr(L_e), B, L,
                                             ptr(out), out.stride(0), ptr(hidden_out), ptr(ws), ws.numel(), mode,
                                             stream_ptr()))
        return out


def _score_similarity(self, P_e: torch.Tensor, L_e: torch.Tensor, temperature: float, mode: int = PN_STRICT):
    """cosine similarity / temperature for every pair, [B, L / k] fp32 (ProtNote.py:281-284)."""
    dev = self.packed.device
    B, L = P_e.shape[0], L_e.shape[0]
    k = self.cfg.descriptions_per_label
    if L % k != 0:
        raise ValueError(f"{L} label rows is not a multiple of inference_descriptions_per_label={k}")
    out = torch.empty(B, L // k, dtype=torch.float32, device=dev)
    if B == 0 or L == 0:
        return out
    ws = scratch(dev, "similarity", self.lib.pn_similarity_workspace_bytes(C.byref(self.cfg), B, L))
    with torch.cuda.device(dev):
        check(self.lib.pn_score_similarity(C.byref(self.cfg), ptr(P_e), ptr(L_e), B, L, C.c_float(temperature), ptr(out),
                                           out.stride(0), ptr(ws), ws.numel(), mode, stream_ptr()))
    return out


PackedScorer.score_similarity = _score_similarity


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], mode: int = PN_STRICT) -> torch.Tensor:
    """y = x w^T + bias on the tensor-core engine (ProteInfer.output_layer, protein_encoders.py:120-123)."""
    lib = _lib.load()
    _require_cuda(x, "input")
    x, w = _f32c(x), _f32c(w.detach())
    b = _f32c(bias.detach()) if bias is not None else None
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    if M == 0:
        return y
    ws = scratch(x.device, "linear", lib.pn_linear_workspace_bytes(M, N, K))
    with torch.cuda.device(x.device):
        check(lib.pn_linear(ptr(x), M, K, K, ptr(w), N, ptr(b), ptr(y), N, ptr(ws), ws.numel(), mode, stream_ptr()))
    return y


def conv1d_channels_last(x: torch.Tensor, lengths: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor],
                         dilation: int, mode: int = PN_STRICT) -> torch.Tensor:
    """Masked 'same' Conv1d (MaskedConv1D, protein_encoders.py:8-17): x [B,Cin,T] -> y [B,T,Cout]."""
    lib = _lib.load()
    _require_cuda(x, "input")
    x, w = _f32c(x), _f32c(w.detach())
    b = _f32c(bias.detach()) if bias is not None else None
    lengths = lengths.to(device=x.device, dtype=torch.int64).contiguous()
    B, cin, T = x.shape
    cout, _, taps = w.shape
    y = torch.empty(B, T, cout, dtype=torch.float32, device=x.device)
    ws = scratch(x.device, "conv1d", lib.pn_conv1d_workspace_bytes(B, T, cin, cout, taps))
    with torch.cuda.device(x.device):
        check(lib.pn_conv1d(ptr(x), ptr(lengths), B, cin, T, ptr(w), ptr(b), cout, taps, dilation, ptr(y), ptr(ws),
                            ws.numel(), mode, stream_ptr()))
    return y
