"""TEST INFRASTRUCTURE ONLY - the CPU oracle for the ProtNote scoring hot path.

A functional restatement (plain tensor functions over a reference-format
``state_dict``) of the reference's eval-mode forward:

    one-hot [B,Cin,T] + lengths -> ProteInfer dilated ResNet -> masked mean pool
    -> W_p / W_l projection MLPs -> pairwise fusion -> output MLP -> logits [B, L/k]

It exists so that the `-m gpu` tests, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` have a checker that
travels to the GPU box (the reference itself lives in /root/reference, which
does not exist there).  The product package ``protnote_b200`` never imports it.

Parity status: **pinned**.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so this oracle is pinned against the reference's own
classes, imported unmodified in the build container through
``oracle/ref_import.py``:
  * ``tests/test_oracle_vs_reference.py`` (runs where /root/reference exists),
  * ``tests/golden/*.pt`` - seeded input/output vectors produced by
    ``oracle/make_golden.py`` from the reference classes, checked everywhere.

Every function cites the reference file:line whose arithmetic it follows.
All arithmetic is IEEE floating point in ``dtype`` (float32 = the parity
target; float64 = a higher-precision yardstick used to size tolerances).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class EncoderCfg:
    """protnote/models/protein_encoders.py:70-107 ctor arguments (base_config.yaml:104-112)."""

    input_channels: int = 20
    output_channels: int = 1100
    kernel_size: int = 9
    dilation_base: int = 3
    num_resnet_blocks: int = 5
    bottleneck_factor: float = 0.5
    bn_eps: float = 1e-3  # protein_encoders.py:36,48


@dataclass
class ScorerCfg:
    """protnote/models/ProtNote.py:10-36 ctor arguments that shape the arithmetic."""

    protein_embedding_dim: int = 1100
    label_embedding_dim: int = 1024
    latent_dim: int = 1024
    output_mlp_hidden_dim_scale_factor: float = 3
    output_mlp_num_layers: int = 3
    output_mlp_batchnorm: bool = True
    projection_head_num_layers: int = 4
    projection_head_hidden_dim_scale_factor: int = 3
    feature_fusion: str = "concatenation"
    inference_descriptions_per_label: int = 1
    temperature: float = 0.07
    # ProtNote.py:83-86 wraps W_p / W_l in Sequential(Dropout, MLP) when these are > 0,
    # which renames the state_dict keys to W_p.1.* / W_l.1.*
    sequence_embedding_dropout: float = 0.0
    label_embedding_dropout: float = 0.0
    # OUTPUT_MLP_DROPOUT (base_config.yaml:39): Dropout modules inside W_p / W_l / output_layer; inactive in eval mode
    output_mlp_dropout: float = 0.0
    bn_eps: float = 1e-5  # torch.nn.BatchNorm1d default used by torchvision MLP / get_mlp


# ----------------------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------------------


def pad_mask(lengths: Tensor, T: int) -> Tensor:
    """True where position >= length.  protnote/data/datasets.py:559-563."""
    return torch.arange(T, device=lengths.device)[None, :] >= lengths.reshape(-1, 1)


def zero_padding(x: Tensor, lengths: Tensor) -> Tensor:
    """set_padding_to_sentinel(x, lengths, 0) - protnote/data/datasets.py:535-569."""
    return torch.where(pad_mask(lengths, x.shape[-1])[:, None, :], torch.zeros((), dtype=x.dtype, device=x.device), x)


def masked_conv(x: Tensor, lengths: Tensor, w: Tensor, b: Tensor, dilation: int) -> Tensor:
    """MaskedConv1D.forward - protein_encoders.py:8-17: mask, 'same' conv (stride 1), mask."""
    x = zero_padding(x, lengths)
    k = w.shape[-1]
    total = dilation * (k - 1)
    left = total // 2  # torch 'same' padding: left = total//2, right = total-left
    x = F.pad(x, (left, total - left))
    y = F.conv1d(x, w, b, dilation=dilation)
    return zero_padding(y, lengths)


def batchnorm_eval(x: Tensor, sd: Dict[str, Tensor], prefix: str, eps: float, dtype) -> Tensor:
    """BatchNorm1d in eval mode: (x - mean)/sqrt(var+eps)*gamma + beta (running statistics)."""
    g = sd[prefix + ".weight"].to(dtype)
    b = sd[prefix + ".bias"].to(dtype)
    m = sd[prefix + ".running_mean"].to(dtype)
    v = sd[prefix + ".running_var"].to(dtype)
    shape = (1, -1, 1) if x.dim() == 3 else (1, -1)
    return (x - m.reshape(shape)) / torch.sqrt(v.reshape(shape) + eps) * g.reshape(shape) + b.reshape(shape)


def residual_block(x, lengths, sd, prefix, dilation, cfg: EncoderCfg, dtype):
    """Residual.forward - protein_encoders.py:61-67 (ResNet-v2 pre-activation bottleneck)."""
    out = torch.relu(batchnorm_eval(x, sd, prefix + ".bn_activation_1.0", cfg.bn_eps, dtype))
    out = masked_conv(out, lengths, sd[prefix + ".masked_conv1.weight"].to(dtype),
                      sd[prefix + ".masked_conv1.bias"].to(dtype), dilation)
    out = torch.relu(batchnorm_eval(out, sd, prefix + ".bn_activation_2.0", cfg.bn_eps, dtype))
    out = masked_conv(out, lengths, sd[prefix + ".masked_conv2.weight"].to(dtype),
                      sd[prefix + ".masked_conv2.bias"].to(dtype), 1)
    return out + x


def proteinfer_embeddings(sd: Dict[str, Tensor], x: Tensor, lengths: Tensor, cfg: EncoderCfg,
                          prefix: str = "", dtype=torch.float32) -> Tensor:
    """ProteInfer.get_embeddings - protein_encoders.py:109-118.  x [B,Cin,T] -> [B,C]."""
    x = x.to(dtype)
    f = masked_conv(x, lengths, sd[prefix + "conv1.weight"].to(dtype), sd[prefix + "conv1.bias"].to(dtype), 1)
    for i in range(cfg.num_resnet_blocks):
        f = residual_block(f, lengths, sd, f"{prefix}resnet_blocks.{i}", cfg.dilation_base ** i, cfg, dtype)
    f = zero_padding(f, lengths)
    return f.sum(-1) / lengths.reshape(-1, 1).to(dtype)


def proteinfer_logits(sd, x, lengths, cfg: EncoderCfg, prefix: str = "", dtype=torch.float32) -> Tensor:
    """ProteInfer.forward - protein_encoders.py:120-123."""
    e = proteinfer_embeddings(sd, x, lengths, cfg, prefix, dtype)
    return e @ sd[prefix + "output_layer.weight"].to(dtype).T + sd[prefix + "output_layer.bias"].to(dtype)


# ----------------------------------------------------------------------------------------
# projection heads + pair scorer
# ----------------------------------------------------------------------------------------


def projection_head(sd, prefix: str, x: Tensor, cfg: ScorerCfg, dtype) -> Tensor:
    """torchvision.ops.MLP as built at ProtNote.py:63-81: [Linear(no bias), BN, ReLU, Dropout]x(n-1),
    Linear(no bias), Dropout.  Module indices are 4*i for the Linear, 4*i+1 for its BN."""
    n = cfg.projection_head_num_layers
    for i in range(n):
        x = x @ sd[f"{prefix}.{4 * i}.weight"].to(dtype).T
        if i < n - 1:
            x = torch.relu(batchnorm_eval(x, sd, f"{prefix}.{4 * i + 1}", cfg.bn_eps, dtype))
    return x


def output_mlp_layout(cfg: ScorerCfg):
    """Module indices produced by get_mlp - ProtNote.py:337-378.
    Returns ([(linear_idx, bn_idx or None)] for hidden layers, final_linear_idx)."""
    idx, hidden = 0, []
    for layer in range(cfg.output_mlp_num_layers):
        lin = idx
        idx += 1
        bn = None
        if cfg.output_mlp_batchnorm:
            bn = idx
            idx += 1
        idx += 1  # ReLU
        if layer < cfg.output_mlp_num_layers - 1:
            idx += 1  # Dropout
        hidden.append((lin, bn))
    return hidden, idx


def joint_features(P_e: Tensor, L_e: Tensor, fusion: str) -> Tensor:
    """ProtNote._get_joint_embeddings - ProtNote.py:112-152: every (protein,label) pair, protein-major."""
    B, L = P_e.shape[0], L_e.shape[0]
    p = P_e[:, None, :].expand(B, L, P_e.shape[1])
    t = L_e[None, :, :].expand(B, L, L_e.shape[1])
    parts = [p, t]
    if fusion == "concatenation_diff":
        parts.append(p - t)
    elif fusion == "concatenation_prod":
        parts.append(p * t)
    elif fusion != "concatenation":
        raise ValueError(fusion)
    return torch.cat(parts, dim=2).reshape(B * L, -1)


def output_mlp(sd, prefix: str, joint: Tensor, cfg: ScorerCfg, dtype, return_hidden: bool = False) -> Tensor:
    """get_mlp forward - ProtNote.py:337-378 (hidden Linear has a bias only when batch_norm is off)."""
    hidden, last = output_mlp_layout(cfg)
    x = joint
    for lin, bn in hidden:
        x = x @ sd[f"{prefix}.{lin}.weight"].to(dtype).T
        if f"{prefix}.{lin}.bias" in sd:
            x = x + sd[f"{prefix}.{lin}.bias"].to(dtype)
        if bn is not None:
            x = batchnorm_eval(x, sd, f"{prefix}.{bn}", cfg.bn_eps, dtype)
        x = torch.relu(x)
    if return_hidden:
        return x
    return x @ sd[f"{prefix}.{last}.weight"].to(dtype).T + sd[f"{prefix}.{last}.bias"].to(dtype)


def ensemble_logits(logits: Tensor, B: int, L: int, k: int) -> Tensor:
    """ProtNote.py:308-322: reshape, or logit(mean_k sigmoid(x), eps=1e-7) over k consecutive rows."""
    if k == 1:
        return logits.reshape(B, L)
    p = torch.sigmoid(logits).reshape(B, L // k, k).mean(-1)
    return torch.special.logit(p, eps=1e-7)


def score_pairs(sd, P_f: Tensor, L_f: Tensor, cfg: ScorerCfg, dtype=torch.float32,
                pair_chunk: int = 1 << 16) -> Tensor:
    """ProtNote.forward from the projections on - ProtNote.py:270-322 (eval mode).
    The [B*L, 2d] joint tensor is produced in protein chunks so large L stays in host memory;
    pairs are independent, so chunking does not change any value."""
    wp = "W_p.1" if cfg.sequence_embedding_dropout > 0 else "W_p"
    wl = "W_l.1" if cfg.label_embedding_dropout > 0 else "W_l"
    P_e = projection_head(sd, wp, P_f.to(dtype), cfg, dtype)
    L_e = projection_head(sd, wl, L_f.to(dtype), cfg, dtype)
    B, L = P_e.shape[0], L_e.shape[0]
    if cfg.feature_fusion == "similarity":  # ProtNote.py:281-284
        logits = F.normalize(P_e, dim=-1, p=2) @ F.normalize(L_e, dim=-1, p=2).T / cfg.temperature
        return ensemble_logits(logits, B, L, cfg.inference_descriptions_per_label)
    rows = max(1, pair_chunk // max(L, 1))
    out = []
    for s in range(0, B, rows):
        j = joint_features(P_e[s:s + rows], L_e, cfg.feature_fusion)
        out.append(output_mlp(sd, "output_layer", j, cfg, dtype).reshape(-1, L))
    logits = torch.cat(out, 0)
    return ensemble_logits(logits, B, L, cfg.inference_descriptions_per_label)


def protnote_forward(sd, onehots: Tensor, lengths: Tensor, label_embeddings: Tensor,
                     ecfg: EncoderCfg, scfg: ScorerCfg, dtype=torch.float32) -> Tensor:
    """ProtNote.forward(sequence_onehots=, sequence_lengths=, label_embeddings=) - ProtNote.py:168-334,
    eval mode, cached label embeddings.  Returns logits [B, L/k]."""
    with torch.no_grad():
        P_f = proteinfer_embeddings(sd, onehots, lengths, ecfg, "sequence_encoder.", dtype)
        return score_pairs(sd, P_f, label_embeddings, scfg, dtype)


# ----------------------------------------------------------------------------------------
# seeded synthetic problems
# ----------------------------------------------------------------------------------------


def _calibrate(sd, ecfg: EncoderCfg, scfg: ScorerCfg, g: torch.Generator, logit_std: float, last: int,
               calib_T: int):
    """Make the random network behave like a trained one: walk a small calibration batch through the
    model and set every BatchNorm's running statistics to (a jittered copy of) the statistics it
    actually sees, so activations stay O(1) and pair-to-pair variation is not crushed layer after
    layer (with untouched running stats a 10-layer random MLP maps all pairs to nearly the same
    logit, and rescaling that to a useful std only amplifies fp32 rounding noise).  Finally the
    output neuron is made orthogonal to the mean hidden activation and scaled to `logit_std`."""
    f32 = torch.float32

    def set_bn(name, x, dims):
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        sd[name + ".running_mean"] = mean + 0.1 * var.sqrt() * torch.randn(mean.shape, generator=g)
        sd[name + ".running_var"] = (var * (0.7 + 0.6 * torch.rand(var.shape, generator=g))).clamp_min(1e-4)

    with torch.no_grad():
        x, lengths, labels = synth_inputs(24, calib_T, 96, ecfg, scfg, ragged=True, seed=977)
        valid = (~pad_mask(lengths, calib_T))[:, None, :]
        p = "sequence_encoder."
        f = masked_conv(x, lengths, sd[p + "conv1.weight"], sd[p + "conv1.bias"], 1)
        for i in range(ecfg.num_resnet_blocks):
            q = f"{p}resnet_blocks.{i}"
            set_bn(q + ".bn_activation_1.0", f.permute(1, 0, 2)[:, valid.expand_as(f).permute(1, 0, 2)[0]], (1,))
            out = torch.relu(batchnorm_eval(f, sd, q + ".bn_activation_1.0", ecfg.bn_eps, f32))
            out = masked_conv(out, lengths, sd[q + ".masked_conv1.weight"], sd[q + ".masked_conv1.bias"],
                              ecfg.dilation_base ** i)
            set_bn(q + ".bn_activation_2.0", out.permute(1, 0, 2)[:, valid.expand_as(out).permute(1, 0, 2)[0]], (1,))
            out = torch.relu(batchnorm_eval(out, sd, q + ".bn_activation_2.0", ecfg.bn_eps, f32))
            out = masked_conv(out, lengths, sd[q + ".masked_conv2.weight"], sd[q + ".masked_conv2.bias"], 1)
            f = out + f
        P_f = zero_padding(f, lengths).sum(-1) / lengths[:, None].float()
        embs = {}
        for name, xin, drop in (("W_p", P_f, scfg.sequence_embedding_dropout),
                                ("W_l", labels, scfg.label_embedding_dropout)):
            pre = name + ".1" if drop > 0 else name
            h = xin
            for i in range(scfg.projection_head_num_layers):
                h = h @ sd[f"{pre}.{4 * i}.weight"].T
                if i < scfg.projection_head_num_layers - 1:
                    set_bn(f"{pre}.{4 * i + 1}", h, (0,))
                    h = torch.relu(batchnorm_eval(h, sd, f"{pre}.{4 * i + 1}", scfg.bn_eps, f32))
            embs[name] = h
        if not scfg.feature_fusion.startswith("concatenation"):
            return
        hidden, _ = output_mlp_layout(scfg)
        h = joint_features(embs["W_p"], embs["W_l"], scfg.feature_fusion)
        for lin, bnidx in hidden:
            h = h @ sd[f"output_layer.{lin}.weight"].T
            if f"output_layer.{lin}.bias" in sd:
                h = h + sd[f"output_layer.{lin}.bias"]
            if bnidx is not None:
                set_bn(f"output_layer.{bnidx}", h, (0,))
                h = batchnorm_eval(h, sd, f"output_layer.{bnidx}", scfg.bn_eps, f32)
            h = torch.relu(h)
        m = h.mean(0)
        w = sd[f"output_layer.{last}.weight"][0]
        w = w - (w @ m) / (m @ m) * m
        s = float((h @ w).std())
        if s > 0:
            w = w * (logit_std / s)
        sd[f"output_layer.{last}.weight"] = w[None, :].contiguous()
        sd[f"output_layer.{last}.bias"] = torch.full((1,), -0.25)


def synth_state_dict(ecfg: EncoderCfg, scfg: ScorerCfg, seed: int = 42, num_labels_encoder: int = 8,
                     logit_std: float = 2.0, calib_T: int = 128) -> Dict[str, Tensor]:
    """Random-init weights in the reference's state_dict format, with every BatchNorm randomised so that
    folding is exercised, and the final output neuron rescaled so logits have std ~ `logit_std`
    (PyTorch default init gives std ~1e-3, which would make top-k comparisons meaningless)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def linear(name, out_f, in_f, bias):
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = uniform((out_f, in_f), bound)
        if bias:
            sd[name + ".bias"] = uniform((out_f,), bound)

    def conv(name, out_c, in_c, k):
        bound = 1.0 / math.sqrt(in_c * k)
        sd[name + ".weight"] = uniform((out_c, in_c, k), bound)
        sd[name + ".bias"] = uniform((out_c,), bound)

    def bn(name, c):
        sd[name + ".weight"] = torch.rand(c, generator=g) + 0.5
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    C = ecfg.output_channels
    Cb = int(math.floor(C * ecfg.bottleneck_factor))
    p = "sequence_encoder."
    conv(p + "conv1", C, ecfg.input_channels, ecfg.kernel_size)
    for i in range(ecfg.num_resnet_blocks):
        q = f"{p}resnet_blocks.{i}."
        bn(q + "bn_activation_1.0", C)
        conv(q + "masked_conv1", Cb, C, ecfg.kernel_size)
        bn(q + "bn_activation_2.0", Cb)
        conv(q + "masked_conv2", C, Cb, 1)
    linear(p + "output_layer", num_labels_encoder, C, True)

    hid = scfg.latent_dim * scfg.projection_head_hidden_dim_scale_factor
    for name, in_dim, drop in (("W_p", scfg.protein_embedding_dim, scfg.sequence_embedding_dropout),
                               ("W_l", scfg.label_embedding_dim, scfg.label_embedding_dropout)):
        if drop > 0:
            name = name + ".1"
        dims = [in_dim] + [hid] * (scfg.projection_head_num_layers - 1) + [scfg.latent_dim]
        for i in range(scfg.projection_head_num_layers):
            linear(f"{name}.{4 * i}", dims[i + 1], dims[i], False)
            if i < scfg.projection_head_num_layers - 1:
                bn(f"{name}.{4 * i + 1}", dims[i + 1])

    if scfg.feature_fusion.startswith("concatenation"):
        H = int(round(scfg.output_mlp_hidden_dim_scale_factor * scfg.latent_dim))
        in_dim = scfg.latent_dim * (2 if scfg.feature_fusion == "concatenation" else 3)
        hidden, last = output_mlp_layout(scfg)
        for li, (lin, bnidx) in enumerate(hidden):
            linear(f"output_layer.{lin}", H, in_dim if li == 0 else H, not scfg.output_mlp_batchnorm)
            if bnidx is not None:
                bn(f"output_layer.{bnidx}", H)
        linear(f"output_layer.{last}", 1, H, True)
        _calibrate(sd, ecfg, scfg, g, logit_std, last, calib_T)
    return sd


def synth_inputs(B: int, T: int, L: int, ecfg: EncoderCfg, scfg: ScorerCfg, ragged: bool = True,
                 seed: int = 1234):
    """Token ids -> float32 one-hot [B,Cin,T] (padding columns all-zero like collators.py:123-133),
    lengths int64 [B] (lengths[0] == T), label embeddings randn [L, label_dim]."""
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randint(0, ecfg.input_channels, (B, T), generator=g)
    if ragged:
        lengths = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        lengths[0] = T
    else:
        lengths = torch.full((B,), T, dtype=torch.long)
    onehots = F.one_hot(tokens, ecfg.input_channels).permute(0, 2, 1).to(torch.float32).contiguous()
    onehots = onehots * (~pad_mask(lengths, T))[:, None, :]
    g2 = torch.Generator().manual_seed(seed + 3087)
    labels = torch.randn(L, scfg.label_embedding_dim, generator=g2)
    return onehots, lengths.to(torch.long), labels
