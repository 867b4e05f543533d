"""TEST INFRASTRUCTURE ONLY - CPU oracle of the TRAINING-mode scoring step (travels to the GPU box).

Restates, with plain torch autograd, what the reference computes for one training step of the trainable part of
ProtNote when the sequence embeddings are given (frozen encoder, TRAIN_SEQUENCE_ENCODER False):

    P_f, L_f -> W_p / W_l (torchvision MLP, BatchNorm1d in training mode, ProtNote.py:63-81,270-271)
             -> joint [B*L, 2d] (ProtNote.py:112-126) -> output MLP (get_mlp, ProtNote.py:337-378, BN training mode)
             -> logits [B, L] (ProtNote.py:308-309) -> BCE-with-logits (utils/losses.py:270-294 'BCE') -> backward

Parity status: pinned against the reference's own `ProtNote` class in train mode (imported unmodified in the build
container) by tests/test_train_cpu.py::test_train_oracle_matches_reference, and through tests/golden/train_*.pt
(generated from the reference class by oracle/make_golden_train.py).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .protnote_oracle import ScorerCfg, joint_features, output_mlp_layout


def _bn_train(x, sd, prefix, eps, new_stats, momentum=0.1):
    """BatchNorm1d training forward; records the updated running statistics (momentum 0.1: torch default,
    which torchvision MLP / get_mlp do not override)."""
    rm = sd[prefix + ".running_mean"].detach().clone().to(x.dtype)
    rv = sd[prefix + ".running_var"].detach().clone().to(x.dtype)
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, eps)
    new_stats[prefix + ".running_mean"] = rm
    new_stats[prefix + ".running_var"] = rv
    return y


def focal_loss(logits, target, alpha: float, gamma: float, label_smoothing: float = 0.0, reduction: str = "mean"):
    """FocalLoss.forward - protnote/utils/losses.py:193-213 (the default LOSS_FN, configs/base_config.yaml:61): optional label
    smoothing of the targets, element-wise BCE-with-logits, modulation (1 - exp(-BCE))^gamma, alpha_t weighting when
    alpha >= 0, mean / sum.  Pinned against the reference class by tests/test_train_cpu.py."""
    if label_smoothing > 0:
        target = target * (1.0 - label_smoothing) + (1 - target) * label_smoothing
    bce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")
    pt = torch.exp(-bce)
    loss = ((1 - pt) ** gamma) * bce
    if alpha >= 0:
        loss = (alpha * target + (1 - alpha) * (1 - target)) * loss
    return loss.mean() if reduction == "mean" else loss.sum()


def loss_oracle(logits, targets, loss="bce", pos_weight=None, gamma=2.0, alpha=-1.0, label_smoothing=0.0, reduction="mean"):
    """The reference's get_loss choices that are per-element functions of (logit, target) - utils/losses.py:270-294:
    'BCE' = torch.nn.BCEWithLogitsLoss(reduction='mean', pos_weight=...), 'FocalLoss' = focal_loss above."""
    if loss in ("bce", "BCE"):
        return F.binary_cross_entropy_with_logits(logits, targets, reduction=reduction, pos_weight=pos_weight)
    if loss in ("focal", "FocalLoss"):
        return focal_loss(logits, targets, alpha, gamma, label_smoothing, reduction)
    raise NotImplementedError(loss)


def train_step_oracle(sd: Dict[str, torch.Tensor], P_f, L_f, targets, cfg: ScorerCfg, dtype=torch.float64, masks=None,
                      **loss_kw):
    """Returns (logits [B, L], loss, {state_dict key: gradient}, {running-stat key: updated value}).
    loss_kw: arguments of loss_oracle (default: BCE-with-logits, mean).
    masks: OUTPUT_MLP_DROPOUT as explicit multipliers (keep / (1 - p)) - {("p" | "l", i): after layer i of W_p / W_l (after
    its ReLU; after the bare last Linear: torchvision MLP, ProtNote.py:63-81), ("o", j): after the ReLU of hidden layer j of
    output_layer (get_mlp, ProtNote.py:369-371)}; a dropout module is exactly this multiplication with a random mask."""
    masks = masks or {}
    if not (cfg.feature_fusion.startswith("concatenation") or cfg.feature_fusion == "similarity"):
        raise NotImplementedError(cfg.feature_fusion)
    p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()
         if v.is_floating_point() and not k.startswith("sequence_encoder.") and "running_" not in k}
    full = dict(sd)
    full.update(p)
    new_stats: Dict[str, torch.Tensor] = {}

    def head(prefix, x):
        n = cfg.projection_head_num_layers
        for i in range(n):
            x = x @ full[f"{prefix}.{4 * i}.weight"].T
            if i < n - 1:
                x = torch.relu(_bn_train(x, full, f"{prefix}.{4 * i + 1}", cfg.bn_eps, new_stats))
            if (prefix[2], i) in masks:
                x = x * masks[(prefix[2], i)].to(dtype)
        return x

    # ProtNote.py:83-86: with embedding dropouts the heads are Sequential(Dropout, MLP) and their keys move to W_p.1.* /
    # W_l.1.*.  The dropout itself is a random draw: this oracle takes P_f / L_f AFTER it (the caller replays the masks).
    P_e = head("W_p.1" if cfg.sequence_embedding_dropout > 0 else "W_p", P_f.to(dtype))
    L_e = head("W_l.1" if cfg.label_embedding_dropout > 0 else "W_l", L_f.to(dtype))
    if cfg.feature_fusion == "similarity":      # ProtNote.py:281-284
        logits = F.normalize(P_e, dim=-1, p=2) @ F.normalize(L_e, dim=-1, p=2).t() / cfg.temperature
        return _finish(logits, targets, p, new_stats, dtype, loss_kw)
    x = joint_features(P_e, L_e, cfg.feature_fusion)
    hidden, last = output_mlp_layout(cfg)
    for j, (lin, bn) in enumerate(hidden):
        x = x @ full[f"output_layer.{lin}.weight"].T
        if f"output_layer.{lin}.bias" in full:
            x = x + full[f"output_layer.{lin}.bias"]
        if bn is not None:
            x = _bn_train(x, full, f"output_layer.{bn}", cfg.bn_eps, new_stats)
        x = torch.relu(x)
        if ("o", j) in masks:
            x = x * masks[("o", j)].to(dtype)
    logits = (x @ full[f"output_layer.{last}.weight"].T + full[f"output_layer.{last}.bias"]).reshape(
        P_e.shape[0], L_e.shape[0])
    return _finish(logits, targets, p, new_stats, dtype, loss_kw)


def _finish(logits, targets, p, new_stats, dtype, loss_kw):
    if "pos_weight" in loss_kw and loss_kw["pos_weight"] is not None:
        loss_kw = dict(loss_kw, pos_weight=loss_kw["pos_weight"].to(dtype))
    loss = loss_oracle(logits, targets.to(dtype), **loss_kw)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items() if v.grad is not None}
    return logits.detach(), loss.detach(), grads, new_stats


def synth_targets(B: int, L: int, seed: int = 99, density: float = 0.05):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, L, generator=g) < density).float()


def proteinfer_embeddings_train(sd: Dict[str, torch.Tensor], x, lengths, cfg, prefix: str = "sequence_encoder.",
                                dtype=torch.float64, momentum: float = 0.01):
    """ProteInfer.get_embeddings with the module in .train() mode (protein_encoders.py:109-118 with BatchNorm1d using batch
    statistics over all B x T positions of the masked tensors, :35-37,47-50,61-67).  Returns (embeddings [B, C],
    {running-stat key: updated value}).  Pinned against the reference class by tests/test_train_cpu.py."""
    from .protnote_oracle import masked_conv, zero_padding
    new_stats: Dict[str, torch.Tensor] = {}

    def bn(t, name):
        rm = sd[name + ".running_mean"].detach().clone().to(dtype)
        rv = sd[name + ".running_var"].detach().clone().to(dtype)
        y = F.batch_norm(t, rm, rv, sd[name + ".weight"].to(dtype), sd[name + ".bias"].to(dtype), True, momentum, cfg.bn_eps)
        new_stats[name + ".running_mean"], new_stats[name + ".running_var"] = rm, rv
        return y

    with torch.no_grad():
        f = masked_conv(x.to(dtype), lengths, sd[prefix + "conv1.weight"].to(dtype), sd[prefix + "conv1.bias"].to(dtype), 1)
        for i in range(cfg.num_resnet_blocks):
            q = f"{prefix}resnet_blocks.{i}"
            out = torch.relu(bn(f, q + ".bn_activation_1.0"))
            out = masked_conv(out, lengths, sd[q + ".masked_conv1.weight"].to(dtype), sd[q + ".masked_conv1.bias"].to(dtype),
                              cfg.dilation_base ** i)
            out = torch.relu(bn(out, q + ".bn_activation_2.0"))
            out = masked_conv(out, lengths, sd[q + ".masked_conv2.weight"].to(dtype), sd[q + ".masked_conv2.bias"].to(dtype), 1)
            f = out + f
        f = zero_padding(f, lengths)
        return f.sum(-1) / lengths.reshape(-1, 1).to(dtype), new_stats
