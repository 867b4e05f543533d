"""Training-mode forward + backward of the ProtNote scoring path (reference: `ProtNote.forward` with
`self.training == True`, protnote/models/ProtNote.py:168-334, driven by `ProtNoteTrainer.train_one_epoch`,
protnote/models/ProtNoteTrainer.py:675-825).

What training mode changes with respect to the eval path (csrc/pn_api.cu):
  * every BatchNorm1d uses BATCH statistics - over the B proteins in W_p, over the L label rows in W_l and over all
    B*L (protein, label) pairs in the output MLP (ProtNote.py:63-81,337-378) - and updates its running statistics;
  * the result has to be differentiable with respect to W_p, W_l and output_layer (the sequence encoder is frozen:
    TRAIN_SEQUENCE_ENCODER False, base_config.yaml:71; ProtNoteTrainer.py:211-214).

The arithmetic runs on the sm_100a library through the `pn_t_*` primitives of include/protnote_b200.h (tensor-core GEMMs
on fp16 hi/lo planes + HBM-streaming reduction / normalisation kernels).  This file is launch logic only: it sequences
those primitives, and - when the label axis is sharded over ranks - all-reduces the few per-column statistics that couple
the shards (BatchNorm sums in forward, the two BatchNorm-backward sums in backward).  The same sequencing is exercised on
the CPU by the tests with a torch stand-in for the primitives (oracle/train_ops.py, test infrastructure only).

Configuration coverage (each checked on the CPU against the reference class's autograd and on B200 against the oracle):
every FEATURE_FUSION ('concatenation' / '_diff' folded into the two layer-1 factors, '_prod' with the product block as a
real GEMM, 'similarity'), OUTPUT_MLP_BATCHNORM True / False, OUTPUT_MLP_NUM_LAYERS >= 1, SEQUENCE_ / LABEL_EMBEDDING_DROPOUT
(torch's draw, the reference's RNG order), OUTPUT_MLP_DROPOUT (device-side counter-based masks), the label noise.

Exact identities used (so that the [B*L, 2d] joint tensor never exists in training either):
  layer 1 of the output MLP is linear in [p; t]:  z1[b,l] = a[b] + c[l],  a = P_e W1p^T, c = L_e W1l^T
  its batch statistics over the full B x L grid follow from the factors:
      mean(z1) = mean_b(a) + mean_l(c),     var(z1) = var_b(a) + var_l(c)      (the cross term sums to zero)
  and its backward reduces to  da[b] = sum_l g_z1[b,l],  dc[l] = sum_b g_z1[b,l].
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn


class Comm:
    """The collectives the label-sharded training step needs.  `None` group / world 1 -> no-ops."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0

    def sum_(self, t: torch.Tensor):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def max_(self, t: torch.Tensor):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def bcast_(self, t: torch.Tensor):
        """The first rank's value on every rank (for quantities that are replicated over the ranks but random)."""
        if self.world > 1:
            src = self.dist.get_global_rank(self.group, 0) if self.group is not None else 0
            self.dist.broadcast(t, src=src, group=self.group)
        return t

    def sum_async(self, t: torch.Tensor):
        """Starts the all-reduce of `t` on NCCL's own stream and returns the work handle (None at world 1): the caller keeps
        computing and waits once, so a parameter gradient crosses NVLink while the backward of the next layer runs."""
        if self.world > 1:
            return self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        return None


class _NoComm:
    world = 1
    rank = 0

    def sum_(self, t):
        return t

    def max_(self, t):
        return t

    def bcast_(self, t):
        return t

    def sum_async(self, t):
        return None


class LossSpec:
    """The loss the fused last-layer kernel evaluates (pn_t_bn_relu_dot_loss): the reference's `LOSS_FN` choices that are
    plain per-element functions of (logit, target) - 'BCE' (torch.nn.BCEWithLogitsLoss with optional pos_weight,
    protnote/utils/losses.py:270-272) and 'FocalLoss' (losses.py:171-213; the default of configs/base_config.yaml:61) -
    with reduction 'mean' or 'sum' over the WHOLE batch (B x L_total pairs, all label shards)."""
    KINDS = {"bce": 1, "BCE": 1, "focal": 2, "FocalLoss": 2}

    def __init__(self, kind="bce", pos_weight=None, gamma=2.0, alpha=-1.0, label_smoothing=0.0, reduction="mean"):
        if kind not in self.KINDS:
            raise ValueError(f"fused loss '{kind}' is not implemented (BCE and FocalLoss are); compute it from the logits")
        if reduction not in ("mean", "sum"):
            raise ValueError("reduction must be 'mean' or 'sum'")
        self.kind_id = self.KINDS[kind]
        if self.kind_id == 2 and pos_weight is not None:
            raise ValueError("FocalLoss takes no pos_weight")
        self.pos_weight, self.gamma, self.alpha = pos_weight, float(gamma), float(alpha)
        self.label_smoothing, self.reduction = float(label_smoothing), reduction
        self.grad_scale = 1.0     # set per batch by forward_train


def _unwrap(seq: nn.Module):
    """(modules, p) of a head: ProtNote.py:83-86 wraps W_p / W_l in Sequential(Dropout(p), MLP) when
    SEQUENCE_EMBEDDING_DROPOUT / LABEL_EMBEDDING_DROPOUT > 0 - p is that input dropout, 0 without the wrapper."""
    mods = list(seq)
    if len(mods) == 2 and isinstance(mods[0], nn.Dropout) and isinstance(mods[1], nn.Sequential):
        return list(mods[1]), float(mods[0].p)
    return mods, 0.0


def input_dropout(seq: nn.Module) -> float:
    return _unwrap(seq)[1]


def _split_sequential(seq: nn.Module) -> List[Tuple[nn.Linear, Optional[nn.BatchNorm1d], float]]:
    """[(Linear, BatchNorm1d or None, p)] of a torchvision-MLP-shaped Sequential (ProtNote.py:63-81,337-378); p is the
    probability of the Dropout module that follows the layer (after its BatchNorm / ReLU; after the bare Linear at the end of
    a projection head) - OUTPUT_MLP_DROPOUT, 0 by default (base_config.yaml:39).  The input dropout of a
    Sequential(Dropout, MLP) wrapper is not part of this list (forward_train applies it)."""
    mods, _ = _unwrap(seq)
    out = []
    for i, m in enumerate(mods):
        if isinstance(m, nn.Dropout) and m.p > 0:
            if not out:
                raise NotImplementedError("a Dropout in front of the first Linear (get_mlp's input_dropout) is not implemented "
                                          "on the sm_100a training path; the reference never sets it (ProtNote.py:94-102)")
            out[-1] = (out[-1][0], out[-1][1], float(m.p))
        if isinstance(m, nn.Linear):
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
            out.append((m, bn, 0.0))
    return out


def mlp_layers(model):
    """(W_p layers, W_l layers, output_layer layers) as _split_sequential lists; output_layer is [] for 'similarity'."""
    out = getattr(model, "output_layer", None)
    return _split_sequential(model.W_p), _split_sequential(model.W_l), [] if out is None else _split_sequential(out)


def dropout_sites(model, base_seed: int, rank: int = 0):
    """{(tag, layer index): (seed, p, width)} of a module's active OUTPUT_MLP_DROPOUT sites for one step seeded
    with `base_seed` (tests restate the masks from this; tag 'p' / 'l' / 'o' = W_p / W_l / output_layer)."""
    wp, wl, mods = mlp_layers(model)
    layers = {"p": wp, "l": wl, "o": mods}
    return {(t, i): (seed, p, layers[t][i][0].weight.shape[0])
            for (t, i), (seed, p) in dropout_plan(wp, wl, mods, base_seed, rank).items()}


_SITE_STRIDE = 0xD1B54A32D192ED03
_RANK_STRIDE = 0x9E3779B97F4A7C15
_U64 = (1 << 64) - 1


def dropout_plan(wp, wl, mods, base_seed: int, rank: int = 0):
    """{site: (seed, p)} for every active Dropout inside W_p ('p', i), W_l ('l', i) and output_layer ('o', j).  The mask of a
    site is a pure function of (seed, row, column) (pn_t_dropout_planes): one seed per site from the step's base seed.
    Sites whose rows are this rank's label rows or pairs get a rank-dependent seed; W_p's rows (all B proteins) are
    replicated on every rank and must see the same mask everywhere."""
    plan = {}
    for tag, layers, off, salted in (("p", wp, 0, False), ("l", wl, 1000, True), ("o", mods, 2000, True)):
        for i, (_, _, p) in enumerate(layers):
            if p > 0:
                seed = base_seed + (off + i + 1) * _SITE_STRIDE + (rank * _RANK_STRIDE if salted else 0)
                plan[(tag, i)] = (seed & _U64, float(p))
    return plan


# --------------------------------------------------------------------------------------------------------------------
# projection heads W_p / W_l  (ProtNote.py:63-81):  [Linear(no bias), BN, ReLU] x (n-1), Linear(no bias)
# --------------------------------------------------------------------------------------------------------------------
def head_forward(ops, comm, x_f32, layers, rows_total: int, sharded: bool, update_running: bool, drops=None):
    """drops: {layer index: (seed, p)} of the Dropout modules inside the head (after ReLU of a hidden layer, after the
    final Linear: torchvision MLP order, ProtNote.py:63-81)."""
    drops = drops or {}
    if rows_total <= 1 and any(bn is not None for _, bn, _ in layers):
        # same error as torch.nn.BatchNorm1d in training mode, which the reference module raises here
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x_f32.shape)}")
    x = ops.split(x_f32, want_T=True)
    saved, out = [], None
    for i, (lin, bn, _) in enumerate(layers):
        if lin.bias is not None:
            raise NotImplementedError("projection heads are bias-free in the reference (ProtNote.py:67,77)")
        W = ops.pack(lin.weight)
        drop = drops.get(i)
        if bn is None:
            out = ops.linear(x, W, out_f32=True)
            if drop:
                out = ops.dropout_f32(out, drop)
            saved.append((x, None, None, lin, None, drop))
        else:
            z = ops.linear(x, W, out_f32=False)
            stats = ops.col_stats(z)
            if sharded:
                comm.sum_(stats)
            st = ops.bn_finalize(stats, rows_total, bn, update_running)
            saved.append((x, z, st, lin, bn, drop))
            x = _relu_dropout(ops, ops.bn_relu(z, st, want_T=not drop), drop)
    return out, saved


def _relu_dropout(ops, h, drop):
    """Dropout after a hidden layer's ReLU: the masked / rescaled copy (with the transposed planes the next wgrad reads)."""
    return ops.dropout(h, drop, want_T=True) if drop else h


def head_backward(ops, comm, g_f32, saved, rows_total: int, sharded: bool, grads: Dict):
    if saved and saved[-1][5]:                   # Dropout after the head's last Linear: same mask on the incoming gradient
        g_f32 = ops.dropout_f32(g_f32, saved[-1][5])
    g = ops.split(g_f32, want_T=True, autoscale=True)
    for idx in range(len(saved) - 1, -1, -1):
        x, z, st, lin, bn, drop = saved.pop()    # consumed: activations are released layer by layer
        if st is not None:                       # g is the gradient w.r.t. this layer's output (after ReLU and Dropout)
            if drop:
                g = ops.dropout(g, drop)
            s = ops.bwd_stats(g, z, st)
            grads[bn.weight], grads[bn.bias] = ops.bn_param_grads(s)        # this rank's rows only (before the sum)
            if sharded:
                comm.sum_(s.sums)
            g = ops.bwd_apply(g, z, st, s, rows_total, want_T=True)         # gradient w.r.t. the Linear output
        grads[lin.weight] = ops.wgrad(g, x)
        if idx > 0:
            g = ops.dgrad(g, ops.pack(lin.weight, transposed=True))


# --------------------------------------------------------------------------------------------------------------------
# pair scorer (ProtNote.py:112-126,293,337-378)
# --------------------------------------------------------------------------------------------------------------------
def layer1_factors(lin1: nn.Linear, d: int, fusion: str):
    """The two [H, d] matrices that make layer 1 of the output MLP a sum of a protein term and a label term
    (ProtNote.py:112-152): z1[b,l] = P_e[b] Wp^T + L_e[l] Wl^T.
      'concatenation'       [p; t]        Wp = W[:, :d],               Wl = W[:, d:2d]
      'concatenation_diff'  [p; t; p - t] Wp = W[:, :d] + W[:, 2d:],   Wl = W[:, d:2d] - W[:, 2d:]
      'concatenation_prod'  [p; t; p * t] Wp = W[:, :d],               Wl = W[:, d:2d]   plus the product block, which is
                            not of this form: x[b,l] = (P_e[b] * L_e[l]) W[:, 2d:]^T is added by pairs_forward
    ([H, d] parameter arithmetic, once per step.)"""
    W = lin1.weight.detach()
    if fusion == "concatenation" and W.shape[1] == 2 * d:
        return W[:, :d], W[:, d:]
    if fusion == "concatenation_diff" and W.shape[1] == 3 * d:
        return W[:, :d] + W[:, 2 * d:], W[:, d:2 * d] - W[:, 2 * d:]
    if fusion == "concatenation_prod" and W.shape[1] == 3 * d:
        return W[:, :d], W[:, d:2 * d]
    raise NotImplementedError(f"the training path implements FEATURE_FUSION 'concatenation' (base_config.yaml:44), "
                              f"'concatenation_diff' and 'concatenation_prod'; got '{fusion}' with {W.shape[1]} input "
                              f"features for latent_dim {d}")


def _layer_state(ops, comm, z, lin, bn, count, sharded, update_running):
    """BatchNorm state of one hidden layer from its batch statistics - or, without BatchNorm (OUTPUT_MLP_BATCHNORM False:
    get_mlp then gives the Linear a bias, ProtNote.py:352-365), the fixed affine map relu(z + bias)."""
    if bn is None:
        return ops.affine_state(lin.bias, lin.weight.shape[0], lin.weight.device)
    stats = ops.col_stats(z)
    if sharded:
        comm.sum_(stats)
    return ops.bn_finalize(stats, count, bn, update_running)


def pairs_forward(ops, comm, P_e, L_e, hidden, final: nn.Linear, L_total: int, sharded: bool, update_running: bool,
                  loss: Optional[LossSpec] = None, targets=None, fusion: str = "concatenation", drops=None):
    """drops: {hidden layer index: (seed, p)} of the Dropout modules get_mlp puts after every ReLU but the last
    (ProtNote.py:369-371)."""
    drops = drops or {}
    B, d = P_e.shape
    lin1, bn1 = hidden[0][0], hidden[0][1]
    if drops.get(len(hidden) - 1):
        raise NotImplementedError("a Dropout between the last hidden layer and the output neuron is not implemented "
                                  "(get_mlp has none there, ProtNote.py:369-371)")
    W1p, W1l = layer1_factors(lin1, d, fusion)
    pe = ops.split(P_e, want_T=True)
    le = ops.split(L_e, want_T=True)
    a = ops.linear(pe, ops.pack(W1p), out_f32=True)                         # [B, H] raw protein term
    c = ops.linear(le, ops.pack(W1l), out_f32=True)                         # [L_local, H] raw label term
    want_T1 = len(hidden) > 1 and not drops.get(0)
    prod = None
    if fusion == "concatenation_prod":
        # layer 1 with the product block: z1 = x + a[b] + c[l], x = (P_e[b] * L_e[l]) Wx^T a real GEMM over all pairs; with
        # z1 as planes it is an ordinary layer (statistics pass, BN + ReLU), factorised only in its two linear blocks
        W1x = lin1.weight.detach()[:, 2 * d:]
        q = ops.pair_product(P_e, L_e, want_T=True)                         # [B * L_local, d]
        z1 = ops.pair_add(ops.linear(q, ops.pack(W1x), out_f32=False), a, c)
        st1 = _layer_state(ops, comm, z1, lin1, bn1, B * L_total, sharded, update_running)
        h = ops.bn_relu(z1, st1, want_T=want_T1)
        prod = {"q": q, "z1": z1, "W1x": W1x, "P_e": P_e, "L_e": L_e}
    else:
        if bn1 is None:
            st1 = ops.affine_state(lin1.bias, lin1.weight.shape[0], lin1.weight.device)
        else:
            sa = ops.col_stats_f32(a)
            sc = ops.col_stats_f32(c)
            if sharded:
                comm.sum_(sc)
            st1 = ops.bn_finalize_pair(sa, B, sc, L_total, bn1, update_running)
        h = ops.pair_hidden(a, c, st1, want_T=want_T1)                      # [B * L_local, H]
    h = _relu_dropout(ops, h, drops.get(0))
    ctx = {"drops": drops, "prod": prod, "pe": pe, "le": le, "a": a, "c": c, "st1": st1, "layers": [], "B": B, "d": d, "hidden": hidden,
           "final": final, "count": B * L_total, "fusion": fusion, "W1p": W1p, "W1l": W1l}
    logits = None
    if len(hidden) == 1:
        # OUTPUT_MLP_NUM_LAYERS 1: the output neuron follows layer 1 directly.  h1 >= 0, so relu(h1 * 1 + 0) = h1 and the
        # dot kernel of the last hidden layer applies to the h1 planes with the identity state
        ident = ops.affine_state(None, lin1.weight.shape[0], lin1.weight.device)
        ctx["one_layer"] = (h, ident)
        if loss is None:
            logits = ops.bn_relu_dot(h, ident, final.weight, final.bias)
        else:
            loss.grad_scale = 1.0 / float(B * L_total) if loss.reduction == "mean" else 1.0
            logits, ctx["g_seed"], ctx["loss_sum"] = ops.bn_relu_dot_loss(h, ident, final.weight, final.bias, targets,
                                                                          L_e.shape[0], loss)
    for j in range(1, len(hidden)):
        lin, bn = hidden[j][0], hidden[j][1]
        z = ops.linear(h, ops.pack(lin.weight), out_f32=False)
        st = _layer_state(ops, comm, z, lin, bn, B * L_total, sharded, update_running)
        ctx["layers"].append((h, z, st, lin, bn))
        if j + 1 < len(hidden):
            h = _relu_dropout(ops, ops.bn_relu(z, st, want_T=not drops.get(j)), drops.get(j))
        elif loss is None:                                                  # the last hidden layer is never stored:
            logits = ops.bn_relu_dot(z, st, final.weight, final.bias)       # relu(BN(z)) . w_out + b_out per pair
        else:                                                               # ... and the loss + its gradient seed are fused in
            loss.grad_scale = 1.0 / float(B * L_total) if loss.reduction == "mean" else 1.0
            logits, ctx["g_seed"], ctx["loss_sum"] = ops.bn_relu_dot_loss(z, st, final.weight, final.bias, targets,
                                                                          L_e.shape[0], loss)
    return logits, ctx


class _Grads(dict):
    """{parameter: gradient}.  With `reduce_comm` set every gradient is all-reduced (sum over the label-sharded ranks) as
    soon as it is stored, asynchronously: the 303 MB of parameter gradients cross NVLink behind the dgrad / wgrad GEMMs of
    the layers still to come instead of after the backward as one blocking cat -> all_reduce -> copy."""

    def __init__(self, reduce_comm=None):
        super().__init__()
        self.reduce_comm, self.pending = reduce_comm, []

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        if self.reduce_comm is not None and v is not None:
            w = self.reduce_comm.sum_async(v)
            if w is not None:
                self.pending.append(w)

    def wait(self):
        for w in self.pending:
            w.wait()
        self.pending = []


def _affine_grads(ops, comm, s, lin, bn, sharded: bool, grads: Dict):
    """Parameter gradients of the normalisation that follows `lin`, from the two column sums of the backward statistics
    pass (sum g_y, sum g_y * xhat).  With BatchNorm they are d beta / d gamma, and the sums - completed over the label
    shards - go on into the BatchNorm backward.  Without BatchNorm the first sum is the gradient of the Linear's bias and
    nothing is subtracted from g_y: the sums are cleared, which turns bwd_apply into the plain ReLU backward."""
    if bn is not None:
        grads[bn.weight], grads[bn.bias] = ops.bn_param_grads(s)
        if sharded:
            comm.sum_(s.sums)
        return
    if lin.bias is not None:
        grads[lin.bias] = ops.bn_param_grads(s)[1]
    s.sums.zero_()


def pairs_backward(ops, comm, ctx, g_logit, sharded: bool, grads: Dict):
    hidden, final, count = ctx["hidden"], ctx["final"], ctx["count"]
    layers, drops = ctx["layers"], ctx["drops"]
    # ---- last hidden layer: the incoming gradient is the outer product g_logit (x) w_out, generated on the fly
    go = ops.outer(g_logit, final.weight)
    if "one_layer" in ctx:
        # the same pass over the h1 planes with the identity state: d w_out = sum g_logit * h1, d b_out, and the gradient
        # w.r.t. h1 - masked where h1 = 0, which is the mask layer 1's own backward applies to it anyway
        h, ident = ctx.pop("one_layer")
        s = ops.bwd_stats(go, h, ident)
        grads[final.weight], grads[final.bias] = ops.final_param_grads(s)
        s.sums.zero_()
        g = ops.bwd_apply(go, h, ident, s, count)
        del h
    else:
        h_prev, z, st, lin, bn = layers.pop()    # consumed: activations are released layer by layer
        s = ops.bwd_stats(go, z, st)
        grads[final.weight], grads[final.bias] = ops.final_param_grads(s)
        _affine_grads(ops, comm, s, lin, bn, sharded, grads)
        g = ops.bwd_apply(go, z, st, s, count, want_T=True)
        grads[lin.weight] = ops.wgrad(g, h_prev)
        del h_prev, z
        g = ops.dgrad(g, ops.pack(lin.weight, transposed=True))
    # ---- middle layers
    while layers:
        h_prev, z, st, lin, bn = layers.pop()
        if drops.get(len(layers) + 1):           # g is the gradient w.r.t. the dropped output of hidden layer len(layers) + 1
            g = ops.dropout(g, drops[len(layers) + 1])
        s = ops.bwd_stats(g, z, st)
        _affine_grads(ops, comm, s, lin, bn, sharded, grads)
        g = ops.bwd_apply(g, z, st, s, count, want_T=True)
        grads[lin.weight] = ops.wgrad(g, h_prev)
        del h_prev, z
        g = ops.dgrad(g, ops.pack(lin.weight, transposed=True))
    # ---- layer 1: z1 = a[b] + c[l] is regenerated, its gradient is reduced to the two factors
    lin1, bn1 = hidden[0][0], hidden[0][1]
    d = ctx["d"]
    if drops.get(0):
        g = ops.dropout(g, drops[0])
    if ctx["prod"] is not None:
        return _layer1_backward_prod(ops, comm, ctx, g, sharded, grads)
    zp = ops.pair_source(ctx["a"], ctx["c"])
    s = ops.bwd_stats(g, zp, ctx["st1"])
    _affine_grads(ops, comm, s, lin1, bn1, sharded, grads)
    da, dc = ops.bwd_apply_pair(g, zp, ctx["st1"], s, count)                # fp32 [B, H], [L_local, H]
    dW1 = torch.empty_like(lin1.weight)
    ga = ops.split(da, want_T=True, autoscale=True)
    gc = ops.split(dc, want_T=True, autoscale=True)
    ops.wgrad(ga, ctx["pe"], out=dW1[:, :d])                                # d z1 / d Wp = da^T P_e
    ops.wgrad(gc, ctx["le"], out=dW1[:, d:2 * d])                           # d z1 / d Wl = dc^T L_e
    if ctx["fusion"] == "concatenation_diff":                               # Wp = W_p + W_d, Wl = W_t - W_d (layer1_factors)
        torch.sub(dW1[:, :d], dW1[:, d:2 * d], out=dW1[:, 2 * d:])
    grads[lin1.weight] = dW1
    dPe = ops.dgrad(ga, ops.pack(ctx["W1p"], transposed=True), out_f32=True)
    dLe = ops.dgrad(gc, ops.pack(ctx["W1l"], transposed=True), out_f32=True)
    return dPe, dLe


def _layer1_backward_prod(ops, comm, ctx, g, sharded: bool, grads: Dict):
    """Layer 1 of the output MLP on [p; t; p * t]: an ordinary layer backward on the z1 planes, then the gradient w.r.t.
    z1[b,l] = a[b] + c[l] + q[b,l] Wx^T goes three ways: its two marginals to the linear blocks (da, dc), d Wx = g_z1^T q, and
    through g_q = g_z1 Wx to d P_e[b] += sum_l g_q[b,l] * L_e[l], d L_e[l] += sum_b g_q[b,l] * P_e[b]."""
    lin1, bn1 = ctx["hidden"][0][0], ctx["hidden"][0][1]
    d, B, pr = ctx["d"], ctx["B"], ctx["prod"]
    L = ctx["c"].shape[0]
    s = ops.bwd_stats(g, pr["z1"], ctx["st1"])
    _affine_grads(ops, comm, s, lin1, bn1, sharded, grads)
    gz = ops.bwd_apply(g, pr["z1"], ctx["st1"], s, ctx["count"], want_T=True)
    dW1 = torch.empty_like(lin1.weight)
    ops.wgrad(gz, pr["q"], out=dW1[:, 2 * d:])
    da, dc = ops.pair_marginals(gz, B, L)                                   # fp32 [B, H], [L_local, H]
    gq = ops.dgrad(gz, ops.pack(pr["W1x"], transposed=True))                # [B * L_local, d]
    dPe, dLe = ops.pair_marginals(gq, B, L, wb=pr["P_e"], wl=pr["L_e"])
    ga = ops.split(da, want_T=True, autoscale=True)
    gc = ops.split(dc, want_T=True, autoscale=True)
    ops.wgrad(ga, ctx["pe"], out=dW1[:, :d])
    ops.wgrad(gc, ctx["le"], out=dW1[:, d:2 * d])
    grads[lin1.weight] = dW1
    ops.dgrad(ga, ops.pack(ctx["W1p"], transposed=True), out_f32=True, accumulate_into=dPe)
    ops.dgrad(gc, ops.pack(ctx["W1l"], transposed=True), out_f32=True, accumulate_into=dLe)
    return dPe, dLe


# --------------------------------------------------------------------------------------------------------------------
# FEATURE_FUSION similarity (ProtNote.py:281-284): logits = normalize(P_e) normalize(L_e)^T / temperature, no output MLP
# --------------------------------------------------------------------------------------------------------------------
def similarity_forward(ops, P_e, L_e, temperature: float):
    scale = 1.0 / float(temperature)
    Pn, ip = ops.normalize_rows(P_e, scale)                                 # the 1 / temperature rides on the protein rows
    Ln, il = ops.normalize_rows(L_e, 1.0)
    logits = ops.linear(ops.split(Pn), ops.pack(Ln), out_f32=True)          # [B, L_local]
    return logits, {"Pn": Pn, "ip": ip, "Ln": Ln, "il": il, "scale": scale}


def similarity_backward(ops, sctx, g_logits):
    """g_logits [B, L_local] -> (d P_e, d L_e): d Pn = G Ln, d Ln = G^T Pn (tensor-core engine), then the row-normalisation's
    backward.  Linear in G, so per-rank label slabs add up like every other gradient of the sharded step."""
    g = ops.split(g_logits, want_T=True, autoscale=True)
    dPn = ops.dgrad(g, ops.pack(sctx["Ln"], transposed=True), out_f32=True)             # [B, d]
    dLn = ops.wgrad(g, ops.split(sctx["Pn"], want_T=True))                               # [L_local, d]
    return (ops.normalize_rows_bwd(sctx["Pn"], sctx["ip"], dPn, sctx["scale"]),
            ops.normalize_rows_bwd(sctx["Ln"], sctx["il"], dLn, 1.0))


# --------------------------------------------------------------------------------------------------------------------
# whole step
# --------------------------------------------------------------------------------------------------------------------
def trainable_parameters(model) -> List[nn.Parameter]:
    ps = []
    for part in (model.W_p, model.W_l, getattr(model, "output_layer", None)):
        if part is None:        # FEATURE_FUSION similarity has no output MLP (ProtNote.py:93)
            continue
        ps += [p for p in part.parameters()]
    return ps


def forward_train(ops, comm, model, P_f, L_f, L_total: Optional[int] = None, update_running: bool = True,
                  loss: Optional[LossSpec] = None, targets=None, drop_seed: Optional[int] = None):
    """P_f [B, protein_dim] (all proteins), L_f [L_local, label_dim] (this rank's label rows) -> logits [B, L_local].
    With `loss` (and targets [B, L_local]) the last kernel also leaves ctx['pairs']['loss_sum'] / ['g_seed']."""
    comm = comm or _NoComm()
    sharded = comm.world > 1
    L_total = int(L_total if L_total is not None else L_f.shape[0])
    B = P_f.shape[0]
    wp, wl = _split_sequential(model.W_p), _split_sequential(model.W_l)
    similarity = getattr(model, "feature_fusion", "concatenation") == "similarity"
    if similarity and loss is not None:
        raise NotImplementedError("the fused loss lives in the output MLP's last kernel; with FEATURE_FUSION 'similarity' "
                                  "take train_logits() and compute the loss from the logits")
    mods = [] if similarity else _split_sequential(model.output_layer)
    hidden, final = (mods[:-1], mods[-1][0]) if mods else ([], None)
    # SEQUENCE_EMBEDDING_DROPOUT / LABEL_EMBEDDING_DROPOUT (ProtNote.py:83-86): RNG-stream dependent like the label noise,
    # so they stay the reference's own torch op, drawn in the reference's order (W_p before W_l, ProtNote.py:270-271).
    # Every rank holds all proteins: the dropped P_f is the first rank's; label rows are per rank.
    if input_dropout(model.W_p) > 0:
        P_f = comm.bcast_(torch.nn.functional.dropout(P_f, input_dropout(model.W_p), True).contiguous())
    if input_dropout(model.W_l) > 0:
        L_f = torch.nn.functional.dropout(L_f, input_dropout(model.W_l), True)
    # OUTPUT_MLP_DROPOUT: one base seed per step from torch's CPU generator (so torch.manual_seed governs it), the first
    # rank's on every rank; per-site seeds in dropout_plan
    plan = {}
    if any(p > 0 for _, _, p in wp + wl + mods):
        if drop_seed is None:
            drop_seed = int(torch.randint(0, 1 << 62, (1,), dtype=torch.int64))
        if sharded:
            drop_seed = int(comm.bcast_(torch.tensor([drop_seed], dtype=torch.int64, device=P_f.device)))
        plan = dropout_plan(wp, wl, mods, drop_seed, comm.rank if sharded else 0)
    P_e, saved_p = head_forward(ops, comm, P_f, wp, B, False, update_running,
                                {i: v for (t, i), v in plan.items() if t == "p"})
    L_e, saved_l = head_forward(ops, comm, L_f, wl, L_total, sharded, update_running,
                                {i: v for (t, i), v in plan.items() if t == "l"})
    if similarity:
        logits, pctx = similarity_forward(ops, P_e, L_e, model.temperature)
    else:
        logits, pctx = pairs_forward(ops, comm, P_e, L_e, hidden, final, L_total, sharded, update_running, loss, targets,
                                     getattr(model, "feature_fusion", "concatenation"),
                                     {i: v for (t, i), v in plan.items() if t == "o"})
    ctx = {"saved_p": saved_p, "saved_l": saved_l, "pairs": pctx, "B": B, "L_total": L_total, "sharded": sharded,
           "drop_plan": plan, "similarity": similarity}
    if update_running:
        for part in (model.W_p, model.W_l) + (() if similarity else (model.output_layer,)):
            for m in part.modules():
                if isinstance(m, nn.BatchNorm1d) and m.num_batches_tracked is not None:
                    m.num_batches_tracked += 1
        # the kernels updated the running statistics through raw pointers (torch's version counters did not move):
        # the eval-mode weight pack, which folds them, and the cached label projection are stale from here on
        if hasattr(model, "_packed_key"):
            model._packed_key = None
        if hasattr(model, "_label_cache"):
            model._label_cache = None
    return logits.reshape(B, L_f.shape[0]), ctx


def backward_train(ops, comm, ctx, g_logits, reduce_gradients: bool = False) -> Dict[nn.Parameter, torch.Tensor]:
    """g_logits [B, L_local] -> {parameter: gradient} for this rank's label rows (sum over ranks = full gradient).
    reduce_gradients: all-reduce every gradient over the ranks inside the backward, overlapped with it (see _Grads); the
    returned gradients are then those of the whole batch and `allreduce_gradients` must not be called again."""
    comm = comm or _NoComm()
    sharded = ctx["sharded"]
    grads = _Grads(comm if (reduce_gradients and sharded) else None)
    if ctx["similarity"]:
        dPe, dLe = similarity_backward(ops, ctx["pairs"], g_logits)
    else:
        dPe, dLe = pairs_backward(ops, comm, ctx["pairs"], g_logits.reshape(-1), sharded, grads)
    head_backward(ops, comm, dLe, ctx["saved_l"], ctx["L_total"], sharded, grads)      # the larger head first
    head_backward(ops, comm, dPe, ctx["saved_p"], ctx["B"], False, grads)
    grads.wait()
    return grads


class _TrainFunction(torch.autograd.Function):
    """Connects the primitives to autograd: the trainable parameters are inputs, so `loss.backward()` in the caller
    (ProtNoteTrainer.py:738) delivers their gradients exactly as it does for the reference module."""

    @staticmethod
    def forward(fctx, ops, comm, model, P_f, L_f, L_total, reduce_gradients, *params):
        logits, ctx = forward_train(ops, comm, model, P_f, L_f, L_total)
        fctx.ops, fctx.comm, fctx.tctx, fctx.params, fctx.reduce = ops, comm, ctx, params, reduce_gradients
        return logits

    @staticmethod
    def backward(fctx, g_logits):
        grads = backward_train(fctx.ops, fctx.comm, fctx.tctx, g_logits.contiguous(), fctx.reduce)
        fctx.tctx = None
        out = tuple(grads.get(p) if p.requires_grad else None for p in fctx.params)
        return (None, None, None, None, None, None, None) + out


def train_logits(model, P_f, L_f, ops=None, comm=None, L_total=None, reduce_gradients=False):
    """Differentiable training-mode logits [B, L_local] of a protnote_b200.ProtNote module.
    reduce_gradients=True (label-sharded ranks): the backward all-reduces the parameter gradients itself, overlapped with
    the remaining backward GEMMs; do not call `allreduce_gradients` afterwards."""
    if ops is None:
        from .train_native import NativeOps
        ops = NativeOps(model.precision)
    params = trainable_parameters(model)
    return _TrainFunction.apply(ops, comm, model, P_f.detach(), L_f.detach(), L_total, bool(reduce_gradients), *params)


class _TrainLossFunction(torch.autograd.Function):
    """Loss fused into the last kernel of the forward: returns (loss, logits); the logits carry no graph
    (they are for metrics), `loss.backward()` seeds the backward with the d loss / d logit the forward kernel left behind."""

    @staticmethod
    def forward(fctx, ops, comm, model, P_f, L_f, targets, spec, L_total, reduce_gradients, *params):
        logits, ctx = forward_train(ops, comm, model, P_f, L_f, L_total, loss=spec, targets=targets)
        pairs = ctx["pairs"]
        total = float(ctx["B"] * ctx["L_total"]) if spec.reduction == "mean" else 1.0
        loss = (pairs.pop("loss_sum") / total).to(torch.float32).reshape(())
        fctx.ops, fctx.comm, fctx.tctx, fctx.params, fctx.reduce = ops, comm, ctx, params, reduce_gradients
        fctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(fctx, g_loss, _g_logits):
        pairs = fctx.tctx["pairs"]
        g_logits = pairs.pop("g_seed") * g_loss.to(torch.float32)
        grads = backward_train(fctx.ops, fctx.comm, fctx.tctx, g_logits, fctx.reduce)
        fctx.tctx = None
        out = tuple(grads.get(p) if p.requires_grad else None for p in fctx.params)
        return (None,) * 9 + out


def train_loss(model, P_f, L_f, targets, loss="bce", pos_weight=None, gamma=2.0, alpha=-1.0, label_smoothing=0.0,
               reduction="mean", ops=None, comm=None, L_total=None, reduce_gradients=False):
    """(loss, logits [B, L_local]) of one training-mode forward with the loss evaluated inside the last kernel.
    `loss`: 'bce' (BCEWithLogitsLoss, optional pos_weight [L_local] or scalar) or 'focal' (FocalLoss(alpha, gamma,
    label_smoothing), protnote/utils/losses.py:171-213).  On label-sharded ranks the returned loss is this rank's share
    (sum of its pairs' losses / (B * L_total) for 'mean'): the sum over ranks is the loss of the whole batch, and
    `loss.backward()` produces exactly the gradients the unsharded step would."""
    if ops is None:
        from .train_native import NativeOps
        ops = NativeOps(model.precision)
    spec = LossSpec(loss, pos_weight, gamma, alpha, label_smoothing, reduction)
    params = trainable_parameters(model)
    return _TrainLossFunction.apply(ops, comm, model, P_f.detach(), L_f.detach(), targets.detach(), spec, L_total,
                                    bool(reduce_gradients), *params)


def allreduce_gradients(model, comm: Comm):
    """Label-sharded step: every rank holds the gradient of ITS label rows; the sum over ranks is the gradient of the
    whole B x L batch (backward is linear in the incoming gradient once the BatchNorm sums have been shared)."""
    ps = [p for p in trainable_parameters(model) if p.grad is not None]
    if comm.world == 1 or not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    comm.sum_(flat)
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
