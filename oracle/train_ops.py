"""TEST INFRASTRUCTURE ONLY - a torch (CPU) stand-in for the `pn_t_*` training primitives.

`protnote_b200/train.py` sequences a small set of primitives (GEMMs on fp16 planes, column statistics, BatchNorm
normalise / backward kernels).  On the product path those primitives are the sm_100a kernels bound in
`protnote_b200/train_native.py`.  This class implements the SAME primitive contracts with plain torch tensor
arithmetic so that
  * the sequencing logic (analytic layer-1 statistics, BatchNorm backward through shared sums, label sharding) can be
    checked on the CPU against the reference's autograd (tests/test_train_cpu.py, tests/test_train_gloo.py), and
  * every CUDA primitive has an independent statement of what it must compute (tests/test_gpu_train.py).
The product package never imports this file.

BatchNorm1d training semantics follow torch.nn.BatchNorm1d as used by the reference (ProtNote.py:63-81 via
torchvision.ops.MLP, ProtNote.py:364-365): biased batch variance for normalisation, unbiased for running_var,
running = (1 - momentum) * running + momentum * batch.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class TAct:
    val: torch.Tensor            # [rows, cols] = true value * sc
    sc: float = 1.0
    has_T: bool = False


@dataclass
class TPacked:
    w: torch.Tensor              # [N, K] (already transposed when packed with transposed=True)


@dataclass
class TBN:
    scale: torch.Tensor
    shift: torch.Tensor
    mean: torch.Tensor
    invstd: torch.Tensor


@dataclass
class TOuter:
    g_logit: torch.Tensor        # [rows]
    w: torch.Tensor              # [cols]


@dataclass
class TPair:
    a: torch.Tensor              # [B, H]
    c: torch.Tensor              # [L, H]


@dataclass
class TBwdStats:
    sums: torch.Tensor           # float64 [2, cols]: sum g_y, sum g_y * xhat   (true scale)
    maxes: torch.Tensor          # [2]: max |g| (bounds |g_y|), max |z| (bounds |xhat| through invstd and mean)
    local: Optional[torch.Tensor] = None
    dw: Optional[torch.Tensor] = None
    db: Optional[torch.Tensor] = None


def pow2_scale(absmax: float) -> float:
    """power of two that puts `absmax` into [2^5, 2^6) - the rule of pn_t_autoscale / pn_t_bwd_scale."""
    if not (absmax > 0) or not math.isfinite(absmax):
        return 1.0
    _, e = math.frexp(absmax)    # absmax = f * 2^e, f in [0.5, 1)
    return math.ldexp(1.0, 6 - e)


def dropout_multiplier(seed: int, rows: int, cols: int, p: float) -> torch.Tensor:
    """Statement of the keep mask of pn_t_dropout_planes / pn_t_dropout_f32 (csrc/pn_train.cuh: drop_keep8), as the fp64
    multiplier keep * 65536 / (65536 - thr), [rows, cols].  Counter-based: element (r, c) takes 16 bits of the splitmix64
    output of counter 2 * (r * ceil(cols / 8) + c // 8) + 1 + (c % 8) // 4; keep <=> bits >= thr = round(p * 65536)."""
    import numpy as np
    thr = min(int(np.rint(np.float32(p) * np.float32(65536.0))), 65535)
    scale = float(np.float32(65536.0) / np.float32(65536 - thr))
    groups = (cols + 7) // 8
    r = np.arange(rows, dtype=np.uint64)[:, None]
    c = np.arange(groups * 8, dtype=np.uint64)[None, :]
    ctr = np.uint64(2) * (r * np.uint64(groups) + c // np.uint64(8)) + np.uint64(1) + (c % np.uint64(8)) // np.uint64(4)
    with np.errstate(over="ignore"):
        x = np.uint64(seed & ((1 << 64) - 1)) + ctr * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    bits = (x >> (np.uint64(16) * (c % np.uint64(4)))) & np.uint64(0xFFFF)
    keep = (bits >= np.uint64(thr))[:, :cols]
    return torch.from_numpy(keep.astype(np.float64)) * scale


class TorchOps:
    def __init__(self, dtype=torch.float64):
        self.dtype = dtype

    # ------------------------------------------------------------------ dropout inside the MLPs
    def dropout(self, x: TAct, drop, want_T=False):
        seed, p = drop
        return TAct(x.val * dropout_multiplier(seed, x.val.shape[0], x.val.shape[1], p).to(self.dtype), x.sc, want_T)

    def dropout_f32(self, x, drop):
        seed, p = drop
        return x * dropout_multiplier(seed, x.shape[0], x.shape[1], p).to(x.dtype)

    # ------------------------------------------------------------------ operands
    def split(self, x, want_T=False, autoscale=False):
        x = x.detach().to(self.dtype)
        sc = pow2_scale(float(x.abs().max())) if autoscale and x.numel() else 1.0
        return TAct(x * sc, sc, want_T)

    def pack(self, W, transposed=False):
        W = W.detach().to(self.dtype)
        return TPacked(W.t().contiguous() if transposed else W.contiguous())

    # ------------------------------------------------------------------ GEMMs
    def linear(self, x: TAct, W: TPacked, out_f32=False, accumulate_into=None):
        y = (x.val / x.sc) @ W.w.t()
        if accumulate_into is not None:
            accumulate_into += y.to(accumulate_into.dtype)
            return accumulate_into
        return y if out_f32 else TAct(y, 1.0, False)

    def dgrad(self, g: TAct, WT: TPacked, out_f32=False, accumulate_into=None):
        y = (g.val / g.sc) @ WT.w.t()
        if accumulate_into is not None:
            accumulate_into += y.to(accumulate_into.dtype)
            return accumulate_into
        return y if out_f32 else TAct(y * g.sc, g.sc, False)

    # ------------------------------------------------------------------ FEATURE_FUSION similarity
    def normalize_rows(self, x, scale=1.0):
        x = x.detach().to(self.dtype)
        inv = 1.0 / x.norm(dim=1).clamp_min(1e-12)
        return x * inv[:, None] * scale, inv

    def normalize_rows_bwd(self, y, inv_norm, dy, scale=1.0):
        u = y / scale
        dy = dy.to(self.dtype)
        return scale * inv_norm[:, None] * (dy - u * (u * dy).sum(1, keepdim=True))

    # ------------------------------------------------------------------ FEATURE_FUSION concatenation_prod
    def pair_product(self, P_e, L_e, want_T=False):
        P_e, L_e = P_e.detach().to(self.dtype), L_e.detach().to(self.dtype)
        return TAct((P_e[:, None, :] * L_e[None, :, :]).reshape(-1, P_e.shape[1]), 1.0, want_T)

    def pair_add(self, x: TAct, a, c):
        z = (a.to(self.dtype)[:, None, :] + c.to(self.dtype)[None, :, :]).reshape(-1, a.shape[1])
        return TAct(x.val / x.sc + z, 1.0, False)

    def pair_marginals(self, g: TAct, B, L, wb=None, wl=None):
        G = (g.val / g.sc).reshape(B, L, -1)
        gb = G if wl is None else G * wl.detach().to(self.dtype)[None, :, :]
        gl = G if wb is None else G * wb.detach().to(self.dtype)[:, None, :]
        return gb.sum(1), gl.sum(0)

    def wgrad(self, g: TAct, x: TAct, out=None):
        assert g.has_T and x.has_T, "wgrad contracts over rows: both operands need their transposed planes"
        dW = (g.val / g.sc).t() @ (x.val / x.sc)
        if out is not None:
            out.copy_(dW.to(out.dtype))
            return out
        return dW

    # ------------------------------------------------------------------ statistics / BatchNorm forward
    def col_stats(self, z: TAct):
        v = (z.val / z.sc).double()
        return torch.stack([v.sum(0), (v * v).sum(0)])

    def col_stats_f32(self, x):
        v = x.double()
        return torch.stack([v.sum(0), (v * v).sum(0)])

    def _finalize(self, mean, var, count, bn, update_running):
        eps = bn.eps
        invstd = 1.0 / torch.sqrt(var + eps)
        scale = bn.weight.detach().double() * invstd
        shift = bn.bias.detach().double() - mean * scale
        if update_running and bn.track_running_stats:
            m = bn.momentum
            with torch.no_grad():
                bn.running_mean.mul_(1 - m).add_((m * mean).to(bn.running_mean.dtype))
                unbiased = var * (count / max(count - 1.0, 1.0))
                bn.running_var.mul_(1 - m).add_((m * unbiased).to(bn.running_var.dtype))
        t = self.dtype
        return TBN(scale.to(t), shift.to(t), mean.to(t), invstd.to(t))

    def bn_finalize(self, stats, count, bn, update_running=True):
        mean = stats[0] / count
        var = (stats[1] / count - mean * mean).clamp_min(0)
        return self._finalize(mean, var, float(count), bn, update_running)

    def bn_finalize_pair(self, sa, B, sc, L, bn, update_running=True):
        ma, mc = sa[0] / B, sc[0] / L
        va = (sa[1] / B - ma * ma).clamp_min(0)
        vc = (sc[1] / L - mc * mc).clamp_min(0)
        return self._finalize(ma + mc, va + vc, float(B) * float(L), bn, update_running)

    def affine_state(self, bias, cols, device=None):
        """State of a layer without BatchNorm: relu(z * 1 + bias); mean 0 / invstd 1 make xhat = z, which the backward
        only ever multiplies with a cleared sum."""
        t = self.dtype
        b = torch.zeros(cols, dtype=t) if bias is None else bias.detach().to(t).reshape(-1)
        return TBN(torch.ones(cols, dtype=t), b, torch.zeros(cols, dtype=t), torch.ones(cols, dtype=t))

    def bn_relu(self, z: TAct, st: TBN, want_T=False):
        return TAct(torch.relu(z.val * st.scale + st.shift), 1.0, want_T)

    def bn_relu_dot(self, z: TAct, st: TBN, w, b):
        h = torch.relu(z.val * st.scale + st.shift)
        return h @ w.detach().to(self.dtype).reshape(-1) + b.detach().to(self.dtype).reshape(())

    def bn_relu_dot_loss(self, z: TAct, st: TBN, w, b, targets, L, spec):
        """Statement of pn_t_bn_relu_dot_loss: logits, grad_scale * d loss / d logit, sum of per-pair losses - with the
        reference's own loss formulas (protnote/utils/losses.py:171-213 FocalLoss, :270-272 BCE) under autograd."""
        with torch.enable_grad():       # (called from inside autograd.Function.forward, where grad mode is off)
            return self._loss_and_seed(self.bn_relu_dot(z, st, w, b).detach().requires_grad_(True), targets, L, spec)

    def _loss_and_seed(self, x, targets, L, spec):
        t = targets.detach().to(self.dtype).reshape(-1)
        if spec.kind_id == 1:
            pw = None
            if spec.pos_weight is not None:
                pw = spec.pos_weight.detach().to(self.dtype).reshape(-1)
                pw = (pw.expand(L) if pw.numel() == 1 else pw).repeat(x.numel() // L)
            per = torch.nn.functional.binary_cross_entropy_with_logits(x, t, reduction="none", pos_weight=pw)
        else:
            ts = t * (1.0 - spec.label_smoothing) + (1 - t) * spec.label_smoothing if spec.label_smoothing > 0 else t
            bce = torch.nn.functional.binary_cross_entropy_with_logits(x, ts, reduction="none")
            per = ((1 - torch.exp(-bce)) ** spec.gamma) * bce
            if spec.alpha >= 0:
                per = (spec.alpha * ts + (1 - spec.alpha) * (1 - ts)) * per
        total = per.sum()
        (g,) = torch.autograd.grad(total, x)
        return x.detach(), g * spec.grad_scale, total.detach().reshape(1)

    def pair_hidden(self, a, c, st: TBN, want_T=False):
        z = (a[:, None, :] + c[None, :, :]).reshape(-1, a.shape[1])
        return TAct(torch.relu(z * st.scale + st.shift), 1.0, want_T)

    # ------------------------------------------------------------------ BatchNorm + ReLU backward
    def outer(self, g_logit, w):
        return TOuter(g_logit.detach().to(self.dtype).reshape(-1), w.detach().to(self.dtype).reshape(-1))

    def pair_source(self, a, c):
        return TPair(a, c)

    def _g(self, g):
        if isinstance(g, TOuter):
            return g.g_logit[:, None] * g.w[None, :]
        return g.val / g.sc

    def _z(self, z):
        if isinstance(z, TPair):
            return (z.a[:, None, :] + z.c[None, :, :]).reshape(-1, z.a.shape[1])
        return z.val / z.sc

    def bwd_stats(self, g, z, st: TBN):
        G, Z = self._g(g), self._z(z)
        pre = Z * st.scale + st.shift
        gy = G * (pre > 0)
        xh = (Z - st.mean) * st.invstd
        # maxes: max |g| before the mask (a bound of max |g_y|) and max |z| (bwd_scale bounds |xhat| with it)
        s = TBwdStats(torch.stack([gy.double().sum(0), (gy * xh).double().sum(0)]),
                      torch.stack([G.abs().max(), Z.abs().max()]))
        s.local = s.sums.clone()
        if isinstance(g, TOuter):
            s.dw = (g.g_logit[:, None] * torch.relu(pre)).double().sum(0)
            s.db = g.g_logit.double().sum().reshape(1)
        return s

    def bn_param_grads(self, s: TBwdStats):
        return s.local[1].clone(), s.local[0].clone()       # d gamma = sum g_y xhat, d beta = sum g_y

    def final_param_grads(self, s: TBwdStats):
        return s.dw.reshape(1, -1).clone(), s.db.clone()

    def _gz(self, g, z, st, s, count):
        G, Z = self._g(g), self._z(z)
        gy = G * ((Z * st.scale + st.shift) > 0)
        xh = (Z - st.mean) * st.invstd
        s1 = (s.sums[0] / count).to(self.dtype)
        s2 = (s.sums[1] / count).to(self.dtype)
        return st.scale * (gy - s1 - xh * s2)

    def bwd_apply(self, g, z, st: TBN, s: TBwdStats, count, want_T=False):
        gz = self._gz(g, z, st, s, float(count))
        xmax = float(st.invstd.abs().max()) * (float(s.maxes[1]) + float(st.mean.abs().max()))
        bound = float(st.scale.abs().max()) * (float(s.maxes[0]) + float(s.sums[0].abs().max()) / count
                                               + xmax * float(s.sums[1].abs().max()) / count)
        sc = pow2_scale(bound)
        return TAct(gz * sc, sc, want_T)

    def bwd_apply_pair(self, g, zp: TPair, st: TBN, s: TBwdStats, count):
        gz = self._gz(g, zp, st, s, float(count))
        B, L, H = zp.a.shape[0], zp.c.shape[0], zp.a.shape[1]
        gz = gz.reshape(B, L, H)
        return gz.sum(1), gz.sum(0)
