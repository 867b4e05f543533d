"""The training primitives of protnote_b200/train.py bound to the sm_100a library (`pn_t_*`, include/protnote_b200.h).

As in native.py, PyTorch owns the device memory and names the stream; every value on the training path is computed by
the library's kernels.  There is no other implementation of these primitives in the product package.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import PN_STRICT, BwdSrc, check, ptr, stream_ptr
from .native import MODES, _require_cuda


def _up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Act:
    """[rows, cols] tensor as fp16 planes (+ transposed planes) with an optional power-of-two device scale."""
    __slots__ = ("hi", "lo", "hiT", "loT", "rows", "cols", "ld", "ldT", "sc")

    def __init__(self, rows, cols, device, strict, want_T):
        self.rows, self.cols = int(rows), int(cols)
        self.ld, self.ldT = _up(self.cols, 64), _up(self.rows, 64) // 64     # ldT = 64-row blocks of the transposed planes
        self.hi = torch.empty(self.rows, self.ld, dtype=torch.float16, device=device)
        self.lo = torch.empty(self.rows, self.ld, dtype=torch.float16, device=device) if strict else None
        self.hiT = self.loT = None
        if want_T:
            # K-blocked transposed planes [blocks][cols][64] (include/protnote_b200.h)
            self.hiT = torch.empty(self.ldT, self.cols, 64, dtype=torch.float16, device=device)
            self.loT = torch.empty(self.ldT, self.cols, 64, dtype=torch.float16, device=device) if strict else None
        self.sc = None

    @property
    def has_T(self):
        return self.hiT is not None


class Packed:
    __slots__ = ("hi", "lo", "N", "K", "ld", "ws")


class BwdStats:
    __slots__ = ("sums", "maxes", "local", "dw", "db", "gyl")


class Outer:
    def __init__(self, g_logit, w):
        self.g_logit, self.w = g_logit, w


class PairSrc:
    def __init__(self, a, c):
        self.a, self.c = a, c
        self.a_stats = None       # column sums of a (fp64 [2, H]), filled lazily for the analytic dc


class NativeOps:
    def __init__(self, precision: str = "strict"):
        self.lib = _lib.load()
        self.mode = MODES[precision]
        self.strict = self.mode == PN_STRICT
        # K elements summed in the tensor-core accumulator between fp32 promotions (see csrc/pn_gemm.cuh)
        # strict: 32 for the projection heads and the layer-1 halves (few rows, their error is amplified by every later
        # layer - same setting as the eval path), 256 for the pair GEMMs (millions of rows, error not dominant)
        self.promote_fwd = 256 if self.strict else 0
        self.promote_small = 32 if self.strict else 0
        self.promote_wgrad = 256 if self.strict else 2048
        # wgrad: rows per launch (measured on B200, fast mode, K = 2M rows: 4096 -> 42 ms, 8192 -> 35 ms, 16384 -> 31.5 ms, 32768 -> 30.4 ms per wgrad)
        self.split_wgrad = 16384 if self.strict else 32768
        import os
        if os.environ.get("PN_SPLIT_WGRAD"):          # experiment switch
            self.split_wgrad = int(os.environ["PN_SPLIT_WGRAD"])

    # ------------------------------------------------------------------ operands
    def _f32(self, x, name):
        _require_cuda(x, name)
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        return x.contiguous()

    def split(self, x, want_T=False, autoscale=False) -> Act:
        x = self._f32(x, "input")
        rows, cols = x.shape
        act = Act(rows, cols, x.device, self.strict, want_T)
        with torch.cuda.device(x.device):
            if autoscale:
                act.sc = torch.empty(2, dtype=torch.float32, device=x.device)
                check(self.lib.pn_t_autoscale(ptr(x), rows, cols, cols, ptr(act.sc), stream_ptr()))
            check(self.lib.pn_t_split(ptr(x), rows, cols, cols, ptr(act.sc), ptr(act.hi), ptr(act.lo), act.ld,
                                      ptr(act.hiT), ptr(act.loT), act.ldT, stream_ptr()))
        return act

    def pack(self, W, transposed=False) -> Packed:
        _require_cuda(W, "weight")
        W = W.detach()
        if W.dtype != torch.float32:
            W = W.float()
        if W.stride(1) != 1 and W.stride(0) != 1:
            W = W.contiguous()
        sn, sk = W.stride(0), W.stride(1)
        N, K = W.shape
        if transposed:
            N, K, sn, sk = K, N, sk, sn
        pk = Packed()
        pk.N, pk.K, pk.ld = N, K, _up(K, 64)
        pk.hi = torch.empty(N, pk.ld, dtype=torch.float16, device=W.device)
        pk.lo = torch.empty(N, pk.ld, dtype=torch.float16, device=W.device)   # packing always writes both planes
        pk.ws = torch.empty(2, dtype=torch.float32, device=W.device)
        with torch.cuda.device(W.device):
            check(self.lib.pn_t_pack_weight(ptr(W), N, K, sn, sk, ptr(pk.hi), ptr(pk.lo), pk.ld, ptr(pk.ws), stream_ptr()))
        return pk

    # ------------------------------------------------------------------ GEMMs
    def _gemm(self, a_hi, a_lo, M, K, lda, b_hi, b_lo, N, ldb, scales, out_f32=None, accumulate=False, out_act=None,
              promote=0, split_k=0, k_blocked=False):
        dev = a_hi.device
        scratch = torch.empty(N, dtype=torch.float32, device=dev)
        s = list(scales) + [None] * (3 - len(scales))
        with torch.cuda.device(dev):
            check(self.lib.pn_t_gemm(ptr(a_hi), ptr(a_lo), M, K, lda, ptr(b_hi), ptr(b_lo), N, ldb, ptr(s[0]), ptr(s[1]),
                                     ptr(s[2]), ptr(scratch), ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
                                     int(accumulate), ptr(out_act.hi) if out_act else None,
                                     ptr(out_act.lo) if out_act else None, out_act.ld if out_act else 0, self.mode,
                                     int(promote), int(split_k), int(k_blocked), stream_ptr()))

    def _promote(self, rows):
        return self.promote_small if rows <= (1 << 17) else self.promote_fwd

    def linear(self, x: Act, W: Packed, out_f32=False, accumulate_into=None):
        assert x.cols == W.K
        if out_f32:
            out = accumulate_into
            if out is None:
                out = torch.empty(x.rows, W.N, dtype=torch.float32, device=x.hi.device)
            assert out.dtype == torch.float32 and tuple(out.shape) == (x.rows, W.N) and out.stride(1) == 1
            self._gemm(x.hi, x.lo, x.rows, x.cols, x.ld, W.hi, W.lo, W.N, W.ld, [W.ws, x.sc], out_f32=out,
                       accumulate=accumulate_into is not None, promote=self._promote(x.rows))
            return out
        z = Act(x.rows, W.N, x.hi.device, self.strict, False)
        z.sc = x.sc
        self._gemm(x.hi, x.lo, x.rows, x.cols, x.ld, W.hi, W.lo, W.N, W.ld, [W.ws], out_act=z,
                   promote=self._promote(x.rows))
        return z

    def dgrad(self, g: Act, WT: Packed, out_f32=False, accumulate_into=None):
        # g_x = g_z W = g_z (W^T)^T; planes keep g's scale; accumulate_into: fp32 tensor the result is added to
        return self.linear(g, WT, out_f32=out_f32, accumulate_into=accumulate_into)

    def wgrad(self, g: Act, x: Act, out: Optional[torch.Tensor] = None):
        if not (g.has_T and x.has_T):
            raise _lib.ProtnoteB200Error("wgrad contracts over rows: both operands need their transposed planes")
        assert g.rows == x.rows
        if out is None:
            out = torch.empty(g.cols, x.cols, dtype=torch.float32, device=g.hi.device)
        assert out.stride(1) == 1 and tuple(out.shape) == (g.cols, x.cols)
        self._gemm(g.hiT, g.loT, g.cols, g.rows, 0, x.hiT, x.loT, x.cols, 0, [g.sc, x.sc], out_f32=out,
                   promote=self.promote_wgrad, split_k=self.split_wgrad, k_blocked=True)
        return out

    # ------------------------------------------------------------------ statistics / BatchNorm forward
    def col_stats(self, z: Act):
        out = torch.empty(2, z.cols, dtype=torch.float64, device=z.hi.device)
        with torch.cuda.device(out.device):
            check(self.lib.pn_t_col_stats(ptr(z.hi), ptr(z.lo), None, z.rows, z.cols, z.ld, ptr(out), stream_ptr()))
        return out

    def col_stats_f32(self, x):
        x = self._f32(x, "input")
        out = torch.empty(2, x.shape[1], dtype=torch.float64, device=x.device)
        with torch.cuda.device(out.device):
            check(self.lib.pn_t_col_stats(None, None, ptr(x), x.shape[0], x.shape[1], x.shape[1], ptr(out), stream_ptr()))
        return out

    def _finalize(self, stats, count, stats2, count2, bn, update_running):
        if bn.momentum is None:
            raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative average) is not implemented")
        cols = stats.shape[1]
        state = torch.empty(4, cols, dtype=torch.float32, device=stats.device)
        upd = update_running and bn.track_running_stats
        with torch.cuda.device(stats.device):
            check(self.lib.pn_t_bn_finalize(ptr(stats), float(count), ptr(stats2), float(count2), ptr(bn.weight.detach()),
                                            ptr(bn.bias.detach()), float(bn.eps), float(bn.momentum),
                                            ptr(bn.running_mean) if upd else None, ptr(bn.running_var) if upd else None,
                                            cols, ptr(state), stream_ptr()))
        return state

    def bn_finalize(self, stats, count, bn, update_running=True):
        return self._finalize(stats, count, None, 0.0, bn, update_running)

    def bn_finalize_pair(self, sa, B, sc, L, bn, update_running=True):
        return self._finalize(sa, B, sc, L, bn, update_running)

    def affine_state(self, bias, cols, device):
        """Layer state [4, cols] (scale, shift, mean, invstd - csrc/pn_train.cuh) of a hidden layer WITHOUT BatchNorm:
        relu(z * 1 + bias).  mean 0 / invstd 1 make xhat = z in the backward kernels, where it only ever multiplies a
        cleared sum (train._affine_grads).  Assembled from constants and the bias vector, no arithmetic."""
        state = torch.zeros(4, int(cols), dtype=torch.float32, device=device)
        state[0].fill_(1.0)
        state[3].fill_(1.0)
        if bias is not None:
            state[1].copy_(self._f32(bias, "bias").reshape(-1))
        return state

    # ------------------------------------------------------------------ dropout inside the MLPs (OUTPUT_MLP_DROPOUT)
    def dropout(self, x: Act, drop, want_T=False) -> Act:
        """x * keep / (1 - p) as new planes (+ transposed planes); drop = (seed, p).  The mask depends on (seed, row, column)
        only, so the backward calls this on the incoming gradient with the forward's seed.  The tensor's scale is kept."""
        seed, p = drop
        out = Act(x.rows, x.cols, x.hi.device, x.lo is not None, want_T)
        out.sc = x.sc
        with torch.cuda.device(x.hi.device):
            check(self.lib.pn_t_dropout_planes(ptr(x.hi), ptr(x.lo), x.rows, x.cols, x.ld, C.c_ulonglong(seed), C.c_float(p),
                                               ptr(out.hi), ptr(out.lo), out.ld, ptr(out.hiT), ptr(out.loT), out.ldT,
                                               stream_ptr()))
        return out

    def dropout_f32(self, x, drop):
        seed, p = drop
        x = self._f32(x, "input")
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(self.lib.pn_t_dropout_f32(ptr(x), x.shape[0], x.shape[1], x.stride(0), C.c_ulonglong(seed), C.c_float(p),
                                            ptr(out), out.stride(0), stream_ptr()))
        return out

    def bn_relu(self, z: Act, st, want_T=False) -> Act:
        h = Act(z.rows, z.cols, z.hi.device, self.strict, want_T)
        with torch.cuda.device(z.hi.device):
            check(self.lib.pn_t_bn_relu(ptr(z.hi), ptr(z.lo), z.rows, z.cols, z.ld, ptr(st), ptr(h.hi), ptr(h.lo), h.ld,
                                        ptr(h.hiT), ptr(h.loT), h.ldT, stream_ptr()))
        return h

    def bn_relu_dot(self, z: Act, st, w, b):
        out = torch.empty(z.rows, dtype=torch.float32, device=z.hi.device)
        w = self._f32(w, "output weight").reshape(-1)
        b = self._f32(b, "output bias").reshape(-1)
        with torch.cuda.device(out.device):
            check(self.lib.pn_t_bn_relu_dot(ptr(z.hi), ptr(z.lo), z.rows, z.cols, z.ld, ptr(st), ptr(w), ptr(b), ptr(out),
                                            stream_ptr()))
        return out

    def bn_relu_dot_loss(self, z: Act, st, w, b, targets, L, spec):
        """bn_relu_dot with the loss of `spec` (train.LossSpec) fused in: returns (logits [rows], g_seed [rows] = grad_scale *
        d loss / d logit, loss_sum fp64 [1] = sum of the per-pair losses of this rank's rows)."""
        dev = z.hi.device
        out = torch.empty(z.rows, dtype=torch.float32, device=dev)
        g = torch.empty(z.rows, dtype=torch.float32, device=dev)
        loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        w = self._f32(w, "output weight").reshape(-1)
        b = self._f32(b, "output bias").reshape(-1)
        t = self._f32(targets, "targets").reshape(-1)
        if t.numel() != z.rows:
            raise ValueError(f"targets hold {t.numel()} entries, the batch has {z.rows} (protein, label) pairs")
        pw = None
        if spec.pos_weight is not None:
            pw = self._f32(spec.pos_weight, "pos_weight").reshape(-1)
            if pw.numel() == 1:
                pw = pw.expand(L).contiguous()
            if pw.numel() != L:
                raise ValueError(f"pos_weight has {pw.numel()} entries for {L} label rows")
        with torch.cuda.device(dev):
            check(self.lib.pn_t_bn_relu_dot_loss(ptr(z.hi), ptr(z.lo), z.rows, z.cols, z.ld, ptr(st), ptr(w), ptr(b), ptr(out),
                                                 ptr(t), int(L), ptr(pw), spec.kind_id, C.c_float(spec.gamma),
                                                 C.c_float(spec.alpha), C.c_float(spec.label_smoothing),
                                                 C.c_float(spec.grad_scale), ptr(g), ptr(loss_sum), stream_ptr()))
        return out, g, loss_sum

    def pair_hidden(self, a, c, st, want_T=False) -> Act:
        B, H = a.shape
        L = c.shape[0]
        h = Act(B * L, H, a.device, self.strict, want_T)
        with torch.cuda.device(a.device):
            check(self.lib.pn_t_pair_hidden(ptr(a), B, ptr(c), L, H, ptr(st), ptr(h.hi), ptr(h.lo), h.ld, ptr(h.hiT),
                                            ptr(h.loT), h.ldT, stream_ptr()))
        return h

    # ------------------------------------------------------------------ FEATURE_FUSION concatenation_prod
    def pair_product(self, P_e, L_e, want_T=False) -> Act:
        """q[b*L + l] = P_e[b] * L_e[l] (elementwise) as planes [B*L, d]"""
        P_e, L_e = self._f32(P_e, "sequence embeddings"), self._f32(L_e, "label embeddings")
        (B, d), L = P_e.shape, L_e.shape[0]
        q = Act(B * L, d, P_e.device, self.strict, want_T)
        with torch.cuda.device(P_e.device):
            check(self.lib.pn_t_pair_product(ptr(P_e), B, ptr(L_e), L, d, ptr(q.hi), ptr(q.lo), q.ld, ptr(q.hiT), ptr(q.loT),
                                             q.ldT, stream_ptr()))
        return q

    def pair_add(self, x: Act, a, c) -> Act:
        """z[b*L + l] = x[b*L + l] + a[b] + c[l] as planes"""
        assert x.sc is None and x.rows == a.shape[0] * c.shape[0] and x.cols == a.shape[1] == c.shape[1]
        a, c = self._f32(a, "protein term"), self._f32(c, "label term")
        z = Act(x.rows, x.cols, x.hi.device, x.lo is not None, False)
        with torch.cuda.device(x.hi.device):
            check(self.lib.pn_t_pair_add(ptr(x.hi), ptr(x.lo), x.ld, ptr(a), a.shape[0], ptr(c), c.shape[0], x.cols, ptr(z.hi),
                                         ptr(z.lo), z.ld, stream_ptr()))
        return z

    def pair_marginals(self, g: Act, B, L, wb=None, wl=None):
        """(sum_l g[b,l] * wl[l], sum_b g[b,l] * wb[b]) of a pair-grid tensor g [B*L, cols] (true scale); unit weights if None"""
        assert g.rows == B * L
        dev = g.hi.device
        wb = None if wb is None else self._f32(wb, "protein weights")
        wl = None if wl is None else self._f32(wl, "label weights")
        out_b = torch.empty(B, g.cols, dtype=torch.float32, device=dev)
        out_l = torch.empty(L, g.cols, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.pn_t_pair_marginals(ptr(g.hi), ptr(g.lo), g.ld, ptr(g.sc), B, L, g.cols, ptr(wb), ptr(wl),
                                               ptr(out_b), ptr(out_l), stream_ptr()))
        return out_b, out_l

    # ------------------------------------------------------------------ FEATURE_FUSION similarity
    def normalize_rows(self, x, scale=1.0):
        """(x * scale / max(|x|, 1e-12) per row, 1 / max(|x|, 1e-12) [rows])"""
        x = self._f32(x, "embeddings")
        y = torch.empty_like(x)
        inv = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(self.lib.pn_t_normalize_rows(ptr(x), x.shape[0], x.shape[1], C.c_float(scale), ptr(y), ptr(inv),
                                               stream_ptr()))
        return y, inv

    def normalize_rows_bwd(self, y, inv_norm, dy, scale=1.0):
        y, dy = self._f32(y, "normalised embeddings"), self._f32(dy, "gradient")
        dx = torch.empty_like(y)
        with torch.cuda.device(y.device):
            check(self.lib.pn_t_normalize_rows_bwd(ptr(y), ptr(inv_norm), ptr(dy), y.shape[0], y.shape[1], C.c_float(scale),
                                                   ptr(dx), stream_ptr()))
        return dx

    # ------------------------------------------------------------------ BatchNorm + ReLU backward
    def outer(self, g_logit, w):
        return Outer(self._f32(g_logit, "logit gradient").reshape(-1), self._f32(w, "output weight").reshape(-1))

    def pair_source(self, a, c):
        return PairSrc(a, c)

    def _src(self, g, z, st):
        s = BwdSrc()
        keep = [st]
        s.state = st.data_ptr()
        if isinstance(g, Outer):
            s.kind = 1
            s.rows, s.cols = g.g_logit.numel(), g.w.numel()
            s.g_logit, s.w = g.g_logit.data_ptr(), g.w.data_ptr()
            keep += [g.g_logit, g.w]
        else:
            s.kind = 2 if isinstance(z, PairSrc) else 0
            s.rows, s.cols = g.rows, g.cols
            s.g_hi = g.hi.data_ptr()
            s.g_lo = g.lo.data_ptr() if g.lo is not None else None
            s.ld_g = g.ld
            s.g_sc = g.sc.data_ptr() if g.sc is not None else None
        if isinstance(z, PairSrc):
            s.a, s.c, s.L = z.a.data_ptr(), z.c.data_ptr(), z.c.shape[0]
        else:
            assert z.rows == s.rows and z.cols == s.cols
            s.z_hi = z.hi.data_ptr()
            s.z_lo = z.lo.data_ptr() if z.lo is not None else None
            s.ld_z = z.ld
        return s, keep

    def bwd_stats(self, g, z, st) -> BwdStats:
        src, _keep = self._src(g, z, st)
        dev = st.device
        out = BwdStats()
        out.sums = torch.empty(2, src.cols, dtype=torch.float64, device=dev)
        out.maxes = torch.empty(2, dtype=torch.float32, device=dev)
        out.dw = out.db = out.gyl = None
        if src.kind == 2:          # by-product of the pass: per-label masked gradient sums (see bwd_apply_pair)
            out.gyl = torch.empty(z.c.shape[0], src.cols, dtype=torch.float32, device=dev)
        if src.kind == 1:
            out.dw = torch.empty(src.cols, dtype=torch.float64, device=dev)
            out.db = torch.empty(1, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.pn_t_bwd_stats(C.byref(src), ptr(out.sums), ptr(out.maxes), ptr(out.dw), ptr(out.db),
                                          ptr(out.gyl), stream_ptr()))
        out.local = out.sums.clone()
        return out

    def bn_param_grads(self, s: BwdStats):
        return s.local[1].float(), s.local[0].float()       # d gamma, d beta (this rank's rows)

    def final_param_grads(self, s: BwdStats):
        return s.dw.float().reshape(1, -1), s.db.float()

    def bwd_apply(self, g, z, st, s: BwdStats, count, want_T=False) -> Act:
        src, _keep = self._src(g, z, st)
        dev = st.device
        out = Act(src.rows, src.cols, dev, self.strict, want_T)
        out.sc = torch.empty(2, dtype=torch.float32, device=dev)
        means = torch.empty(2, src.cols, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.pn_t_bwd_scale(ptr(s.sums), ptr(s.maxes), ptr(st), float(count), src.cols, ptr(out.sc),
                                          ptr(means), stream_ptr()))
            check(self.lib.pn_t_bwd_apply(C.byref(src), ptr(means), ptr(out.sc), ptr(out.hi), ptr(out.lo), out.ld,
                                          ptr(out.hiT), ptr(out.loT), out.ldT, stream_ptr()))
        return out

    def bwd_apply_pair(self, g, zp: PairSrc, st, s: BwdStats, count):
        src, _keep = self._src(g, zp, st)
        dev = st.device
        B, H = zp.a.shape
        L = zp.c.shape[0]
        da64 = torch.empty(B, H, dtype=torch.float64, device=dev)
        da = torch.empty(B, H, dtype=torch.float32, device=dev)
        dc = torch.empty(L, H, dtype=torch.float32, device=dev)
        means = torch.empty(2, src.cols, dtype=torch.float32, device=dev)
        gyl = getattr(s, "gyl", None)
        if gyl is not None and zp.a_stats is None:
            zp.a_stats = self.col_stats_f32(zp.a)
        with torch.cuda.device(dev):
            check(self.lib.pn_t_bwd_scale(ptr(s.sums), ptr(s.maxes), ptr(st), float(count), src.cols, None, ptr(means),
                                          stream_ptr()))
            check(self.lib.pn_t_bwd_apply_pair(C.byref(src), ptr(means), B, ptr(gyl),
                                               ptr(zp.a_stats) if gyl is not None else None, ptr(da64), ptr(da), ptr(dc),
                                               stream_ptr()))
        return da, dc
