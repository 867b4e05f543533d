"""ctypes binding of include/protnote_b200.h.  There is no fallback: if the library is missing the import of any
compute entry point raises (the product path must fail loudly, never drop to PyTorch or CPU arithmetic)."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libprotnote_b200.so")

PN_STRICT = 3
PN_FAST = 1
FUSIONS = {"concatenation": 0, "concatenation_diff": 1, "concatenation_prod": 2, "similarity": 3}


class EncoderCfg(C.Structure):
    _fields_ = [("input_channels", C.c_int), ("channels", C.c_int), ("bottleneck", C.c_int),
                ("kernel_size", C.c_int), ("dilation_base", C.c_int), ("num_blocks", C.c_int),
                ("bn_eps", C.c_float)]


class ScorerCfg(C.Structure):
    _fields_ = [("protein_dim", C.c_int), ("label_dim", C.c_int), ("latent_dim", C.c_int),
                ("proj_hidden", C.c_int), ("proj_layers", C.c_int), ("out_hidden", C.c_int),
                ("out_layers", C.c_int), ("out_batchnorm", C.c_int), ("fusion", C.c_int),
                ("descriptions_per_label", C.c_int), ("bn_eps", C.c_float)]


class BwdSrc(C.Structure):
    """pn_bwd_src of include/protnote_b200.h"""
    _fields_ = [("kind", C.c_int), ("rows", C.c_longlong), ("cols", C.c_int),
                ("g_hi", C.c_void_p), ("g_lo", C.c_void_p), ("ld_g", C.c_longlong), ("g_sc", C.c_void_p),
                ("g_logit", C.c_void_p), ("w", C.c_void_p),
                ("z_hi", C.c_void_p), ("z_lo", C.c_void_p), ("ld_z", C.c_longlong),
                ("a", C.c_void_p), ("c", C.c_void_p), ("L", C.c_longlong),
                ("state", C.c_void_p)]


_P = C.c_void_p
_LL = C.c_longlong
_SZ = C.c_size_t
_I = C.c_int

# name -> (restype, argtypes); every symbol include/protnote_b200.h declares
# host callback of pn_encoder_forward_train_sharded: (device pointer, count, user, stream) -> 0 on success
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)

SIGNATURES = {
    "pn_version": (_I, []),
    "pn_last_error": (C.c_char_p, []),
    "pn_device_check": (_I, [_I]),
    "pn_set_option": (_I, [C.c_char_p, _LL]),
    "pn_launch_count": (_LL, []),
    "pn_gemm_timing": (_I, [_I]),
    "pn_gemm_timing_read": (_I, [C.POINTER(C.c_double), C.POINTER(_LL), C.POINTER(C.c_double)]),
    "pn_encoder_packed_bytes": (_SZ, [C.POINTER(EncoderCfg)]),
    "pn_encoder_pack": (_I, [C.POINTER(EncoderCfg), C.POINTER(_P), _I, _P, _SZ, _P]),
    "pn_encoder_workspace_bytes": (_SZ, [C.POINTER(EncoderCfg), _I, _I]),
    "pn_encoder_forward": (_I, [C.POINTER(EncoderCfg), _P, _P, _P, _I, _I, _P, _P, _SZ, _I, _P]),
    "pn_encoder_pack_raw": (_I, [C.POINTER(EncoderCfg), C.POINTER(_P), _I, _P, _SZ, _P]),
    "pn_encoder_train_workspace_bytes": (_SZ, [C.POINTER(EncoderCfg), _I, _I]),
    "pn_encoder_forward_train": (_I, [C.POINTER(EncoderCfg), _P, _P, _P, _I, _I, C.POINTER(_P), _I, C.c_float, _I, _P, _P,
                                      _SZ, _I, _P]),
    "pn_encoder_forward_train_sharded": (_I, [C.POINTER(EncoderCfg), _P, _P, _P, _I, _I, C.POINTER(_P), _I, C.c_float, _I, _P,
                                              _P, _SZ, _I, C.c_double, _P, REDUCE_FN, _P, _P]),
    "pn_encoder_forward_tokens": (_I, [C.POINTER(EncoderCfg), _P, _P, _P, _I, _I, _P, _P, _SZ, _I, _P]),
    "pn_postprocess": (_I, [_P, _LL, _LL, _LL, _P, _I, _LL, C.c_float, _P, _LL, _P, _P, _P, _I, _P, _P, _P]),
    "pn_scorer_num_params": (_I, [C.POINTER(ScorerCfg)]),
    "pn_scorer_packed_bytes": (_SZ, [C.POINTER(ScorerCfg)]),
    "pn_scorer_pack": (_I, [C.POINTER(ScorerCfg), C.POINTER(_P), _I, _P, _SZ, _P]),
    "pn_project_workspace_bytes": (_SZ, [C.POINTER(ScorerCfg), _LL]),
    "pn_project_sequences": (_I, [C.POINTER(ScorerCfg), _P, _P, _LL, _P, _P, _P, _SZ, _I, _P]),
    "pn_project_labels": (_I, [C.POINTER(ScorerCfg), _P, _P, _LL, _P, _P, _P, _SZ, _I, _P]),
    "pn_scorer_min_workspace_bytes": (_SZ, [C.POINTER(ScorerCfg)]),
    "pn_scorer_workspace_bytes": (_SZ, [C.POINTER(ScorerCfg), _LL, _LL]),
    "pn_score_pairs": (_I, [C.POINTER(ScorerCfg), _P, _P, _P, _P, _P, _LL, _LL, _P, _LL, _P, _SZ, _I, _P]),
    "pn_score_pairs_ex": (_I, [C.POINTER(ScorerCfg), _P, _P, _P, _P, _P, _LL, _LL, _P, _LL, _P, _P, _SZ, _I, _P]),
    "pn_similarity_workspace_bytes": (_SZ, [C.POINTER(ScorerCfg), _LL, _LL]),
    "pn_score_similarity": (_I, [C.POINTER(ScorerCfg), _P, _P, _LL, _LL, C.c_float, _P, _LL, _P, _SZ, _I, _P]),
    "pn_linear_workspace_bytes": (_SZ, [_LL, _LL, _LL]),
    "pn_linear": (_I, [_P, _LL, _LL, _LL, _P, _LL, _P, _P, _LL, _P, _SZ, _I, _P]),
    "pn_conv1d_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "pn_conv1d": (_I, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _SZ, _I, _P]),
    # training primitives
    "pn_t_autoscale": (_I, [_P, _LL, _LL, _LL, _P, _P]),
    "pn_t_split": (_I, [_P, _LL, _LL, _LL, _P, _P, _P, _LL, _P, _P, _LL, _P]),
    "pn_t_pack_weight": (_I, [_P, _LL, _LL, _LL, _LL, _P, _P, _LL, _P, _P]),
    "pn_t_gemm": (_I, [_P, _P, _LL, _LL, _LL, _P, _P, _LL, _LL, _P, _P, _P, _P, _P, _LL, _I, _P, _P, _LL, _I, _I, _LL, _I, _P]),
    "pn_t_col_stats": (_I, [_P, _P, _P, _LL, _I, _LL, _P, _P]),
    "pn_t_bn_finalize": (_I, [_P, C.c_double, _P, C.c_double, _P, _P, C.c_float, C.c_float, _P, _P, _I, _P, _P]),
    "pn_t_bn_relu": (_I, [_P, _P, _LL, _I, _LL, _P, _P, _P, _LL, _P, _P, _LL, _P]),
    "pn_t_bn_relu_dot": (_I, [_P, _P, _LL, _I, _LL, _P, _P, _P, _P, _P]),
    "pn_t_bn_relu_dot_loss": (_I, [_P, _P, _LL, _I, _LL, _P, _P, _P, _P, _P, _LL, _P, _I, C.c_float, C.c_float, C.c_float,
                                   C.c_float, _P, _P, _P]),
    "pn_t_dropout_planes": (_I, [_P, _P, _LL, _I, _LL, C.c_ulonglong, C.c_float, _P, _P, _LL, _P, _P, _LL, _P]),
    "pn_t_dropout_f32": (_I, [_P, _LL, _I, _LL, C.c_ulonglong, C.c_float, _P, _LL, _P]),
    "pn_t_pair_hidden": (_I, [_P, _LL, _P, _LL, _I, _P, _P, _P, _LL, _P, _P, _LL, _P]),
    "pn_t_pair_product": (_I, [_P, _LL, _P, _LL, _I, _P, _P, _LL, _P, _P, _LL, _P]),
    "pn_t_pair_add": (_I, [_P, _P, _LL, _P, _LL, _P, _LL, _I, _P, _P, _LL, _P]),
    "pn_t_pair_marginals": (_I, [_P, _P, _LL, _P, _LL, _LL, _I, _P, _P, _P, _P, _P]),
    "pn_t_normalize_rows": (_I, [_P, _LL, _I, C.c_float, _P, _P, _P]),
    "pn_t_normalize_rows_bwd": (_I, [_P, _P, _P, _LL, _I, C.c_float, _P, _P]),
    "pn_t_bwd_stats": (_I, [C.POINTER(BwdSrc), _P, _P, _P, _P, _P, _P]),
    "pn_t_bwd_scale": (_I, [_P, _P, _P, C.c_double, _I, _P, _P, _P]),
    "pn_t_bwd_apply": (_I, [C.POINTER(BwdSrc), _P, _P, _P, _P, _LL, _P, _P, _LL, _P]),
    "pn_t_bwd_apply_pair": (_I, [C.POINTER(BwdSrc), _P, _LL, _P, _P, _P, _P, _P, _P]),
}

_lib = None


class ProtnoteB200Error(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it has not been built: there is no other compute path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ProtnoteB200Error(
            f"{LIB_PATH} is missing: build it with `python -m protnote_b200.build` (nvcc, sm_100a). "
            "protnote_b200 has no PyTorch/CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise ProtnoteB200Error(load().pn_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pointer_array(tensors):
    arr = (_P * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
