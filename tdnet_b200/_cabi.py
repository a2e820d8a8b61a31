"""ctypes binding of libtdnet_b200.so (the C-ABI declared in include/tdnet_b200.h).

The library is built in-tree by __graft_entry__.build() (nvcc, sm_100a).  There is no fallback: if
the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtdnet_b200.so")

TDN_F32, TDN_SPLIT16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_lo", C.c_void_p), ("dtype", C.c_int32),
                ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("stride_n", C.c_int64), ("stride_h", C.c_int64), ("stride_w", C.c_int64)]


class Conv2dDesc(C.Structure):
    _fields_ = [("in_", Tensor), ("out", Tensor), ("residual", Tensor),
                ("weight", C.c_void_p), ("scale", C.c_void_p), ("bias", C.c_void_p),
                ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("stride", C.c_int32), ("pad", C.c_int32), ("dilation", C.c_int32),
                ("act", C.c_int32), ("leaky_slope", C.c_float), ("weight_kn", C.c_int32),
                ("batch", C.c_int32),
                ("in_batch_stride", C.c_int64), ("out_batch_stride", C.c_int64),
                ("residual_batch_stride", C.c_int64), ("weight_batch_stride", C.c_int64)]


class TcConvDesc(C.Structure):
    _fields_ = [("in_", Tensor), ("out", Tensor), ("residual", Tensor),
                ("out_f32_copy", C.c_void_p), ("weight_hi", C.c_void_p), ("weight_lo", C.c_void_p),
                ("weight_ld", C.c_int64), ("weight_batch_stride", C.c_int64),
                ("weight_batched", C.c_int32), ("bias_along_m", C.c_int32),
                ("scale", C.c_void_p), ("bias", C.c_void_p),
                ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("dilation", C.c_int32),
                ("act", C.c_int32), ("leaky_slope", C.c_float), ("range_flag", C.c_void_p),
                ("stride", C.c_int32), ("variant", C.c_int32), ("flags", C.c_int32)]


TC_AUTO, TC_BASE, TC_HALO, TC_PAIR, TC_PAIR_TAIL, TC_PAIR_QUAD, TC_HALO_SW, TC_BASE_TS, TC_PAIR_BAND = range(9)   # tdn_tc_conv_desc.variant


class AttentionDesc(C.Structure):
    _fields_ = [("q_hi", C.c_void_p), ("q_lo", C.c_void_p), ("q_ld", C.c_int64), ("q_batch_stride", C.c_int64),
                ("k_hi", C.c_void_p), ("k_lo", C.c_void_p), ("k_ld", C.c_int64), ("k_batch_stride", C.c_int64),
                ("vt_hi", C.c_void_p), ("vt_lo", C.c_void_p), ("vt_ld", C.c_int64), ("vt_batch_stride", C.c_int64),
                ("out", Tensor), ("residual", Tensor),
                ("n", C.c_int32), ("pq", C.c_int32), ("pk", C.c_int32), ("d_k", C.c_int32), ("d_v", C.c_int32),
                ("range_flag", C.c_void_p), ("flags", C.c_int32)]


class PspProjection(C.Structure):
    """tdn_psp_projection (include/tdnet_b200.h)."""
    _fields_ = [("w", C.c_void_p), ("dst_hi", C.c_void_p), ("dst_lo", C.c_void_p), ("ld", C.c_int64),
                ("batch_stride", C.c_int64), ("cout", C.c_int32), ("reserved", C.c_int32)]


PSP_MAX_PROJECTIONS = 4
TC_FLAG_FAST = 1   # tdn_tc_conv_desc.flags / tdn_attention_desc.flags: one fp16 product per K step (opt-in, not fp32-faithful)

# symbol -> (restype, argtypes); tests/test_cabi.py checks this list against include/tdnet_b200.h
_TP = C.POINTER(Tensor)
SIGNATURES = {
    "tdn_abi_version": (C.c_int, []),
    "tdn_strerror": (C.c_char_p, [C.c_int]),
    "tdn_last_error": (C.c_char_p, []),
    "tdn_device_arch": (C.c_int, []),
    "tdn_conv2d": (C.c_int, [C.POINTER(Conv2dDesc), C.c_void_p]),
    "tdn_conv2d_tc": (C.c_int, [C.POINTER(TcConvDesc), C.c_void_p]),
    "tdn_attention_tc": (C.c_int, [C.POINTER(AttentionDesc), C.c_void_p]),
    "tdn_attention_tc_launches": (C.c_int, [C.POINTER(AttentionDesc), C.POINTER(C.c_int32)]),
    "tdn_split16": (C.c_int, [_TP, _TP, C.c_void_p]),
    "tdn_merge16": (C.c_int, [_TP, _TP, C.c_void_p]),
    "tdn_image_to_nhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _TP, C.c_void_p]),
    "tdn_stem_conv_pool": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     _TP, C.c_void_p]),
    "tdn_stem_conv_pool_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, _TP, C.c_void_p]),
    "tdn_stem_conv_pool_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, _TP, C.c_void_p, C.c_void_p]),
    "tdn_stem_conv_pool_tc_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_void_p, _TP, C.c_int32, C.c_float, C.c_void_p,
                                            C.c_void_p]),
    "tdn_maxpool3x3s2": (C.c_int, [_TP, _TP, C.c_void_p]),
    "tdn_psp_pool": (C.c_int, [_TP, _TP, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tdn_psp_pool_workspace_bytes": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32]),
    "tdn_bilinear_nhwc": (C.c_int, [_TP, _TP, C.c_void_p]),
    "tdn_psp_branch_convs": (C.c_int, [_TP, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]),
    "tdn_psp_branch_project": (C.c_int, [_TP, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                         C.c_int32, C.POINTER(C.c_void_p), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "tdn_psp_concat": (C.c_int, [_TP, C.POINTER(C.c_void_p), C.c_int32, _TP, C.c_void_p]),
    "tdn_copy_nhwc": (C.c_int, [_TP, _TP, C.c_void_p]),
    "tdn_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_float, C.c_void_p]),
    "tdn_softmax_rows_split16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_float, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "tdn_layernorm_hw_stats": (C.c_int, [_TP, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tdn_layernorm_hw_workspace_bytes": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "tdn_layernorm_hw_apply": (C.c_int, [_TP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _TP, C.c_void_p]),
    "tdn_upsample_argmax": (C.c_int, [_TP, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "tdn_upsample_logits": (C.c_int, [_TP, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "tdn_upsample_argmax_sampled": (C.c_int, [_TP, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                              C.c_int32, C.c_void_p]),
    "tdn_resize_linear_u8": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "tdn_pointwise_linear": (C.c_int, [_TP, C.c_void_p, C.c_void_p, C.c_void_p, _TP, C.c_void_p]),
    "tdn_fa_context": (C.c_int, [_TP, _TP, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "tdn_fa_context_workspace_bytes": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "tdn_fa_apply": (C.c_int, [_TP, C.c_void_p, _TP, C.c_float, C.c_void_p, C.c_void_p]),
    "tdn_add_upsampled": (C.c_int, [_TP, _TP, _TP, _TP, C.c_void_p]),
    "tdn_sm_clock_probe": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
}

_lib = None


def load():
    """dlopen the in-tree library and declare every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"tdnet_b200: CUDA library not built ({LIB_PATH} missing). Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` in the repo root. "
            "There is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype, fn.argtypes = res, args
    if lib.tdn_abi_version() != 2:
        raise RuntimeError("tdnet_b200: ABI version mismatch between _cabi.py and the library")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        lib = load()
        raise RuntimeError(f"tdnet_b200 {what}: {lib.tdn_strerror(rc).decode()} ({rc}): "
                           f"{lib.tdn_last_error().decode()}")
