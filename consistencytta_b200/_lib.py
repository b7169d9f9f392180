"""ctypes binding of libctta.so (declared in include/ctta.h).  No CPU fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTTA_LIB") or os.path.join(HERE, "libctta.so")   # CTTA_LIB: A/B runs of two builds on one box

F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_TANH, ACT_LRELU = 0, 1, 2, 3, 4
A_ROWS, A_CONV1D, A_CONV2D = 0, 1, 2
MAX_TAPS = 16


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_mode", C.c_int32), ("ab_dtype", C.c_int32), ("c", C.c_int32), ("a_ld", C.c_int32),
        ("n_img", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("rows_per_img", C.c_int32),
        ("ntaps", C.c_int32), ("tap_d0", C.c_int16 * MAX_TAPS), ("tap_d1", C.c_int16 * MAX_TAPS),
        ("wgt", C.c_void_p), ("n", C.c_int32),
        ("bias", C.c_void_p), ("rowadd", C.c_void_p), ("rowadd_ld", C.c_int32), ("rowadd_rows", C.c_int32),
        ("act", C.c_int32), ("act_slope", C.c_float),
        ("residual", C.c_void_p), ("res_dtype", C.c_int32), ("res_ld", C.c_int32),
        ("accumulate", C.c_int32), ("out_scale", C.c_float),
        ("out", C.c_void_p), ("out_dtype", C.c_int32), ("out_ld", C.c_int32),
        ("out2", C.c_void_p), ("out2_ld", C.c_int32), ("act2", C.c_int32), ("act2_slope", C.c_float),
        ("out_rows_per_img", C.c_int32), ("out_stride", C.c_int32), ("out_off", C.c_int32),
        ("stats", C.c_void_p), ("stats_groups", C.c_int32), ("stats_rows_per_img", C.c_int32),
        ("res_neg_scale", C.c_float), ("wgt_img_stride", C.c_int64),
        ("out_up_phase", C.c_int32), ("stats_keep", C.c_int32),
    ]


# every exported symbol of include/ctta.h with its argument types (tests check that each one resolves)
_I32, _I64, _F, _P = C.c_int32, C.c_int64, C.c_float, C.c_void_p
SIGNATURES = {
    "ctta_last_error": (C.c_char_p, []),
    "ctta_version": (C.c_int, []),
    "ctta_launch_count": (C.c_longlong, []),
    "ctta_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "ctta_groupnorm_stats": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "ctta_groupnorm_apply": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P,
                                       _F, _I32, _I32, _P, _I32, _I32, _P, _I32, _P]),
    "ctta_layernorm": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _F, _P, _I32, _I32, _P]),
    "ctta_attention": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I64, _I64, _I64, _I64, _I64,
                                 _I64, _I64, _I64, _P, _F, _P]),
    "ctta_attention_debug": (C.c_int, [C.POINTER(C.c_longlong)]),
    "ctta_softmax_rows": (C.c_int, [_P, _I32, _I32, _I64, _F, _P, _I32, _I64, _P]),
    "ctta_im2col_s2": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P]),
    "ctta_nchw_to_nhwc": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _I32, _F, _P, _P]),
    "ctta_nhwc_to_nchw": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P]),
    "ctta_time_features": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P]),
    "ctta_small_linear": (C.c_int, [_P, _I32, _I32, _P, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "ctta_wave_minmax": (C.c_int, [_P, _I64, _P, _P]),
    "ctta_wave_to_int16": (C.c_int, [_P, _I64, _P, _P, _P]),
    "ctta_lrelu_cast": (C.c_int, [_P, _I64, _F, _P, _I32, _P]),
    "ctta_cfg_mix": (C.c_int, [_P, _I64, _F, _P, _P]),
    "ctta_tap_sum": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, C.POINTER(C.c_int16), C.POINTER(C.c_int16), _P, _I32, _P, _P,
                               _I32, _P]),
    "ctta_resblock_pair_supported": (C.c_int, [_I32, _I32, _I32, _I32]),
    "ctta_resblock_pair": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _I32, _I32, _F, _P]),
    "ctta_mrf_combine": (C.c_int, [C.POINTER(C.c_void_p), _I32, _I64, _I32, _F, _F, _F, _P, _P]),
}

_lib = None


class CttaError(RuntimeError):
    pass


def lib():
    """Loads libctta.so (building it first if nvcc is available and the library is missing/stale)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.environ.get("CTTA_REBUILD"):
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise CttaError("libctta.so not found at %s: run `python -m consistencytta_b200.build`" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise CttaError("libctta error %d: %s" % (rc, lib().ctta_last_error().decode()))
