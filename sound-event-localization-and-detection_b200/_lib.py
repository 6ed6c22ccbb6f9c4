"""ctypes binding of libseldq.so (C ABI: include/seldq.h).  No torch types cross this boundary:
only raw device pointers, sizes and a cudaStream_t.

The library is built in-tree by csrc/build.sh (nvcc, sm_100a).  There is no CPU fallback: if
the shared object is missing this module raises, and compute calls on a host without a CUDA
device return SELDQ_ERR_CUDA which `check` turns into RuntimeError.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libseldq.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

ABI_VERSION = 5                                   # SELDQ_ABI_VERSION of include/seldq.h this binding was written against
ALG_REAL, ALG_Q, ALG_DQ, ALG_DQ_LINEAR = 0, 1, 2, 3
ALG_Q_LINEAR_IO, ALG_DQ_LINEAR_IO = 4, 5          # linear layers as 1x1 convolutions on their own (in, out) tensors
PREC_FP32, PREC_BF16 = 0, 1
PASS_FWD, PASS_DGRAD, PASS_WGRAD = 0, 1, 2
ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = -1, -2, -3, -4


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "algebra", "precision", "ndim", "batch", "cin", "cout", "in_h", "in_w", "k_h", "k_w",
        "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w")]


class CnnTailDesc(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("c", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("pool", ctypes.c_int32), ("drop_p", ctypes.c_float), ("salt", ctypes.c_uint32)]


class BnRef(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("sums", "gamma", "beta", "running_mean", "running_var")]


class TcnGlue(ctypes.Structure):
    """seldq_tcn_glue_t (include/seldq.h)."""
    _fields_ = [("n", ctypes.c_int32), ("c", ctypes.c_int32), ("t", ctypes.c_int32), ("c2", ctypes.c_int32),
                ("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("drop_p", ctypes.c_float),
                ("salt", ctypes.c_uint32), ("count", ctypes.c_double), ("bn", BnRef * 2),
                ("inp", ctypes.c_void_p * 5), ("out32", ctypes.c_void_p), ("out_cl", ctypes.c_void_p * 2),
                ("out_t16", ctypes.c_void_p * 2), ("dsums", ctypes.c_void_p), ("stats_out", ctypes.c_void_p * 2),
                ("accum", ctypes.c_void_p), ("seed", ctypes.c_void_p), ("flag", ctypes.c_int32),
                ("sync", ctypes.c_void_p), ("out32b", ctypes.c_void_p)]


TCN_PREACT_FWD, TCN_ROW_STATS, TCN_GATE_FWD, TCN_RESIDUAL_FWD = 0, 1, 2, 3
TCN_GATE_BWD_REDUCE, TCN_GATE_BWD_APPLY, TCN_PREACT_BWD_REDUCE, TCN_PREACT_BWD_APPLY = 4, 5, 6, 7
TCN_GATE_BWD, TCN_PREACT_BWD, TCN_GATE_FWD_STATS, TCN_RESIDUAL_PREACT_FWD = 8, 9, 10, 11


class ConvEpilogue(ctypes.Structure):
    """seldq_conv_epilogue_t (include/seldq.h)."""
    _fields_ = [("mode", ctypes.c_int32), ("addend", ctypes.c_void_p), ("stats", ctypes.c_void_p)]


EPI_STORE, EPI_ACCUMULATE, EPI_ADD = 0, 1, 2
ACT_RELU, ACT_TANH = 0, 1
QOP_HAMILTON, QOP_HAMILTON_CONJ_B, QOP_HAMILTON_CONJ_A, QOP_NORMALIZE, QOP_NORMALIZE_BWD, QOP_EXP, QOP_EXP_BWD = range(7)


class AttentionDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("batch", "heads", "seq", "head_dim")]


class StftOptions(ctypes.Structure):
    """seldq_stft_options_t (include/seldq.h)."""
    _fields_ = [("input_int16", ctypes.c_int32), ("mean", ctypes.c_float * 2), ("inv_std", ctypes.c_float * 2),
                ("stats", ctypes.c_void_p)]


class LinearDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("algebra", "precision", "rows", "in_features", "out_features")]


def build(force=False):
    """Compile libseldq.so in-tree (cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))
            if f.endswith((".cu", ".cuh", ".h", ".sh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "seldq.h"))
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["bash", BUILD_SCRIPT, LIB_PATH])
    return LIB_PATH


_P = ctypes.c_void_p
_PROTOS = {
    "seldq_abi_version": (ctypes.c_int, []),
    "seldq_last_error": (ctypes.c_char_p, []),
    "seldq_device_count": (ctypes.c_int, []),
    "seldq_conv_out_shape": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ctypes.c_int32),
                                            ctypes.POINTER(ctypes.c_int32)]),
    "seldq_conv_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(ConvDesc), ctypes.c_int32]),
    "seldq_conv_operand_info": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32,
                                               ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                               ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "seldq_stage_operand": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32, _P, _P, _P, _P]),
    "seldq_conv_packed_bytes": (ctypes.c_size_t, [ctypes.POINTER(ConvDesc), ctypes.c_int32]),
    "seldq_conv_pack_weights": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32, ctypes.POINTER(_P), _P, _P]),
    "seldq_conv_pack_table_entry_bytes": (ctypes.c_size_t, []),
    "seldq_conv_pack_table_fill": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32, ctypes.POINTER(_P), _P, _P,
                                                  ctypes.POINTER(ctypes.c_int32)]),
    "seldq_conv_pack_table_run": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P]),
    "seldq_conv_fwd": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, ctypes.POINTER(_P), _P, _P, _P, _P,
                                      ctypes.c_size_t, _P]),
    "seldq_conv_fwd_bf16": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, ctypes.POINTER(_P), _P, _P, _P, _P,
                                           ctypes.c_size_t, _P]),
    "seldq_bn_stats": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, _P, _P]),
    "seldq_bn_finalize": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, ctypes.c_double, ctypes.c_float, ctypes.c_float,
                                         _P, _P, _P, _P]),
    "seldq_cnn_tail_fwd": (ctypes.c_int, [ctypes.POINTER(CnnTailDesc), ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P,
                                          _P, _P, _P]),
    "seldq_cnn_tail_bwd": (ctypes.c_int, [ctypes.POINTER(CnnTailDesc), ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P,
                                          _P, _P, _P, _P]),
    "seldq_cnn_first_bwd_supported": (ctypes.c_int, [ctypes.POINTER(CnnTailDesc), ctypes.POINTER(ConvDesc)]),
    "seldq_cnn_first_bwd_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(ConvDesc)]),
    "seldq_cnn_first_bwd": (ctypes.c_int, [ctypes.POINTER(CnnTailDesc), ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P,
                                           _P, ctypes.POINTER(_P), ctypes.c_int32, _P, ctypes.c_size_t, _P]),
    "seldq_tcn_glue": (ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(TcnGlue), ctypes.POINTER(ConvDesc), ctypes.c_int32,
                                      _P]),
    "seldq_tcn_glue_fused_supported": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "seldq_conv_dgrad": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, ctypes.POINTER(_P), _P, _P, _P,
                                        ctypes.c_size_t, _P]),
    "seldq_conv_wgrad": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, ctypes.POINTER(_P), _P,
                                        ctypes.c_int32, _P, ctypes.c_size_t, _P]),
    "seldq_conv_epi": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32, _P, _P, _P, ctypes.POINTER(ConvEpilogue),
                                      _P]),
    "seldq_conv_pair_supported": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32]),
    "seldq_conv_pair": (ctypes.c_int, [ctypes.POINTER(ConvDesc), ctypes.c_int32, _P, _P, _P, _P, _P, _P,
                                       ctypes.POINTER(ConvEpilogue), ctypes.POINTER(ConvEpilogue), _P]),
    "seldq_conv_wgrad_pair": (ctypes.c_int, [ctypes.POINTER(ConvDesc), _P, _P, _P, ctypes.POINTER(_P),
                                             ctypes.POINTER(_P), ctypes.c_int32, _P]),
    "seldq_linear_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(LinearDesc), ctypes.c_int32]),
    "seldq_linear_fwd": (ctypes.c_int, [ctypes.POINTER(LinearDesc), _P, ctypes.POINTER(_P), _P, _P, _P,
                                        ctypes.c_size_t, _P]),
    "seldq_linear_dgrad": (ctypes.c_int, [ctypes.POINTER(LinearDesc), _P, ctypes.POINTER(_P), _P, _P,
                                          ctypes.c_size_t, _P]),
    "seldq_linear_wgrad": (ctypes.c_int, [ctypes.POINTER(LinearDesc), _P, _P, ctypes.POINTER(_P), _P,
                                          ctypes.c_int32, _P, ctypes.c_size_t, _P]),
    "seldq_stft_features": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.POINTER(StftOptions), _P, _P]),
    "seldq_act_pool1d_fwd": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "seldq_act_pool1d_bwd": (ctypes.c_int, [_P, _P, _P, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                            _P, _P]),
    "seldq_adam_step": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, _P, _P]),
    "seldq_adam_step_part": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_double, _P, ctypes.c_int32, _P]),
    "seldq_seld_events": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_float, _P, _P, _P]),
    "seldq_debug_fprop_trace": (ctypes.c_int, [_P]),
    "seldq_rotation_weight": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "seldq_rotation_weight_bwd": (ctypes.c_int, [ctypes.POINTER(_P), _P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                 ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(_P), _P]),
    "seldq_quaternion_pointwise": (ctypes.c_int, [ctypes.c_int32, _P, _P, _P, ctypes.c_int64, ctypes.c_int64, _P]),
    "seldq_attention_supported": (ctypes.c_int, [ctypes.POINTER(AttentionDesc)]),
    "seldq_attention_saved_bytes": (ctypes.c_size_t, [ctypes.POINTER(AttentionDesc)]),
    "seldq_attention_bwd_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(AttentionDesc)]),
    "seldq_attention_fwd": (ctypes.c_int, [ctypes.POINTER(AttentionDesc), _P, _P, _P, _P, _P, _P, _P]),
    "seldq_attention_bwd": (ctypes.c_int, [ctypes.POINTER(AttentionDesc), _P, _P, _P, _P, _P, _P, _P, _P,
                                           ctypes.c_size_t, _P]),
    "seldq_cast_bf16": (ctypes.c_int, [_P, _P, ctypes.c_size_t, _P]),
    "seldq_stft_shape": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int32)]),
    "seldq_stft_magphase": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libseldq.so is missing (%s). Build it with __graft_entry__.build() or csrc/build.sh; "
                "this package has no CPU or PyTorch fallback." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        if h.seldq_abi_version() != ABI_VERSION:
            raise RuntimeError("libseldq.so reports ABI version %d, this package binds version %d: rebuild it "
                               "(csrc/build.sh)" % (h.seldq_abi_version(), ABI_VERSION))
        _lib = h
    return _lib


def exported_symbols():
    return sorted(_PROTOS)


def check(rc):
    if rc == 0:
        return
    msg = lib().seldq_last_error().decode("utf-8", "replace")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError("seldq: " + msg)
    if rc == ERR_INVALID:
        raise RuntimeError("seldq: " + msg)
    raise RuntimeError("seldq (status %d): %s" % (rc, msg))


def ptr_array(ptrs):
    return (_P * len(ptrs))(*ptrs)
