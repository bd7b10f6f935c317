"""ctypes binding of libsagan_b200.so (C ABI declared in include/sagan_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails,
this module raises.  PyTorch is only used by callers for device memory and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SAGAN_B200_LIB: another build of the same library (A/B kernel measurements on one box); still no fallback
LIB_PATH = os.environ.get("SAGAN_B200_LIB") or os.path.join(_HERE, "libsagan_b200.so")

MATH_FP32_STRICT = 0
MATH_BF16_TC = 1
ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2
CONV_TC_TF32, CONV_TC_SPLIT_BF16 = 1, 2


class SaganError(RuntimeError):
    pass


class SnDesc(C.Structure):
    _fields_ = [("W", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("W_bar", C.c_void_p),
                ("W_bar_bf16", C.c_void_p), ("sigma", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32),
                ("Ip", C.c_int32), ("factor", C.c_float)]


class ConvGeom(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
                ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad_t", C.c_int32),
                ("pad_l", C.c_int32)]


class SnBwdDesc(C.Structure):
    _fields_ = [("dW_bar", C.c_void_p), ("W_bar", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("sigma", C.c_void_p), ("dW", C.c_void_p), ("factor", C.c_float), ("rows", C.c_int32),
                ("cols", C.c_int32)]


class AccDesc(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("n", C.c_longlong)]


class DpPeers(C.Structure):
    _fields_ = [("grads", C.c_void_p * 8), ("params", C.c_void_p * 8), ("flags", C.c_void_p * 8)]      # sagan_dp_peers


_P, _I, _F, _LL, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# every symbol include/sagan_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "sagan_abi_version": (_I, []),
    "sagan_last_error": (C.c_char_p, []),
    "sagan_launch_count": (C.c_ulonglong, []),
    "sagan_deterministic_forward": (_I, [_I]),
    "sagan_sn_plan_create": (_I, [C.POINTER(SnDesc), _I, _I, C.POINTER(_P)]),
    "sagan_sn_plan_run": (_I, [_P, _P]),
    "sagan_sn_plan_refresh": (_I, [_P, _P]),
    "sagan_sn_plan_destroy": (_I, [_P]),
    "sagan_sn_plan_algorithmic_bytes": (C.c_ulonglong, [_P]),
    "sagan_sn_plan_phase_times": (_I, [_P, C.POINTER(C.c_float)]),
    "sagan_sn_backward_workspace_bytes": (_SZ, [_LL]),
    "sagan_sn_backward": (_I, [_P, _P, _P, _P, _P, _F, _P, _I, _I, _P, _SZ, _P]),
    "sagan_sn_backward_multi": (_I, [C.POINTER(SnBwdDesc), _I, _I, _P, _SZ, _P]),
    "sagan_attn_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "sagan_attn_fwd": (_I, [_P] * 13 + [_I, _I, _I, _I, _P, _SZ, _P]),
    "sagan_attn_bwd": (_I, [_P] * 23 + [_I, _I, _I, _I, _P, _SZ, _P]),
    "sagan_attn_pool_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "sagan_attn_pool_fwd": (_I, [_P] * 13 + [_I, _I, _I, _I, _I, _P, _SZ, _P]),
    "sagan_attn_pool_bwd": (_I, [_P] * 23 + [_I, _I, _I, _I, _I, _P, _SZ, _P]),
    "sagan_conv2d_fwd": (_I, [_P, _P, _P, _P, C.POINTER(ConvGeom), _I, _F, _I, _P]),
    "sagan_conv2d_dgrad": (_I, [_P, _P, _P, C.POINTER(ConvGeom), _I, _P]),
    "sagan_conv2d_wgrad": (_I, [_P, _P, _P, _P, C.POINTER(ConvGeom), _I, _P]),
    "sagan_conv_tc_precision": (_I, [_I]),
    "sagan_act_bwd": (_I, [_P, _P, _P, _LL, _I, _F, _P]),
    "sagan_ew_fwd": (_I, [_P, _P, _P, _P, _LL, _I, _I, _F, _P]),
    "sagan_colsum": (_I, [_P, _P, _LL, _I, _P]),
    "sagan_bn_workspace_bytes": (_SZ, [_I]),
    "sagan_bn_lrelu_fwd": (_I, [_P] * 8 + [_LL, _I, _F, _F, _F, _P, _SZ, _P]),
    "sagan_bn_lrelu_bwd": (_I, [_P] * 9 + [_LL, _I, _F, _P, _SZ, _P]),
    "sagan_bn_lrelu_infer": (_I, [_P] * 6 + [_LL, _I, _F, _F, _P]),
    "sagan_wn_fwd": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "sagan_wn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "sagan_u8_to_f32": (_I, [_P, _P, _LL, _F, _F, _P]),
    "sagan_hinge_d": (_I, [_P, _P, _LL, _F, _P, _P, _P, _P]),
    "sagan_hinge_g": (_I, [_P, _LL, _F, _P, _P, _P]),
    "sagan_adam_step": (_I, [_P, _P, _P, _P, _LL, _P, _F, _P]),
    "sagan_accumulate_multi": (_I, [C.POINTER(AccDesc), _I, _P]),
    "sagan_adam_schedule": (_I, [_P, _P, C.c_double, C.c_double, _LL, C.c_double, C.c_double, C.c_double, _P]),
    "sagan_dp_max_world": (_I, []),
    "sagan_dp_set_timeout_ms": (_I, [_LL]),
    "sagan_dp_flag_bytes": (_SZ, []),
    "sagan_dp_sum_adam": (_I, [C.POINTER(DpPeers), _I, _I, _LL, _P, _P, _P, _P, _P]),
    "sagan_dp_sum_adam_losses": (_I, [C.POINTER(DpPeers), _I, _I, _LL, _P, _P, _P, _P, C.POINTER(_P), _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises SaganError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SaganError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            f"(or `make -C self-attention-gan_b200/csrc`). There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.sagan_abi_version() != 1:
        raise SaganError(f"ABI version mismatch: library reports {lib.sagan_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().sagan_last_error()
        msg = msg.decode() if msg else ""
        kind = "bad argument" if rc == -1 else "unsupported" if rc == -2 else "workspace" if rc == -3 else f"cuda error {rc}"
        if rc == -1 and "power iterations" in msg:
            raise ValueError(msg)
        raise SaganError(f"{what} failed ({kind}): {msg}")


def launch_count():
    return int(load().sagan_launch_count())
