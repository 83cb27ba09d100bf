"""ctypes binding of libdesta_b200.so (C ABI in include/desta_b200.h).  Fails loudly when the library is absent."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdesta_b200.so")


class GemmEpilogue(Structure):
    _fields_ = [("out_f32", c_void_p), ("ld_f32", c_int), ("out_bf16", c_void_p), ("ld_bf16", c_int),
                ("bias", c_void_p), ("resid", c_void_p), ("ld_resid", c_int), ("pos", c_void_p), ("pos_rows", c_int),
                ("scale", c_float), ("scale_cols", c_int), ("gelu", c_int)]


class Dropout(Structure):
    """desta_dropout: keep iff rnd16 >= thr16, kept values * scale; (seed, site, element) -> Philox counter."""
    _fields_ = [("thr16", c_uint32), ("scale", c_float), ("seed", c_uint64), ("site", c_uint32)]


class SpliceDesc(Structure):
    _fields_ = [("prompt", c_int32), ("start", c_int32), ("n_trans", c_int32), ("trans_off", c_int32)]


EXPORTS = {
    "desta_version": (c_int, []),
    "desta_last_error": (c_char_p, []),
    "desta_logmel_plan_bytes": (c_size_t, [c_int]),
    "desta_logmel_plan_build": (c_int, [c_int, c_void_p, c_size_t]),
    "desta_logmel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "desta_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, POINTER(GemmEpilogue), c_void_p]),
    "desta_layernorm": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_int, c_void_p,
                                c_int, c_void_p]),
    "desta_attention": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_float, c_void_p, POINTER(Dropout), c_void_p]),
    "desta_gemm_bf16_wgrad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "desta_attention_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                    c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_float, POINTER(Dropout), c_void_p]),
    "desta_dropout_add": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, POINTER(Dropout), c_void_p, c_int, c_void_p,
                                  c_int, c_void_p]),
    "desta_dropout_mask": (c_int, [POINTER(Dropout), c_int, c_int, c_void_p, c_void_p]),
    "desta_layernorm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_float, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "desta_gelu_fwd": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "desta_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "desta_colsum": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "desta_layer_mix_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "desta_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "desta_splice_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "desta_im2col_k3": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "desta_layer_mix": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "desta_splice": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, POINTER(SpliceDesc), c_int, c_void_p, c_void_p, c_int,
                             c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


class DestaError(RuntimeError):
    pass


def lib():
    """Load libdesta_b200.so (built in-tree by ``python desta2.5-audio_b200/build.py`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DestaError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)      # AttributeError if the ABI drifted from include/desta_b200.h
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().desta_last_error().decode(errors="replace")
        if rc == -3:
            raise ValueError(f"{what}: {msg}")
        raise DestaError(f"{what} failed ({rc}): {msg}")
