"""ctypes binding of libdesta_b200.so (C ABI in include/desta_b200.h).  Fails loudly when the library is absent."""
import ctypes
import os
from ctypes import CFUNCTYPE, POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DESTA_B200_LIB") or os.path.join(HERE, "libdesta_b200.so")   # override: A/B runs of two builds


class GemmEpilogue(Structure):
    _fields_ = [("out_f32", c_void_p), ("ld_f32", c_int), ("out_bf16", c_void_p), ("ld_bf16", c_int),
                ("bias", c_void_p), ("resid", c_void_p), ("ld_resid", c_int), ("pos", c_void_p), ("pos_rows", c_int),
                ("scale", c_float), ("scale_cols", c_int), ("gelu", c_int),
                ("ln_stats_out", c_void_p), ("ln_stats_in", c_void_p), ("ln_groups", c_int), ("ln_cols", c_int),
                ("ln_eps", c_float), ("ln_colsum", c_void_p)]


class Dropout(Structure):
    """desta_dropout: keep iff rnd16 >= thr16, kept values * scale; (seed, site, element) -> Philox counter."""
    _fields_ = [("thr16", c_uint32), ("scale", c_float), ("seed", c_uint64), ("site", c_uint32), ("row_keep_ratio", c_void_p)]


class Dims(Structure):
    _fields_ = [("n_mels", c_int), ("d_model", c_int), ("n_heads", c_int), ("ffn", c_int), ("n_layers", c_int),
                ("n_taps", c_int), ("tap_layers", c_int * 8), ("qf_layers", c_int), ("qf_ffn", c_int), ("prompt_size", c_int),
                ("d_llm", c_int), ("max_pos", c_int)]


class EncLayer(Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "o_w", "o_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b",
                                        "fc2_w", "fc2_b", "qkv_w_ln", "qkv_s", "qkv_c", "fc1_w_ln", "fc1_s", "fc1_c")]


class EncoderWeights(Structure):
    _fields_ = [("conv1_w", c_void_p), ("conv1_b", c_void_p), ("conv2_w", c_void_p), ("conv2_b", c_void_p), ("pos", c_void_p),
                ("layers", POINTER(EncLayer))]


class QfLayer(Structure):
    _fields_ = [(n, c_void_p) for n in ("s_qkv_w", "s_qkv_b", "s_o_w", "s_o_b", "s_ln_w", "s_ln_b", "c_q_w", "c_q_b", "c_o_w",
                                        "c_o_b", "c_ln_w", "c_ln_b", "f1_w", "f1_b", "f2_w", "f2_b", "f_ln_w", "f_ln_b")]


class ConnectorWeights(Structure):
    _fields_ = [("layer_weights", c_void_p), ("prompts", c_void_p), ("kv_w", c_void_p), ("kv_b", c_void_p),
                ("proj_ln_w", c_void_p), ("proj_ln_b", c_void_p), ("proj_w", c_void_p), ("proj_b", c_void_p),
                ("layers", POINTER(QfLayer))]


class DropoutPlan(Structure):
    _fields_ = [("p_hidden", c_float), ("p_attn", c_float), ("seed", c_uint64)]


class QfLayerT(Structure):
    _fields_ = [(n, c_void_p) for n in ("s_qkv", "s_o", "c_q", "c_o", "f1", "f2")]


class ConnectorWeightsT(Structure):
    _fields_ = [("proj", c_void_p), ("layers", POINTER(QfLayerT))]


GRAD_READY_FN = CFUNCTYPE(None, c_void_p, c_size_t, c_size_t)


class SpliceDesc(Structure):
    _fields_ = [("prompt", c_int32), ("start", c_int32), ("n_trans", c_int32), ("trans_off", c_int32)]


class ResamplePlan(Structure):
    _fields_ = [("L", c_int32), ("M", c_int32), ("P", c_int32), ("reserved", c_int32)]


class PcmClip(Structure):
    _fields_ = [("pcm", c_void_p), ("n_frames", c_int64), ("n_out", c_int64), ("out_row", c_int64)]


EXPORTS = {
    "desta_version": (c_int, []),
    "desta_last_error": (c_char_p, []),
    "desta_logmel_plan_bytes": (c_size_t, [c_int]),
    "desta_logmel_plan_build": (c_int, [c_int, c_void_p, c_size_t]),
    "desta_logmel": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "desta_resample_plan_bytes": (c_size_t, [c_int, c_int, POINTER(ResamplePlan)]),
    "desta_resample_plan_build": (c_int, [c_int, c_int, c_void_p, c_size_t]),
    "desta_decode_resample": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(ResamplePlan), c_void_p, c_void_p, c_int64, c_int64,
                                      c_void_p]),
    "desta_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, POINTER(GemmEpilogue), c_void_p]),
    "desta_layernorm": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_int, c_void_p,
                                c_int, c_void_p]),
    "desta_attention": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_float, c_void_p, POINTER(Dropout), c_void_p]),
    "desta_attention_variant": (c_int, [c_int]),
    "desta_colsum_heads": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
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
    "desta_row_hash": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "desta_splice_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "desta_im2col_k3": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "desta_layer_mix": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "desta_encoder_workspace_bytes": (c_size_t, [POINTER(Dims), c_int]),
    "desta_encoder_fwd": (c_int, [POINTER(Dims), POINTER(EncoderWeights), c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                                  c_void_p, c_void_p]),
    "desta_connector_workspace_bytes": (c_size_t, [POINTER(Dims), c_int]),
    "desta_connector_fwd": (c_int, [POINTER(Dims), POINTER(ConnectorWeights), c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                                    c_void_p]),
    "desta_connector_train_workspace_bytes": (c_size_t, [POINTER(Dims), c_int]),
    "desta_connector_grad_floats": (c_size_t, [POINTER(Dims)]),
    "desta_connector_train_fwd": (c_int, [POINTER(Dims), POINTER(ConnectorWeights), c_void_p, c_int, c_void_p, c_size_t,
                                          POINTER(DropoutPlan), c_void_p, c_void_p]),
    "desta_connector_bwd": (c_int, [POINTER(Dims), POINTER(ConnectorWeights), POINTER(ConnectorWeightsT), c_void_p, c_int, c_void_p,
                                    c_size_t, POINTER(DropoutPlan), c_void_p, c_void_p, c_size_t, GRAD_READY_FN, c_void_p, c_void_p]),
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
