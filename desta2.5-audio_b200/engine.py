"""Device-side schedule of the hot path: conv stem -> Whisper encoder layers with taps -> Q-Former connector.

This is host code that only sequences launches of libdesta_b200.so kernels over caller-owned buffers (torch is used
for device memory and the stream).  Arithmetic contract: bf16 GEMM operands, fp32
accumulation, fp32 residual stream, fp32 LayerNorm / softmax statistics; taps handed to the connector in bf16.

ref: desta/models/modeling_desta25.py:544-608 (forward_whisper), $TF/models/whisper/modeling_whisper.py:380-414,
$TF/models/bert/modeling_bert.py:379-421.
"""
from dataclasses import dataclass
from typing import Dict, List

import torch

import ctypes

from . import ops
from ._lib import ConnectorWeights, Dims, EncLayer, EncoderWeights, QfLayer, check, lib

LOG2E, LN2 = 1.4426950408889634, 0.6931471805599453
ENC = "whisper.model.encoder."
CON = "connector."


@dataclass
class PathDims:
    n_mels: int
    d_model: int
    n_heads: int
    ffn: int
    n_layers: int
    tap_layers: List[int]
    qf_layers: int
    qf_ffn: int
    prompt_size: int
    d_llm: int
    max_pos: int = 1500

    @property
    def head_dim(self):
        return self.d_model // self.n_heads

    def c_struct(self) -> Dims:
        taps = (ctypes.c_int * 8)(*self.tap_layers)
        return Dims(self.n_mels, self.d_model, self.n_heads, self.ffn, self.n_layers, len(self.tap_layers), taps, self.qf_layers,
                    self.qf_ffn, self.prompt_size, self.d_llm, self.max_pos)


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


class EncoderPack:
    """Frozen Whisper encoder weights repacked once for the kernels (bf16 GEMM operands, fp32 bias / LN / positions)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: PathDims, device):
        d = dims.d_model
        dev = torch.device(device)

        def get(name):
            return sd[ENC + name].to(dev)

        # Conv1d weight [out, in, k] -> GEMM weight [out, k*in] matching im2col column order (tap-major)
        self.conv1_w = _bf16(get("conv1.weight").permute(0, 2, 1).reshape(d, 3 * dims.n_mels))
        self.conv1_b = _f32(get("conv1.bias"))
        self.conv2_w = _bf16(get("conv2.weight").permute(0, 2, 1).reshape(d, 3 * d))
        self.conv2_b = _f32(get("conv2.bias"))
        self.pos = _f32(get("embed_positions.weight")[: dims.max_pos])
        # final encoder LayerNorm: NOT on the tap path (modeling_desta25.py:577-598 never applies it) but needed for the
        # HF-compatible encoder output that the ASR pass (whisper.generate, :1586) consumes
        self.final_ln_w = _f32(get("layer_norm.weight")) if ENC + "layer_norm.weight" in sd else None
        self.final_ln_b = _f32(get("layer_norm.bias")) if ENC + "layer_norm.bias" in sd else None
        self.layers = []
        for i in range(dims.n_layers):
            p = f"layers.{i}."
            qw, kw, vw = get(p + "self_attn.q_proj.weight"), get(p + "self_attn.k_proj.weight"), get(p + "self_attn.v_proj.weight")
            qb, vb = get(p + "self_attn.q_proj.bias"), get(p + "self_attn.v_proj.bias")
            L = dict(
                ln1_w=_f32(get(p + "self_attn_layer_norm.weight")), ln1_b=_f32(get(p + "self_attn_layer_norm.bias")),
                qkv_w=_bf16(torch.cat([qw, kw, vw], 0)),
                qkv_b=_f32(torch.cat([qb, torch.zeros_like(qb), vb], 0)),       # k_proj has no bias
                o_w=_bf16(get(p + "self_attn.out_proj.weight")), o_b=_f32(get(p + "self_attn.out_proj.bias")),
                ln2_w=_f32(get(p + "final_layer_norm.weight")), ln2_b=_f32(get(p + "final_layer_norm.bias")),
                fc1_w=_bf16(get(p + "fc1.weight")), fc1_b=_f32(get(p + "fc1.bias")),
                fc2_w=_bf16(get(p + "fc2.weight")), fc2_b=_f32(get(p + "fc2.bias")),
            )
            # LayerNorm folded into the GEMM that consumes it (desta_gemm_epilogue ln_*):  LN(x) W^T + b  =
            # rstd (x (gain o W)^T - mean colsum) + (W beta + b), colsum taken over the bf16 operand the tensor cores see
            for name, wcat, bcat, ln in (("qkv", torch.cat([qw, kw, vw], 0), L["qkv_b"], "ln1"), ("fc1", get(p + "fc1.weight"), L["fc1_b"], "ln2")):
                gain, beta = L[ln + "_w"], L[ln + "_b"]
                L[name + "_w_ln"] = _bf16(wcat.float() * gain[None, :])
                L[name + "_s"] = L[name + "_w_ln"].float().sum(1).contiguous()
                L[name + "_c"] = (wcat.float() @ beta + bcat).contiguous()
            self.layers.append(L)
        # desta_encoder_weights for the native schedule (desta_encoder_fwd); the tensors above keep the memory alive
        self._c_layers = (EncLayer * dims.n_layers)(*[
            EncLayer(*[L[k].data_ptr() for k in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "o_w", "o_b", "ln2_w", "ln2_b", "fc1_w",
                                                 "fc1_b", "fc2_w", "fc2_b", "qkv_w_ln", "qkv_s", "qkv_c", "fc1_w_ln", "fc1_s",
                                                 "fc1_c")]) for L in self.layers])
        self.c_weights = EncoderWeights(self.conv1_w.data_ptr(), self.conv1_b.data_ptr(), self.conv2_w.data_ptr(),
                                        self.conv2_b.data_ptr(), self.pos.data_ptr(), self._c_layers)


class ConnectorPack:
    """bf16 / fp32 kernel-side copies of the (trainable) connector parameters; rebuilt when a parameter changes."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: PathDims, device):
        dev = torch.device(device)
        d = dims.d_model

        def get(name):
            return sd[CON + name].to(dev)

        self.layer_weights = _f32(get("layer_weights"))
        self.prompts = _f32(torch.cat([get(f"layer_prompts.{k}").reshape(dims.prompt_size, d) for k in range(len(dims.tap_layers))], 0))
        self.layers = []
        kv_w, kv_b = [], []
        for l in range(dims.qf_layers):
            p = f"qformer.layer.{l}."
            sa, ca = p + "attention.", p + "crossattention."
            L = dict(
                s_qkv_w=_bf16(torch.cat([get(sa + "self.query.weight"), get(sa + "self.key.weight"), get(sa + "self.value.weight")], 0)),
                s_qkv_b=_f32(torch.cat([get(sa + "self.query.bias"), get(sa + "self.key.bias"), get(sa + "self.value.bias")], 0)),
                s_o_w=_bf16(get(sa + "output.dense.weight")), s_o_b=_f32(get(sa + "output.dense.bias")),
                s_ln_w=_f32(get(sa + "output.LayerNorm.weight")), s_ln_b=_f32(get(sa + "output.LayerNorm.bias")),
                c_q_w=_bf16(get(ca + "self.query.weight")), c_q_b=_f32(get(ca + "self.query.bias")),
                c_o_w=_bf16(get(ca + "output.dense.weight")), c_o_b=_f32(get(ca + "output.dense.bias")),
                c_ln_w=_f32(get(ca + "output.LayerNorm.weight")), c_ln_b=_f32(get(ca + "output.LayerNorm.bias")),
                f1_w=_bf16(get(p + "intermediate.dense.weight")), f1_b=_f32(get(p + "intermediate.dense.bias")),
                f2_w=_bf16(get(p + "output.dense.weight")), f2_b=_f32(get(p + "output.dense.bias")),
                f_ln_w=_f32(get(p + "output.LayerNorm.weight")), f_ln_b=_f32(get(p + "output.LayerNorm.bias")),
            )
            self.layers.append(L)
            kv_w += [get(ca + "self.key.weight"), get(ca + "self.value.weight")]
            kv_b += [get(ca + "self.key.bias"), get(ca + "self.value.bias")]
        # all layers' cross-attention K/V projections of a tap as ONE GEMM (78 % of connector FLOPs)
        self.kv_w = _bf16(torch.cat(kv_w, 0))          # [L*2d, d]
        self.kv_b = _f32(torch.cat(kv_b, 0))
        self.proj_ln_w, self.proj_ln_b = _f32(get("proj.0.weight")), _f32(get("proj.0.bias"))
        self.proj_w, self.proj_b = _bf16(get("proj.1.weight")), _f32(get("proj.1.bias"))
        names = ("s_qkv_w", "s_qkv_b", "s_o_w", "s_o_b", "s_ln_w", "s_ln_b", "c_q_w", "c_q_b", "c_o_w", "c_o_b", "c_ln_w", "c_ln_b",
                 "f1_w", "f1_b", "f2_w", "f2_b", "f_ln_w", "f_ln_b")
        self._c_layers = (QfLayer * dims.qf_layers)(*[QfLayer(*[L[k].data_ptr() for k in names]) for L in self.layers])
        self.c_weights = ConnectorWeights(self.layer_weights.data_ptr(), self.prompts.data_ptr(), self.kv_w.data_ptr(),
                                          self.kv_b.data_ptr(), self.proj_ln_w.data_ptr(), self.proj_ln_b.data_ptr(),
                                          self.proj_w.data_ptr(), self.proj_b.data_ptr(), self._c_layers)


class Workspace:
    """Caller-owned device buffers for one batch size (the library allocates nothing): the outputs of the encoder stage
    (taps, residual stream), one byte workspace per native stage entry point, and - allocated only when the per-op Python
    schedule is used (profiling / training) - the individual activation buffers."""

    _PY = ("cols1", "c1", "cols2", "y", "ln_stats", "qkv", "ao", "ff", "kv", "x_f32", "x_bf", "q_qkv", "q_ao", "q_tmp", "a_f32", "a_bf", "q_q",
           "c_f32", "c_bf", "q_ff", "mix", "mix_bf")

    def __init__(self, dims: PathDims, N: int, device, keep_for_backward: bool = False):
        d, T, Q = dims.d_model, dims.max_pos, dims.prompt_size
        K = len(dims.tap_layers)
        self.dims, self.N, self.device = dims, N, device
        self.R = K * N * Q
        self.h = torch.empty(N * T, d, dtype=torch.float32, device=device)
        self.taps = torch.empty(K, N * T, d, dtype=torch.bfloat16, device=device)
        self._c_dims = dims.c_struct()
        self._enc_ws = None
        self._con_ws = None
        self.generation = 0         # bumped by every forward that overwrites the buffers (checked by the training backward)
        self._pin = None            # weakref to the token of the training graph that still needs these buffers
        self.h_valid = False

    @property
    def pinned(self):
        return self._pin is not None and self._pin() is not None

    def pin(self, token):
        import weakref
        self._pin = weakref.ref(token)

    def unpin(self):
        self._pin = None

    def native_ws(self, which: str) -> torch.Tensor:
        attr = "_enc_ws" if which == "encoder" else "_con_ws"
        if getattr(self, attr) is None:
            fn = lib().desta_encoder_workspace_bytes if which == "encoder" else lib().desta_connector_workspace_bytes
            nbytes = fn(ctypes.byref(self._c_dims), self.N)
            if nbytes == 0:
                raise ValueError(f"unsupported dims for the native {which} schedule: {lib().desta_last_error().decode()}")
            setattr(self, attr, torch.empty(nbytes, dtype=torch.uint8, device=self.device))
        return getattr(self, attr)

    def __getattr__(self, name):
        if name in Workspace._PY:                       # first use of the per-op schedule: allocate its buffers
            self._alloc_py()
            return self.__dict__[name]
        raise AttributeError(name)

    def _alloc_py(self):
        dims, N, device = self.dims, self.N, self.device
        d, T, Q = dims.d_model, dims.max_pos, dims.prompt_size
        M = N * T
        K = len(dims.tap_layers)
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *s, dt=bf: torch.empty(*s, dtype=dt, device=device)  # noqa: E731
        self.cols1 = e(N * 2 * T, 3 * dims.n_mels)
        self.c1 = e(N * 2 * T, d)
        self.cols2 = e(M, 3 * d)
        self.y = e(M, d)
        self.ln_stats = e(M, (d + 127) // 128, 2, dt=f32)
        self.qkv = e(M, 3 * d)
        self.ao = e(M, d)
        self.ff = e(M, dims.ffn)
        self.kv = e(K * M, dims.qf_layers * 2 * d)
        R = self.R
        self.x_f32 = e(R, d, dt=f32)
        self.x_bf = e(R, d)
        self.q_qkv = e(R, 3 * d)
        self.q_ao = e(R, d)
        self.q_tmp = e(R, d, dt=f32)
        self.a_f32 = e(R, d, dt=f32)
        self.a_bf = e(R, d)
        self.q_q = e(R, d)
        self.c_f32 = e(R, d, dt=f32)
        self.c_bf = e(R, d)
        self.q_ff = e(R, dims.qf_ffn)
        self.mix = e(N * Q, d, dt=f32)
        self.mix_bf = e(N * Q, d)


def _native_enabled():
    return ops._PROFILE is None        # the per-family CUDA-event profile of bench.py needs the per-op schedule


def encoder_forward(enc: EncoderPack, dims: PathDims, ws: Workspace, mel_tm: torch.Tensor, taps=None, h=None, mel_floor=None):
    """mel_tm: bf16 time-major log-mel [n, 3000, n_mels] -> taps[k] (bf16 [n*1500, d]) and h (fp32 [n*1500, d]).  One call
    of the native schedule desta_encoder_fwd (same launches, same order as encoder_forward_ops below).
    By default n == ws.N and the outputs are ws.taps / ws.h.  A SUBSET run (scope row n2: only the clips whose encoder
    state is not cached) passes n < ws.N clips with its own compact `taps` [K, n*1500, d] / `h` [n*1500, d] and borrows
    the byte workspace of the full batch.  ``mel_floor``: clip-max words of ``ops.logmel(want_f32=False)`` when mel_tm still
    lacks the per-clip log-mel floor (applied by the conv1 im2col on read)."""
    n = mel_tm.shape[0]
    taps = ws.taps if taps is None else taps
    h = ws.h if h is None else h
    ops._cuda(mel_tm, torch.bfloat16), ops._cuda(taps, torch.bfloat16), ops._cuda(h, torch.float32)
    assert mel_tm.is_contiguous() and mel_tm.shape == (n, 2 * dims.max_pos, dims.n_mels) and 0 < n <= ws.N
    assert taps.is_contiguous() and taps.shape == (len(dims.tap_layers), n * dims.max_pos, dims.d_model)
    assert h.is_contiguous() and h.shape == (n * dims.max_pos, dims.d_model)
    ops.STATS["encoder_clips"] = ops.STATS.get("encoder_clips", 0) + n
    if not _native_enabled():
        if n != ws.N or taps is not ws.taps or h is not ws.h:          # per-op schedule: buffers are sized per batch
            sub = Workspace(dims, n, ws.device)
            encoder_forward_ops(enc, dims, sub, mel_tm, mel_floor)
            taps.copy_(sub.taps)
            h.copy_(sub.h)
            return
        return encoder_forward_ops(enc, dims, ws, mel_tm, mel_floor)
    wsb = ws.native_ws("encoder")
    ops.STATS["launches"] += 4 + (5 if n * dims.max_pos > 128 else 7) * dims.n_layers
    with ops.device_guard(ops._dev(mel_tm, taps, h, wsb, enc.pos)):
        check(lib().desta_encoder_fwd(ctypes.byref(ws._c_dims), ctypes.byref(enc.c_weights), ops._p(mel_tm), ops._p(mel_floor), n,
                                      ops._p(wsb), wsb.numel(), ops._p(taps), ops._p(h), ops._stream()), "desta_encoder_fwd")


def connector_forward(con: ConnectorPack, dims: PathDims, ws: Workspace, out_f32: torch.Tensor):
    """Q-Former over ws.taps -> out_f32 [N*Q, d_llm] through the native schedule desta_connector_fwd (eval mode)."""
    if not _native_enabled():
        return connector_forward_ops(con, dims, ws, out_f32)
    ops._cuda(out_f32, torch.float32)
    assert out_f32.is_contiguous() and out_f32.shape == (ws.N * dims.prompt_size, dims.d_llm)
    wsb = ws.native_ws("connector")
    ops.STATS["launches"] += 2 + 11 * dims.qf_layers + 3
    with ops.device_guard(ops._dev(ws.taps, wsb, out_f32, con.proj_b)):
        check(lib().desta_connector_fwd(ctypes.byref(ws._c_dims), ctypes.byref(con.c_weights), ops._p(ws.taps), ws.N, ops._p(wsb),
                                        wsb.numel(), ops._p(out_f32), ops._stream()), "desta_connector_fwd")
    return out_f32


def encoder_forward_ops(enc: EncoderPack, dims: PathDims, ws: Workspace, mel_tm: torch.Tensor, mel_floor=None):
    """Per-op schedule of the encoder (one C-ABI call per kernel; profiling / reference for the native schedule)."""
    N, d, T = ws.N, dims.d_model, dims.max_pos
    hd = dims.head_dim
    ops.im2col_k3(mel_tm, N, 2 * T, dims.n_mels, 1, ws.cols1, mel_floor)
    ops.gemm(ws.cols1, enc.conv1_w, enc.conv1_b, out_bf16=ws.c1, gelu=True)
    ops.im2col_k3(ws.c1, N, 2 * T, d, 2, ws.cols2)
    if N * T > 128:
        # LayerNorm folded into the GEMMs around it (same launches as desta_encoder_fwd in schedule.cu)
        stats = ws.ln_stats
        ops.gemm(ws.cols2, enc.conv2_w, enc.conv2_b, out_f32=ws.h, out_bf16=ws.y, gelu=True, pos=enc.pos, ln_stats_out=stats)
        hb, tap_idx = ws.y, 0
        for i, L in enumerate(enc.layers):
            ops.gemm(hb, L["qkv_w_ln"], L["qkv_c"], out_bf16=ws.qkv, scale=LOG2E * hd ** -0.5, scale_cols=d, ln_in=(stats, L["qkv_s"], d, 1e-5))
            ops.attention(ws.qkv[:, :d], ws.qkv[:, d:2 * d], ws.qkv[:, 2 * d:], ws.ao, N, dims.n_heads, T, T, LN2)
            ops.gemm(ws.ao, L["o_w"], L["o_b"], out_f32=ws.h, resid=ws.h, out_bf16=ws.y, ln_stats_out=stats)
            ops.gemm(ws.y, L["fc1_w_ln"], L["fc1_c"], out_bf16=ws.ff, gelu=True, ln_in=(stats, L["fc1_s"], d, 1e-5))
            hb = ws.y
            if i in dims.tap_layers:
                hb = ws.taps[tap_idx]
                tap_idx += 1
            ops.gemm(ws.ff, L["fc2_w"], L["fc2_b"], out_f32=ws.h, resid=ws.h, out_bf16=hb, ln_stats_out=stats)
        assert tap_idx == len(dims.tap_layers)
        return
    ops.gemm(ws.cols2, enc.conv2_w, enc.conv2_b, out_f32=ws.h, gelu=True, pos=enc.pos)
    tap_idx = 0
    for i, L in enumerate(enc.layers):
        ops.layernorm(ws.h, L["ln1_w"], L["ln1_b"], 1e-5, out_bf16=ws.y)
        # head_dim^-1/2 * log2(e) folded into q, attention scale ln 2: same arithmetic as desta_encoder_fwd (schedule.cu)
        ops.gemm(ws.y, L["qkv_w"], L["qkv_b"], out_bf16=ws.qkv, scale=LOG2E * hd ** -0.5, scale_cols=d)
        ops.attention(ws.qkv[:, :d], ws.qkv[:, d:2 * d], ws.qkv[:, 2 * d:], ws.ao, N, dims.n_heads, T, T, LN2)
        ops.gemm(ws.ao, L["o_w"], L["o_b"], out_f32=ws.h, resid=ws.h)
        ops.layernorm(ws.h, L["ln2_w"], L["ln2_b"], 1e-5, out_bf16=ws.y)
        ops.gemm(ws.y, L["fc1_w"], L["fc1_b"], out_bf16=ws.ff, gelu=True)
        if i in dims.tap_layers:
            ops.gemm(ws.ff, L["fc2_w"], L["fc2_b"], out_f32=ws.h, resid=ws.h, out_bf16=ws.taps[tap_idx])
            tap_idx += 1
        else:
            ops.gemm(ws.ff, L["fc2_w"], L["fc2_b"], out_f32=ws.h, resid=ws.h)
    assert tap_idx == len(dims.tap_layers)


def connector_forward_ops(con: ConnectorPack, dims: PathDims, ws: Workspace, out_f32: torch.Tensor):
    """Per-op schedule of the connector: taps in ws.taps (all taps batched as K*N sequences) -> out_f32 [N*Q, d_llm]."""
    N, d, T, Q = ws.N, dims.d_model, dims.max_pos, dims.prompt_size
    K = len(dims.tap_layers)
    hd = dims.head_dim
    B = K * N
    scale = hd ** -0.5
    # cross-attention K/V of every Q-Former layer for every tap: one GEMM over [K*N*1500, d]
    ops.gemm(ws.taps.view(K * N * T, d), con.kv_w, con.kv_b, out_bf16=ws.kv)
    # x0 = layer_prompts[k] broadcast over clips
    ws.x_f32.view(K, N, Q, d).copy_(con.prompts.view(K, 1, Q, d).expand(K, N, Q, d))
    ws.x_bf.copy_(ws.x_f32)
    for l, L in enumerate(con.layers):
        ops.gemm(ws.x_bf, L["s_qkv_w"], L["s_qkv_b"], out_bf16=ws.q_qkv)
        ops.attention(ws.q_qkv[:, :d], ws.q_qkv[:, d:2 * d], ws.q_qkv[:, 2 * d:], ws.q_ao, B, dims.n_heads, Q, Q, scale)
        ops.gemm(ws.q_ao, L["s_o_w"], L["s_o_b"], out_f32=ws.q_tmp, resid=ws.x_f32)
        ops.layernorm(ws.q_tmp, L["s_ln_w"], L["s_ln_b"], 1e-12, out_bf16=ws.a_bf, out_f32=ws.a_f32)
        ops.gemm(ws.a_bf, L["c_q_w"], L["c_q_b"], out_bf16=ws.q_q)
        kcol = l * 2 * d
        ops.attention(ws.q_q, ws.kv[:, kcol:kcol + d], ws.kv[:, kcol + d:kcol + 2 * d], ws.q_ao, B, dims.n_heads, Q, T, scale)
        ops.gemm(ws.q_ao, L["c_o_w"], L["c_o_b"], out_f32=ws.q_tmp, resid=ws.a_f32)
        ops.layernorm(ws.q_tmp, L["c_ln_w"], L["c_ln_b"], 1e-12, out_bf16=ws.c_bf, out_f32=ws.c_f32)
        ops.gemm(ws.c_bf, L["f1_w"], L["f1_b"], out_bf16=ws.q_ff, gelu=True)
        ops.gemm(ws.q_ff, L["f2_w"], L["f2_b"], out_f32=ws.q_tmp, resid=ws.c_f32)
        ops.layernorm(ws.q_tmp, L["f_ln_w"], L["f_ln_b"], 1e-12, out_bf16=ws.x_bf, out_f32=ws.x_f32)
    ops.layer_mix(ws.x_f32.view(K, N * Q, d), con.layer_weights, N, Q, ws.mix)
    ops.layernorm(ws.mix, con.proj_ln_w, con.proj_ln_b, 1e-5, out_bf16=ws.mix_bf)
    ops.gemm(ws.mix_bf, con.proj_w, con.proj_b, out_f32=out_f32)
    return out_f32
