"""Connector training path: forward that keeps the activations the backward needs, and the hand-written backward
(row a6 of the scope table: autograd of desta/models/modeling_desta25.py:587-606 over $TF/models/bert/modeling_bert.py).

Only the connector's parameters receive gradients: the Whisper encoder is frozen (modeling_desta25.py:1451-1453), so
taps are constants and the K/V projections need a weight gradient but no data gradient.

Train-mode dropout of the Q-Former (BertConfig hidden_dropout_prob / attention_probs_dropout_prob = 0.1,
$TF/models/bert/modeling_bert.py:198-201, 296, 354) is applied with a counter-based Philox mask: the forward draws one
64-bit seed per step from torch's default generator (so torch.manual_seed governs it) and the backward regenerates
every mask from (seed, site) instead of storing it.  In eval mode nothing is dropped, exactly like nn.Dropout.

The autograd.Function exposes every connector parameter as an input so that HF Trainer / Adafactor / DDP see ordinary
`.grad`s on the reference's own parameter names.
"""
from typing import Dict, List

import torch

import ctypes

from . import engine, ops
from ._lib import ConnectorWeightsT, DropoutPlan, GRAD_READY_FN, QfLayerT, check, lib
from .engine import CON, PathDims


def _bf16_t(w):
    return w.detach().to(torch.bfloat16).t().contiguous()


class ConnectorTPack:
    """Transposed bf16 weights for the data-gradient GEMMs (dX = dY · W is run as a TN GEMM against W^T)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: PathDims, device):
        dev = torch.device(device)

        def get(name):
            return sd[CON + name].to(dev)

        self.layers = []
        for l in range(dims.qf_layers):
            p = f"qformer.layer.{l}."
            sa, ca = p + "attention.", p + "crossattention."
            self.layers.append(dict(
                s_qkv=_bf16_t(torch.cat([get(sa + "self.query.weight"), get(sa + "self.key.weight"), get(sa + "self.value.weight")], 0)),
                s_o=_bf16_t(get(sa + "output.dense.weight")),
                c_q=_bf16_t(get(ca + "self.query.weight")),
                c_o=_bf16_t(get(ca + "output.dense.weight")),
                f1=_bf16_t(get(p + "intermediate.dense.weight")),
                f2=_bf16_t(get(p + "output.dense.weight")),
            ))
        self.proj = _bf16_t(get("proj.1.weight"))
        names = ("s_qkv", "s_o", "c_q", "c_o", "f1", "f2")
        self._c_layers = (QfLayerT * dims.qf_layers)(*[QfLayerT(*[L[k].data_ptr() for k in names]) for L in self.layers])
        self.c_weights = ConnectorWeightsT(self.proj.data_ptr(), self._c_layers)


SITE_SELF_PROBS, SITE_SELF_OUT, SITE_CROSS_PROBS, SITE_CROSS_OUT, SITE_FFN_OUT = range(5)


class DropPlan:
    """Dropout sites of one training step: site id = layer * 8 + position (the five nn.Dropout calls of a BertLayer)."""

    def __init__(self, p_hidden: float = 0.0, p_attn: float = 0.0, seed: int = 0):
        self.p_hidden, self.p_attn, self.seed = float(p_hidden), float(p_attn), int(seed)

    @property
    def active(self):
        return self.p_hidden > 0 or self.p_attn > 0

    def site(self, layer: int, which: int, row_keep_ratio=None):
        p = self.p_attn if which in (SITE_SELF_PROBS, SITE_CROSS_PROBS) else self.p_hidden
        return ops.make_dropout(p, self.seed, layer * 8 + which, row_keep_ratio)

    def c_struct(self) -> DropoutPlan:
        return DropoutPlan(self.p_hidden, self.p_attn, self.seed & 0xFFFFFFFFFFFFFFFF)


NO_DROP = DropPlan()


class TrainState:
    """Gradient arena + (lazily) the saved activations of one batch size: one byte workspace for the native stage entry
    points (desta_connector_train_fwd / desta_connector_bwd), or - when the per-op schedule runs (profiling) - the
    individual activation buffers."""

    _PY = ("layers", "x_out", "x_out_bf", "mix", "mix_bf", "dkv", "g_f32_a", "g_f32_b", "g_f32_c", "g_bf_a", "g_bf_b", "g_qkv",
           "g_ff", "g_h1", "g_out_bf", "g_mix", "g_mixn", "g_x")

    def __getattr__(self, name):
        if name in TrainState._PY:                      # first use of the per-op schedule: allocate its buffers
            self._alloc_py()
            return self.__dict__[name]
        raise AttributeError(name)

    @property
    def pinned(self):
        return self._pin is not None and self._pin() is not None

    def pin(self, token):
        import weakref
        self._pin = weakref.ref(token)

    def unpin(self):
        self._pin = None

    def native_ws(self) -> torch.Tensor:
        if self._native_ws is None:
            nbytes = lib().desta_connector_train_workspace_bytes(ctypes.byref(self._c_dims), self.N)
            if nbytes == 0:
                raise ValueError(f"unsupported dims for the native training schedule: {lib().desta_last_error().decode()}")
            self._native_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._native_ws

    def _alloc_py(self):
        dims, N, device = self.dims, self.N, self.device
        d, Q, T = dims.d_model, dims.prompt_size, dims.max_pos
        K, L, H = len(dims.tap_layers), dims.qf_layers, dims.n_heads
        R, B = K * N * Q, K * N
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *s, dt=bf: torch.empty(*s, dtype=dt, device=device)  # noqa: E731
        self.layers = [dict(x_bf=e(R, d), x_f32=e(R, d, dt=f32), qkv=e(R, 3 * d), lse_s=e(B * H * Q, dt=f32), ao_s=e(R, d),
                            t1=e(R, d, dt=f32), a_bf=e(R, d), a_f32=e(R, d, dt=f32), q_c=e(R, d), lse_c=e(B * H * Q, dt=f32),
                            rr_c=e(B * H * Q, dt=f32),
                            ao_c=e(R, d), t2=e(R, d, dt=f32), c_bf=e(R, d), c_f32=e(R, d, dt=f32), h1=e(R, dims.qf_ffn),
                            f=e(R, dims.qf_ffn), t3=e(R, d, dt=f32)) for _ in range(L)]
        self.x_out = e(R, d, dt=f32)
        self.x_out_bf = e(R, d)
        self.mix = e(N * Q, d, dt=f32)
        self.mix_bf = e(N * Q, d)
        # backward scratch
        self.dkv = e(K * N * T, L * 2 * d)
        self.g_f32_a = e(R, d, dt=f32)
        self.g_f32_b = e(R, d, dt=f32)
        self.g_f32_c = e(R, d, dt=f32)
        self.g_bf_a = e(R, d)
        self.g_bf_b = e(R, d)
        self.g_qkv = e(R, 3 * d)
        self.g_ff = e(R, dims.qf_ffn)
        self.g_h1 = e(R, dims.qf_ffn)
        self.g_out_bf = e(N * Q, dims.d_llm)
        self.g_mix = e(N * Q, d, dt=f32)
        self.g_mixn = e(N * Q, d, dt=f32)
        self.g_x = e(R, d, dt=f32)

    def __init__(self, dims: PathDims, N: int, device):
        d, Q, T = dims.d_model, dims.prompt_size, dims.max_pos
        K, L, H = len(dims.tap_layers), dims.qf_layers, dims.n_heads
        f32 = torch.float32
        self.dims, self.N, self.device = dims, N, device
        self._c_dims = dims.c_struct()
        self._native_ws = None
        self.generation = 0         # bumped by every training forward that overwrites the saved activations
        self._pin = None            # weakref to the token of the autograd graph that still needs them (see _ConnectorFn)
        # gradient buffers (fp32, fused like the packed weights): views of ONE arena, so that zeroing them and handing a
        # private copy to autograd are single launches instead of one per tensor (165 parameters at the flagship size)
        shapes = dict(layer_weights=(Q, K), prompts=(K * Q, d), kv_w=(L * 2 * d, d), kv_b=(L * 2 * d,),
                      proj_ln_w=(d,), proj_ln_b=(d,), proj_w=(dims.d_llm, d), proj_b=(dims.d_llm,))
        lshapes = dict(s_qkv_w=(3 * d, d), s_qkv_b=(3 * d,), s_o_w=(d, d), s_o_b=(d,), s_ln_w=(d,), s_ln_b=(d,),
                       c_q_w=(d, d), c_q_b=(d,), c_o_w=(d, d), c_o_b=(d,), c_ln_w=(d,), c_ln_b=(d,),
                       f1_w=(dims.qf_ffn, d), f1_b=(dims.qf_ffn,), f2_w=(d, dims.qf_ffn), f2_b=(d,), f_ln_w=(d,), f_ln_b=(d,))

        def numel(shape):
            n = 1
            for v in shape:
                n *= v
            return (n + 3) // 4 * 4                     # keep every buffer 16-byte aligned

        total = sum(numel(v) for v in shapes.values()) + L * sum(numel(v) for v in lshapes.values())
        native_total = lib().desta_connector_grad_floats(ctypes.byref(self._c_dims))
        assert native_total in (0, total), (native_total, total)     # the C schedule carves the same layout
        self.arena = torch.zeros(total, dtype=f32, device=device)
        off = 0

        def carve(shape):
            nonlocal off
            n = 1
            for v in shape:
                n *= v
            t = self.arena[off: off + n].view(*shape)
            off += numel(shape)
            return t

        self.grads = {k: carve(v) for k, v in shapes.items()}
        self.global_slice = (0, off)                        # arena range of the non-layer gradients (complete last)
        self.lgrads, self.layer_slices = [], []
        for _ in range(L):
            start = off
            self.lgrads.append({k: carve(v) for k, v in lshapes.items()})
            self.layer_slices.append((start, off))          # arena range of one Q-Former layer's gradients

    def zero_grads(self):
        self.arena.zero_()


def _dense_res(a_bf, w, b, t, resid, drop):
    """t = dropout(a W^T + b) + resid  (BertSelfOutput / BertOutput before their LayerNorm)."""
    if drop is None:
        ops.gemm(a_bf, w, b, out_f32=t, resid=resid)
    else:
        ops.gemm(a_bf, w, b, out_f32=t)
        ops.dropout_add(t, resid, drop, out_f32=t)


def connector_forward_train(con, dims: PathDims, ws, st: TrainState, out_f32, plan: DropPlan = NO_DROP):
    N, d, T, Q = ws.N, dims.d_model, dims.max_pos, dims.prompt_size
    K, H = len(dims.tap_layers), dims.n_heads
    B = K * N
    scale = dims.head_dim ** -0.5
    ops.gemm(ws.taps.view(K * N * T, d), con.kv_w, con.kv_b, out_bf16=ws.kv)
    x_f32, x_bf = st.layers[0]["x_f32"], st.layers[0]["x_bf"]
    x_f32.view(K, N, Q, d).copy_(con.prompts.view(K, 1, Q, d).expand(K, N, Q, d))
    x_bf.copy_(x_f32)
    for l, (L, S) in enumerate(zip(con.layers, st.layers)):
        ops.gemm(S["x_bf"], L["s_qkv_w"], L["s_qkv_b"], out_bf16=S["qkv"])
        ops.attention(S["qkv"][:, :d], S["qkv"][:, d:2 * d], S["qkv"][:, 2 * d:], S["ao_s"], B, H, Q, Q, scale, lse=S["lse_s"],
                      drop=plan.site(l, SITE_SELF_PROBS))
        _dense_res(S["ao_s"], L["s_o_w"], L["s_o_b"], S["t1"], S["x_f32"], plan.site(l, SITE_SELF_OUT))
        ops.layernorm(S["t1"], L["s_ln_w"], L["s_ln_b"], 1e-12, out_bf16=S["a_bf"], out_f32=S["a_f32"])
        ops.gemm(S["a_bf"], L["c_q_w"], L["c_q_b"], out_bf16=S["q_c"])
        kc = l * 2 * d
        # rr_c: per query row, sum of the dropped probabilities over their plain sum - what the value-bias gradient needs
        ops.attention(S["q_c"], ws.kv[:, kc:kc + d], ws.kv[:, kc + d:kc + 2 * d], S["ao_c"], B, H, Q, T, scale, lse=S["lse_c"],
                      drop=plan.site(l, SITE_CROSS_PROBS, S["rr_c"]))
        _dense_res(S["ao_c"], L["c_o_w"], L["c_o_b"], S["t2"], S["a_f32"], plan.site(l, SITE_CROSS_OUT))
        ops.layernorm(S["t2"], L["c_ln_w"], L["c_ln_b"], 1e-12, out_bf16=S["c_bf"], out_f32=S["c_f32"])
        ops.gemm(S["c_bf"], L["f1_w"], L["f1_b"], out_bf16=S["h1"])          # pre-activation kept for the backward
        ops.gelu_fwd(S["h1"], S["f"])
        _dense_res(S["f"], L["f2_w"], L["f2_b"], S["t3"], S["c_f32"], plan.site(l, SITE_FFN_OUT))
        nxt = st.layers[l + 1] if l + 1 < len(st.layers) else None
        ops.layernorm(S["t3"], L["f_ln_w"], L["f_ln_b"], 1e-12,
                      out_bf16=nxt["x_bf"] if nxt else st.x_out_bf, out_f32=nxt["x_f32"] if nxt else st.x_out)
    ops.layer_mix(st.x_out.view(K, N * Q, d), con.layer_weights, N, Q, st.mix)
    ops.layernorm(st.mix, con.proj_ln_w, con.proj_ln_b, 1e-5, out_bf16=st.mix_bf)
    ops.gemm(st.mix_bf, con.proj_w, con.proj_b, out_f32=out_f32)
    return out_f32


def connector_backward(con, conT: ConnectorTPack, dims: PathDims, ws, st: TrainState, grad_out, plan: DropPlan = NO_DROP,
                       on_ready=None):
    """grad_out fp32 [N*Q, d_llm] -> fills st.grads / st.lgrads (+=; call st.zero_grads() first).
    `on_ready(arena_slice)` is called as soon as a contiguous part of the gradient arena is final (one Q-Former layer at a
    time, last layer first; the remaining tensors at the end) so that a data-parallel all-reduce can overlap the rest of
    the backward (desta_b200.dist.GradSync)."""
    N, d, T, Q = ws.N, dims.d_model, dims.max_pos, dims.prompt_size
    K, H = len(dims.tap_layers), dims.n_heads
    B, R = K * N, K * N * Q
    scale = dims.head_dim ** -0.5
    G = st.grads
    # ---- proj: out = LN(mix) Wp^T + bp ----
    ops.cast_bf16(grad_out, st.g_out_bf)
    ops.gemm_wgrad(st.g_out_bf, st.mix_bf, G["proj_w"])
    ops.colsum(grad_out, G["proj_b"])
    ops.gemm(st.g_out_bf, conT.proj, None, out_f32=st.g_mixn)                       # d(LN out)
    ops.layernorm_bwd(st.g_mixn, st.mix, con.proj_ln_w, 1e-5, G["proj_ln_w"], G["proj_ln_b"], dx_f32=st.g_mix)
    # ---- layer mix ----
    ops.layer_mix_bwd(st.g_mix, st.x_out.view(K, N * Q, d), con.layer_weights, N, Q, st.g_x.view(K, N * Q, d), G["layer_weights"])
    dx = st.g_x                                                                      # grad wrt the layer output x' (fp32)
    for l in range(len(con.layers) - 1, -1, -1):
        L, LT, S, g = con.layers[l], conT.layers[l], st.layers[l], st.lgrads[l]
        # x' = LN(t3)
        ops.layernorm_bwd(dx, S["t3"], L["f_ln_w"], 1e-12, g["f_ln_w"], g["f_ln_b"], dx_f32=st.g_f32_a, dx_bf16=st.g_bf_a)
        # t3 = dropout(f W2^T + b2) + c : the dense branch sees the masked gradient, the residual the full one
        gd = _drop_grad(st, plan.site(l, SITE_FFN_OUT))
        ops.gemm_wgrad(st.g_bf_a, S["f"], g["f2_w"])
        ops.colsum(gd, g["f2_b"])
        ops.gemm(st.g_bf_a, LT["f2"], None, out_bf16=st.g_ff)                       # df
        ops.gelu_bwd(st.g_ff, S["h1"], st.g_h1)                                     # dh1
        # h1 = c W1^T + b1
        ops.gemm_wgrad(st.g_h1, S["c_bf"], g["f1_w"])
        ops.colsum(st.g_h1, g["f1_b"])
        ops.gemm(st.g_h1, LT["f1"], None, out_f32=st.g_f32_b, resid=st.g_f32_a)     # dc = dt3 + dh1 W1
        # c = LN(t2)
        ops.layernorm_bwd(st.g_f32_b, S["t2"], L["c_ln_w"], 1e-12, g["c_ln_w"], g["c_ln_b"], dx_f32=st.g_f32_a, dx_bf16=st.g_bf_a)
        # t2 = dropout(ao_c Wco^T + b) + a
        gd = _drop_grad(st, plan.site(l, SITE_CROSS_OUT))
        ops.gemm_wgrad(st.g_bf_a, S["ao_c"], g["c_o_w"])
        ops.colsum(gd, g["c_o_b"])
        ops.gemm(st.g_bf_a, LT["c_o"], None, out_bf16=st.g_bf_b)                    # d(ao_c)
        kc = l * 2 * d
        # value bias of the K/V projection: O = P~ (X Wv^T + 1 bv^T)  ->  d bv = sum_q (row sum of P~)_q dO_q: a 21 MB reduction
        # over dO instead of a column sum over the 490 MB dV of this layer (row sums: 1 without dropout, rr_c with it)
        ops.colsum_heads(st.g_bf_b, S["rr_c"] if plan.site(l, SITE_CROSS_PROBS) is not None else None, B, H, Q,
                         G["kv_b"][kc + d:kc + 2 * d])
        ops.attention_bwd(S["q_c"], ws.kv[:, kc:kc + d], ws.kv[:, kc + d:kc + 2 * d], S["ao_c"], st.g_bf_b, S["lse_c"],
                          st.g_qkv[:, :d], st.dkv[:, kc:kc + d], st.dkv[:, kc + d:kc + 2 * d], B, H, Q, T, scale,
                          drop=plan.site(l, SITE_CROSS_PROBS))
        dq_c = st.g_qkv[:, :d]
        # q_c = a Wq^T + bq
        ops.gemm_wgrad(dq_c, S["a_bf"], g["c_q_w"])
        ops.colsum(dq_c, g["c_q_b"])
        ops.gemm(dq_c, LT["c_q"], None, out_f32=st.g_f32_b, resid=st.g_f32_a)        # da = dt2 + dq_c Wq
        # a = LN(t1)
        ops.layernorm_bwd(st.g_f32_b, S["t1"], L["s_ln_w"], 1e-12, g["s_ln_w"], g["s_ln_b"], dx_f32=st.g_f32_a, dx_bf16=st.g_bf_a)
        # t1 = dropout(ao_s Wso^T + b) + x
        gd = _drop_grad(st, plan.site(l, SITE_SELF_OUT))
        ops.gemm_wgrad(st.g_bf_a, S["ao_s"], g["s_o_w"])
        ops.colsum(gd, g["s_o_b"])
        ops.gemm(st.g_bf_a, LT["s_o"], None, out_bf16=st.g_bf_b)                    # d(ao_s)
        ops.attention_bwd(S["qkv"][:, :d], S["qkv"][:, d:2 * d], S["qkv"][:, 2 * d:], S["ao_s"], st.g_bf_b, S["lse_s"],
                          st.g_qkv[:, :d], st.g_qkv[:, d:2 * d], st.g_qkv[:, 2 * d:], B, H, Q, Q, scale,
                          drop=plan.site(l, SITE_SELF_PROBS))
        # qkv = x Wqkv^T + b
        ops.gemm_wgrad(st.g_qkv, S["x_bf"], g["s_qkv_w"])
        ops.colsum(st.g_qkv, g["s_qkv_b"])
        ops.gemm(st.g_qkv, LT["s_qkv"], None, out_f32=st.g_x, resid=st.g_f32_a)      # dx = dt1 + dqkv Wqkv
        dx = st.g_x
        if on_ready is not None:
            a0, a1 = st.layer_slices[l]
            on_ready(st.arena[a0:a1])
    # ---- learned prompts: x0[k, n] = layer_prompts[k] for every clip n ----
    for k in range(K):
        ops.colsum(dx.view(K, N, Q * d)[k], G["prompts"].view(K, Q * d)[k])
    # ---- cross-attention K/V projections of all layers: one weight-gradient GEMM over all taps' rows ----
    ops.gemm_wgrad(st.dkv, ws.taps.view(K * N * T, d), G["kv_w"])
    # Bias gradients of those projections need no pass over the 5.9 GB of dK / dV: a key bias adds q . b_k to every score of
    # a row, which the softmax ignores, so its gradient is identically zero (the reference's autograd produces rounding noise
    # there; the arena is zero already), and the value-bias gradient was taken from dO per layer above (colsum_heads).
    if on_ready is not None:
        a0, a1 = st.global_slice
        on_ready(st.arena[a0:a1])


def _drop_grad(st: TrainState, drop):
    """Gradient entering a dropped dense branch: masks st.g_f32_a into (st.g_f32_c fp32, st.g_bf_a bf16) and returns the
    fp32 tensor for the bias gradient; st.g_f32_a keeps the residual branch's unmasked gradient."""
    if drop is None:
        return st.g_f32_a
    ops.dropout_add(st.g_f32_a, None, drop, out_f32=st.g_f32_c, out_bf16=st.g_bf_a)
    return st.g_f32_c


def connector_forward_train_native(con, dims: PathDims, ws, st: TrainState, out_f32, plan: DropPlan = NO_DROP):
    """The same schedule as connector_forward_train as ONE call (desta_connector_train_fwd)."""
    wsb = st.native_ws()
    cplan = plan.c_struct()
    ops.STATS["launches"] += 2 + 13 * dims.qf_layers + 3
    check(lib().desta_connector_train_fwd(ctypes.byref(st._c_dims), ctypes.byref(con.c_weights), ops._p(ws.taps), ws.N, ops._p(wsb),
                                          wsb.numel(), ctypes.byref(cplan), ops._p(out_f32), ops._stream()), "desta_connector_train_fwd")
    return out_f32


def connector_backward_native(con, conT: ConnectorTPack, dims: PathDims, ws, st: TrainState, grad_out, plan: DropPlan = NO_DROP,
                              on_ready=None):
    """The same schedule as connector_backward as ONE call (desta_connector_bwd); `on_ready` as there."""
    wsb = st.native_ws()
    cplan = plan.c_struct()
    cb = GRAD_READY_FN(lambda _user, first, count: on_ready(st.arena[first:first + count])) if on_ready is not None else GRAD_READY_FN()
    ops.STATS["launches"] += 6 + 26 * dims.qf_layers + len(dims.tap_layers) + 2
    check(lib().desta_connector_bwd(ctypes.byref(st._c_dims), ctypes.byref(con.c_weights), ctypes.byref(conT.c_weights), ops._p(ws.taps),
                                    ws.N, ops._p(wsb), wsb.numel(), ctypes.byref(cplan), ops._p(grad_out), ops._p(st.arena),
                                    st.arena.numel(), cb, None, ops._stream()), "desta_connector_bwd")


def param_order(connector) -> List[str]:
    return [n for n, _ in connector.named_parameters()]


def grads_by_name(st: TrainState, dims: PathDims) -> Dict[str, torch.Tensor]:
    d, Q = dims.d_model, dims.prompt_size
    G = st.grads
    out = {"layer_weights": G["layer_weights"], "proj.0.weight": G["proj_ln_w"], "proj.0.bias": G["proj_ln_b"],
           "proj.1.weight": G["proj_w"], "proj.1.bias": G["proj_b"]}
    for k in range(len(dims.tap_layers)):
        out[f"layer_prompts.{k}"] = G["prompts"][k * Q:(k + 1) * Q].view(1, Q, d)
    for l, g in enumerate(st.lgrads):
        p = f"qformer.layer.{l}."
        for i, nm in enumerate(("query", "key", "value")):
            out[p + f"attention.self.{nm}.weight"] = g["s_qkv_w"][i * d:(i + 1) * d]
            out[p + f"attention.self.{nm}.bias"] = g["s_qkv_b"][i * d:(i + 1) * d]
        out[p + "attention.output.dense.weight"], out[p + "attention.output.dense.bias"] = g["s_o_w"], g["s_o_b"]
        out[p + "attention.output.LayerNorm.weight"], out[p + "attention.output.LayerNorm.bias"] = g["s_ln_w"], g["s_ln_b"]
        out[p + "crossattention.self.query.weight"], out[p + "crossattention.self.query.bias"] = g["c_q_w"], g["c_q_b"]
        kc = l * 2 * d
        out[p + "crossattention.self.key.weight"] = G["kv_w"][kc:kc + d]
        out[p + "crossattention.self.key.bias"] = G["kv_b"][kc:kc + d]
        out[p + "crossattention.self.value.weight"] = G["kv_w"][kc + d:kc + 2 * d]
        out[p + "crossattention.self.value.bias"] = G["kv_b"][kc + d:kc + 2 * d]
        out[p + "crossattention.output.dense.weight"], out[p + "crossattention.output.dense.bias"] = g["c_o_w"], g["c_o_b"]
        out[p + "crossattention.output.LayerNorm.weight"], out[p + "crossattention.output.LayerNorm.bias"] = g["c_ln_w"], g["c_ln_b"]
        out[p + "intermediate.dense.weight"], out[p + "intermediate.dense.bias"] = g["f1_w"], g["f1_b"]
        out[p + "output.dense.weight"], out[p + "output.dense.bias"] = g["f2_w"], g["f2_b"]
        out[p + "output.LayerNorm.weight"], out[p + "output.LayerNorm.bias"] = g["f_ln_w"], g["f_ln_b"]
    return out


class _GraphToken:
    """Lives as long as the autograd node of one training forward; the workspaces it pins are not reused meanwhile."""


class _ConnectorFn(torch.autograd.Function):
    """Differentiable connector.  The activations the backward needs (taps, K/V, per-layer states, dropout plan) live in
    the perception module's Workspace / TrainState, not in autograd saved tensors, so autograd's version counters cannot
    protect them.  Two guards do instead: (1) from this forward until the end of its backward both objects are PINNED - a
    second forward before the backward (two grad-enabled forwards in one loss, an eval / generate call in between) gets
    fresh buffers instead of overwriting these; (2) both carry a generation counter that every overwriting forward bumps;
    the backward raises if it no longer matches what the forward stamped (e.g. a second backward through a retained graph
    after the buffers were released and reused) rather than returning wrong gradients."""

    @staticmethod
    def forward(ctx, perception, ws, st, *params):
        N = ws.N
        enc, con = perception._packs()
        out = torch.empty(N * perception.dims.prompt_size, perception.dims.d_llm, dtype=torch.float32, device=ws.h.device)
        plan = perception._drop_plan()
        native = engine._native_enabled()
        st.generation += 1
        with ops.device_guard(ws.h.device):
            (connector_forward_train_native if native else connector_forward_train)(con, perception.dims, ws, st, out, plan)
        ctx.native = native
        ctx.perception, ctx.N, ctx.plan = perception, N, plan
        ctx.ws, ctx.st, ctx.stamp = ws, st, (ws.generation, st.generation)
        ctx.token = _GraphToken()
        ws.pin(ctx.token)
        st.pin(ctx.token)
        return out.view(N, perception.dims.prompt_size, perception.dims.d_llm)

    @staticmethod
    def backward(ctx, grad_out):
        p, N = ctx.perception, ctx.N
        ws, st = ctx.ws, ctx.st
        if (ws.generation, st.generation) != ctx.stamp:
            raise RuntimeError(
                "desta_b200: the activations saved by this connector forward were overwritten by a later forward "
                f"(stamp {ctx.stamp}, now {(ws.generation, st.generation)}); a graph can be differentiated again only while no "
                "other forward of the same batch size has run since its first backward")
        _, con = p._packs()
        conT = p._t_pack()
        st.zero_grads()
        g = grad_out.to(torch.float32).contiguous().view(N * p.dims.prompt_size, p.dims.d_llm)
        sync = getattr(p, "grad_sync", None)       # data-parallel all-reduce overlapped with the backward (dist.GradSync)
        with ops.device_guard(ws.h.device):
            (connector_backward_native if ctx.native else connector_backward)(
                con, conT, p.dims, ws, st, g, ctx.plan, on_ready=sync.reduce if sync is not None else None)
        if sync is not None:
            sync.finish(st.arena)
        # one private copy of the whole gradient arena; the per-parameter gradients are views of it (same offsets)
        priv = st.arena.clone()
        base = st.arena.data_ptr()
        by_name = grads_by_name(st, p.dims)
        grads = []
        for name, prm in p.connector.named_parameters():
            if not prm.requires_grad:
                grads.append(None)
                continue
            v = by_name[name]
            gv = torch.as_strided(priv, v.shape, v.stride(), (v.data_ptr() - base) // 4).reshape(prm.shape)
            grads.append(gv if prm.dtype == torch.float32 else gv.to(prm.dtype))
        ws.unpin()                                 # the buffers may serve the next step (same objects, no reallocation)
        st.unpin()
        return (None, None, None, *grads)


def connector_training_forward(perception, input_features):
    """Frozen encoder forward (no graph) + differentiable connector."""
    N = input_features.size(0)
    with torch.no_grad():
        ws = perception._run_encoder(input_features)      # bumps ws.generation; never a pinned workspace
    st = perception._train_state(N)
    params = [prm for _, prm in perception.connector.named_parameters()]
    return _ConnectorFn.apply(perception, ws, st, *params)
