"""Thin torch-tensor wrappers over the C ABI (device pointers + the current CUDA stream).  No torch compute here."""
import ctypes

import numpy as np
import torch

from ._lib import Dropout, GemmEpilogue, PcmClip, ResamplePlan, SpliceDesc, check, lib

N_SAMPLES = 480000
N_FRAMES = 3000


# ---- instrumentation (bench.py / profiling only): launch counter and optional per-family CUDA-event timing ----
STATS = {"launches": 0}
_PROFILE = None     # when a list: (family, work, unit, start_event, end_event) per op


def start_profile():
    global _PROFILE
    _PROFILE = []


def stop_profile():
    """Returns {family: {"ms": total device ms, "work": total flops/bytes, "unit": ..., "calls": n}}."""
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for fam, work, unit, e0, e1, nbytes in rec:
        d = out.setdefault(fam, {"ms": 0.0, "work": 0.0, "unit": unit, "calls": 0, "bytes": 0.0})
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
        d["bytes"] += nbytes
        d["calls"] += 1
    return out


def _dev(*tensors):
    """The one CUDA device all tensor arguments of a call live on (None entries are skipped)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        assert t.is_cuda, "desta_b200 ops run on CUDA tensors only (no CPU fallback)"
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"desta_b200: tensor arguments on different devices ({dev} and {t.device})")
    return dev


class _op:
    """Accounting + DEVICE GUARD of one C-ABI call: the library launches on the CURRENT device (cudaGetDevice: kernel
    attributes, SM count, the launch itself), so the call is made with the tensors' device current and on that device's
    current stream - a model moved to cuda:1 without torch.cuda.set_device(1) must not launch on cuda:0."""

    def __init__(self, family, launches, work=0.0, unit="flop", nbytes=0.0, dev=None):
        self.family, self.launches, self.work, self.unit = family, launches, work, unit
        self.nbytes = nbytes if unit == "flop" else work      # algorithmic DRAM bytes of the launch
        self.guard = torch.cuda.device(dev) if dev is not None and dev.index is not None else None

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()
        STATS["launches"] += self.launches
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _PROFILE is not None and exc[0] is None:
            self.e1.record()
            _PROFILE.append((self.family, self.work, self.unit, self.e0, self.e1, self.nbytes))
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def device_guard(dev):
    """Guard for the stage-level native calls (engine / autograd): same rule as `_op`."""
    return torch.cuda.device(dev)


def _stream():
    # evaluated inside the `_op` guard: the current device IS the tensors' device here
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _cuda(t, dtype=None):
    assert t.is_cuda, "desta_b200 ops run on CUDA tensors only (no CPU fallback)"
    if dtype is not None:
        assert t.dtype == dtype, (t.dtype, dtype)
    return t


def gemm(a, w, bias=None, *, out_bf16=None, out_f32=None, resid=None, pos=None, scale=1.0, scale_cols=0, gelu=False,
         ln_stats_out=None, ln_in=None):
    """C = epi(a[M,K] @ w[N,K]^T).  a, w bf16 row-major (last dim contiguous); outputs written in place.
    LayerNorm hand-over (desta_gemm_epilogue): ``ln_stats_out`` fp32 [M, ceil(N/128), 2] on a residual (fp32-output) GEMM;
    ``ln_in = (stats, colsum, cols, eps)`` on the GEMM that consumes the bf16 copy with gain-folded weights."""
    _cuda(a, torch.bfloat16), _cuda(w, torch.bfloat16)
    M, K = a.shape
    N, K2 = w.shape
    assert K == K2 and a.stride(1) == 1 and w.stride(1) == 1
    ep = GemmEpilogue()
    if out_f32 is not None:
        _cuda(out_f32, torch.float32)
        assert out_f32.shape == (M, N) and out_f32.stride(1) == 1
        ep.out_f32, ep.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        _cuda(out_bf16, torch.bfloat16)
        assert out_bf16.shape == (M, N) and out_bf16.stride(1) == 1
        ep.out_bf16, ep.ld_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if bias is not None:
        _cuda(bias, torch.float32)
        assert bias.numel() == N and bias.is_contiguous()
        ep.bias = bias.data_ptr()
    if resid is not None:
        _cuda(resid, torch.float32)
        assert resid.shape == (M, N) and resid.stride(1) == 1
        ep.resid, ep.ld_resid = resid.data_ptr(), resid.stride(0)
    if pos is not None:
        _cuda(pos, torch.float32)
        assert pos.shape[1] == N and pos.is_contiguous()
        ep.pos, ep.pos_rows = pos.data_ptr(), pos.shape[0]
    ep.scale, ep.scale_cols, ep.gelu = float(scale), int(scale_cols), int(bool(gelu))
    if ln_stats_out is not None:
        _cuda(ln_stats_out, torch.float32)
        assert out_f32 is not None and ln_stats_out.is_contiguous() and ln_stats_out.numel() == M * ((N + 127) // 128) * 2
        ep.ln_stats_out = ln_stats_out.data_ptr()
    if ln_in is not None:
        stats, colsum, cols, eps = ln_in
        _cuda(stats, torch.float32), _cuda(colsum, torch.float32)
        groups = (cols + 127) // 128
        assert out_bf16 is not None and bias is not None and stats.is_contiguous() and stats.numel() == M * groups * 2
        assert colsum.numel() == N and colsum.is_contiguous()
        ep.ln_stats_in, ep.ln_groups, ep.ln_cols, ep.ln_eps, ep.ln_colsum = stats.data_ptr(), groups, int(cols), float(eps), colsum.data_ptr()
    # algorithmic bytes: A and W read once, every output written once, the residual read once
    nbytes = 2.0 * (M * K + N * K) + M * N * ((4 if out_f32 is not None else 0) + (2 if out_bf16 is not None else 0)
                                              + (4 if resid is not None else 0))
    with _op("gemm", 1, 2.0 * M * N * K, nbytes=nbytes, dev=_dev(a, w, bias, out_f32, out_bf16, resid, pos)):
        check(lib().desta_gemm_bf16(_p(a), a.stride(0), _p(w), w.stride(0), M, N, K, ctypes.byref(ep), _stream()), "desta_gemm_bf16")


def layernorm(x, gamma, beta, eps, *, out_bf16=None, out_f32=None):
    _cuda(x, torch.float32)
    M, D = x.shape
    assert x.stride(1) == 1
    nbytes = M * D * (4 + (2 if out_bf16 is not None else 0) + (4 if out_f32 is not None else 0))
    with _op("layernorm", 1, nbytes, "byte", dev=_dev(x, gamma, beta, out_bf16, out_f32)):
        check(lib().desta_layernorm(_p(x), x.stride(0), _p(gamma), _p(beta), float(eps), M, D,
                                    _p(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0,
                                    _p(out_f32), out_f32.stride(0) if out_f32 is not None else 0, _stream()), "desta_layernorm")


def make_dropout(p, seed, site, row_keep_ratio=None):
    """desta_dropout for probability p (None when p == 0): thr16 = round(p * 65536), scale = 1 / (1 - thr16 / 65536).
    row_keep_ratio (attention forward only): fp32 [batch*heads*Tq] that receives sum_k(dropped P) / sum_k(P) per query row;
    the caller keeps the tensor alive."""
    thr = int(round(float(p) * 65536.0))
    if thr <= 0:
        return None
    thr = min(thr, 65535)
    if row_keep_ratio is not None:
        _cuda(row_keep_ratio, torch.float32)
        assert row_keep_ratio.is_contiguous()
    return Dropout(thr, 1.0 / (1.0 - thr / 65536.0), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site),
                   row_keep_ratio.data_ptr() if row_keep_ratio is not None else None)


def _dp(drop):
    return ctypes.byref(drop) if drop is not None else None


def attention(q, k, v, out, batch, heads, Tq, Tk, scale, lse=None, drop=None):
    """q/k/v/out: 2-D bf16 views [batch*T, >= heads*64] (row stride = leading dim; may be slices of a fused buffer)."""
    for t in (q, k, v, out):
        _cuda(t, torch.bfloat16)
        assert t.stride(1) == 1
    assert q.shape[0] == batch * Tq and k.shape[0] == batch * Tk and v.shape[0] == batch * Tk and out.shape[0] == batch * Tq
    with _op("attention", 1, 4.0 * batch * heads * Tq * Tk * 64, dev=_dev(q, k, v, out, lse)):
        check(lib().desta_attention(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
                                    batch, heads, Tq, Tk, float(scale), _p(lse), _dp(drop), _stream()), "desta_attention")


def attention_variant(v=-1):
    """Select (v >= 0) or query the kernel variant behind the encoder's attention calls; returns the previous value."""
    return lib().desta_attention_variant(int(v))


def attention_bwd(q, k, v, o, dout, lse, dq, dk, dv, batch, heads, Tq, Tk, scale, drop=None):
    for t in (q, k, v, o, dout, dq, dk, dv):
        _cuda(t, torch.bfloat16)
        assert t.stride(1) == 1
    _cuda(lse, torch.float32)
    with _op("attention_bwd", 1, 10.0 * batch * heads * Tq * Tk * 64, dev=_dev(q, k, v, o, dout, lse, dq, dk, dv)):
        check(lib().desta_attention_bwd(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0),
                                        _p(dout), dout.stride(0), _p(lse), _p(dq), dq.stride(0), _p(dk), dk.stride(0),
                                        _p(dv), dv.stride(0), batch, heads, Tq, Tk, float(scale), _dp(drop), _stream()),
              "desta_attention_bwd")


def dropout_add(x, resid, drop, *, out_f32=None, out_bf16=None):
    """out = dropout(x) (+ resid) on an fp32 [M, N] grid (N % 8 == 0); out_f32 may alias x."""
    _cuda(x, torch.float32)
    M, N = x.shape
    assert x.stride(1) == 1 and (out_f32 is not None or out_bf16 is not None)
    if resid is not None:
        _cuda(resid, torch.float32)
        assert resid.shape == (M, N) and resid.stride(1) == 1
    nbytes = M * N * (4 + (4 if resid is not None else 0) + (4 if out_f32 is not None else 0) + (2 if out_bf16 is not None else 0))
    with _op("dropout", 1, nbytes, "byte", dev=_dev(x, resid, out_f32, out_bf16)):
        check(lib().desta_dropout_add(_p(x), x.stride(0), _p(resid), resid.stride(0) if resid is not None else 0, M, N, _dp(drop),
                                      _p(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
                                      _p(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, _stream()),
              "desta_dropout_add")


def dropout_mask(drop, rows, cols, device):
    """Test hook: the keep mask (bool [rows, cols]) the kernels generate for `drop` over a [rows, cols] grid."""
    mask = torch.empty(rows, cols, dtype=torch.uint8, device=device)
    with _op("dropout", 1, rows * cols, "byte", dev=_dev(mask)):
        check(lib().desta_dropout_mask(_dp(drop), rows, cols, _p(mask), _stream()), "desta_dropout_mask")
    return mask.bool()


def gemm_wgrad(dy, x, dw):
    """dw[Nout,Kin] (fp32, +=) from dy[M,Nout], x[M,Kin] (bf16, rows = reduction dim)."""
    _cuda(dy, torch.bfloat16), _cuda(x, torch.bfloat16), _cuda(dw, torch.float32)
    M, Nout = dy.shape
    M2, Kin = x.shape
    assert M == M2 and dw.shape == (Nout, Kin) and dy.stride(1) == 1 and x.stride(1) == 1 and dw.stride(1) == 1
    with _op("gemm_wgrad", 1, 2.0 * M * Nout * Kin, nbytes=2.0 * M * (Nout + Kin) + 8.0 * Nout * Kin, dev=_dev(dy, x, dw)):
        check(lib().desta_gemm_bf16_wgrad(_p(dy), dy.stride(0), _p(x), x.stride(0), M, Nout, Kin, _p(dw), dw.stride(0), _stream()),
              "desta_gemm_bf16_wgrad")


def layernorm_bwd(dy, x, gamma, eps, dgamma, dbeta, *, dx_f32=None, dx_bf16=None):
    _cuda(dy, torch.float32), _cuda(x, torch.float32)
    M, D = x.shape
    with _op("layernorm_bwd", 1, M * D * 12, "byte", dev=_dev(dy, x, gamma, dgamma, dbeta, dx_f32, dx_bf16)):
        check(lib().desta_layernorm_bwd(_p(dy), dy.stride(0), _p(x), x.stride(0), _p(gamma), float(eps), M, D,
                                        _p(dx_f32), dx_f32.stride(0) if dx_f32 is not None else 0,
                                        _p(dx_bf16), dx_bf16.stride(0) if dx_bf16 is not None else 0,
                                        _p(dgamma), _p(dbeta), _stream()), "desta_layernorm_bwd")


def gelu_fwd(h, y):
    _cuda(h, torch.bfloat16), _cuda(y, torch.bfloat16)
    assert h.is_contiguous() and y.is_contiguous()
    with _op("gelu", 1, h.numel() * 4, "byte", dev=_dev(h, y)):
        check(lib().desta_gelu_fwd(_p(h), _p(y), h.numel(), _stream()), "desta_gelu_fwd")


def gelu_bwd(dy, h, dh):
    for t in (dy, h, dh):
        _cuda(t, torch.bfloat16)
        assert t.is_contiguous()
    with _op("gelu", 1, h.numel() * 6, "byte", dev=_dev(dy, h, dh)):
        check(lib().desta_gelu_bwd(_p(dy), _p(h), _p(dh), h.numel(), _stream()), "desta_gelu_bwd")


def colsum(x, out):
    """out[j] += sum_m x[m, j]; x fp32 or bf16 2-D (row stride = leading dim)."""
    assert x.is_cuda and x.stride(1) == 1 and x.dtype in (torch.float32, torch.bfloat16)
    _cuda(out, torch.float32)
    M, N = x.shape
    assert out.numel() == N
    with _op("colsum", 1, x.numel() * x.element_size(), "byte", dev=_dev(x, out)):
        check(lib().desta_colsum(_p(x), 0 if x.dtype == torch.float32 else 1, x.stride(0), M, N, _p(out), _stream()), "desta_colsum")


def colsum_heads(x, w, batch, heads, Tq, out):
    """out[h*64+c] += sum_{b,t} w[b,h,t] * x[b*Tq+t, h*64+c]  (x bf16 [batch*Tq, >= heads*64]; w fp32 [batch*heads*Tq] or None = 1):
    value-bias gradient of an attention from dO."""
    _cuda(x, torch.bfloat16)
    _cuda(out, torch.float32)
    assert x.stride(1) == 1 and x.shape[0] == batch * Tq and out.numel() == heads * 64 and out.is_contiguous()
    if w is not None:
        _cuda(w, torch.float32)
        assert w.is_contiguous() and w.numel() == batch * heads * Tq
    with _op("colsum", 1, batch * Tq * heads * 64 * 2, "byte", dev=_dev(x, w, out)):
        check(lib().desta_colsum_heads(_p(x), x.stride(0), _p(w), batch, heads, Tq, _p(out), _stream()), "desta_colsum_heads")


def layer_mix_bwd(dm, x, layer_weights, N, Q, dx, dlw):
    taps, rows, D = x.shape
    for t in (dm, x, layer_weights, dx, dlw):
        _cuda(t, torch.float32)
        assert t.is_contiguous()
    with _op("layer_mix", 1, (2 * x.numel() + dm.numel()) * 4, "byte", dev=_dev(dm, x, layer_weights, dx, dlw)):
        check(lib().desta_layer_mix_bwd(_p(dm), _p(x), _p(layer_weights), taps, N, Q, D, _p(dx), _p(dlw), _stream()), "desta_layer_mix_bwd")


def cast_bf16(x, y):
    _cuda(x, torch.float32), _cuda(y, torch.bfloat16)
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    with _op("cast", 1, x.numel() * 6, "byte", dev=_dev(x, y)):
        check(lib().desta_cast_f32_bf16(_p(x), _p(y), x.numel(), _stream()), "desta_cast_f32_bf16")


def splice_descs_device(descs, device):
    n = len(descs)
    host = (SpliceDesc * max(n, 1))()
    for i, (p_, s_, t_, o_) in enumerate(descs):
        host[i] = SpliceDesc(int(p_), int(s_), int(t_), int(o_))
    return host, torch.frombuffer(bytearray(bytes(host)), dtype=torch.int32).to(device)


def splice_bwd(grad_out, descs, Q, dfeats):
    """dfeats [n,Q,d] fp32 <- rows of grad_out [B,S,d] bf16 under each clip's audio span."""
    _cuda(grad_out, torch.bfloat16), _cuda(dfeats, torch.float32)
    B, S, d = grad_out.shape
    _, dev = splice_descs_device(descs, grad_out.device)
    with _op("splice", 1, dfeats.numel() * 6, "byte", dev=_dev(grad_out, dfeats)):
        check(lib().desta_splice_bwd(_p(grad_out.contiguous()), _p(dev), len(descs), Q, S, d, _p(dfeats), _stream()), "desta_splice_bwd")


def im2col_k3(x, N, T_in, C, stride, cols, mel_floor=None):
    """mel_floor: the clip-max words of ``logmel(..., want_f32=False)`` - the deferred log-mel floor is applied on read."""
    _cuda(x, torch.bfloat16), _cuda(cols, torch.bfloat16)
    assert x.is_contiguous() and cols.is_contiguous() and cols.numel() == N * (T_in // stride) * 3 * C
    if mel_floor is not None:
        _cuda(mel_floor, torch.float32)
        assert mel_floor.is_contiguous() and mel_floor.numel() >= N
    with _op("im2col", 1, cols.numel() * 4, "byte", dev=_dev(x, cols, mel_floor)):
        check(lib().desta_im2col_k3(_p(x), N, T_in, C, stride, _p(cols), _p(mel_floor), _stream()), "desta_im2col_k3")


def layer_mix(x, layer_weights, N, Q, out):
    _cuda(x, torch.float32), _cuda(layer_weights, torch.float32), _cuda(out, torch.float32)
    taps, rows, D = x.shape
    assert rows == N * Q and x.is_contiguous() and out.is_contiguous() and layer_weights.is_contiguous()
    with _op("layer_mix", 1, (x.numel() + out.numel()) * 4, "byte", dev=_dev(x, layer_weights, out)):
        check(lib().desta_layer_mix(_p(x), _p(layer_weights), taps, N, Q, D, _p(out), _stream()), "desta_layer_mix")


_plans = {}


def logmel_plan(n_mels, device):
    key = (n_mels, str(device))
    if key not in _plans:
        nbytes = lib().desta_logmel_plan_bytes(n_mels)
        if nbytes == 0:
            raise ValueError(f"unsupported n_mels {n_mels}")
        host = np.zeros(nbytes // 4, dtype=np.int32)
        check(lib().desta_logmel_plan_build(n_mels, host.ctypes.data_as(ctypes.c_void_p), nbytes), "desta_logmel_plan_build")
        _plans[key] = torch.from_numpy(host).to(device)
    return _plans[key]


def logmel(wave, n_mels, *, n_valid=None, want_bf16_tm=False, want_f32=True):
    """wave [N,480000] f32 CUDA -> (mel f32 [N,n_mels,3000], optional bf16 time-major [N,3000,n_mels]).
    ``want_f32=False`` (product path, one launch): returns (None, mel_tm, clip_max) where mel_tm lacks the per-clip floor
    and clip_max is handed to ``im2col_k3(mel_floor=...)`` / ``engine.encoder_forward(mel_floor=...)``."""
    _cuda(wave, torch.float32)
    assert wave.dim() == 2 and wave.shape[1] == 480000 and wave.is_contiguous()
    assert want_f32 or want_bf16_tm
    N = wave.shape[0]
    plan = logmel_plan(n_mels, wave.device)
    out = torch.empty(N, n_mels, 3000, dtype=torch.float32, device=wave.device) if want_f32 else None
    out_tm = torch.empty(N, 3000, n_mels, dtype=torch.bfloat16, device=wave.device) if want_bf16_tm else None
    cmax = torch.empty(max(N, 1), dtype=torch.float32, device=wave.device)
    if n_valid is not None:
        _cuda(n_valid, torch.int32)
    # algorithmic bytes: waveform in + the log-mel tensor the call hands on (fp32 3.456 MB / clip, bf16 2.688)
    nbytes = N * (480000 * 4 + n_mels * 3000 * (4 if want_f32 else 2))
    STATS["logmel_calls"] = STATS.get("logmel_calls", 0) + 1
    with _op("logmel", 2 if want_f32 else 1, nbytes, "byte", dev=_dev(wave, n_valid, plan)):
        check(lib().desta_logmel(_p(wave), _p(n_valid), N, n_mels, _p(plan), _p(out), _p(out_tm), _p(cmax), _stream()), "desta_logmel")
    if not want_f32:
        return None, out_tm, cmax
    return out, out_tm


_resample_plans = {}


def resample_plan(sr_in, sr_out, device):
    """(ResamplePlan, device coefficient table [L, P] f32) for one rate pair, built once per device."""
    key = (int(sr_in), int(sr_out), str(device))
    if key not in _resample_plans:
        plan = ResamplePlan()
        nbytes = lib().desta_resample_plan_bytes(int(sr_in), int(sr_out), ctypes.byref(plan))
        if nbytes == 0:
            raise ValueError(f"unsupported sample rates {sr_in} -> {sr_out}")
        if nbytes > (256 << 20):
            raise ValueError(f"sample rates {sr_in} -> {sr_out} reduce to {plan.L}/{plan.M}: coefficient table of {nbytes} bytes")
        host = np.zeros(nbytes // 4, dtype=np.float32)
        check(lib().desta_resample_plan_build(int(sr_in), int(sr_out), host.ctypes.data_as(ctypes.c_void_p), nbytes),
              "desta_resample_plan_build")
        _resample_plans[key] = (plan, torch.from_numpy(host).to(device))
    return _resample_plans[key]


def decode_resample(pcm, clips, fmt, channels, sr_in, sr_out, out, general_kernel=False):
    """pcm: uint8 CUDA tensor holding the raw interleaved frames of every clip; clips: list of (byte offset into pcm,
    n_frames, out_row).  Writes out[out_row, :] (f32, zero-padded / truncated to out.shape[1]); returns the valid lengths.
    general_kernel=True forces the any-ratio kernel where the tiled one would run (test hook)."""
    _cuda(pcm, torch.uint8), _cuda(out, torch.float32)
    assert out.dim() == 2 and out.stride(1) == 1
    plan, taps = resample_plan(sr_in, sr_out, out.device)
    if general_kernel:
        plan = ResamplePlan(plan.L, plan.M, plan.P, 1)
    n = len(clips)
    host = (PcmClip * max(n, 1))()
    lengths = []
    for i, (off, n_frames, row) in enumerate(clips):
        if not 0 <= row < out.shape[0]:
            raise ValueError(f"clip {i}: output row {row} outside [0, {out.shape[0]})")
        n_out = -(-int(n_frames) * plan.L // plan.M)
        host[i] = PcmClip(pcm.data_ptr() + int(off), int(n_frames), n_out, int(row))
        lengths.append(min(n_out, out.shape[1]))
    if n == 0:
        return lengths
    dev = torch.frombuffer(bytearray(bytes(host)), dtype=torch.int64).to(out.device)
    frames = sum(c[1] for c in clips)
    nbytes = frames * channels * (1, 2, 3, 4, 4, 8)[fmt] + n * out.shape[1] * 4
    with _op("decode_resample", 1, nbytes, "byte", dev=_dev(pcm, out, taps)):
        check(lib().desta_decode_resample(_p(dev), n, int(fmt), int(channels), ctypes.byref(plan), _p(taps), _p(out), out.stride(0),
                                          out.shape[1], _stream()), "desta_decode_resample")
    return lengths


_err_words = {}


def _err_word(device):
    """int64[2] in pinned host memory that the splice kernel writes through UVA (count of rows with a token id outside the
    embedding table, one offending id).  The host can read it at any time without touching the stream."""
    key = str(device)
    if key not in _err_words:
        _err_words[key] = torch.zeros(2, dtype=torch.int64).pin_memory()
    return _err_words[key]


def check_token_ids(device=None, sync=False):
    """Raise IndexError if a splice on `device` (all devices if None) saw a token id outside the embedding table - the
    reference's `embed_tokens(...)` raises there (modeling_desta25.py:981, :1009; IndexError on the CPU, a device-side
    assert on CUDA).  Like CUDA's own asynchronous errors the report surfaces at the next call on that device (every
    `splice` checks first), or immediately with sync=True."""
    for key, w in list(_err_words.items()):
        if device is not None and key != str(torch.device(device)):
            continue
        if sync:
            torch.cuda.synchronize(torch.device(key))
        n = int(w[0])
        if n:
            bad = int(w[1])
            w.zero_()
            raise IndexError(f"index out of range in self: {n} row(s) of the <|AUDIO|> splice on {key} had a token id outside the "
                             f"embedding table (e.g. {bad})")


def row_hash(x):
    """64-bit content hash of every row of a contiguous CUDA tensor [n, ...] (any 4-byte-multiple row size) -> int64 [n]."""
    assert x.is_cuda and x.is_contiguous() and x.dim() >= 2
    n = x.shape[0]
    row_bytes = x[0].numel() * x.element_size()
    out = torch.empty(max(n, 1), dtype=torch.int64, device=x.device)
    with _op("row_hash", 1, n * row_bytes, "byte", dev=x.device):
        check(lib().desta_row_hash(_p(x), n, row_bytes, row_bytes, _p(out), _stream()), "desta_row_hash")
    return out[:n]


def splice(input_ids, table, descs, trans_ids, feats, Q, out=None):
    """descs: list of (prompt, start, n_trans, trans_off); trans_ids: 1-D int64 CUDA tensor (concatenated) or None."""
    _cuda(input_ids, torch.int64), _cuda(table, torch.bfloat16)
    check_token_ids(input_ids.device)                  # report of an earlier splice on this device, if any
    B, S = input_ids.shape
    V, d = table.shape
    n = len(descs)
    host = (SpliceDesc * max(n, 1))()
    for i, (p_, s_, t_, o_) in enumerate(descs):
        host[i] = SpliceDesc(int(p_), int(s_), int(t_), int(o_))
    dev = torch.frombuffer(bytearray(bytes(host)), dtype=torch.int32).to(input_ids.device, non_blocking=False)
    if out is None:
        out = torch.empty(B, S, d, dtype=torch.bfloat16, device=input_ids.device)
    feat_dtype = 0
    if feats is not None:
        assert feats.is_contiguous() and feats.shape[-1] == d
        feat_dtype = 0 if feats.dtype == torch.float32 else 1
        assert feats.dtype in (torch.float32, torch.bfloat16)
    fb = 4 if feat_dtype == 0 else 2
    n_audio_rows = n * Q
    nbytes = B * S * 8 + (B * S - n_audio_rows) * d * 2 + n_audio_rows * d * fb + B * S * d * 2
    err = _err_word(input_ids.device)
    with _op("splice", 1, nbytes, "byte", dev=_dev(input_ids, table, trans_ids, feats, out)):
        check(lib().desta_splice(_p(input_ids.contiguous()), _p(table), V, _p(dev), host, n, _p(trans_ids), _p(feats), feat_dtype,
                                 Q, _p(out), B, S, d, ctypes.c_void_p(err.data_ptr()), _stream()), "desta_splice")
    return out
