"""Kernel micro-benchmarks on one B200 (CUDA events, L2-exceeding working sets). Usage: python tools/kbench.py [gemm] [attn] [mel] [splice]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from desta_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_gemm():
    M = 96000
    for name, N, K, kw in [("qkv", 3840, 1280, dict(scale=0.125, scale_cols=1280)), ("qkv+ln_in", 3840, 1280, dict(scale=0.125, scale_cols=1280, ln_in=True)),
                           ("out_proj+res", 1280, 1280, dict(resid=True)), ("out_proj+res+ln", 1280, 1280, dict(resid=True, ln_out=True)),
                           ("fc1+gelu", 5120, 1280, dict(gelu=True)), ("fc1+gelu+ln_in", 5120, 1280, dict(gelu=True, ln_in=True)),
                           ("fc2+res", 1280, 5120, dict(resid=True)), ("fc2+res+ln", 1280, 5120, dict(resid=True, ln_out=True)),
                           ("kvproj", 15360, 1280, {}), ("conv2", 1280, 3840, dict(gelu=True)), ("qf_small", 1280, 1280, dict(M=16384))]:
        m = kw.pop("M", M)
        a = torch.randn(m, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
        b = torch.randn(N, device="cuda")
        ln_out, ln_in = kw.pop("ln_out", False), kw.pop("ln_in", False)
        if kw.pop("resid", False):
            h = torch.randn(m, N, device="cuda")
            if ln_out:
                hb = torch.empty(m, N, device="cuda", dtype=torch.bfloat16)
                st = torch.empty(m, (N + 127) // 128, 2, device="cuda")
                fn = lambda: ops.gemm(a, w, b, out_f32=h, resid=h, out_bf16=hb, ln_stats_out=st)  # noqa: E731
            else:
                fn = lambda: ops.gemm(a, w, b, out_f32=h, resid=h)  # noqa: E731
        else:
            o = torch.empty(m, N, device="cuda", dtype=torch.bfloat16)
            if ln_in:
                st = torch.rand(m, (K + 127) // 128, 2, device="cuda") + 1.0
                cs = torch.randn(N, device="cuda")
                fn = lambda: ops.gemm(a, w, b, out_bf16=o, ln_in=(st, cs, K, 1e-5), **kw)  # noqa: E731
            else:
                fn = lambda: ops.gemm(a, w, b, out_bf16=o, **kw)  # noqa: E731
        ms = timeit(fn)
        print(f"gemm {name:16s} M={m} N={N} K={K}: {ms:.3f} ms  {2.0 * m * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
        del a, w


def bench_attn():
    for name, B, H, Tq, Tk, scale in [("encoder", 64, 20, 1500, 1500, 1.0), ("enc_unit", 64, 20, 1500, 1500, 0.6931471805599453),
                                      ("qf_cross", 256, 20, 64, 1500, 0.125), ("qf_self", 256, 20, 64, 64, 0.125)]:
        d = H * 64
        qkv = (torch.randn(B * max(Tq, Tk), 3 * d, device="cuda") * 0.5).to(torch.bfloat16)   # logits ~ N(0, 2): few rescales, like the model
        out = torch.empty(B * Tq, d, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.attention(qkv[:B * Tq, :d], qkv[:B * Tk, d:2 * d], qkv[:B * Tk, 2 * d:], out, B, H, Tq, Tk, scale)  # noqa: E731
        ms = timeit(fn)
        print(f"attn {name:9s} B={B} H={H} Tq={Tq} Tk={Tk}: {ms:.3f} ms  {4.0 * B * H * Tq * Tk * 64 / ms / 1e9:.0f} TFLOP/s", flush=True)


def bench_attn_bwd():
    for name, B, H, Tq, Tk, scale in [("qf_cross", 128, 20, 64, 1500, 0.125), ("qf_self", 128, 20, 64, 64, 0.125)]:
        d = H * 64
        q = (torch.randn(B * Tq, d, device="cuda") * 0.5).to(torch.bfloat16)
        kv = (torch.randn(B * Tk, 2 * d, device="cuda") * 0.5).to(torch.bfloat16)
        o, do = torch.empty_like(q), torch.randn_like(q)
        lse = torch.empty(B * H * Tq, device="cuda")
        ops.attention(q, kv[:, :d], kv[:, d:], o, B, H, Tq, Tk, scale, lse=lse)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        fn = lambda: ops.attention_bwd(q, kv[:, :d], kv[:, d:], o, do, lse, dq, dkv[:, :d], dkv[:, d:], B, H, Tq, Tk, scale)  # noqa: E731
        ms = timeit(fn)
        print(f"attn_bwd {name:9s} B={B} H={H} Tq={Tq} Tk={Tk}: {ms:.3f} ms  {10.0 * B * H * Tq * Tk * 64 / ms / 1e9:.0f} TFLOP/s", flush=True)


def bench_train_ew():
    """HBM-bound kernels of the connector backward at the training shape (32 clips: R = 4 * 32 * 64 = 8192 rows of 1280)."""
    R, D, K, N, Q = 8192, 1280, 4, 32, 64
    dy, x = torch.randn(R, D, device="cuda"), torch.randn(R, D, device="cuda")
    gamma = torch.rand(D, device="cuda") + 0.5
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dx, dxb = torch.empty(R, D, device="cuda"), torch.empty(R, D, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.layernorm_bwd(dy, x, gamma, 1e-12, dg, db, dx_f32=dx, dx_bf16=dxb), iters=20)
    print(f"layernorm_bwd R={R} D={D}: {ms:.4f} ms  {R * D * 14 / ms / 1e6:.0f} GB/s (dy, x read; dx fp32 + bf16 written)", flush=True)
    dm, xs = torch.randn(N * Q, D, device="cuda"), torch.randn(K, N * Q, D, device="cuda")
    lw = torch.randn(Q, K, device="cuda")
    gx, dlw = torch.empty(K, N * Q, D, device="cuda"), torch.zeros(Q, K, device="cuda")
    ms = timeit(lambda: ops.layer_mix_bwd(dm, xs, lw, N, Q, gx, dlw), iters=20)
    print(f"layer_mix_bwd K={K} N={N} Q={Q} D={D}: {ms:.4f} ms  {N * Q * D * 4 * (1 + 2 * K) / ms / 1e6:.0f} GB/s", flush=True)


def bench_mel():
    for N in (64, 256):
        wave = torch.clamp(0.1 * torch.randn(N, 480000, device="cuda"), -1, 1)
        ms = timeit(lambda: ops.logmel(wave, 128, want_bf16_tm=True))
        print(f"logmel N={N} fp32 + bf16 outputs (2 launches): {ms:.3f} ms  {N * (480000 * 4 + 128 * 3000 * 4) / ms / 1e6:.0f} GB/s (algorithmic in + fp32 out)", flush=True)
        ms = timeit(lambda: ops.logmel(wave, 128, want_bf16_tm=True, want_f32=False))
        print(f"logmel N={N} product path (1 launch, bf16 out): {ms:.3f} ms  {N * (480000 * 4 + 128 * 3000 * 2) / ms / 1e6:.0f} GB/s (algorithmic in + bf16 out)", flush=True)


def bench_splice():
    for B, S, d, V in ((64, 128, 4096, 128256), (256, 512, 4096, 128256)):
        table = (torch.randn(V, d, device="cuda") * 0.02).to(torch.bfloat16)
        ids = torch.randint(0, V, (B, S), device="cuda")
        feats = torch.randn(B, 64, d, device="cuda")
        descs = [(i, 20, 0, 0) for i in range(B)]
        out = torch.empty(B, S, d, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.splice(ids, table, descs, None, feats, 64, out=out))
        nbytes = B * S * 8 + (B * S - B * 64) * d * 2 + B * 64 * d * 4 + B * S * d * 2
        print(f"splice B={B} S={S} d={d}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s", flush=True)


def bench_resample():
    import numpy as np
    N = 64
    for sr, ch in ((22050, 1), (24000, 1), (32000, 1), (44100, 2), (44100, 1), (48000, 2), (8000, 1), (16000, 2)):
        n = sr * 30
        pcm = torch.randint(-20000, 20000, (N, n * ch), dtype=torch.int16, device="cuda").view(torch.uint8).reshape(-1)
        clips = [(i * n * ch * 2, n, i) for i in range(N)]
        out = torch.empty(N, 480000, device="cuda")
        ms = timeit(lambda: ops.decode_resample(pcm, clips, 1, ch, sr, 16000, out))
        plan, _ = ops.resample_plan(sr, 16000, out.device)
        nbytes = N * (n * ch * 2 + 480000 * 4)
        print(f"decode_resample {sr} Hz x{ch}ch s16 -> 16 kHz, N={N}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s algorithmic, "
              f"{2.0 * N * 480000 * plan.P / ms / 1e9:.2f} TFLOP/s fp32 (P={plan.P})", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "mel", "splice"]
    for w in which:
        {"gemm": bench_gemm, "attn": bench_attn, "attn_bwd": bench_attn_bwd, "mel": bench_mel, "train_ew": bench_train_ew, "splice": bench_splice, "resample": bench_resample}[w]()
