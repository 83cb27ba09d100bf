"""One launch per flagship kernel shape (after one warm-up launch each) for `ncu --set full`:
   python tools/prof_shapes.py        # plain
   ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/prof_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from desta_b200 import ops  # noqa: E402

M = 96000
g = torch.Generator(device="cuda").manual_seed(0)


def r(*s, scale=1.0, dt=torch.bfloat16):
    return (torch.randn(*s, device="cuda", generator=g) * scale).to(dt)


def gemm(name, N, K, m=M, resid=False, ln_out=False, ln_in=False, **kw):
    a, w, b = r(m, K), r(N, K, scale=0.02), r(N, dt=torch.float32)
    for _ in range(2):
        if resid:
            h = r(m, N, dt=torch.float32)
            if ln_out:       # round 2: residual GEMM + bf16 copy of the rows + per-row partial sums (LayerNorm hand-over)
                hb = torch.empty(m, N, device="cuda", dtype=torch.bfloat16)
                st = torch.empty(m, (N + 127) // 128, 2, device="cuda")
                ops.gemm(a, w, b, out_f32=h, resid=h, out_bf16=hb, ln_stats_out=st)
            else:
                ops.gemm(a, w, b, out_f32=h, resid=h)
        else:
            o = torch.empty(m, N, device="cuda", dtype=torch.bfloat16)
            if ln_in:        # round 2: LayerNorm applied in the epilogue of the consuming GEMM
                st = torch.rand(m, (K + 127) // 128, 2, device="cuda") + 1.0
                ops.gemm(a, w, b, out_bf16=o, ln_in=(st, r(N, dt=torch.float32), K, 1e-5), **kw)
            else:
                ops.gemm(a, w, b, out_bf16=o, **kw)
    torch.cuda.synchronize()
    print("ran", name, flush=True)


gemm("qkv+ln_in", 3840, 1280, ln_in=True, scale=0.125, scale_cols=1280)
gemm("out_proj+res+ln", 1280, 1280, resid=True, ln_out=True)
gemm("fc1+gelu+ln_in", 5120, 1280, ln_in=True, gelu=True)
gemm("fc2+res+ln", 1280, 5120, resid=True, ln_out=True)
gemm("kvproj", 15360, 1280)
for name, B, H, Tq, Tk, scale in [("attn encoder (UNIT mode)", 64, 20, 1500, 1500, 0.6931471805599453), ("attn qf_cross", 256, 20, 64, 1500, 0.125)]:
    d = H * 64
    qkv = r(B * max(Tq, Tk), 3 * d, scale=0.5)
    out = torch.empty(B * Tq, d, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.attention(qkv[:B * Tq, :d], qkv[:B * Tk, d:2 * d], qkv[:B * Tk, 2 * d:], out, B, H, Tq, Tk, scale)
    torch.cuda.synchronize()
    print("ran", name, flush=True)
x = r(M, 1280, dt=torch.float32)
y = torch.empty(M, 1280, device="cuda", dtype=torch.bfloat16)
gam, bet = r(1280, dt=torch.float32), r(1280, dt=torch.float32)
for _ in range(2):
    ops.layernorm(x, gam, bet, 1e-5, out_bf16=y)
wave = torch.clamp(0.1 * torch.randn(64, 480000, device="cuda", generator=g), -1, 1)
for _ in range(2):
    ops.logmel(wave, 128, want_bf16_tm=True, want_f32=False)
torch.cuda.synchronize()
print("done")
