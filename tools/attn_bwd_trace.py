"""Cycle accounting of attn_bwd_t64_kernel (debug build: python tools/attn_trace.py build; then run this on the GPU box)."""
import ctypes, os, sys, torch
ROOT = os.getcwd()
L = ctypes.CDLL(os.path.join(ROOT, "tools/_bin/libdesta_trace.so"))
B, H, Tq, Tk = 128, 20, 64, 1536
d = H * 64
vp = ctypes.c_void_p
q = (torch.randn(B * Tq, d, device="cuda") * 0.5).to(torch.bfloat16)
kv = (torch.randn(B * Tk, 2 * d, device="cuda") * 0.5).to(torch.bfloat16)
o, do = torch.empty_like(q), torch.randn_like(q)
lse = torch.empty(B * H * Tq, device="cuda")
st = vp(torch.cuda.current_stream().cuda_stream)
assert L.desta_attention(vp(q.data_ptr()), d, vp(kv.data_ptr()), 2 * d, vp(kv.data_ptr() + 2 * d), 2 * d, vp(o.data_ptr()), d, B, H, Tq, Tk, ctypes.c_float(0.125), vp(lse.data_ptr()), None, st) == 0
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
args = [vp(q.data_ptr()), d, vp(kv.data_ptr()), 2 * d, vp(kv.data_ptr() + 2 * d), 2 * d, vp(o.data_ptr()), d, vp(do.data_ptr()), d, vp(lse.data_ptr()),
        vp(dq.data_ptr()), d, vp(dkv.data_ptr()), 2 * d, vp(dkv.data_ptr() + 2 * d), 2 * d, B, H, Tq, Tk, ctypes.c_float(0.125), None, st]
for _ in range(3): assert L.desta_attention_bwd(*args) == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8)()
L.desta_attention_bwd_trace(buf, 1)
assert L.desta_attention_bwd(*args) == 0
torch.cuda.synchronize()
L.desta_attention_bwd_trace(buf, 0)
n = B * H; nb = Tk // 128
names = ["MMA: wait kv_full", "MMA: wait pds_ready", "MMA: wait dkv_empty", "compute w2: wait sdp_full", "compute w2: wait dkv_full", "MMA warp lifetime", "compute w2 lifetime"]
for i, nm in enumerate(names):
    print(f"{nm:28s} {buf[i] / n:10.0f} cycles per CTA   {buf[i] / n / nb:8.0f} per block")
