"""Cycle accounting of the attention kernel's warp roles (debug build with -DDESTA_ATTN_TRACE).
Build here:  python tools/attn_trace.py build      Run on the GPU box:  python tools/attn_trace.py"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "desta2.5-audio_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_bin", "libdesta_trace.so")
SOURCES = ["api.cu", "schedule.cu", "gemm.cu", "gemm2.cu", "attention.cu", "attention_bwd.cu", "elementwise.cu", "elementwise_bwd.cu", "logmel.cu", "splice.cu", "resample.cu"]


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-DDESTA_ATTN_TRACE", "-I", os.path.join(ROOT, "include"), "-I", CSRC, *[os.path.join(CSRC, s) for s in SOURCES],
           "-o", OUT, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    subprocess.run(cmd, check=True)
    print(OUT)


def run():
    import torch
    L = ctypes.CDLL(OUT)
    B, H, T = 64, 20, 1500
    d = H * 64
    vp = ctypes.c_void_p
    names = ["MMA warp: wait kv_full", "MMA warp: wait p_ready", "softmax warp 2: wait s_full", "MMA warp: total loop", "softmax warp 2: total loop",
             "softmax warp 4 (SMSP 0, with TMA warp): wait s_full", "softmax warp 5 (SMSP 1, with MMA warp): wait s_full",
             "softmax warp 3 (SMSP 3): wait s_full"]
    ctas, blocks = 12 * H * B, 24
    # layouts: (label, row pitch in elements, junk bytes allocated before the buffer to shift its address)
    LN2 = 0.6931471805599453
    for label, pitch, junk, scale in [("general path (scale 1.0)", 3 * d, 0, 1.0), ("UNIT path (scale ln 2)", 3 * d, 0, LN2)]:
        pad = torch.empty(junk, dtype=torch.uint8, device="cuda") if junk else None
        buf_t = torch.empty(B * T, pitch, device="cuda", dtype=torch.bfloat16)
        buf_t.normal_(0, 0.5)
        out = torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
        args = [vp(buf_t.data_ptr()), pitch, vp(buf_t.data_ptr() + 2 * d), pitch, vp(buf_t.data_ptr() + 4 * d), pitch, vp(out.data_ptr()), d,
                B, H, T, T, ctypes.c_float(scale), None, None, vp(torch.cuda.current_stream().cuda_stream)]
        for _ in range(3):
            assert L.desta_attention(*args) == 0
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 8)()
        L.desta_attention_trace(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            assert L.desta_attention(*args) == 0
        e1.record()
        torch.cuda.synchronize()
        L.desta_attention_trace(buf, 0)
        print(f"{label}: {e0.elapsed_time(e1) / 5:.3f} ms  (qkv at 0x{buf_t.data_ptr():x})  per key block per CTA: "
              f"kv_full wait {buf[0] / ctas / blocks / 5:.0f}, p_ready wait {buf[1] / ctas / blocks / 5:.0f}, MMA loop {buf[3] / ctas / blocks / 5:.0f}, "
              f"softmax s_full wait {buf[2] / ctas / blocks / 5:.0f} (w4 {buf[5] / ctas / blocks / 5:.0f}, w5 {buf[6] / ctas / blocks / 5:.0f}, w3 {buf[7] / ctas / blocks / 5:.0f}), "
              f"softmax loop {buf[4] / ctas / blocks / 5:.0f}", flush=True)
        tl = (ctypes.c_longlong * (16 * 32))()
        if os.environ.get("DESTA_ATTN_UNIT8", "1") != "0" and scale != 1.0:
            L.desta_attention_timeline(tl)
            ev = [[tl[e * 32 + j] for j in range(24)] for e in range(16)]
            t0 = ev[13][0]
            rel = lambda v: (v - t0) if v > 0 else -1
            print("   unit8 timeline of CTA (5,7,30), cycles from kernel entry: init done %d, Q landed %d, S_0 seen %d, shift published %d, "
                  "loop end %d, last PV done %d, stores issued %d" % tuple(rel(ev[13][k]) for k in range(1, 8)))
            print("    j | MMA: p_ready seen, PV issued, QK_j issued (s_free seen, K_j seen) | step start w2..w5 (cols 0-31) | w6..w9 (cols 32-63) | w2: got S_j, announced P_j")
            for j in range(24):
                print(f"   {j:2d} | {rel(ev[0][j]):6d} {rel(ev[1][j]):6d} {rel(ev[2][j]):6d} ({rel(ev[14][j]):6d} {rel(ev[15][j]):6d}) | " + " ".join(f"{rel(ev[3 + k][j]):6d}" for k in range(4)) + " | "
                      + " ".join(f"{rel(ev[7 + k][j]):6d}" for k in range(4)) + f" | {rel(ev[11][j]):6d} {rel(ev[12][j]):6d}")
        elif hasattr(L, "desta_attention_timeline") and L.desta_attention_timeline(tl) == 0:
            ev = [[tl[e * 32 + j] for j in range(24)] for e in range(12)]
            t0 = min(v for row in ev for v in row if v > 0)
            rel = lambda v: (v - t0) if v > 0 else -1
            print("   timeline of CTA (5,7,30), cycles from its first event; per block j:")
            print("    j | MMA: p_ready seen, PV issued, QK_j issued | softmax step start w2 w3 w4 w5 | got S_j w2 w3 w4 w5")
            for j in range(24):
                print(f"   {j:2d} | {rel(ev[0][j]):6d} {rel(ev[1][j]):6d} {rel(ev[2][j]):6d} | {rel(ev[3][j]):6d} {rel(ev[4][j]):6d} {rel(ev[5][j]):6d} {rel(ev[6][j]):6d} |"
                      f" {rel(ev[7][j]):6d} {rel(ev[8][j]):6d} {rel(ev[9][j]):6d} {rel(ev[10][j]):6d}")
        del buf_t, out, pad


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else run()
