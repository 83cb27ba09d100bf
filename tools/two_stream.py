"""Experiment: does running two half-batches on two streams (phase-shifted, so that one half's HBM-bound LayerNorms overlap
the other half's GEMMs / attention) beat one full batch?  python tools/two_stream.py [clips]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
from bench import WORKLOADS, synth_wave_batch  # noqa: E402
from desta_b200 import engine, ops  # noqa: E402
from desta_b200.engine import Workspace  # noqa: E402
from desta_b200.synthetic import build_perception  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = WORKLOADS["large_v3_llama"]
dev = torch.device("cuda", 0)
p = build_perception(w["n_mels"], w["d_model"], w["n_heads"], w["ffn"], w["n_layers"], w["tap_layers"], w["qf_layers"],
                     w["prompt_size"], w["d_llm"], dev, seed=0)
wave = synth_wave_batch(n, 1234).to(dev)
enc, con = p._packs()


def run(ws, wv, out):
    _, mel_tm, mel_floor = ops.logmel(wv, p.dims.n_mels, want_bf16_tm=True, want_f32=False)
    engine.encoder_forward(enc, p.dims, ws, mel_tm, mel_floor=mel_floor)
    engine.connector_forward(con, p.dims, ws, out)


def timeit(fn, iters=5, warm=3):
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


with torch.no_grad():
    ws_full = Workspace(p.dims, n, dev)
    out_full = torch.empty(n * p.dims.prompt_size, p.dims.d_llm, device=dev)
    t1 = timeit(lambda: run(ws_full, wave, out_full))
    print(f"one stream, {n} clips: {t1:.2f} ms", flush=True)
    del ws_full
    h = n // 2
    ws = [Workspace(p.dims, h, dev) for _ in range(2)]
    outs = [torch.empty(h * p.dims.prompt_size, p.dims.d_llm, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]

    def two():
        cur = torch.cuda.current_stream(dev)
        for i in range(2):
            streams[i].wait_stream(cur)
            with torch.cuda.stream(streams[i]):
                run(ws[i], wave[i * h:(i + 1) * h], outs[i])
        for i in range(2):
            cur.wait_stream(streams[i])

    t2 = timeit(two)
    print(f"two streams, 2 x {h} clips: {t2:.2f} ms  ({t1 / t2:.3f}x)", flush=True)
    t3 = timeit(lambda: (run(ws[0], wave[:h], outs[0]), run(ws[1], wave[h:], outs[1])))
    print(f"one stream, 2 x {h} clips back to back: {t3:.2f} ms", flush=True)
