"""Context number (SURVEY.md §8d "GPU library baseline"): the stock HF Whisper-large-v3 ENCODER (random init, bf16, SDPA
attention, output_hidden_states for the four taps) on the same B200 - the library path the reference itself would run on
a GPU, without the Q-Former connector, the log-mel or the splice.  python tools/library_baseline.py [clips]"""
import sys
import time

import torch
from transformers import WhisperConfig
from transformers.models.whisper.modeling_whisper import WhisperEncoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = WhisperConfig(num_mel_bins=128, d_model=1280, encoder_layers=32, encoder_attention_heads=20, encoder_ffn_dim=5120,
                    max_source_positions=1500, attn_implementation="sdpa")
torch.manual_seed(0)
enc = WhisperEncoder(cfg).to("cuda", torch.bfloat16).eval()
mel = torch.randn(n, 128, 3000, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(2):
        enc(mel, output_hidden_states=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        enc(mel, output_hidden_states=True)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"HF WhisperEncoder large-v3, bf16 weights + SDPA, {n} clips: {ms:.1f} ms per step = {30.0 * n / ms * 1e3:.0f} audio-s/s "
      f"(encoder only: 2.274 TFLOP/clip -> {2.2738e12 * n / ms / 1e9:.0f} TFLOP/s)")
