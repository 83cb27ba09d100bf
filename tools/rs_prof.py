import os, sys, torch
sys.path.insert(0, "/root/repo")
import __graft_entry__; __graft_entry__.build()
from desta_b200 import ops
N = 64
for sr, ch in ((22050, 1), (24000, 1), (44100, 2)):
    n = sr * 30
    pcm = torch.randint(-20000, 20000, (N, n * ch), dtype=torch.int16, device="cuda").view(torch.uint8).reshape(-1)
    clips = [(i * n * ch * 2, n, i) for i in range(N)]
    out = torch.empty(N, 480000, device="cuda")
    for _ in range(2):
        ops.decode_resample(pcm, clips, 1, ch, sr, 16000, out)
    torch.cuda.synchronize()
