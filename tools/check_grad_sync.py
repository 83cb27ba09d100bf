"""torchrun check (N >= 2 GPUs): gradients from the backward-overlapped all-reduce (dist.GradSync) equal the average of the
per-rank gradients obtained with a plain all-reduce after the backward.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_grad_sync.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import __graft_entry__  # noqa: E402

if rank == 0:
    __graft_entry__.build()
dist.barrier()
from desta_b200.dist import GradSync  # noqa: E402
from desta_b200.synthetic import build_perception  # noqa: E402

p = build_perception(80, 384, 6, 1536, 4, [0, 1, 2, 3], 2, 64, 1024, dev, seed=0).eval()      # same weights on every rank
g = torch.Generator(device=dev).manual_seed(100 + rank)                                         # different clips per rank
wave = torch.clamp(0.1 * torch.randn(3, 480000, device=dev, generator=g), -1, 1)
gout = torch.randn(3, 64, 1024, device=dev, generator=g)


def grads(sync):
    p.grad_sync = sync
    for prm in p.connector.parameters():
        prm.grad = None
    feats, _ = p(input_features=wave)
    (feats * gout).sum().backward()
    return [prm.grad.clone() for prm in p.connector.parameters()]


ref = grads(None)
for t in ref:
    dist.all_reduce(t)
    t.div_(world)
got = grads(GradSync())
worst = max((a - b).abs().max().item() / (b.abs().max().item() + 1e-30) for a, b in zip(got, ref))
print(f"rank {rank}: {len(got)} gradient tensors, worst relative difference {worst:.3e}", flush=True)
assert worst < 1e-6, worst
dist.destroy_process_group()
