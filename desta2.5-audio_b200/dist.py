"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink; gloo on CPU for tests).

Clips are independent units: rank r of G owns a contiguous block of clips and replicated weights; the
data path needs NO collective.  Collectives exist only at the two real exchange points of the reference:
  * training: the connector gradients are summed across ranks exactly like the reference's DDP bucketed all-reduce
    (HF Trainer wraps the model in DistributedDataParallel, ref: examples/train/run_desta_qwen3_4b.sbatch:68);
  * inference: embeddings of all ranks may be gathered to the rank(s) hosting the LLM.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of ceil(n/world) items for `rank` (the last ranks may get fewer / none)."""
    per = (n_items + world - 1) // world
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def shard_batch(batch_features, batch_transcription_ids, batch_start_positions, rank: int, world: int):
    """Shard a batch by CLIP, keeping each clip's (prompt, start) descriptor with it."""
    s, e = shard_range(len(batch_start_positions), rank, world)
    return batch_features[s:e], batch_transcription_ids[s:e], batch_start_positions[s:e], (s, e)


def gather_embeddings(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank audio features [n_local, Q, d] into [n_total, Q, d] in shard order (ranks are padded to the
    common block size for the collective and trimmed afterwards)."""
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


class GradBucket:
    """One flat fp32 buffer for all connector gradients: a single NCCL all-reduce per step (131.5 M elements at the
    flagship size), averaged like DDP."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()

    def allreduce(self, group=None, async_op: bool = False) -> Optional[dist.Work]:
        # gather the gradients into the flat buffer with ONE multi-tensor copy (165 tensors at the flagship size: a
        # Python loop of small copies costs more host time than the collective itself)
        have = [(p, v) for p, v in zip(self.params, self.views) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        if async_op:
            return work
        work.wait()
        self.finish(group)
        return None

    def finish(self, group=None):
        """Average like DDP and expose the result as views of the flat buffer (DDP's gradient_as_bucket_view): no copy out."""
        world = dist.get_world_size(group)
        self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            p.grad = v


class GradSync:
    """Gradient exchange overlapped with the connector backward.  Assign an instance to `perception.grad_sync`: the
    backward then all-reduces each Q-Former layer's slice of the gradient arena as soon as it is final (asynchronously,
    on NCCL's stream) while the remaining layers are still being differentiated, and `.grad`s come out already averaged
    over the data-parallel group - the role DDP's bucketed hooks play for the reference (no further all-reduce needed)."""

    def __init__(self, group=None):
        self.group = group
        self.works = []

    def reduce(self, arena_slice: torch.Tensor):
        self.works.append(dist.all_reduce(arena_slice, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self, arena: torch.Tensor):
        for w in self.works:
            w.wait()
        self.works = []
        arena.div_(dist.get_world_size(self.group))


def allreduce_connector_grads(connector: torch.nn.Module, group=None, bucket: Optional[GradBucket] = None) -> GradBucket:
    bucket = bucket or GradBucket(list(connector.parameters()))
    bucket.allreduce(group)
    return bucket
