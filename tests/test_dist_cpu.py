"""N > 1 host logic on CPU: world_size-2 gloo process group (clip sharding, embedding gather, gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from desta_b200.dist import GradBucket, gather_embeddings, shard_batch, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 65, 257):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                s, e = shard_range(n, r, world)
                assert 0 <= s <= e <= n
                seen += list(range(s, e))
            assert seen == list(range(n))


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 4 * 8, dtype=torch.float32).view(n_total, 4, 8)
        trans = [torch.full((1, i % 3), i) for i in range(n_total)]
        pos = [(i // 2, 10 * i) for i in range(n_total)]
        feats, t, p, (s, e) = shard_batch(full, trans, pos, rank, world)
        assert len(t) == len(p) == feats.shape[0] == e - s and p == pos[s:e]
        gathered = gather_embeddings(feats * 1.0, n_total)
        assert torch.equal(gathered, full)
        # gradient all-reduce == mean over ranks of per-rank grads (DDP semantics)
        torch.manual_seed(0)
        lin = torch.nn.Linear(8, 3)
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank))
        lin(x).pow(2).sum().backward()
        mine = [p_.grad.clone() for p_ in lin.parameters()]
        bucket = GradBucket(list(lin.parameters()))
        bucket.allreduce()
        allg = [torch.zeros_like(mine[0]) for _ in range(world)]
        dist.all_gather(allg, mine[0])
        assert torch.allclose(lin.weight.grad, torch.stack(allg).mean(0), atol=1e-6)
        out.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        out.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _gradsync_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from desta_b200.dist import GradSync
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    arena = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    sync = GradSync()
    for a0, a1 in ((600, 1000), (300, 600), (0, 300)):          # slices become final back to front, like the backward
        sync.reduce(arena[a0:a1])
    sync.finish(arena)
    q.put((rank, arena.clone()))
    dist.destroy_process_group()


def test_gloo_world2_gradsync_overlapped_slices():
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650
    procs = [ctx.Process(target=_gradsync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    want = torch.arange(1000, dtype=torch.float32) * 1.5        # mean of 1x and 2x
    assert torch.equal(res[0], want) and torch.equal(res[1], want)
