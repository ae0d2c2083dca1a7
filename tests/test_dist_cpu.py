"""CPU, world_size 2, gloo: the data-parallel host logic (flat gradient buffer, backward-order bucketing,
mean all-reduce, SyncBN statistic reduction)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from unet_torch_b200.dist import DataParallelContext

    ctx = DataParallelContext.enable(sync_bn=True, bucket_mb=0.001)  # tiny buckets -> several all-reduces
    assert DataParallelContext.current() is ctx and ctx.world_size == world
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (700,), (64, 3, 3, 3), (2,), (1000,))]
    flat = ctx.make_flat_grads(params)
    assert len(flat.buckets) >= 3
    views = [flat.view_for(p) for p in params]
    for i, (p, v) in enumerate(zip(params, views)):
        assert v.shape == p.shape
        v.fill_(float(rank + 1) * (i + 1))
    # gradients become ready in backward order, possibly several at a time
    flat.mark_ready(params[:1])
    flat.mark_ready(params[1:3])
    flat.mark_ready(params[3:])
    flat.finish()
    ok = all(torch.allclose(v, torch.full_like(v, (1 + 2) / 2 * (i + 1))) for i, v in enumerate(views))
    # out of declared order (an attention gate's gradients are finished before the ConvTranspose2d of its Up block, which comes
    # first in the declared order): buckets only leave once the contiguous prefix is ready, every value is still averaged once
    flat2 = ctx.make_flat_grads(params)
    views2 = [flat2.view_for(p) for p in params]
    for i, v in enumerate(views2):
        v.fill_(float(rank + 1) * (i + 2))
    flat2.mark_ready(params[2:4])
    ok = ok and flat2.next_bucket == 0          # nothing may be reduced before params[0] is there
    flat2.mark_ready(params[4:])
    flat2.mark_ready(params[:2])
    flat2.finish()
    ok = ok and all(torch.allclose(v, torch.full_like(v, (1 + 2) / 2 * (i + 2))) for i, v in enumerate(views2))
    # SyncBN: all-reduced [sum, sum^2] with the global count reproduce whole-batch statistics
    g = torch.Generator().manual_seed(7)
    full = torch.randn(4, 8, 6, 6, generator=g, dtype=torch.float64)
    mine = full[rank * 2:(rank + 1) * 2]
    sums = torch.cat([mine.sum((0, 2, 3)), (mine * mine).sum((0, 2, 3))])
    ctx.all_reduce_sum(sums)
    cnt = mine.numel() / 8 * world
    mean = sums[:8] / cnt
    var = sums[8:] / cnt - mean * mean
    ok = ok and torch.allclose(mean, full.mean((0, 2, 3))) and torch.allclose(var, full.var((0, 2, 3), unbiased=False))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_grads_and_syncbn_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
