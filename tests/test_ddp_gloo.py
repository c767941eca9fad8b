"""world_size-2 gloo test of the data-parallel gradient reducer (host logic of the N>1 path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coocc_b200.ddp import GradReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv3d(4, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv3d(8, 2, 1))
    # one parameter in channels_last_3d, like the conv weights of the hot path
    net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last_3d)
    # reference first (no hooks installed yet): average of the per-rank gradients
    ref = [torch.zeros_like(p) for p in net.parameters()]
    for r in range(world):
        net.zero_grad()
        xr = torch.randn(1, 4, 5, 5, 3, generator=torch.Generator().manual_seed(100 + r))
        net(xr).square().mean().backward()
        for acc, p in zip(ref, net.parameters()):
            acc += p.grad / world
    net.zero_grad()
    red = GradReducer(net.parameters(), bucket_bytes=64)
    assert len(red.buckets) >= 2
    x = torch.randn(1, 4, 5, 5, 3, generator=torch.Generator().manual_seed(100 + rank))
    net(x).square().mean().backward()
    nbytes = red.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(grads, ref))
    red.remove()
    # arena mode: p.grad are views of one flat buffer, buckets are slices of it (no flatten / copy back); a second
    # backward before finish() accumulates (gradient accumulation) and a step that raised leaves no stale state
    from coocc_b200.ddp import GradArena
    net.zero_grad(set_to_none=True)
    arena = GradArena(list(net.parameters()))
    red2 = GradReducer(None, bucket_bytes=64, arena=arena)
    ok = ok and len(red2.buckets) >= 2 and all(p.grad.data_ptr() == p._coocc_grad.data_ptr() for p in net.parameters())
    ok = ok and net[0].weight.grad.stride() == net[0].weight.stride()
    red2.mark_ready(net[2].bias)          # a step that died after one gradient ...
    red2.begin()                          # ... is forgotten at the start of the next one
    net(x).square().mean().backward()
    nbytes2 = red2.finish()
    ok = ok and all(torch.allclose(p.grad, b, atol=1e-6) for p, b in zip(net.parameters(), ref))
    ok = ok and all(p.grad.data_ptr() == p._coocc_grad.data_ptr() for p in net.parameters())     # still the arena
    ok = ok and nbytes2 >= nbytes
    q.put((rank, ok, nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(nb > 0 for _, _, nb in res)


def _syncbn_worker(rank, world, port, q):
    """SyncBN bookkeeping on CPU/gloo: the statistics all-reduce of _BNActFn is exercised through its
    pure-python pieces (the CUDA kernels themselves are covered by the GPU tests)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coocc_b200 import functional as CF
    ok = CF._sync_group() is not None
    CF.SYNC_BN["enabled"] = False
    ok = ok and CF._sync_group() is None
    CF.SYNC_BN["enabled"] = True
    # the reduction the forward performs on the epilogue sums
    stats = torch.full((2, 4), float(rank + 1))
    CF._sync_group().all_reduce(stats)
    ok = ok and bool((stats == 3.0).all())
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_group_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
