"""Peer-memory all-reduce kernel (csrc/peer_reduce.cu, co-occ_b200/peer.py): equal to an NCCL all-reduce, bit-identical
on all ranks, correct across slot wrap-around, inside a CUDA graph, and (2 GPUs) faster than the NCCL call it replaces.
Runs on min(2, device_count) ranks: on a single-GPU box the kernel pushes into its own buffer (world 1), with
`gpurun --gpus 2` over NVLink."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from coocc_b200.peer import PeerExchange
        px = PeerExchange(nslots=4, slot_floats=2048)
        ok = True
        g = torch.Generator().manual_seed(10 + rank)
        for it in range(37):                                    # 37 calls over 4 slots: many wrap-arounds
            n = [2048, 256, 2, 1000][it % 4]
            t = torch.randn(n, generator=g).to(dev)
            ref = t.clone()
            dist.all_reduce(ref)
            if it % 5 == 0:
                px.begin_step()
            px.all_reduce(t)
            ok = ok and torch.allclose(t, ref, rtol=1e-6, atol=1e-6)
            both = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(both, t)
            ok = ok and all(torch.equal(both[0], b) for b in both)          # bit-identical replicas
        # inside a CUDA graph
        px.begin_step()
        buf = torch.zeros(512, device=dev)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            px.all_reduce(buf)
            px.all_reduce(buf)
        for it in range(3):
            buf.fill_(float(rank + 1 + it))
            graph.replay()
            torch.cuda.synchronize()
            want = float(sum(r + 1 + it for r in range(world)) * world)   # two reductions: sum, then sum of sums
            ok = ok and bool((buf == want).all())
        # latency against the NCCL call it replaces
        t = torch.zeros(2048, device=dev)
        for fn in (lambda: dist.all_reduce(t), lambda: px.all_reduce(t)):
            for _ in range(20):
                fn()
        torch.cuda.synchronize()
        res = []
        for fn in (lambda: dist.all_reduce(t), lambda: px.all_reduce(t)):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 200 * 1e3)
        q.put((rank, ok, res))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_allreduce():
    world = min(2, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    print("[peer] us per call (nccl, peer):", [r[2] for r in res])
    assert all(ok for _, ok, _ in res), res
    if world > 1:
        for _, _, (t_nccl, t_peer) in res:
            assert t_peer < t_nccl, (t_nccl, t_peer)
