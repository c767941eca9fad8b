"""Peer-memory exchange for the latency-bound collectives of the data-parallel path (csrc/peer_reduce.cu).

`PeerExchange` owns one exchange buffer per rank, allocated and mapped into every process with
torch.distributed._symmetric_memory (CUDA IPC / fabric handles over NVLink -- plumbing), and hands the peer pointers to
`coocc_peer_allreduce`, a single-CTA push / flag / sum kernel.  It replaces the per-BatchNorm NCCL all-reduces of
SyncBN (72 per step, each ~20 us of launch + protocol latency on the critical path) with ~5 us kernels that are
ordinary nodes of the step's CUDA graph.  Results are summed in rank order on every rank: replicas stay bit-identical.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class PeerExchange:
    def __init__(self, group=None, nslots=8, slot_floats=4096, device=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.nslots, self.slot_floats = int(nslots), int(slot_floats)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        L = _lib.lib()
        nbytes = int(L.coocc_peer_buffer_bytes(self.world, self.nslots, self.slot_floats))
        if nbytes < 0:
            raise ValueError("PeerExchange: unsupported world size / slot geometry")
        name = self.group.group_name
        try:
            if not symm.is_symm_mem_enabled_for_group(name):
                symm.enable_symm_mem_for_group(name)
        except Exception:  # noqa: BLE001 -- newer torch enables groups lazily
            pass
        self.buf = symm.empty(nbytes // 4, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr()
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.epoch = torch.zeros(self.nslots, device=dev, dtype=torch.int32)
        self.counter = 0
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)            # every buffer is zeroed before any rank pushes into it

    def begin_step(self):
        """Slots are assigned by the position of a call inside the step, so that an eagerly launched step and a step
        replayed from a CUDA graph (slot numbers baked at capture) agree on every rank."""
        self.counter = 0

    def fits(self, t):
        return t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and 0 < t.numel() <= self.slot_floats

    def all_reduce(self, t):
        """in-place sum over the ranks of a small contiguous fp32 CUDA tensor"""
        assert self.fits(t)
        slot = self.counter % self.nslots
        self.counter += 1
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.coocc_peer_allreduce(ctypes.c_void_p(t.data_ptr()), t.numel(), self.ptrs, self.rank, self.world,
                                          slot, self.nslots, self.slot_floats, ctypes.c_void_p(self.epoch.data_ptr()),
                                          st), "peer_allreduce")
        return t
