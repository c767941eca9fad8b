"""Data-parallel gradient all-reduce for the hot path.

The reference wraps the detector in MMDistributedDataParallel(broadcast_buffers=False)
(P/coocc/apis/mmdet_train.py:76-80): one process per GPU, replicas only (the voxel grid never
shards, `assert B == 1` at coocc_ray.py:365), bucketed gradient all-reduce (mean) overlapped with
backward.  GradReducer does the same over torch.distributed (NCCL over NVLink on the B200 box,
gloo in the CPU tests): parameters are bucketed in reverse registration order (the order their
gradients become ready), a bucket is flattened and all-reduced asynchronously as soon as its last
gradient has been accumulated, and `finish()` waits and scatters the averaged values back.
"""
import torch
import torch.distributed as dist


def _flat_view(t):
    """1-D view of a dense tensor in its own memory order (no copy for contiguous or
    channels_last_3d tensors; every rank uses the same layout so element order agrees)."""
    if t.is_contiguous():
        return t.reshape(-1)
    if t.dim() == 5 and t.permute(0, 2, 3, 4, 1).is_contiguous():
        return t.permute(0, 2, 3, 4, 1).reshape(-1)
    return None


class GradReducer:
    def __init__(self, params, bucket_bytes=64 << 20, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # list of lists of params
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._bucket_of[id(p)] = bi
        self._ready = [0] * len(self.buckets)
        self._inflight = []
        self._hooks = []
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p):
        bi = self._bucket_of[id(p)]
        self._ready[bi] += 1
        if self._ready[bi] == len(self.buckets[bi]):
            self._launch(bi)

    def _launch(self, bi):
        grads = [p.grad for p in self.buckets[bi]]
        views = []
        for g in grads:
            v = _flat_view(g)
            views.append(v if v is not None else g.contiguous().reshape(-1))
        flat = torch.cat(views)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((bi, flat, work))

    def finish(self):
        """Wait for all buckets and write the averaged gradients back.  Returns bytes reduced."""
        if self.world == 1:
            return 0
        # buckets whose hooks did not all fire (unused parameters) are reduced here
        launched = {bi for bi, _, _ in self._inflight}
        for bi, b in enumerate(self.buckets):
            if bi not in launched:
                for p in b:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(bi)
        total = 0
        for bi, flat, work in self._inflight:
            work.wait()
            flat.div_(self.world)
            total += flat.numel() * flat.element_size()
            off = 0
            for p in self.buckets[bi]:
                n = p.numel()
                v = _flat_view(p.grad)
                if v is not None:
                    v.copy_(flat[off:off + n])
                else:
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        self._inflight = []
        self._ready = [0] * len(self.buckets)
        return total

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
