"""AdamW for the hot-path parameters as ONE multi-tensor kernel (csrc/adamw.cu) that also refreshes the bf16 operand
copy ("shadow") of every parameter, so that the bf16 mode no longer converts 148 M weights per step in separate
passes, and -- with a ddp.GradArena -- clears the gradients in the same pass.  Update rule = torch.optim.AdamW (the
reference trains with AdamW, coocc_multi_r50_256x704.py:263-276), including what its optimizer config asks for:

  * paramwise_cfg norm_decay_mult = 0 (:276): per-parameter weight-decay / lr multipliers (`param_mults`);
  * optimizer_config grad_clip max_norm = 5 (:279): `max_norm`, the clip coefficient is computed on the device and
    read by the kernel at run time;
  * lr_config step policy (:282-285): `set_lr_scale()` writes a device scalar the kernel multiplies the learning
    rate with, so a step replayed from a CUDA graph follows the schedule.

Verified on the CPU against torch.optim.AdamW (tests/test_adamw_emul.py, kernel body compiled for the host) and on
the GPU (tests/test_gpu_adamw.py: eager, under several CUDA graphs and with an eager step in between).
"""
import ctypes
import struct

import torch

from . import _lib

CHUNK = 16384      # elements per scheduling chunk (multiple of 4)


def build_tables(sizes, chunk=CHUNK):
    """(chunk_tensor, chunk_index) int32 lists covering tensors of the given element counts."""
    ct, ci = [], []
    for t, n in enumerate(sizes):
        for c in range((n + chunk - 1) // chunk):
            ct.append(t)
            ci.append(c)
    return ct, ci


def _dense(t):
    """flat view of a dense tensor in its own memory order (contiguous or channels_last_3d)."""
    if t.is_contiguous():
        return t.view(-1)
    if t.dim() == 5 and t.permute(0, 2, 3, 4, 1).is_contiguous():
        return t.permute(0, 2, 3, 4, 1).reshape(-1)
    raise ValueError("FusedAdamW needs dense parameters")


def pack_mults(wd_mult, lr_mult):
    """the two float32 multipliers of a table entry as one int64 (little endian: wd_mult first)"""
    return struct.unpack("<q", struct.pack("<ff", float(wd_mult), float(lr_mult)))[0]


def norm_decay_mults(model, norm_decay_mult=0.0):
    """{parameter: (wd_mult, lr_mult)} for mmcv's `paramwise_cfg=dict(norm_decay_mult=...)`: the weights and biases of
    normalisation layers get weight_decay * norm_decay_mult."""
    import torch.nn as nn
    norms = (nn.modules.batchnorm._BatchNorm, nn.GroupNorm, nn.LayerNorm)
    out = {}
    for m in model.modules():
        if isinstance(m, norms):
            for p in m.parameters(recurse=False):
                out[p] = (norm_decay_mult, 1.0)
    return out


class FusedAdamW:
    """step() / zero_grad() like a torch optimizer for a fixed list of fp32 CUDA parameters.

    shadow=True keeps `p._coocc_bf16` (a flat bf16 buffer in p's memory order) equal to bf16(p) after every step;
    functional.weight_operand uses it instead of converting.
    arena = ddp.GradArena: gradients live in the arena (`p.grad` are views of it, bound for good); the kernel clears
    them after use and zero_grad() becomes a no-op after a step.  Without an arena every parameter must have a
    gradient at step() and the pointer table follows the gradient tensors autograd hands out.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, shadow=True, arena=None,
                 param_mults=None, max_norm=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.arena, self.max_norm = arena, max_norm
        self.mults = [tuple((param_mults or {}).get(p, (1.0, 1.0))) for p in self.params]
        dev = self.params[0].device
        self.m = [torch.zeros_like(_dense(p)) for p in self.params]
        self.v = [torch.zeros_like(_dense(p)) for p in self.params]
        self.step_t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.dyn = torch.ones(2, device=dev, dtype=torch.float32)       # {lr multiplier, gradient scale}
        self.grad_norm = torch.zeros((), device=dev, dtype=torch.float32)
        self.shadow = shadow
        if shadow:
            for p in self.params:
                p._coocc_bf16 = _dense(p.detach()).to(torch.bfloat16)
                p._coocc_bf16_version = p._version
                # the shadow is a flat copy in p's memory order *now*: a later relayout (p.data = ...) keeps the
                # version counter, so the storage address and strides are part of the validity key
                p._coocc_bf16_layout = (p.data_ptr(), tuple(p.stride()))
        ct, ci = build_tables([p.numel() for p in self.params])
        self.chunk_tensor = torch.tensor(ct, device=dev, dtype=torch.int32)
        self.chunk_index = torch.tensor(ci, device=dev, dtype=torch.int32)
        # one (pinned host, device) pointer table per distinct set of gradient addresses.  A table is written once
        # and never changed afterwards: the upload of a table first seen during a CUDA-graph capture becomes a
        # memcpy node that re-reads its own pinned buffer on every replay, so the buffer must stay intact while
        # other captures / eager steps (different gradient addresses) create their own tables.  With an arena the
        # addresses never change: one table, uploaded here.
        self._tables = {}
        self._grads_clean = arena is not None
        if arena is not None:
            arena.bind()
            self._arena_table = self._table_for([_dense(p._coocc_grad).data_ptr() for p in self.params])
            torch.cuda.synchronize(dev)
        self.state = {"initialised": True}      # torch-optimizer-like attribute (graph.GraphedStep checks it is non-empty)

    def set_lr_scale(self, scale):
        """learning-rate schedule: the kernel multiplies lr with this device scalar (also inside a replayed graph)"""
        self.dyn[0:1].fill_(float(scale))

    def zero_grad(self, set_to_none=True):
        if self.arena is not None:
            self.arena.bind()
            if not self._grads_clean:
                self.arena.flat.zero_()
            self._grads_clean = False          # the coming backward accumulates into the arena
            return
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _table_for(self, ptrs):
        key = tuple(ptrs)
        hit = self._tables.get(key)
        if hit is not None:
            return hit[1]
        rows = []
        for p, g, m, v, (wd_mult, lr_mult) in zip(self.params, ptrs, self.m, self.v, self.mults):
            sh = p._coocc_bf16.data_ptr() if self.shadow else 0
            rows += [_dense(p.detach()).data_ptr(), g, m.data_ptr(), v.data_ptr(), sh, p.numel(),
                     pack_mults(wd_mult, lr_mult)]
        host = torch.tensor(rows, dtype=torch.int64).pin_memory()
        devt = torch.empty(len(rows), device=self.params[0].device, dtype=torch.int64)
        devt.copy_(host, non_blocking=True)
        if len(self._tables) >= 64:          # gradient buffers keep moving (no graph, caching allocator churn)
            self._tables.pop(next(iter(self._tables)))
        self._tables[key] = (host, devt)
        return devt

    def _refresh_table(self):
        if self.arena is not None:
            return self._arena_table
        ptrs = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("FusedAdamW.step(): a parameter has no gradient")
            ptrs.append(_dense(p.grad).data_ptr())
        return self._table_for(ptrs)

    def _clip(self):
        """clip_grad_norm_(max_norm): total L2 norm over all gradients -> dyn[1] = min(1, max_norm / (norm + 1e-6))"""
        if self.arena is not None:
            norm = torch.linalg.vector_norm(self.arena.flat)
        else:
            norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad) for p in self.params]))
        self.grad_norm = norm
        self.dyn[1:2].copy_((self.max_norm / (norm + 1e-6)).clamp(max=1.0).reshape(1))

    @torch.no_grad()
    def step(self):
        L = _lib.lib()
        table = self._refresh_table()
        if self.max_norm is not None:
            self._clip()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.coocc_adamw_step(ctypes.c_void_p(table.data_ptr()), len(self.params),
                                      ctypes.c_void_p(self.chunk_tensor.data_ptr()),
                                      ctypes.c_void_p(self.chunk_index.data_ptr()), self.chunk_tensor.numel(), CHUNK,
                                      float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                      float(self.weight_decay), ctypes.c_void_p(self.step_t.data_ptr()),
                                      1 if self.arena is not None else 0, ctypes.c_void_p(self.dyn.data_ptr()), st),
                   "adamw_step")
        self._grads_clean = self.arena is not None
        # the shadow was rewritten in the same pass as p: mark it valid for p's current version (the raw-pointer update
        # does not move the autograd version counter; any later in-place change of p does, and invalidates it)
        if self.shadow:
            for p in self.params:
                p._coocc_bf16_version = p._version
