"""The work-item body of csrc/adamw.cuh (multi-tensor AdamW + bf16 shadow) executed on the CPU through
tests/emul/adamw_emul.cpp -- the same source the CUDA launcher compiles -- against torch.optim.AdamW."""
import ctypes
import os
import subprocess

import pytest
import torch

from coocc_b200 import optim as CO

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "adamw_emul.cpp")
SO = os.path.join(HERE, "emul", "libadamw_emul.so")


@pytest.fixture(scope="module")
def E():
    hdr = os.path.join(HERE, "..", "co-occ_b200", "csrc", "adamw.cuh")
    if not os.path.isfile(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    return ctypes.CDLL(SO)


def test_tables_cover_every_element_once():
    sizes = [5, 16384, 16385, 40000, 3]
    ct, ci = CO.build_tables(sizes)
    covered = [0] * len(sizes)
    for t, c in zip(ct, ci):
        covered[t] += max(0, min(CO.CHUNK, sizes[t] - c * CO.CHUNK))
    assert covered == sizes and len(ct) == 1 + 1 + 2 + 3 + 1


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adamw_matches_torch_over_several_steps(E, wd):
    gen = torch.Generator().manual_seed(0)
    shapes = [(7,), (64, 33), (16385,), (8, 4, 3, 3, 3)]
    ps = [torch.randn(*s, generator=gen).requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in ps]
    opt = torch.optim.AdamW(ref, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    flat = [p.detach().reshape(-1) for p in ps]                  # views: the harness updates p in place
    m = [torch.zeros_like(f) for f in flat]
    v = [torch.zeros_like(f) for f in flat]
    sh = [torch.zeros(f.numel(), dtype=torch.int16) for f in flat]
    ct, ci = CO.build_tables([f.numel() for f in flat])
    ctt, cit = torch.tensor(ct, dtype=torch.int32), torch.tensor(ci, dtype=torch.int32)
    step = torch.zeros(1)
    for it in range(5):
        grads = [torch.randn(f.numel(), generator=gen) * (0.1 + it) for f in flat]
        for r, g in zip(ref, grads):
            r.grad = g.reshape(r.shape).clone()
        opt.step()
        rows = []
        for f, g, mm, vv, s in zip(flat, grads, m, v, sh):
            rows += [f.data_ptr(), g.data_ptr(), mm.data_ptr(), vv.data_ptr(), s.data_ptr(), f.numel(), CO.pack_mults(1.0, 1.0)]
        table = torch.tensor(rows, dtype=torch.int64)
        E.emul_adamw_step(ctypes.c_void_p(table.data_ptr()), len(flat), ctypes.c_void_p(ctt.data_ptr()),
                          ctypes.c_void_p(cit.data_ptr()), len(ct), CO.CHUNK, ctypes.c_float(1e-2), ctypes.c_float(0.9),
                          ctypes.c_float(0.99), ctypes.c_float(1e-8), ctypes.c_float(wd),
                          ctypes.c_void_p(step.data_ptr()), 1, None)
        assert float(step) == it + 1
        for p, r, g, s in zip(ps, ref, grads, sh):
            assert torch.allclose(p.detach(), r.detach(), rtol=2e-5, atol=1e-7), it
            assert float(g.abs().max()) == 0.0                                        # zero_grad folded into the pass
            assert torch.equal(s.view(torch.bfloat16), p.detach().reshape(-1).to(torch.bfloat16))   # shadow = bf16(p), RNE


def test_param_multipliers_lr_scale_and_grad_scale(E):
    """wd_mult / lr_mult per tensor (mmcv paramwise_cfg), the device-side lr multiplier and gradient scale: equal to
    torch.optim.AdamW with the corresponding param groups fed pre-scaled gradients."""
    gen = torch.Generator().manual_seed(1)
    ps = [torch.randn(n, generator=gen).requires_grad_(True) for n in (50, 300)]
    ref = [p.detach().clone().requires_grad_(True) for p in ps]
    lr, wd, lr_scale, gscale = 1e-2, 0.05, 0.1, 0.25
    opt = torch.optim.AdamW([dict(params=[ref[0]], weight_decay=0.0, lr=lr * lr_scale),
                             dict(params=[ref[1]], weight_decay=wd, lr=lr * lr_scale * 2.0)], betas=(0.9, 0.99))
    mults = [(0.0, 1.0), (1.0, 2.0)]
    flat = [p.detach().reshape(-1) for p in ps]
    m = [torch.zeros_like(f) for f in flat]
    v = [torch.zeros_like(f) for f in flat]
    ct, ci = CO.build_tables([f.numel() for f in flat])
    ctt, cit = torch.tensor(ct, dtype=torch.int32), torch.tensor(ci, dtype=torch.int32)
    step = torch.zeros(1)
    dyn = torch.tensor([lr_scale, gscale])
    for it in range(3):
        grads = [torch.randn(f.numel(), generator=gen) for f in flat]
        for r, g in zip(ref, grads):
            r.grad = (g * gscale).clone()
        opt.step()
        rows = []
        for f, g, mm, vv, mu in zip(flat, grads, m, v, mults):
            rows += [f.data_ptr(), g.data_ptr(), mm.data_ptr(), vv.data_ptr(), 0, f.numel(), CO.pack_mults(*mu)]
        table = torch.tensor(rows, dtype=torch.int64)
        E.emul_adamw_step(ctypes.c_void_p(table.data_ptr()), len(flat), ctypes.c_void_p(ctt.data_ptr()),
                          ctypes.c_void_p(cit.data_ptr()), len(ct), CO.CHUNK, ctypes.c_float(lr), ctypes.c_float(0.9),
                          ctypes.c_float(0.99), ctypes.c_float(1e-8), ctypes.c_float(wd),
                          ctypes.c_void_p(step.data_ptr()), 0, ctypes.c_void_p(dyn.data_ptr()))
        for p, r in zip(ps, ref):
            assert torch.allclose(p.detach(), r.detach(), rtol=2e-5, atol=1e-7), it
