"""trilinear mix kernels alone: per-voxel vs column form (values compared, CUDA-event times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import functional as CF, _lib
L = _lib.lib()
dev = "cuda"


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for dt in (torch.bfloat16, torch.float32):
    # OccHead level fusion: 4 levels of 128 channels -> 200x200x16, softmax weights
    dims = [(200, 200, 16), (100, 100, 8), (50, 50, 4), (25, 25, 2)]
    srcs = [torch.randn(d[0] * d[1] * d[2], 128, device=dev).to(dt).requires_grad_(True) for d in dims]
    w = torch.softmax(torch.randn(640000, 4, device=dev), 1).requires_grad_(True)
    g = torch.randn(640000, 128, device=dev).to(dt)
    res = {}
    for flags, name in ((3, "per-voxel"), (1, "column")):
        L.coocc_trilinear_tune(flags)
        out = CF.resize_mix(srcs, dims, dims[0], wts=w)
        t_f = timeit(lambda: CF.resize_mix(srcs, dims, dims[0], wts=w))
        grads = torch.autograd.grad(out, srcs + [w], g)
        t_b = timeit(lambda: torch.autograd.grad(CF.resize_mix(srcs, dims, dims[0], wts=w), srcs + [w], g)) - t_f
        res[name] = (out.float(), [x.float() for x in grads])
        print("head mix %s %-9s fwd %.3f ms  bwd %.3f ms" % (str(dt)[6:], name, t_f, t_b), flush=True)
    a, b = res["per-voxel"], res["column"]
    print("   max |diff| out %.3e  grads %s" % ((a[0] - b[0]).abs().max().item(), ["%.2e" % (x - y).abs().max().item() for x, y in zip(a[1], b[1])]))
    # FPN top-down add: base + upsample(256 channels)
    src = torch.randn(80000, 256, device=dev).to(dt).requires_grad_(True)
    base = torch.randn(640000, 256, device=dev).to(dt).requires_grad_(True)
    g2 = torch.randn(640000, 256, device=dev).to(dt)
    res = {}
    for flags, name in ((3, "per-voxel"), (1, "column")):
        L.coocc_trilinear_tune(flags)
        out = CF.resize_mix([src], [dims[1]], dims[0], base=base)
        t_f = timeit(lambda: CF.resize_mix([src], [dims[1]], dims[0], base=base))
        res[name] = out.float()
        print("fpn add  %s %-9s fwd %.3f ms" % (str(dt)[6:], name, t_f), flush=True)
    print("   max |diff| out %.3e" % (res["per-voxel"] - res["column"]).abs().max().item())
L.coocc_trilinear_tune(1)
