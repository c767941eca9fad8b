"""Times the BatchNorm kernels alone (CUDA events, inputs larger than L2) at the north-star layer shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib()
dev = "cuda"
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
peak = 6552.0


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for V, C in ((640000, 128), (640000, 256), (80000, 256), (10000, 512)):
    for bf in (1, 0):
        dt = torch.bfloat16 if bf else torch.float32
        e = 2 if bf else 4
        x = torch.randn(V, C, device=dev).to(dt)
        res = torch.randn(V, C, device=dev).to(dt)
        out = torch.empty_like(x)
        dout = torch.randn(V, C, device=dev).to(dt)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x)
        mi = torch.stack([torch.zeros(C, device=dev), torch.ones(C, device=dev)]).contiguous()
        g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
        sums = torch.zeros(2, C, device=dev)
        for has_res in (False, True):
            r = res if has_res else None
            o = out if has_res else None
            t = timeit(lambda: L.coocc_bn_act_fwd(P(x), C, V, C, P(mi), P(g), P(b), P(r), C if has_res else 0, 1, P(out), C, bf, st()))
            by = e * V * C * (3 if has_res else 2)
            print("V=%d C=%d %s res=%d  fwd    %.3f ms  %.0f GB/s (%.2f)" % (V, C, "bf16" if bf else "fp32", has_res, t, by / t / 1e6, by / t / 1e6 / peak))
            t = timeit(lambda: L.coocc_bn_act_bwd_reduce(P(dout), C, P(o), C if has_res else 0, P(x), C, V, C, P(mi), P(g), P(b), 1, P(sums), bf, st()))
            by = e * V * C * (3 if has_res else 2)
            print("V=%d C=%d %s res=%d  reduce %.3f ms  %.0f GB/s (%.2f)" % (V, C, "bf16" if bf else "fp32", has_res, t, by / t / 1e6, by / t / 1e6 / peak))
            t = timeit(lambda: L.coocc_bn_act_bwd_apply(P(dout), C, P(o), C if has_res else 0, P(x), C, V, C, P(mi), P(g), P(b), 1, P(sums), V, P(dx), C, bf, P(dres) if has_res else None, C if has_res else 0, st()))
            by = e * V * C * (5 if has_res else 3)
            print("V=%d C=%d %s res=%d  apply  %.3f ms  %.0f GB/s (%.2f)" % (V, C, "bf16" if bf else "fp32", has_res, t, by / t / 1e6, by / t / 1e6 / peak))
    # reference: torch copy of the same tensor
    t = timeit(lambda: out.copy_(x))
    print("V=%d C=%d torch copy_ %.3f ms %.0f GB/s" % (V, C, t, 2 * x.numel() * x.element_size() / t / 1e6))
