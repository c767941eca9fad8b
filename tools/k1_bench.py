"""1x1x1 convolution launches of the north-star step, timed back to back with CUDA events (HBM-bound layers:
GB/s against the algorithmic bytes)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib(); dev = "cuda"
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
X, Y, Z = 200, 200, 16
V = X * Y * Z
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()                                   # L2 flush between launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


for (Cin, Cout) in ((128, 64), (128, 128), (128, 256), (64, 17)):
    xc = torch.randn(V, Cin, device=dev).to(torch.bfloat16)
    wc = (torch.randn(Cout, Cin, device=dev) * 0.05).to(torch.bfloat16)
    ldy = (Cout + 7) // 8 * 8
    gy = torch.randn(V, ldy, device=dev).to(torch.bfloat16)
    y = torch.empty(V, ldy, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(V, Cin, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(Cout, Cin, device=dev); stats = torch.zeros(2, Cout, device=dev)
    d = _lib.ConvDesc(X, Y, Z, Cin, Cout, 1, 1, 1, Cin, ldy, 1)
    f = lambda: L.coocc_conv3d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), y.data_ptr(), ldy, None, 0, stats.data_ptr(), st())
    g = lambda: L.coocc_conv3d_dgrad(ctypes.byref(d), gy.data_ptr(), wc.data_ptr(), dx.data_ptr(), Cin, st())
    w = lambda: L.coocc_conv3d_wgrad(ctypes.byref(d), xc.data_ptr(), gy.data_ptr(), dw.data_ptr(), st())
    by = 2.0 * V * (Cin + Cout)
    for name, fn in (("fwd+stats", f), ("dgrad", g), ("wgrad", w)):
        us = timed(fn)
        print("%3d->%3d %-9s %7.1f us  %5.0f GB/s" % (Cin, Cout, name, us, by / us / 1e3))
