"""A few launches of tc_conv_kernel on the 1x1x1 layer shapes of the north-star step (for ncu captures)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib(); dev = "cuda"
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
X, Y, Z = 200, 200, 16
for (Cin, Cout) in ((128, 64), (128, 256)):
    xc = torch.randn(X * Y * Z, Cin, device=dev).to(torch.bfloat16)
    wc = (torch.randn(Cout, Cin, device=dev) * 0.05).to(torch.bfloat16)
    gy = torch.randn(X * Y * Z, Cout, device=dev).to(torch.bfloat16)
    y = torch.empty(X * Y * Z, Cout, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(X * Y * Z, Cin, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(Cout, Cin, device=dev); stats = torch.zeros(2, Cout, device=dev)
    d = _lib.ConvDesc(X, Y, Z, Cin, Cout, 1, 1, 1, Cin, Cout, 1)
    for _ in range(2):
        L.coocc_conv3d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), y.data_ptr(), Cout, None, 0, stats.data_ptr(), st())
        L.coocc_conv3d_dgrad(ctypes.byref(d), gy.data_ptr(), wc.data_ptr(), dx.data_ptr(), Cin, st())
        L.coocc_conv3d_wgrad(ctypes.byref(d), xc.data_ptr(), gy.data_ptr(), dw.data_ptr(), st())
    torch.cuda.synchronize()
print("done")
