"""dgrad of the con_enc convolution (512->256, 3x3x3, 200x200x16) with fp32 and with bf16 output rows."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib(); dev = "cuda"
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
X, Y, Z, Cin, Cout = 200, 200, 16, 512, 256
V = X * Y * Z
wc = (torch.randn(Cout, 27 * Cin, device=dev) * 0.01).to(torch.bfloat16)
gy = torch.randn(V, Cout, device=dev).to(torch.bfloat16)
for out_bf16 in (0, 1):
    dx = torch.empty(V, Cin, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    d = _lib.ConvDesc(X, Y, Z, Cin, Cout, 3, 1, 1, Cin, Cout, out_bf16)
    f = lambda: L.coocc_conv3d_dgrad(ctypes.byref(d), gy.data_ptr(), wc.data_ptr(), dx.data_ptr(), Cin, st())
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("dgrad 512->256 out_bf16=%d: %.3f ms  %.0f TFLOP/s" % (out_bf16, ms, 2.0 * V * 27 * Cin * Cout / ms / 1e9))
