"""transposed trilinear resize (coocc_trilinear_bwd): separable axis passes vs the direct gather, values + time."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib(); dev = "cuda"
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (o, s_, C) in (((200, 200, 16), (100, 100, 8), 128), ((200, 200, 16), (50, 50, 4), 128), ((200, 200, 16), (100, 100, 8), 256),
                   ((200, 200, 16), (25, 25, 2), 128)):
    V, Vs = o[0] * o[1] * o[2], s_[0] * s_[1] * s_[2]
    dout = torch.randn(V, C, device=dev).to(torch.bfloat16)
    w = torch.rand(V, 4, device=dev)
    res = {}
    for flags in (0, 1):
        L.coocc_trilinear_tune(flags)
        dsrc = torch.empty(Vs, C, device=dev, dtype=torch.bfloat16)
        f = lambda: L.coocc_trilinear_bwd(dout.data_ptr(), C, o[0], o[1], o[2], C, w.data_ptr(), 4, dsrc.data_ptr(), C,
                                          s_[0], s_[1], s_[2], 1, st())
        assert f() == 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            f()
        b.record(); torch.cuda.synchronize()
        res[flags] = dsrc.float().clone()
        print("%s <- %s C=%d %-9s %.3f ms" % (s_, o, C, "separable" if flags else "direct", a.elapsed_time(b) / 10))
    print("   max |diff| %.3e (max |value| %.3e)" % ((res[0] - res[1]).abs().max().item(), res[0].abs().max().item()))
L.coocc_trilinear_tune(1)
