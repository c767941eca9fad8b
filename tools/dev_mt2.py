import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib(); dev = "cuda"
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for dtype in (1, 0):
    for (X, Y, Z, Cin, Cout, k) in ((200, 200, 16, 128, 64, 1), (200, 200, 16, 64, 4, 1), (200, 200, 16, 64, 17, 1), (200, 200, 16, 256, 128, 3), (100, 100, 8, 128, 128, 3)):
        tdt = torch.bfloat16 if dtype == 1 else torch.float32
        V = X * Y * Z
        xc = torch.randn(V, Cin, device=dev).to(tdt)
        wc = (torch.randn(Cout, k ** 3 * Cin, device=dev) * 0.05).to(tdt)
        ldo = (Cout + 3) // 4 * 4
        y = torch.zeros(V, ldo, device=dev)
        stats = torch.zeros(2, Cout, device=dev)
        d = _lib.ConvDesc(X, Y, Z, Cin, Cout, k, 1, dtype, Cin, (Cout + 7) // 8 * 8)
        rc = L.coocc_conv3d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), y.data_ptr(), ldo, None, 0, stats.data_ptr(), st())
        try:
            torch.cuda.synchronize()
            # spot check a few rows against torch for k=1
            msg = ""
            if k == 1:
                ref = xc[:4096].float() @ wc.float().t()
                msg = "err %.2e" % ((y[:4096, :Cout] - ref).abs().max() / ref.abs().max()).item()
                ref2 = xc[-4096:].float() @ wc.float().t()
                msg += " tail %.2e" % ((y[-4096:, :Cout] - ref2).abs().max() / ref2.abs().max()).item()
            print("ok", dtype, (X, Y, Z, Cin, Cout, k), rc, msg, flush=True)
        except Exception as e:
            print("FAIL", dtype, (X, Y, Z, Cin, Cout, k), rc, str(e)[:80], flush=True)
            sys.exit(1)
