"""GPU dev check + A/B timing of the ky-fused 3x3x3 conv path (coocc_conv_tune) against the plain
im2col path and torch (cuDNN fp32, TF32 disabled)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from dev_conv_check import run, L, stream, dev
from coocc_b200 import _lib

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    cases = [
        (32, 40, 16, 64, 64, 3, 1),       # Z=16, Y % 8 == 0
        (40, 30, 16, 96, 80, 3, 1),       # ragged Y (30 -> 4 blocks of 8), ragged channels
        (24, 100, 8, 128, 128, 3, 1),     # Z=8, 100 -> 7 blocks of 16
        (64, 40, 16, 128, 256, 3, 1),     # BN=256, MT=2 single accumulator
        (40, 40, 16, 64, 512, 3, 1),      # two N tiles
        (64, 40, 16, 256, 128, 3, 1),     # 4 k-blocks per tap
    ]
    what = ("fwd", "dgrad", "wgrad")
    for mt in (2, 4):
        L.coocc_conv_tune(1, mt)
        print("== ky path, MT=%d" % mt)
        for dtype in (1, 0):
            for c in cases:
                try:
                    run(*c, dtype, what=what)
                except Exception as e:  # noqa
                    print("EXC", c, dtype, repr(e)[:300], flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "check":
        sys.exit(0)
    # ---- timing
    shapes = [(200, 200, 16, 128, 128), (200, 200, 16, 512, 256), (200, 200, 16, 256, 256), (200, 200, 16, 256, 128),
              (100, 100, 8, 256, 256), (100, 100, 8, 128, 256), (100, 100, 8, 512, 256)]
    dtype = 1
    for (X, Y, Z, Cin, Cout) in shapes:
        xc = torch.randn(X * Y * Z, Cin, device=dev).to(torch.bfloat16)
        wc = (torch.randn(Cout, 27 * Cin, device=dev) * 0.01).to(torch.bfloat16)
        gyc = torch.randn(X * Y * Z, Cout, device=dev).to(torch.bfloat16)
        y = torch.empty(X * Y * Z, Cout, device=dev)
        dx = torch.empty(X * Y * Z, Cin, device=dev)
        stats = torch.zeros(2, Cout, device=dev)
        dw = torch.zeros(Cout, 27 * Cin, device=dev)
        d = _lib.ConvDesc(X, Y, Z, Cin, Cout, 3, 1, dtype, Cin, Cout)
        fl = 2.0 * X * Y * Z * 27 * Cin * Cout
        for (ky, mt) in ((0, 2), (1, 2), (1, 4)):
            L.coocc_conv_tune(ky, mt)
            out = []
            for name, fn in (("fwd", lambda: L.coocc_conv3d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), y.data_ptr(), Cout, None, 0, stats.data_ptr(), stream())),
                             ("dgrad", lambda: L.coocc_conv3d_dgrad(ctypes.byref(d), gyc.data_ptr(), wc.data_ptr(), dx.data_ptr(), Cin, stream())),
                             ("wgrad", lambda: L.coocc_conv3d_wgrad(ctypes.byref(d), xc.data_ptr(), gyc.data_ptr(), dw.data_ptr(), stream()))):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                out.append("%s %.3f ms %6.0f TF/s" % (name, ms, fl / ms / 1e9))
            print("%dx%dx%d %d->%d ky=%d mt=%d: %s" % (X, Y, Z, Cin, Cout, ky, mt, "  ".join(out)), flush=True)
