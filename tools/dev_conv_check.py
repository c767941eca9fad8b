"""GPU dev check of the tcgen05 conv kernels against torch (cuDNN fp32, TF32 disabled)."""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import coocc_b200
from coocc_b200 import _lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
L = _lib.lib()
dev = "cuda"


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(X, Y, Z, Cin, Cout, k, s, dtype, what=("fwd", "dgrad", "wgrad")):
    g = torch.Generator(device="cpu").manual_seed(X * 7 + Cin + Cout + k)
    x = torch.randn(1, Cin, X, Y, Z, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, k, generator=g) / (Cin * k ** 3) ** 0.5).to(dev)
    tdt = torch.float32 if dtype == 0 else torch.bfloat16
    if dtype == 1:
        x = x.to(tdt).float(); w = w.to(tdt).float()
    x.requires_grad_(True); w.requires_grad_(True)
    y_ref = F.conv3d(x, w, None, s, k // 2)
    oX, oY, oZ = y_ref.shape[2:]
    gy = torch.randn(y_ref.shape, generator=g).to(dev)
    if dtype == 1:
        gy = gy.to(tdt).float()
    y_ref.backward(gy)
    # NDHWC operands
    xc = x.detach().permute(0, 2, 3, 4, 1).reshape(-1, Cin).contiguous().to(tdt)
    wc = w.detach().permute(0, 2, 3, 4, 1).reshape(Cout, -1).contiguous().to(tdt)
    gyc = gy.permute(0, 2, 3, 4, 1).reshape(-1, Cout).contiguous().to(tdt)
    d = _lib.ConvDesc(X, Y, Z, Cin, Cout, k, s, dtype, Cin, Cout)
    tag = "X%dx%dx%d Cin%d Cout%d k%d s%d %s" % (X, Y, Z, Cin, Cout, k, s, "tf32" if dtype == 0 else "bf16")
    res = {}
    if "fwd" in what:
        y = torch.full((oX * oY * oZ, Cout), float("nan"), device=dev)
        stats = torch.zeros(2, Cout, device=dev)
        rc = L.coocc_conv3d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), y.data_ptr(), Cout, None, 0, stats.data_ptr(), stream())
        torch.cuda.synchronize()
        yr = y_ref.detach().permute(0, 2, 3, 4, 1).reshape(-1, Cout)
        err = (y - yr).abs().max().item() / yr.abs().max().item()
        serr = (stats[0] - yr.sum(0)).abs().max().item() / yr.sum(0).abs().max().item()
        qerr = (stats[1] - (yr * yr).sum(0)).abs().max().item() / (yr * yr).sum(0).abs().max().item()
        res["fwd"] = (rc, err, serr, qerr)
        if not (err < 5e-2):
            bad = ((y - yr).abs() > 0.05 * yr.abs().max()) | torch.isnan(y)
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print("   fwd bad rows:", rows[:8].tolist(), "...", rows[-4:].tolist(), "n=", len(rows), " bad cols n=", len(cols), cols[:8].tolist())
            print("   y[0,:4]", y[0, :4].tolist(), "ref", yr[0, :4].tolist())
    if "dgrad" in what and s == 1:
        dx = torch.full((X * Y * Z, Cin), float("nan"), device=dev)
        rc = L.coocc_conv3d_dgrad(ctypes.byref(d), gyc.data_ptr(), wc.data_ptr(), dx.data_ptr(), Cin, stream())
        torch.cuda.synchronize()
        dxr = x.grad.permute(0, 2, 3, 4, 1).reshape(-1, Cin)
        err = (dx - dxr).abs().max().item() / dxr.abs().max().item()
        res["dgrad"] = (rc, err)
        if not (err < 5e-2):
            print("   dx[0,:4]", dx[0, :4].tolist(), "ref", dxr[0, :4].tolist())
    if "wgrad" in what:
        dw = torch.zeros(Cout, k ** 3 * Cin, device=dev)
        rc = L.coocc_conv3d_wgrad(ctypes.byref(d), xc.data_ptr(), gyc.data_ptr(), dw.data_ptr(), stream())
        torch.cuda.synchronize()
        dwr = w.grad.permute(0, 2, 3, 4, 1).reshape(Cout, -1)
        err = (dw - dwr).abs().max().item() / dwr.abs().max().item()
        res["wgrad"] = (rc, err)
        if not (err < 5e-2):
            print("   dw[0,:4]", dw[0, :4].tolist(), "ref", dwr[0, :4].tolist())
    print(tag, " ".join("%s rc=%d err=%s" % (k2, v[0], " ".join("%.2e" % e for e in v[1:])) for k2, v in res.items()), flush=True)


if __name__ == "__main__":  # noqa
    print("lib version", L.coocc_version(), torch.cuda.get_device_name(0))
    cases = [
        (16, 16, 8, 32, 32, 1, 1),      # plain GEMM, single k-block
        (16, 16, 8, 64, 48, 1, 1),
        (16, 16, 8, 32, 32, 3, 1),      # 3x3x3
        (25, 25, 2, 64, 128, 3, 1),     # ragged M tail
        (20, 20, 4, 128, 256, 3, 1),
        (20, 20, 4, 32, 64, 3, 2),      # strided
        (13, 13, 1, 64, 128, 3, 2),
        (20, 20, 4, 64, 128, 1, 2),     # 1x1x1 stride-2 downsample
        (16, 16, 8, 256, 17, 1, 1),     # classifier-like
        (10, 10, 8, 128, 512, 3, 1),    # N > 256 -> two N tiles
    ]
    for dtype in (0, 1):
        for c in cases:
            try:
                run(*c, dtype)
            except Exception as e:  # noqa
                print("EXC", c, dtype, repr(e)[:300], flush=True)
    # timing of a large layer (con_enc-like) in both dtypes
    for dtype in (0, 1):
        X, Y, Z, Cin, Cout = 100, 100, 8, 512, 256
        tdt = torch.float32 if dtype == 0 else torch.bfloat16
        xc = torch.randn(X * Y * Z, Cin, device=dev).to(tdt)
        wc = (torch.randn(Cout, 27 * Cin, device=dev) * 0.01).to(tdt)
        gyc = torch.randn(X * Y * Z, Cout, device=dev).to(tdt)
        y = torch.empty(X * Y * Z, Cout, device=dev)
        dx = torch.empty(X * Y * Z, Cin, device=dev)
        dw = torch.zeros(Cout, 27 * Cin, device=dev)
        d = _lib.ConvDesc(X, Y, Z, Cin, Cout, 3, 1, dtype, Cin, Cout)
        fl = 2.0 * X * Y * Z * 27 * Cin * Cout
        for name, fn in (("fwd", lambda: L.coocc_conv3d_fwd(ctypes.byref(d), xc.data_ptr(), wc.data_ptr(), y.data_ptr(), Cout, None, 0, None, stream())),
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
            print("time %s dtype=%d: %.3f ms  %.1f TFLOP/s" % (name, dtype, ms, fl / ms / 1e9), flush=True)
