"""Per-op GPU checks of the storage-type variants and the fused multi-level resize:
  * resize_mix (csrc/trilinear.cu) against F.interpolate(mode='trilinear', align_corners=False) +
    the weighted sum of P/coocc/dense_heads/occ_head.py:161-165 and the top-down add of
    P/coocc/necks/fpn3d.py:91-94, forward and backward, fp32 and bf16 storage;
  * BatchNorm+ReLU+residual kernels with bf16 storage against the fp32-storage kernels;
  * conv epilogue bf16 output against the fp32 output rounded to bf16 (bit-exact).
Tolerances: fp32 storage 1e-5 (same arithmetic, different summation order); bf16 storage 2^-8
relative to the tensor maximum (one rounding of inputs and one of the output)."""
import pytest
import torch
import torch.nn.functional as F

import coocc_b200
from coocc_b200 import functional as CF
from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rows(x5):
    return CF.to_cl2d(x5)[0]


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("case", ["occhead", "fpn"])
def test_resize_mix_matches_interpolate(case, bf16):
    torch.manual_seed(0)
    C = 64
    d0 = (12, 10, 8)
    dims = [d0, (6, 5, 4), (3, 3, 2), (2, 2, 1)] if case == "occhead" else [(6, 5, 4)]
    dt = torch.bfloat16 if bf16 else torch.float32
    srcs5 = [torch.randn(1, C, *d, device=DEV).to(dt).float().requires_grad_(True) for d in dims]
    V0 = d0[0] * d0[1] * d0[2]
    if case == "occhead":
        wts = torch.softmax(torch.randn(V0, len(dims), device=DEV), 1).requires_grad_(True)
        base5 = None
    else:
        wts = None
        base5 = torch.randn(1, C, *d0, device=DEV).to(dt).float().requires_grad_(True)
    # reference: torch ops in fp32
    ref = 0
    for l, s5 in enumerate(srcs5):
        r = F.interpolate(s5, size=list(d0), mode="trilinear", align_corners=False)
        r2 = _rows(r)
        ref = ref + (r2 * wts[:, l:l + 1] if wts is not None else r2)
    if base5 is not None:
        ref = ref + _rows(base5)
    g = torch.randn(V0, C, device=DEV).to(dt).float()
    (ref * g).sum().backward()
    ref_grads = [s5.grad.clone() for s5 in srcs5]
    ref_gw = wts.grad.clone() if wts is not None else None
    ref_gb = base5.grad.clone() if base5 is not None else None
    # ours
    srcs2 = [_rows(s5.detach()).to(dt).requires_grad_(True) for s5 in srcs5]
    w2 = wts.detach().clone().requires_grad_(True) if wts is not None else None
    b2 = _rows(base5.detach()).to(dt).requires_grad_(True) if base5 is not None else None
    out = CF.resize_mix(srcs2, dims, d0, base=b2, wts=w2)
    assert out.dtype == dt
    (out.float() * g).sum().backward()
    tol = 8e-3 if bf16 else 1e-5
    assert rel_err(out.float(), ref.detach()) < tol
    for l, (s2, rg) in enumerate(zip(srcs2, ref_grads)):
        assert s2.grad.dtype == dt
        assert rel_err(s2.grad.float(), _rows(rg)) < tol, "dsrc level %d" % l
    if w2 is not None:
        assert rel_err(w2.grad, ref_gw) < (2e-2 if bf16 else 1e-5)
    if b2 is not None:
        assert rel_err(b2.grad.float(), _rows(ref_gb)) < tol


@pytest.mark.parametrize("res", [False, True])
def test_bn_act_bf16_storage_matches_fp32_storage(res):
    torch.manual_seed(1)
    V, C = 4096, 64
    x = torch.randn(V, C, device=DEV).to(torch.bfloat16)
    r = torch.randn(V, C, device=DEV).to(torch.bfloat16) if res else None
    gamma = (torch.rand(C, device=DEV) + 0.5).requires_grad_(True)
    beta = torch.randn(C, device=DEV).requires_grad_(True)
    g = torch.randn(V, C, device=DEV).to(torch.bfloat16)

    def run(dt):
        xx = x.to(dt).requires_grad_(True)
        rr = r.to(dt).requires_grad_(True) if res else None
        xf = xx.detach().float()
        stats = torch.stack([xf.sum(0), (xf * xf).sum(0)]).contiguous()
        gm, bt = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
        out = CF.bn_act(xx, stats, gm, bt, residual=rr, relu=True)
        assert out.dtype == dt
        # identical ReLU masks in both runs: take the mask from the fp32-storage result
        out.backward(g.to(dt))
        return out, xx.grad, (rr.grad if res else None), gm.grad, bt.grad

    o32, dx32, dr32, dg32, db32 = run(torch.float32)
    o16, dx16, dr16, dg16, db16 = run(torch.bfloat16)
    assert dx16.dtype == torch.bfloat16
    # elements whose pre-activation is within bf16 rounding of zero may flip the ReLU mask
    assert rel_err(o16.float(), o32) < 8e-3
    bad = ((o16.float() > 0) != (o32 > 0)).float().mean().item()
    assert bad < 5e-3
    assert rel_err(dg16, dg32) < 3e-2 and rel_err(db16, db32) < 3e-2
    from helpers import rel_l2
    assert rel_l2(dx16.float(), dx32) < 3e-2
    if res:
        assert rel_l2(dr16.float(), dr32) < 3e-2


@pytest.mark.parametrize("shape", [(40, 40, 16, 64, 64, 3), (9, 7, 5, 32, 48, 3), (30, 20, 8, 64, 40, 1)])
def test_conv_epilogue_bf16_output_is_rounded_fp32_output(shape):
    X, Y, Z, Cin, Cout, k = shape
    coocc_b200.set_precision("bf16")
    try:
        torch.manual_seed(2)
        x = torch.randn(X * Y * Z, Cin, device=DEV).to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, k, k, k, device=DEV) / (Cin * k ** 3) ** 0.5)
        w = w.contiguous(memory_format=torch.channels_last_3d)
        y32 = CF.conv3d(x.float(), w, (X, Y, Z), k, 1)
        y16, stats = CF.conv3d(x, w, (X, Y, Z), k, 1, want_stats=True, out_bf16=True)
        assert y16.dtype == torch.bfloat16 and y32.dtype == torch.float32
        # split-K layers sum partials in a different order -> allow one bf16 ulp there
        diff = (y16.float() - y32.to(torch.bfloat16).float()).abs().max().item()
        assert diff <= y32.abs().max().item() * 2 ** -7
        # statistics come from the fp32 accumulators, not from the rounded output
        assert rel_err(stats[0], y32.sum(0)) < 1e-3
        assert rel_err(stats[1], (y32 * y32).sum(0)) < 1e-3
    finally:
        coocc_b200.set_precision("tf32")


@pytest.mark.parametrize("shape", [(40, 40, 16, 64, 64, 3, 1), (64, 60, 10, 32, 64, 3, 1), (30, 20, 8, 64, 40, 1, 1),
                                   (26, 22, 8, 32, 64, 3, 2), (100, 100, 8, 128, 128, 3, 1)])
def test_conv_dynamic_tile_scheduler_matches_static(shape):
    """coocc_conv_set_dynamic(1): tiles handed out by an atomic counter instead of the static round-robin.  fwd and
    dgrad without split-K are bit-identical (the same tiles, computed the same way, by other CTAs); statistics and
    wgrad sum in a different order."""
    import ctypes
    from coocc_b200 import _lib
    L = _lib.lib()
    X, Y, Z, Cin, Cout, k, s = shape
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    V = X * Y * Z
    od = [CF.out_dim(n, k, s) for n in (X, Y, Z)]
    Vo = od[0] * od[1] * od[2]
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(V, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, k ** 3 * Cin, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    gy = torch.randn(Vo, Cout, device="cuda", generator=g).to(torch.bfloat16)
    d = _lib.ConvDesc(X, Y, Z, Cin, Cout, k, s, 1, Cin, Cout, 1)
    res = []
    try:
        for dyn in (0, 1, 1):
            L.coocc_conv_set_dynamic(dyn)
            y = torch.empty(Vo, Cout, device="cuda", dtype=torch.bfloat16)
            dx = torch.zeros(V, Cin, device="cuda", dtype=torch.bfloat16)
            dw = torch.zeros(Cout, k ** 3 * Cin, device="cuda")
            stats = torch.zeros(2, Cout, device="cuda")
            _lib.check(L.coocc_conv3d_fwd(ctypes.byref(d), x.data_ptr(), w.data_ptr(), y.data_ptr(), Cout, None, 0,
                                          stats.data_ptr(), st()), "fwd")
            _lib.check(L.coocc_conv3d_dgrad(ctypes.byref(d), gy.data_ptr(), w.data_ptr(), dx.data_ptr(), Cin, st()), "dgrad")
            _lib.check(L.coocc_conv3d_wgrad(ctypes.byref(d), x.data_ptr(), gy.data_ptr(), dw.data_ptr(), st()), "wgrad")
            torch.cuda.synchronize()
            res.append((y.float(), dx.float(), dw, stats))
    finally:
        L.coocc_conv_set_dynamic(0)
    for r in res[1:]:
        assert (r[0] - res[0][0]).abs().max() <= 1e-2 * res[0][0].abs().max()      # (split-K layers: fp32 atomics order)
        assert (r[1] - res[0][1]).abs().max() <= 1e-2 * res[0][1].abs().max()
        assert (r[2] - res[0][2]).abs().max() <= 1e-4 * res[0][2].abs().max()
        assert (r[3] - res[0][3]).abs().max() <= 1e-4 * res[0][3].abs().max()
    if shape[0] == 100:        # big enough for the unsplit persistent path: identical bits
        assert torch.equal(res[1][0], res[0][0]) and torch.equal(res[1][1], res[0][1])
