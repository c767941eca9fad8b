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
