"""The work-item bodies of csrc/fine_stage.cuh (OccHead fine / cascade stage kernels) executed on the CPU through
tests/emul/fine_emul.cpp -- the SAME source the CUDA launchers compile -- against torch's grid_sample / group_norm
and the pinned oracle (oracle/finestage.py).  This checks the kernels' arithmetic (indexing, interpolation weights,
projection, normalisation, gradients) without a GPU; launch geometry and atomics are what remains for the B200 run.
The harness is test infrastructure: the product never loads it."""
import ctypes
import os
import subprocess

import pytest
import torch
import torch.nn.functional as F

from oracle import finestage as OF
from oracle.make_golden import FINE_GRID, fine_head_params, fine_inputs

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "fine_emul.cpp")
SO = os.path.join(HERE, "emul", "libfine_emul.so")
PCR = torch.tensor([-10.0, -10.0, -5.0, 10.0, 10.0, 3.0])


@pytest.fixture(scope="module")
def E():
    hdr = os.path.join(HERE, "..", "co-occ_b200", "csrc", "fine_stage.cuh")
    if not os.path.isfile(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    return ctypes.CDLL(SO)


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def sample3d(E, feats5, coords, S3):
    """feats5 [1,C,X,Y,Z] -> [M,C] through the kernel body."""
    _, C, X, Y, Z = feats5.shape
    rows = feats5.permute(0, 2, 3, 4, 1).reshape(-1, C).contiguous()
    c = coords.to(torch.int32).contiguous()
    M = c.shape[1]
    out = torch.empty(M, C)
    E.emul_sample3d_fwd(P(rows), ctypes.c_longlong(C), X, Y, Z, C, P(c), M, S3[0], S3[1], S3[2], P(out), ctypes.c_longlong(C))
    return out, rows, c


def ref_sample3d(feats5, coords, S3):
    g = coords.float()
    g = torch.stack([(g[i] / (S3[i] - 1) - 0.5) * 2 for i in range(3)], 0)
    grid = g[None, None, None].permute(0, 4, 1, 2, 3)
    v = F.grid_sample(feats5.permute(0, 1, 4, 3, 2), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return v[0, :, :, 0, 0].permute(1, 0)


def test_sample3d_forward_backward(E):
    gen = torch.Generator().manual_seed(0)
    grid = (9, 7, 5)
    S3 = [2 * s for s in grid]
    feats = torch.randn(1, 16, *grid, generator=gen).requires_grad_(True)
    coords = torch.stack([torch.randint(0, S3[i], (300,), generator=gen) for i in range(3)])
    coords[:, :4] = torch.tensor([[0, S3[0] - 1, 0, S3[0] - 1], [0, 0, S3[1] - 1, S3[1] - 1], [0, S3[2] - 1, 0, S3[2] - 1]])
    out, rows, c = sample3d(E, feats.detach(), coords, S3)
    ref = ref_sample3d(feats, coords, S3)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    w = torch.randn(300, 16, generator=gen)
    (ref * w).sum().backward()
    d = torch.zeros_like(rows)
    E.emul_sample3d_bwd(P(w), ctypes.c_longlong(16), *grid, 16, P(c), 300, *S3, P(d), ctypes.c_longlong(16))
    dref = feats.grad.permute(0, 2, 3, 4, 1).reshape(-1, 16)
    assert torch.allclose(d, dref, rtol=1e-4, atol=1e-6)


def project(E, coords, transform, grid_fine):
    rots, trans, intr, prot, ptr, bda = (t[0] for t in transform[:6])
    n = rots.shape[0]
    vs = (PCR[3:] - PCR[:3]) / torch.tensor([grid_fine[0] - 1, grid_fine[1] - 1, grid_fine[2] - 1])
    cam = torch.cat([rots.inverse().reshape(n, 9), trans.reshape(n, 3), intr.reshape(n, 9), prot[:, :2, :2].reshape(n, 4),
                     ptr[:, :2].reshape(n, 2)], 1).contiguous()
    inv_bda = bda.inverse().contiguous()
    c = coords.to(torch.int32).contiguous()
    M = c.shape[1]
    uv = torch.empty(n, M, 2)
    mask = torch.empty(M, n, dtype=torch.uint8)
    f3 = lambda t: (ctypes.c_float * 3)(*[float(v) for v in t.tolist()])
    E.emul_project(P(c), M, n, f3(vs), f3(PCR[:3]), P(inv_bda), P(cam), ctypes.c_float(float(transform[-1][1][0])),
                   ctypes.c_float(float(transform[-1][0][0])), P(uv), P(mask))
    return uv, mask


def test_projection_matches_oracle(E):
    feats, occ, img_feats, transform = fine_inputs()
    gen = torch.Generator().manual_seed(1)
    gf = [2 * s for s in FINE_GRID]
    coords = torch.stack([torch.randint(0, gf[i], (500,), generator=gen) for i in range(3)])
    uv, mask = project(E, coords, transform, gf)
    uv_o, m_o = OF.project_points_on_img(coords[None].permute(0, 2, 1).float().contiguous(), transform[0][0:1],
                                         transform[1][0:1], transform[2][0:1], transform[3][0:1], transform[4][0:1],
                                         transform[5][0:1], PCR, transform[-1][1][0:1], transform[-1][0][0:1], *gf)
    uv_o = uv_o[:, :, 0, :]                       # [ncam, M, 2]
    m_o = m_o.reshape(500, -1)                    # [M, ncam]
    agree = mask.bool() == m_o
    assert agree.float().mean() > 0.995           # a point exactly on an image border may flip with fp32 rounding
    vis = mask.bool() & m_o
    assert vis.sum() > 50
    sel = vis.t()                                 # [ncam, M]
    assert torch.allclose(uv[sel], uv_o[sel], rtol=1e-4, atol=1e-4)


def test_sample2d_forward_backward(E):
    gen = torch.Generator().manual_seed(2)
    n, C, H, W, M = 3, 8, 6, 11, 200
    img = torch.randn(n, C, H, W, generator=gen).requires_grad_(True)
    uv = torch.rand(n, M, 2, generator=gen) * 2.4 - 1.2
    mask = (torch.rand(M, n, generator=gen) < 0.7).to(torch.uint8)
    ref = F.grid_sample(img, uv[:, :, None, :], align_corners=True, mode="bilinear", padding_mode="zeros")   # [n,C,M,1]
    ref = (ref * mask.t()[:, None, :, None]).sum(0)[:, :, 0].permute(1, 0)
    rows = img.detach().permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    out = torch.empty(M, C)
    uvc = uv.contiguous()
    E.emul_sample2d_fwd(P(rows), ctypes.c_longlong(C), n, H, W, C, P(uvc), P(mask), M, P(out), ctypes.c_longlong(C))
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    w = torch.randn(M, C, generator=gen)
    (ref * w).sum().backward()
    d = torch.zeros_like(rows)
    E.emul_sample2d_bwd(P(w), ctypes.c_longlong(C), n, H, W, C, P(uvc), P(mask), M, P(d), ctypes.c_longlong(C))
    assert torch.allclose(d, img.grad.permute(0, 2, 3, 1).reshape(-1, C), rtol=1e-4, atol=1e-6)


def groupnorm(E, x, G, span, gamma, beta, relu=True):
    rows, C = x.shape
    stats = torch.empty(rows // span * G * 2)
    y = torch.empty_like(x)
    E.emul_groupnorm_fwd(P(x), ctypes.c_longlong(C), ctypes.c_longlong(rows), C, G, span, P(gamma), P(beta),
                         ctypes.c_float(1e-5), 1 if relu else 0, P(stats), P(y), ctypes.c_longlong(C))
    return y, stats


def groupnorm_bwd(E, x, G, span, gamma, beta, stats, dy, relu=True):
    rows, C = x.shape
    sums = torch.empty(rows // span * G * 2)
    dx = torch.empty_like(x)
    dg, db = torch.zeros(C), torch.zeros(C)
    E.emul_groupnorm_bwd(P(x), ctypes.c_longlong(C), ctypes.c_longlong(rows), C, G, span, P(gamma), P(beta), 1 if relu else 0,
                         P(stats), P(dy), ctypes.c_longlong(C), P(sums), P(dx), ctypes.c_longlong(C), P(dg), P(db))
    return dx, dg, db


@pytest.mark.parametrize("span", [1, 35])
def test_groupnorm_relu_forward_backward(E, span):
    gen = torch.Generator().manual_seed(3)
    C, G = 64, 16
    nsamp = 40 if span == 1 else 3
    x = torch.randn(nsamp * span, C, generator=gen).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(C, generator=gen)).requires_grad_(True)
    beta = (0.2 * torch.randn(C, generator=gen)).requires_grad_(True)
    if span == 1:
        ref = F.relu(F.group_norm(x, G, gamma, beta, 1e-5))
    else:      # NHWC rows of nsamp feature maps with span pixels each -> [n, C, span]
        ref = F.relu(F.group_norm(x.reshape(nsamp, span, C).permute(0, 2, 1), G, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(-1, C)
    y, stats = groupnorm(E, x.detach().contiguous(), G, span, gamma.detach(), beta.detach())
    assert torch.allclose(y, ref, rtol=1e-4, atol=1e-5)
    w = torch.randn(nsamp * span, C, generator=gen)
    (ref * w).sum().backward()
    dx, dg, db = groupnorm_bwd(E, x.detach().contiguous(), G, span, gamma.detach(), beta.detach(), stats, w.contiguous())
    assert torch.allclose(dx, x.grad, rtol=1e-3, atol=1e-5)
    assert torch.allclose(dg, gamma.grad, rtol=1e-3, atol=1e-4) and torch.allclose(db, beta.grad, rtol=1e-3, atol=1e-4)


def test_whole_fine_stage_through_the_kernel_bodies_matches_reference_fixture(E, golden):
    """occ_head.py:182-237 composed from the kernel bodies (+ torch for the Linear layers, which the product runs on
    the tensor-core conv kernel) against the reference fixture."""
    g = golden("fine")
    feats, occ, img_feats, transform = fine_inputs()
    p = fine_head_params()
    gf = [2 * s for s in FINE_GRID]
    coords = torch.from_numpy(g["fine_coord"]).long()                       # the reference's own subset (randperm draw)
    vox, _, _ = sample3d(E, feats, coords, gf)
    n, Ci, Hf, Wf = img_feats.shape[1:]
    rows = img_feats[0].permute(0, 2, 3, 1).reshape(-1, Ci)
    f2d = F.linear(rows, p["img_mlp_0.0.weight"].reshape(128, Ci), p["img_mlp_0.0.bias"]).contiguous()
    f2d, _ = groupnorm(E, f2d, 16, Hf * Wf, p["img_mlp_0.1.weight"], p["img_mlp_0.1.bias"])
    uv, mask = project(E, coords, transform, gf)
    M = coords.shape[1]
    samp = torch.empty(M, 128)
    E.emul_sample2d_fwd(P(f2d), ctypes.c_longlong(128), n, Hf, Wf, 128, P(uv), P(mask), M, P(samp), ctypes.c_longlong(128))
    a = F.linear(samp, p["img_mlp.0.weight"], p["img_mlp.0.bias"]).contiguous()
    a, _ = groupnorm(E, a, 16, 1, p["img_mlp.1.weight"], p["img_mlp.1.bias"])
    h = F.linear(torch.cat([vox, a], 1), p["fine_mlp.0.weight"], p["fine_mlp.0.bias"]).contiguous()
    h, _ = groupnorm(E, h, 16, 1, p["fine_mlp.1.weight"], p["fine_mlp.1.bias"])
    out = F.linear(h, p["fine_mlp.3.weight"], p["fine_mlp.3.bias"])
    ref = torch.from_numpy(g["fine_output"])
    close = torch.isclose(out, ref, rtol=1e-3, atol=1e-3).all(1)
    assert close.float().mean() > 0.99, close.float().mean()    # points whose projection sits on an image border may differ
