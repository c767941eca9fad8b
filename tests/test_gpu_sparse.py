"""Sparse LiDAR encoder on the GPU (co-occ_b200/sparse_enc.py, csrc/sparse_conv.cu) against the masked-dense oracle
(oracle/sparse_enc.py): the two convolution types on random features (forward + all gradients), the whole
SparseLiDAREnc8x built from the config's dict, and the encoder feeding BiFuser_N."""
import numpy as np
import pytest
import torch

import coocc_b200
from coocc_b200 import functional as CF
from coocc_b200 import registry
from coocc_b200 import sparse_enc as SE
from coocc_b200 import synthetic as S
from helpers import rel_err, rel_l2
from oracle import sparse_enc as OS

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _fp32():
    coocc_b200.set_precision("fp32")
    yield
    coocc_b200.set_precision("tf32")


def _dense_case(seed, dims, n, cin):
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    lin = torch.randperm(D * H * W, generator=g)[:n].sort().values
    zyx = torch.stack([lin // (H * W), (lin // W) % H, lin % W], 1)
    feats = torch.randn(n, cin, generator=g)
    m = torch.zeros(1, 1, D, H, W, dtype=torch.bool)
    m[0, 0, zyx[:, 0], zyx[:, 1], zyx[:, 2]] = True
    coords = torch.cat([torch.zeros(n, 1, dtype=torch.long), zyx], 1).int()
    return feats, coords, m


@pytest.mark.parametrize("strided", [False, True])
def test_sparse_conv_rows_matches_dense_oracle(strided):
    dims, n, cin, cout = (12, 20, 18), 900, 16, 32
    feats, coords, m = _dense_case(5, dims, n, cin)
    g = torch.Generator().manual_seed(9)
    w = torch.randn(cout, 3, 3, 3, cin, generator=g) * 0.1
    # oracle
    fo, wo = feats.clone().requires_grad_(True), w.clone().requires_grad_(True)
    x = torch.zeros(1, cin, *dims)
    x[0][:, m[0, 0]] = fo.t()
    if strided:
        y, m2 = OS.strided_conv(x, m, wo)
    else:
        y, m2 = OS.subm_conv(x, m, wo), m
    ro = OS.rows_of(y, m2)
    gy = torch.randn(ro.shape, generator=g)
    (ro * gy).sum().backward()
    # CUDA
    lvl = SE.SpLevel(coords.to(DEV), dims)
    conv = SE._SpConv(cin, cout, False).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(w)
    f = feats.to(DEV).requires_grad_(True)
    if strided:
        nxt, nbr = lvl.downsample()
        assert np.array_equal(nxt.coords[:, 1:].cpu().numpy(), torch.nonzero(m2[0, 0]).numpy())      # bit-exact sites
    else:
        nbr = lvl.subm_table()
    r = SE.sparse_conv_rows(f, nbr, conv)
    (r * gy.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(r, ro) < 1e-5
    assert rel_l2(f.grad, fo.grad) < 1e-5 and rel_l2(conv.weight.grad, wo.grad) < 1e-5


def test_encoder_from_config_matches_oracle():
    cfg = dict(type='SparseLiDAREnc8x', input_channel=4, base_channel=16, out_channel=128,
               norm_cfg=dict(type='SyncBN', requires_grad=True), sparse_shape_xyz=[48, 40, 16])     # config :127-134
    enc = registry.build_middle_encoder(cfg).to(DEV).train()
    P = S.sparse_enc_params()
    enc.load_state_dict(P, strict=True)
    feats, coors = S.make_lidar_voxels(cfg["sparse_shape_xyz"], 2500, seed=2)
    po = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P.items()}
    out_o, m_o = OS.sparse_encoder_forward(po, feats, coors, cfg["sparse_shape_xyz"])
    w = torch.linspace(-1, 1, out_o.numel()).reshape(out_o.shape)
    (out_o * w).sum().backward()
    out = enc(feats.to(DEV), coors.to(DEV), 1)
    x = out['x']
    assert tuple(x.shape) == (1, 128, 6, 5, 2) == tuple(out_o.shape)          # [B, C, W/8, H/8, D/8]
    (x * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert torch.equal((x[0].abs().sum(0) != 0).cpu(), (out_o[0].abs().sum(0) != 0))              # same active sites
    assert rel_err(x, out_o) < 2e-3
    for k in ("conv_out.0.weight", "conv3.2.net.3.weight", "conv2.0.0.weight", "conv1.1.net.1.weight", "conv_out.1.bias"):
        mod = enc
        for part in k.split(".")[:-1]:
            mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
        got = getattr(mod, k.split(".")[-1]).grad
        assert rel_l2(got, po[k].grad) < 2e-2, k
    assert int(enc.conv1[0][1].num_batches_tracked) == 1


def test_encoder_output_feeds_the_fuser():
    """extract_pts_feat -> occ_fuser (coocc_ray.py:215-253): the encoder's strided [1,C,W,H,D] view is consumed by
    BiFuser_N as is."""
    shape = [200, 200, 64]                               # -> 25 x 25 x 8 at stride 8
    enc = SE.SparseLiDAREnc8x(4, dict(type='SyncBN'), 16, 128, shape).to(DEV).train()
    enc.load_state_dict(S.sparse_enc_params(), strict=True)
    feats, coors = S.make_lidar_voxels(shape, 60000, seed=3)
    pts = enc(feats.to(DEV), coors.to(DEV), 1)['x']
    assert tuple(pts.shape) == (1, 128, 25, 25, 8)
    n_pts = int((pts.sum(1) != 0).sum())
    assert n_pts > 2048
    img = torch.randn(1, 128, 25, 25, 8, device=DEV) * 0.3
    fuser = coocc_b200.BiFuser_N(128, 128, knum=2).to(DEV).train()
    out = fuser(img, pts)
    out.sum().backward()
    assert tuple(out.shape) == (1, 128, 25, 25, 8) and torch.isfinite(out).all()
    assert enc.conv_out[1].bias.grad is not None and float(enc.conv_out[1].bias.grad.abs().sum()) > 0
