"""The masked-dense restatement of spconv's SubMConv3d / SparseConv3d (oracle/sparse_enc.py) against a rulebook-style
evaluation straight from the definition, and the encoder's state_dict layout.  (Parity against spconv itself is
unpinned: the package is third party and not available -- see the oracle's header.)"""
import numpy as np
import torch

import coocc_b200
from coocc_b200 import synthetic as S
from oracle import sparse_enc as OS


def _case(seed, dims=(6, 7, 5), n=40, cin=3, cout=5):
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    lin = torch.randperm(D * H * W, generator=g)[:n].sort().values
    zyx = torch.stack([lin // (H * W), (lin // W) % H, lin % W], 1)
    feats = torch.randn(n, cin, generator=g)
    w = torch.randn(cout, 3, 3, 3, cin, generator=g)
    b = torch.randn(cout, generator=g)
    m = torch.zeros(1, 1, D, H, W, dtype=torch.bool)
    m[0, 0, zyx[:, 0], zyx[:, 1], zyx[:, 2]] = True
    x = torch.zeros(1, cin, D, H, W)
    x[0][:, zyx[:, 0], zyx[:, 1], zyx[:, 2]] = feats.t()
    return dims, zyx, feats, w, b, m, x


def test_subm_matches_rulebook_definition():
    dims, zyx, feats, w, b, m, x = _case(0)
    y = OS.subm_conv(x, m, w, b)
    oc, of = OS.brute_force_conv(feats.numpy(), zyx.numpy(), dims, w.numpy(), 1, 1, True, b.numpy())
    assert np.array_equal(oc, torch.nonzero(m[0, 0]).numpy())            # no dilation: output sites = input sites
    np.testing.assert_allclose(OS.rows_of(y, m).numpy(), of, rtol=1e-4, atol=1e-5)
    assert float((y * ~m).abs().max()) == 0.0


def test_strided_conv_matches_rulebook_definition():
    dims, zyx, feats, w, b, m, x = _case(1, dims=(8, 6, 7), n=30)
    y, m2 = OS.strided_conv(x, m, w)
    oc, of = OS.brute_force_conv(feats.numpy(), zyx.numpy(), dims, w.numpy(), 2, 1, False)
    assert tuple(m2.shape[2:]) == (4, 3, 4)
    assert np.array_equal(oc, torch.nonzero(m2[0, 0]).numpy())           # every site reached by an active input
    np.testing.assert_allclose(OS.rows_of(y, m2).numpy(), of, rtol=1e-4, atol=1e-5)


def test_encoder_state_dict_has_spconv_names_and_layout():
    enc = coocc_b200.sparse_enc.SparseLiDAREnc8x(input_channel=4, norm_cfg=dict(type="SyncBN", requires_grad=True),
                                                 base_channel=16, out_channel=128, sparse_shape_xyz=[800, 800, 64])
    sd = enc.state_dict()
    ref = S.sparse_enc_params()
    assert set(ref) <= set(sd), sorted(set(ref) - set(sd))
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert tuple(sd["conv_input.0.weight"].shape) == (16, 3, 3, 3, 4) and "conv_input.0.bias" in sd
    assert tuple(sd["conv3.2.net.3.weight"].shape) == (128, 3, 3, 3, 128) and "conv3.2.net.3.bias" not in sd
    assert "conv2.0.1.num_batches_tracked" in sd and "conv_out.1.bias" in sd
    enc.load_state_dict(ref, strict=True)


def test_degenerate_input_groupnorm_is_reproduced():
    """Q13: conv_input's nn.GroupNorm(16, 16) sees one channel per group on [N, C] rows -> its output is relu(bias)
    whatever the LiDAR features are; the oracle (like the reference) propagates only the occupancy pattern."""
    p = S.sparse_enc_params()
    f, c = S.make_lidar_voxels([16, 16, 8], 120, seed=1)
    a, _ = OS.sparse_encoder_forward(p, f, c, [16, 16, 8])
    b, _ = OS.sparse_encoder_forward(p, f * 3.0 + 1.0, c, [16, 16, 8])
    # (up to the rounding noise of torch's x * scale + shift form of the normalisation, scale = weight / sqrt(eps) ~ 300)
    assert torch.allclose(a, b, atol=2e-3) and float(a.abs().max()) > 0.1
