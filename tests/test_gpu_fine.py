"""OccHead fine / cascade stage on the GPU (csrc/fine_stage.cu) against the pinned oracle and the reference fixture."""
import numpy as np
import pytest
import torch

import coocc_b200
from coocc_b200 import synthetic as S
from helpers import rel_l2
from oracle import finestage as OF
from oracle.make_golden import FINE_GRID, fine_head_params, fine_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
PCR = [-10.0, -10.0, -5.0, 10.0, 10.0, 3.0]


def _head():
    h = coocc_b200.OccHead(in_channels=[256] * 4, out_channel=17, num_level=4, soft_weights=True,
                           norm_cfg=dict(type="SyncBN", requires_grad=True), cascade_ratio=2, sample_from_voxel=True,
                           sample_from_img=True, final_occ_size=[2 * s for s in FINE_GRID], fine_topk=150,
                           point_cloud_range=PCR).to(DEV).train()
    h.load_state_dict(fine_head_params(), strict=False)
    return h


def test_fine_stage_matches_reference_fixture_and_oracle(golden):
    g = golden("fine")
    coocc_b200.set_precision("fp32")
    try:
        feats, occ, img_feats, transform = fine_inputs()
        head = _head()
        f = feats.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        tr = tuple(t.to(DEV) if torch.is_tensor(t) else t for t in transform)
        torch.manual_seed(123)
        fc, fo = head.forward_fine(f, occ.to(DEV), [img_feats.to(DEV)], tr)
        assert np.array_equal(fc.cpu().numpy().astype(np.int32), g["fine_coord"])          # bit-exact coordinates
        ref = torch.from_numpy(g["fine_output"])
        close = torch.isclose(fo.detach().cpu(), ref, rtol=1e-3, atol=1e-3).all(1)
        assert close.float().mean() > 0.99
        gt = S.make_gt_occ(FINE_GRID, 2, 3).to(DEV)
        ld = head.loss_point(fc, fo, gt, "fine")
        names = ["loss_voxel_ce_fine", "loss_voxel_sem_scal_fine", "loss_voxel_geo_scal_fine", "loss_voxel_lovasz_fine"]
        got = np.array([ld[k].item() for k in names])
        assert np.allclose(got, g["losses"], rtol=1e-3), (got, g["losses"])
        sum(ld.values()).backward()
        # oracle gradients on the same coordinates
        p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fine_head_params().items()}
        fo_ = feats.clone().requires_grad_(True)
        torch.manual_seed(123)
        c2, o2 = OF.fine_forward(p, fo_, occ, img_feats, transform, [2 * s for s in FINE_GRID], torch.tensor(PCR), 2, 150)
        sum(OF.loss_point(c2, o2, gt.cpu()).values()).backward()
        assert rel_l2(f.grad, fo_.grad) < 1e-3
        assert rel_l2(head.fine_mlp[3].weight.grad, p["fine_mlp.3.weight"].grad) < 1e-3
        assert rel_l2(head.img_mlp_0[0].weight.grad.reshape(128, 512), p["img_mlp_0.0.weight"].grad.reshape(128, 512)) < 1e-2
    finally:
        coocc_b200.set_precision("tf32")


@pytest.mark.parametrize("case", [(5000, 64, 16, 1), (3001, 128, 16, 1), (6 * 77, 128, 16, 77), (2 * 640, 64, 16, 640),
                                  (300, 48, 16, 1)])
def test_groupnorm_kernels_match_torch(case):
    """nn.GroupNorm(16) + ReLU on point rows (span 1: the fused one-thread-per-(row, group) kernels) and on NHWC feature
    maps (span = H*W: one block per (sample, group)) against F.group_norm, forward and all three gradients."""
    import torch.nn.functional as F
    import torch.nn as nn
    from coocc_b200 import functional as CF
    rows, C, G, span = case
    gen = torch.Generator().manual_seed(rows + C)
    x0 = torch.randn(rows, C, generator=gen) * 1.5 + 0.3
    gn = nn.GroupNorm(G, C)
    with torch.no_grad():
        gn.weight.copy_(torch.rand(C, generator=gen) + 0.5)
        gn.bias.copy_(torch.randn(C, generator=gen) * 0.2)
    w = torch.randn(rows, C, generator=gen)
    xr = x0.clone().requires_grad_(True)
    if span == 1:
        ref = F.relu(F.group_norm(xr, G, gn.weight, gn.bias, gn.eps))
    else:
        n = rows // span
        ref = F.relu(F.group_norm(xr.reshape(n, span, C).permute(0, 2, 1), G, gn.weight, gn.bias, gn.eps)).permute(0, 2, 1).reshape(rows, C)
    (ref * w).sum().backward()
    ref_dg, ref_db = gn.weight.grad.clone(), gn.bias.grad.clone()
    gn.zero_grad()
    gd = nn.GroupNorm(G, C).to(DEV)
    gd.load_state_dict(gn.state_dict())
    x = x0.to(DEV).requires_grad_(True)
    y = CF.group_norm_rows(x, gd, span=span, relu=True)
    (y * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert rel_l2(y, ref) < 1e-5
    assert rel_l2(x.grad, xr.grad) < 1e-4
    assert rel_l2(gd.weight.grad, ref_dg) < 1e-4 and rel_l2(gd.bias.grad, ref_db) < 1e-4
