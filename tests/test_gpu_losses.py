"""csrc/occ_loss.cu (OccHead.loss_voxel: label vote, CE, sem_scal, geo_scal, Lovasz-softmax) through the
C ABI against the oracle restatement (oracle/losses.py) and the reference fixtures; full-size checks
through properties (the oracle's torch code runs on the GPU there as an independent evaluation)."""
import numpy as np
import pytest
import torch

import coocc_b200
from coocc_b200 import functional as CF
from coocc_b200 import synthetic as S
from helpers import rel_l2, sample, stats
from oracle import losses as OL

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAMES = ["loss_voxel_ce_c_0", "loss_voxel_sem_scal_c_0", "loss_voxel_geo_scal_c_0", "loss_voxel_lovasz_c_0"]


def _c1_inputs():
    grid = S.CONFIGS["c1"]["grid"]
    gen = torch.Generator().manual_seed(2024)
    logits = torch.randn(1, 17, *grid, generator=gen) * 2.0
    logits[:, 0] += 1.5
    return logits, S.make_gt_occ(grid, 2, seed=0)


def _head():
    return coocc_b200.OccHead(in_channels=[64] * 4, out_channel=17, num_level=4, soft_weights=True,
                              norm_cfg=dict(type="SyncBN", requires_grad=True)).to(DEV)


@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8, torch.int32])
def test_label_vote_bit_exact(golden, dtype):
    g = golden("losses")
    logits, gt = _c1_inputs()
    tv = CF.downsample_labels(gt.to(DEV).to(dtype), tuple(logits.shape[2:]), 0)
    assert tv.dtype == torch.int32
    assert np.array_equal(tv.cpu().numpy().astype(np.uint8).reshape(g["tv"].shape), g["tv"])


def test_losses_match_reference_fixture_and_oracle(golden):
    g = golden("losses")
    logits, gt = _c1_inputs()
    head = _head()
    x = logits.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    ld = head.loss(output_voxels=[x], target_voxels=gt.to(DEV))
    got = np.array([ld[k].item() for k in NAMES])
    assert np.allclose(got, g["losses"], rtol=1e-5), (got, g["losses"])          # bound 1e-3, observed ~1e-6
    sum(float(c) * ld[k] for c, k in zip(g["coef"], NAMES)).backward()
    assert np.allclose(stats(x.grad), g["dlogits_stats"], rtol=1e-4)
    assert np.allclose(sample(x.grad), g["dlogits_sample"], rtol=1e-3, atol=1e-8)
    # every loss on its own against the oracle's autograd
    for i, k in enumerate(NAMES):
        xo = logits.clone().requires_grad_(True)
        OL.loss_voxel(xo, gt)[k].backward()
        x2 = logits.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        head.loss(output_voxels=[x2], target_voxels=gt.to(DEV))[k].backward()
        assert rel_l2(x2.grad, xo.grad) < 1e-4, (k, rel_l2(x2.grad, xo.grad))
        assert np.allclose(sample(x2.grad), g["d%d_sample" % i], rtol=1e-3, atol=1e-8), k


def test_edge_cases_absent_classes_all_ignored_rows_and_padded_logits():
    head = _head()
    gen = torch.Generator().manual_seed(3)
    V, C = 5000, 17
    logits = torch.randn(V, C, generator=gen) * 3
    labels = torch.randint(0, 4, (V,), generator=gen)            # classes 4..16 absent
    labels[::7] = 255
    labels[:64] = 255                                            # a whole chunk ignored
    lo = logits.clone().requires_grad_(True)
    pred, tv = lo.t().reshape(1, C, V, 1, 1), labels.reshape(1, V, 1, 1)
    ref = {NAMES[0]: OL.ce_ssc_loss(pred, tv, OL.class_weights().float(), 255), NAMES[1]: OL.sem_scal_loss(pred, tv, 255),
           NAMES[2]: OL.geo_scal_loss(pred, tv, 255, 0), NAMES[3]: OL.lovasz_softmax(pred, tv, 255)}
    sum(ref.values()).backward()
    buf = torch.zeros(V, 20, device=DEV)                         # row stride 20 like the conv output
    buf[:, :C] = logits.to(DEV)
    x = buf[:, :C].requires_grad_(True)
    l4 = CF.occ_voxel_losses(x, labels.to(DEV).to(torch.int32), head.class_weights.to(DEV, torch.float32), 255, 0)
    for a, k in zip(l4.tolist(), NAMES):
        assert abs(a - ref[k].item()) <= 1e-5 * abs(ref[k].item()), (k, a, ref[k].item())
    l4.sum().backward()
    assert rel_l2(x.grad, lo.grad) < 1e-4
    assert float(x.grad[labels.to(DEV) == 255].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["r50", "northstar"])
def test_full_size_against_torch_evaluation_on_gpu(name):
    """V = 80 000 / 640 000 voxels x 17 classes: the oracle's torch code evaluated on the GPU (torch.sort,
    cumsum, ...) is an independent computation of the same losses and gradients."""
    grid = S.CONFIGS[name]["grid"]
    head = _head()
    gen = torch.Generator().manual_seed(11)
    logits = torch.randn(1, 17, *grid, generator=gen) * 2.0
    logits[:, 0] += 2.0
    gt = S.make_gt_occ(grid, 2, seed=3).to(DEV)
    tv = CF.downsample_labels(gt, grid, 0)
    tv_o = OL.downsample_labels(gt, grid[0], 0)
    assert torch.equal(tv.long().reshape(tv_o.shape), tv_o)                      # bit-exact at full size
    xo = logits.to(DEV).requires_grad_(True)
    ref = OL.loss_voxel(xo, gt)
    sum(ref.values()).backward()
    x = logits.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    ld = head.loss(output_voxels=[x], target_voxels=gt)
    for k in NAMES:
        assert abs(ld[k].item() - ref[k].item()) <= 2e-5 * abs(ref[k].item()), (k, ld[k].item(), ref[k].item())
    sum(ld.values()).backward()
    assert rel_l2(x.grad, xo.grad) < 1e-3
    # permutation invariance: the losses do not depend on the voxel order
    V = grid[0] * grid[1] * grid[2]
    perm = torch.randperm(V, device=DEV)
    x2d = x.detach().permute(0, 2, 3, 4, 1).reshape(V, 17)
    cw = head.class_weights.to(DEV, torch.float32)
    a = CF.occ_voxel_losses(x2d.contiguous(), tv, cw, 255, 0)
    b = CF.occ_voxel_losses(x2d[perm].contiguous(), tv[perm].contiguous(), cw, 255, 0)
    assert torch.allclose(a, b, rtol=1e-5)
