"""oracle/finestage.py (OccHead fine / cascade stage + loss_point, occ_head.py:182-237, 295-312;
coordinate_transform.py:3-70) against the unmodified reference head -- fixture from oracle/make_golden.py.
The product does not implement this stage yet; the oracle is pinned ahead of the kernels."""
import numpy as np
import torch

from coocc_b200 import synthetic as S
from helpers import sample, stats
from oracle import finestage as OF
from oracle.make_golden import FINE_GRID, fine_head_params, fine_inputs

NAMES = ["loss_voxel_ce_fine", "loss_voxel_sem_scal_fine", "loss_voxel_geo_scal_fine", "loss_voxel_lovasz_fine"]
PCR = torch.tensor([-10.0, -10.0, -5.0, 10.0, 10.0, 3.0])


def test_fine_stage_matches_reference_fixture(golden):
    g = golden("fine")
    feats, occ, img_feats, transform = fine_inputs()
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fine_head_params().items()}
    f = feats.clone().requires_grad_(True)
    torch.manual_seed(123)                       # the subset of occupied coarse voxels is a torch.randperm draw
    fc, fo = OF.fine_forward(p, f, occ, img_feats, transform, [2 * s for s in FINE_GRID], PCR, 2, 150, 0, True)
    assert np.array_equal(fc.numpy().astype(np.int32), g["fine_coord"])              # coordinates bit-exact
    assert fc.shape[1] == 8 * 150
    assert np.allclose(fo.detach().numpy(), g["fine_output"], rtol=1e-5, atol=1e-6)
    ld = OF.loss_point(fc, fo, S.make_gt_occ(FINE_GRID, 2, 3))
    assert np.allclose(np.array([ld[k].item() for k in NAMES]), g["losses"], rtol=2e-6)
    sum(ld.values()).backward()
    assert np.allclose(stats(f.grad), g["dfeats_stats"], rtol=1e-4)
    assert np.allclose(sample(f.grad), g["dfeats_sample"], rtol=1e-3, atol=1e-8)
    assert np.allclose(p["fine_mlp.3.weight"].grad.numpy(), g["dw_fine3"], rtol=1e-4, atol=1e-7)


def test_coarse_to_fine_without_subset_keeps_every_child():
    c = torch.tensor([[1, 3], [0, 2], [1, 0]])
    out = OF.coarse_to_fine_coordinates(c, 2, topk=100)
    assert out.shape == (3, 16)
    assert out[:, :2].tolist() == [[2, 6], [0, 4], [2, 0]]               # child (0,0,0) of both voxels first
    assert out[:, -2:].tolist() == [[3, 7], [1, 5], [3, 1]]              # child (1,1,1) last
