"""csrc/lss_pool.cu through the C ABI: voxel keys bit-exact against the oracle's index arithmetic, pooled
features / gradients within fp32 summation-order tolerance (the reference's own order is unspecified: its
argsort is unstable), fused lift+splat == pooling of the materialised volume, full-size properties."""
import numpy as np
import pytest
import torch

import coocc_b200
from coocc_b200 import lss as LSS
from helpers import rel_err, rel_l2, sample, stats
from oracle import lss as OLS
from oracle.make_golden import LSS_DATA, LSS_GRID, lss_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dev(i):
    r = {k: v.to(DEV) for k, v in i["rig"].items()}
    return r, {k: i[k].to(DEV) for k in ("frustum", "depth", "feat", "dx", "bx", "nx")}


def test_geometry_matches_oracle_and_fixture(golden):
    g = golden("lss")
    i = lss_inputs()
    r, d = _dev(i)
    geom = LSS.get_geometry(d["frustum"], r["rots"], r["trans"], r["intrins"], r["post_rots"], r["post_trans"], r["bda"])
    ro = i["rig"]
    ref = OLS.get_geometry(i["frustum"], ro["rots"], ro["trans"], ro["intrins"], ro["post_rots"], ro["post_trans"], ro["bda"])
    assert rel_err(geom, ref) < 1e-5                                     # fp32, bound 1e-3
    assert np.allclose(sample(geom), g["geom_sample"], rtol=1e-4, atol=1e-3)


def test_pooling_matches_oracle_fixture_and_fused_path(golden):
    g = golden("lss")
    i = lss_inputs()
    r, d = _dev(i)
    ro = i["rig"]
    geom = OLS.get_geometry(i["frustum"], ro["rots"], ro["trans"], ro["intrins"], ro["post_rots"], ro["post_trans"], ro["bda"])
    # reference-signature path: pooling of the materialised volume
    vol_o = OLS.lift(i["depth"], i["feat"]).clone().requires_grad_(True)
    out_o = OLS.voxel_pooling(geom, vol_o, i["bx"], i["dx"], i["nx"])
    w = torch.linspace(-1, 1, out_o.numel()).reshape(out_o.shape)
    (out_o * w).sum().backward()
    vol = OLS.lift(i["depth"], i["feat"]).to(DEV).requires_grad_(True)
    out = LSS.voxel_pooling(geom.to(DEV), vol, d["bx"], d["dx"], d["nx"])
    assert list(out.shape) == list(g["out_shape"])
    assert torch.equal((out.abs().sum(1) != 0).cpu(), out_o.abs().sum(1) != 0)       # voxel indexing bit-exact
    assert rel_err(out, out_o) < 1e-5
    assert np.allclose(stats(out), g["out_stats"], rtol=2e-4)                          # reference fixture
    (out * w.to(DEV)).sum().backward()
    assert rel_err(vol.grad, vol_o.grad) < 1e-6
    # fused lift + splat (no volume) gives the same result and the gradients of depth / features
    dep_o = i["depth"].clone().requires_grad_(True)
    feat_o = i["feat"].clone().requires_grad_(True)
    (OLS.voxel_pooling(geom, OLS.lift(dep_o, feat_o), i["bx"], i["dx"], i["nx"]) * w).sum().backward()
    dep = d["depth"].clone().requires_grad_(True)
    feat = d["feat"].clone().requires_grad_(True)
    out_f = LSS.lift_splat(geom.to(DEV), dep, feat, d["bx"], d["dx"], d["nx"])
    assert rel_err(out_f, out_o) < 1e-5
    (out_f * w.to(DEV)).sum().backward()
    assert rel_l2(dep.grad, dep_o.grad) < 1e-5 and rel_l2(feat.grad, feat_o.grad) < 1e-5


def test_module_interface_and_edge_cases():
    m = LSS.LSSVoxelPool(LSS_GRID, LSS_DATA, downsample=16).to(DEV)
    assert tuple(m.frustum.shape) == (64, 16, 44, 3) and m.D == 64
    # all points outside the grid -> zeros; points in (-1, 0) cells truncate into cell 0 like `.long()`
    geom = torch.full((1, 1, 2, 1, 2, 3), 1e4, device=DEV)
    x = torch.ones(1, 1, 2, 1, 2, 8, device=DEV)
    assert float(m.voxel_pooling(geom, x).abs().sum()) == 0.0
    lo = (m.bx - m.dx / 2).tolist()
    geom[0, 0, 0, 0, 0] = torch.tensor([lo[0] - 0.5, lo[1] + 0.25, lo[2] - 0.999], device=DEV)   # -> cell (0,0,0)
    geom[0, 0, 1, 0, 1] = torch.tensor([lo[0] - 1.0, lo[1], lo[2]], device=DEV)                   # -> x = -1: dropped
    out = m.voxel_pooling(geom, x)
    assert float(out[0, :, 0, 0, 0].sum()) == 8.0 and float(out.sum()) == 8.0


def test_r101_size_properties():
    """6 x 112 x 56 x 100 = 3.76 M frustum points, C = 128, 100x100x8 grid (the r101 config): total mass is
    conserved, fused == index_add on the GPU, gradients of a linear functional are exact."""
    gen = torch.Generator().manual_seed(0)
    N, D, H, W, C = 6, 112, 56, 100, 128
    grid = dict(xbound=[-50.0, 50.0, 1.0], ybound=[-50.0, 50.0, 1.0], zbound=[-5.0, 3.0, 1.0], dbound=[2.0, 58.0, 0.5])
    m = LSS.LSSVoxelPool(grid, dict(input_size=(896, 1600)), 16).to(DEV)
    rig = {k: v.to(DEV) for k, v in coocc_b200.synthetic.make_camera_rig(N, 1, 896, 1600).items()}
    geom = m.get_geometry(rig["rots"], rig["trans"], rig["intrins"], rig["post_rots"], rig["post_trans"], rig["bda"])
    depth = torch.softmax(torch.randn(N, D, H, W, generator=gen), 1).to(DEV).requires_grad_(True)
    feat = torch.randn(N, C, H, W, generator=gen).to(DEV).requires_grad_(True)
    out = m.lift_splat(geom, depth, feat)
    idx, kept = OLS.voxel_indices(geom, m.bx, m.dx, m.nx)
    vox = ((idx[:, 0] * 100 + idx[:, 1]) * 8 + idx[:, 2])[kept]
    pid = torch.nonzero(kept).flatten()
    n_i, hw_i = pid // (D * H * W), pid % (H * W)
    rows = feat.detach().permute(0, 2, 3, 1).reshape(N * H * W, C)[n_i * H * W + hw_i] * depth.detach().reshape(-1)[pid, None]
    ref = torch.zeros(80000, C, device=DEV).index_add_(0, vox, rows)
    got = out.permute(0, 2, 3, 4, 1).reshape(80000, C)
    assert rel_err(got, ref) < 1e-5
    assert torch.equal(got.abs().sum(1) != 0, ref.abs().sum(1) != 0)
    out.sum().backward()
    # d(sum out)/d depth[i] = sum_c feat[row_i, c] for kept points, 0 otherwise
    exp = torch.zeros(N * D * H * W, device=DEV)
    exp[pid] = feat.detach().permute(0, 2, 3, 1).reshape(N * H * W, C)[n_i * H * W + hw_i].sum(1)
    assert rel_err(depth.grad.reshape(-1), exp) < 1e-5
