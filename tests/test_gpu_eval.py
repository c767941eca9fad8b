"""csrc/eval_hist.cu (COOCC_Ray.evaluation_semantic + fast_hist, coocc_ray.py:659-684, 726-730) through the
C ABI: confusion matrices are integer work -> bit-exact against the reference fixture and the oracle."""
import numpy as np
import pytest
import torch

import coocc_b200
from coocc_b200 import functional as CF
from coocc_b200 import synthetic as S
from oracle import evalpath as OE
from test_oracle_eval import inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("dtype", [torch.int64, torch.uint8])
def test_confusion_matches_reference_fixture(golden, dtype):
    g = golden("eval")
    pred, gt, vis = inputs()
    x2d, dims = CF.to_cl2d(pred.to(DEV).contiguous(memory_format=torch.channels_last_3d))
    ssc, ssc_vis, sc = CF.eval_confusion(x2d, dims, gt.to(DEV).to(dtype), vis.to(DEV))
    assert np.array_equal(ssc.cpu().numpy(), g["ssc"])
    assert np.array_equal(ssc_vis.cpu().numpy(), g["ssc_vis"])
    assert np.array_equal(sc.cpu().numpy(), g["sc"])
    gt1 = gt[:, ::2, ::2, ::2].contiguous()
    ssc1, none, _ = CF.eval_confusion(x2d, dims, gt1.to(DEV).to(dtype))
    assert none is None and np.array_equal(ssc1.cpu().numpy(), g["ssc_r1"])


def test_full_size_against_torch_on_gpu_and_padded_rows():
    """r50 grid, labels 200x200x16 (5.1 M voxels at the north-star size are covered by the same code path):
    torch's own interpolate + argmax + bincount on the GPU is the independent evaluation."""
    grid = S.CONFIGS["r50"]["grid"]
    gen = torch.Generator().manual_seed(3)
    V = grid[0] * grid[1] * grid[2]
    buf = torch.zeros(V, 20, device=DEV)                       # row stride 20, like the head's logits
    buf[:, :17] = (torch.randn(V, 17, generator=gen) * 2).to(DEV)
    x2d = buf[:, :17]
    pred5 = CF.to_5d(x2d, grid)
    gt = S.make_gt_occ(grid, 2, seed=4).to(DEV)
    ssc, _, sc = CF.eval_confusion(x2d, grid, gt)
    up = torch.nn.functional.interpolate(pred5.contiguous(), size=list(gt.shape[1:]), mode="trilinear", align_corners=False)
    p = up[0].argmax(0).reshape(-1)
    g = gt.reshape(-1)
    keep = g != 255
    ref = torch.bincount(17 * g[keep] + p[keep], minlength=289).reshape(17, 17)
    # torch's CUDA interpolation may round differently in the last bit: allow a handful of argmax flips
    assert int((ssc - ref).abs().sum()) <= 8, int((ssc - ref).abs().sum())
    assert int(ssc.sum()) == int(keep.sum()) == int(sc.sum())
    assert int(sc[1, :].sum()) == int(((g != 0) & keep).sum())


def test_simple_test_runs_in_eval_mode():
    cfg = S.CONFIGS["c1"]
    coocc_b200.set_precision("fp32")
    try:
        torch.manual_seed(0)
        model = coocc_b200.HotPath(coocc_b200.model_cfg(cfg["C"], cfg["K"]), cfg["C"]).to(DEV).eval()
        inp = S.make_inputs("c1")
        gt = S.make_gt_occ(cfg["grid"], 2, 0).to(DEV)
        out = model.simple_test(inp["img_voxel_feats"].to(DEV), inp["pts_voxel_feats"].to(DEV), gt)
        sc_o, _ = OE.evaluation_semantic(out["pred_c"].float().cpu(), gt.cpu(), "SC")
        ssc_o, _ = OE.evaluation_semantic(out["pred_c"].float().cpu(), gt.cpu(), "SSC")
        assert np.array_equal(out["SC_metric"], sc_o) and np.array_equal(out["SSC_metric"], ssc_o)
        assert out["SSC_occ_metric"] is None and out["SSC_metric"].sum() == int((gt != 255).sum())
    finally:
        coocc_b200.set_precision("tf32")
