"""oracle/evalpath.py against the reference's evaluation_semantic / fast_hist (coocc_ray.py:659-684,
726-730): fixtures from oracle/make_golden.py, live comparison when /root/reference exists."""
import numpy as np
import pytest
import torch

from coocc_b200 import synthetic as S
from oracle import evalpath as OE
from oracle import refshim


def inputs():
    grid = S.CONFIGS["c1"]["grid"]
    gen = torch.Generator().manual_seed(77)
    pred = torch.randn(1, 17, *grid, generator=gen) * 2.0
    pred[:, 0] += 1.0
    gt = S.make_gt_occ(grid, 2, seed=1)
    vis = (torch.rand(gt.shape, generator=gen) < 0.6).to(torch.uint8)
    return pred, gt, vis


def test_eval_matches_reference_fixture(golden):
    g = golden("eval")
    pred, gt, vis = inputs()
    sc, _ = OE.evaluation_semantic(pred, gt, "SC")
    ssc, ssc_vis = OE.evaluation_semantic(pred, gt, "SSC", vis)
    assert np.array_equal(sc, g["sc"]) and np.array_equal(ssc, g["ssc"]) and np.array_equal(ssc_vis, g["ssc_vis"])
    ssc1, none = OE.evaluation_semantic(pred, gt[:, ::2, ::2, ::2].contiguous(), "SSC")
    assert none is None and np.array_equal(ssc1, g["ssc_r1"])
    assert int(ssc.sum()) == int((gt != 255).sum()) and int(sc.sum()) == int(ssc.sum())


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_eval_live_other_seed():
    gen = torch.Generator().manual_seed(5)
    pred = torch.randn(1, 17, 12, 12, 4, generator=gen)
    gt = S.make_gt_occ((12, 12, 4), 2, seed=9)
    for t in ("SC", "SSC"):
        a, _ = refshim.reference_evaluation_semantic(pred, gt, t)
        b, _ = OE.evaluation_semantic(pred, gt, t)
        assert np.array_equal(np.asarray(a), b)
