"""oracle/lss.py (Lift-Splat voxel pooling, get_geometry) against the reference's own lines
(ViewTransformerLSSVoxel.py:100-123, ViewTransformerLSSBEVDepth.py:117-150) -- fixtures from
oracle/make_golden.py, live when /root/reference exists."""
import numpy as np
import pytest
import torch

from helpers import sample, stats
from oracle import lss as OLS
from oracle import refshim
from oracle.make_golden import lss_inputs


def test_geometry_and_pooling_match_reference_fixture(golden):
    g = golden("lss")
    i = lss_inputs()
    r = i["rig"]
    geom = OLS.get_geometry(i["frustum"], r["rots"], r["trans"], r["intrins"], r["post_rots"], r["post_trans"], r["bda"])
    assert np.allclose(sample(geom), g["geom_sample"], rtol=1e-5, atol=1e-4)
    assert np.allclose(stats(geom), g["geom_stats"], rtol=1e-6)
    idx, kept = OLS.voxel_indices(geom, i["bx"], i["dx"], i["nx"])
    same = idx[::101].numpy().astype(np.int32) == g["idx_sample"]
    assert same.mean() > 0.9995                        # geometry differs in the last bit -> a boundary point may move
    vol = OLS.lift(i["depth"], i["feat"]).clone().requires_grad_(True)
    out = OLS.voxel_pooling(geom, vol, i["bx"], i["dx"], i["nx"])
    assert list(out.shape) == list(g["out_shape"])
    assert np.allclose(stats(out), g["out_stats"], rtol=2e-4)
    assert abs(int((out.abs().sum(1) != 0).sum()) - int(g["occupied"])) <= 2
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    (out * w).sum().backward()
    assert np.allclose(stats(vol.grad), g["dvol_stats"], rtol=2e-4)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_pooling_live_same_geometry_is_exact_up_to_summation_order():
    i = lss_inputs(seed=3, C=8)
    r = i["rig"]
    geom = refshim.reference_get_geometry(i["frustum"], r["rots"], r["trans"], r["intrins"], r["post_rots"],
                                          r["post_trans"], r["bda"])
    vol = OLS.lift(i["depth"], i["feat"])
    a = refshim.reference_voxel_pooling(geom, vol, i["bx"], i["dx"], i["nx"])
    b = OLS.voxel_pooling(geom, vol, i["bx"], i["dx"], i["nx"])
    assert a.shape == b.shape
    assert torch.equal(a.abs().sum(1) != 0, b.abs().sum(1) != 0)             # same occupancy: indices bit-exact
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
