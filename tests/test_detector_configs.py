"""The drop-in boundary at the registry level, CPU only: every `model = dict(...)` of the reference's
projects/configs/coocc_nusc/*.py -- stored verbatim as JSON by oracle/make_config_fixture.py -- builds through the
registry into this package's detectors and modules, with the reference's state_dict names for the hot-path modules."""
import glob
import json
import os

import pytest
import torch

import coocc_b200
from coocc_b200 import detector, registry

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = sorted(glob.glob(os.path.join(HERE, "golden", "configs", "*.model.json")))


def _load(path):
    with open(path) as f:
        return json.load(f)


def test_fixtures_are_the_literal_reference_configs():
    assert len(CONFIGS) == 5
    from oracle import make_config_fixture as M
    if not os.path.isdir(M.CFG_DIR):
        pytest.skip("reference tree not present")
    for path in CONFIGS:
        d = _load(path)
        live = M.load_model_dict(os.path.join(M.REF, d["source"]))
        assert json.loads(json.dumps(live)) == d["model"], d["source"]


@pytest.mark.parametrize("path", CONFIGS, ids=[os.path.basename(p)[:-11] for p in CONFIGS])
def test_literal_config_builds_through_the_registry(path):
    model = _load(path)["model"]
    det = detector.build_detector(model)
    assert type(det).__name__ == model["type"] and isinstance(det, detector.CooccRayHotPath)
    assert det.lidar_only == (model["type"] == "COOCC_Ray_L")
    assert isinstance(det.semantic_encoder, coocc_b200.CustomResNet3D) and isinstance(det.semantic_neck, coocc_b200.FPN3D)
    assert isinstance(det.pts_bbox_head, coocc_b200.OccHead)
    if "occ_fuser" in model:
        assert isinstance(det.occ_fuser, coocc_b200.BiFuser_N) and det.occ_fuser.knum == model["occ_fuser"]["knum"]
    else:
        assert det.occ_fuser is None
    head, hc = det.pts_bbox_head, model["pts_bbox_head"]
    assert head.cascade_ratio == hc["cascade_ratio"] and head.fine_topk == hc["fine_topk"]
    assert head.fine_stage == (hc["cascade_ratio"] != 1 and (hc["sample_from_voxel"] or hc["sample_from_img"]))
    keys = set(det.state_dict().keys())
    # names a reference checkpoint carries for these modules
    want = {"semantic_encoder.input_proj.0.weight", "semantic_encoder.layers.3.1.bn2.running_var",
            "semantic_neck.lateral_convs.0.0.conv.weight", "semantic_neck.fpn_convs.3.0.bn.weight",
            "pts_bbox_head.occ_convs.0.0.weight", "pts_bbox_head.occ_pred_conv.3.weight",
            "pts_bbox_head.voxel_soft_weights.3.weight",
            "sigma_head.hidden_layers.0.weight", "sigma_head.output_layer.bias", "sigma_head.posi_encoder.scales"}
    if head.fine_stage:
        want |= {"pts_bbox_head.fine_mlp.0.weight", "pts_bbox_head.fine_mlp.3.bias"}
        assert head.fine_mlp[0].in_features == (128 if hc["sample_from_voxel"] else 0) + (64 if hc["sample_from_img"] else 0)
    else:                                   # coocc_lidar.py: cascade_ratio=2 but nothing to sample from
        assert not any("fine_mlp" in k for k in keys)
    if hc["sample_from_img"]:
        want |= {"pts_bbox_head.img_mlp_0.0.weight", "pts_bbox_head.img_mlp_0.1.bias", "pts_bbox_head.img_mlp.0.weight"}
    if model["type"] == "COOCC_Ray":
        want |= {"rgb_head.hidden_layers.2.weight", "rgb_head.output_layer.weight"}
    else:
        assert not any(k.startswith("rgb_head") for k in keys)
    if "occ_fuser" in model:
        want |= {"occ_fuser.con_enc.0.weight", "occ_fuser.con_enc.4.running_mean", "occ_fuser.knn_enc.0.bias"}
    assert want <= keys, sorted(want - keys)
    # conv weights are stored [Cout,kx,ky,kz,Cin] from construction on (shape unchanged: checkpoints load in place)
    w = det.semantic_encoder.layers[0][0].conv1.weight
    assert tuple(w.shape) == (128, 128, 3, 3, 3) and w.permute(0, 2, 3, 4, 1).is_contiguous()


def test_registry_refuses_duplicates_without_force_like_mmcv():
    """mmcv's Registry raises KeyError when a name is registered twice unless force=True; the plugin classes are
    registered first (plugin=True), this package afterwards with force=True."""
    reg = registry._LocalRegistry("models")

    class A:
        pass

    class B:
        pass

    reg.register_module(name="BiFuser_N")(A)
    with pytest.raises(KeyError):
        reg.register_module(name="BiFuser_N")(B)
    reg.register_module(name="BiFuser_N", force=True)(B)
    assert reg.get("BiFuser_N") is B
    # the package's own classes went in with force=True and are what the config names resolve to
    for name in ("BiFuser_N", "CustomResNet3D", "FPN3D", "OccHead", "COOCC_Ray", "COOCC_Ray_L"):
        cls = registry.DETECTORS.get(name) or registry.HEADS.get(name)
        assert cls is not None and cls.__module__.startswith("coocc_b200"), name


def test_upstream_tensors_are_required():
    det = detector.build_detector(_load(CONFIGS[-1])["model"])
    with pytest.raises(RuntimeError, match="upstream"):
        det.extract_feat(None, None, None)
