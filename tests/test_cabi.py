"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/coocc_b200.h declares (no compute calls), and the host modules keep the reference's
state_dict keys."""
import os
import re

import pytest
import torch

import coocc_b200
from coocc_b200 import _lib
from coocc_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "coocc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(coocc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.SO_PATH):
        _lib.build()
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "symbol %s declared in the header is not exported" % n
    assert sorted(_lib.exported_symbols()) == names, "ctypes signature table out of sync with the header"
    assert L.coocc_version() >= 100


def test_no_cpu_fallback():
    m = coocc_b200.BiFuser_N(8, 8, knum=1)
    x = torch.zeros(1, 8, 4, 4, 2)
    with pytest.raises(RuntimeError):
        m(x, x)


def test_state_dict_keys_match_reference_layout():
    P = S.make_params("c1")
    C = 32
    nc = dict(type="SyncBN", requires_grad=True)
    mods = {
        "occ_fuser": coocc_b200.BiFuser_N(C, C, knum=4),
        "semantic_encoder": coocc_b200.CustomResNet3D(depth=18, n_input_channels=C, block_inplanes=[C, 2 * C, 4 * C, 8 * C],
                                                      out_indices=(0, 1, 2, 3), norm_cfg=nc),
        "semantic_neck": coocc_b200.FPN3D(with_cp=True, in_channels=[C, 2 * C, 4 * C, 8 * C], out_channels=2 * C, norm_cfg=nc),
        "pts_bbox_head": coocc_b200.OccHead(norm_cfg=nc, soft_weights=True, num_level=4, in_channels=[2 * C] * 4, out_channel=17),
    }
    for k, m in mods.items():
        assert sorted(m.state_dict().keys()) == sorted(P[k].keys()), k
        m.load_state_dict(P[k], strict=True)
    for h, depth, od in (("sigma_head", 1, 1), ("rgb_head", 3, 3)):
        m = coocc_b200.MLP(input_dim=C, output_dim=od, net_depth=depth, skip_layer=None)
        sd = {k[len(h) + 1:]: v for k, v in P["render"].items() if k.startswith(h + ".")}
        assert sorted(m.state_dict().keys()) == sorted(sd.keys())


def test_registry_names():
    from coocc_b200 import registry
    for name in ("BiFuser_N", "CustomResNet3D", "FPN3D", "OccHead"):
        assert registry.FUSION_LAYERS.get(name) is not None or registry.BACKBONES.get(name) is not None
    f = registry.build_fusion_layer(dict(type="BiFuser_N", knum=2, in_channels=16, out_channels=16))
    assert isinstance(f, coocc_b200.BiFuser_N)
