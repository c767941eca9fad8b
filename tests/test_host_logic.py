"""Host-side logic that needs no GPU: Lift-Splat grid constants, the bf16 weight-shadow validity rule, GraphedStep's
bucket keys."""
import torch

import coocc_b200
from coocc_b200 import functional as CF
from coocc_b200 import lss as LSS
from coocc_b200.graph import GraphedStep


def test_gridspec_matches_reference_fp32_arithmetic():
    dx, bx, nx = LSS.gen_dx_bx([-50.0, 50.0, 1.0], [-50.0, 50.0, 1.0], [-5.0, 3.0, 1.0])
    assert nx.tolist() == [100.0, 100.0, 8.0] and bx.tolist() == [-49.5, -49.5, -4.5]
    spec = LSS.GridSpec(bx, dx, nx)
    assert spec.dims == (100, 100, 8)
    lo = (bx - dx / 2.0).tolist()                       # (self.bx - self.dx / 2.) of ViewTransformerLSSVoxel.py:107
    assert [float(v) for v in spec.lo] == lo == [-50.0, -50.0, -5.0]
    dx2, bx2, nx2 = LSS.gen_dx_bx([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5.0, 3.0, 0.8])
    s2 = LSS.GridSpec(bx2, dx2, nx2)
    assert [float(v) for v in s2.lo] == (bx2 - dx2 / 2.0).tolist()      # fp32 rounding kept, not recomputed in double


def test_weight_shadow_is_used_only_while_valid():
    w = torch.nn.Parameter(torch.randn(16, 8, 3, 3, 3).contiguous(memory_format=torch.channels_last_3d))
    rows = CF.weight_rows(w)
    assert rows.shape == (16, 27 * 8) and rows.data_ptr() == w.data_ptr()          # zero-copy view
    a = CF.weight_operand(w, CF.DT_BF16)
    assert a.dtype == torch.bfloat16 and torch.equal(a, rows.detach().to(torch.bfloat16))
    w._coocc_bf16 = rows.detach().reshape(-1).to(torch.bfloat16) + 1                # a recognisable shadow
    w._coocc_bf16_version = w._version
    assert CF.weight_operand(w, CF.DT_BF16).data_ptr() != w._coocc_bf16.data_ptr()  # no layout key: not trusted
    w._coocc_bf16_layout = (w.data_ptr(), tuple(w.stride()))
    b = CF.weight_operand(w, CF.DT_BF16)
    assert b.data_ptr() == w._coocc_bf16.data_ptr()                                 # used as is
    assert CF.weight_operand(w, CF.DT_TF32).dtype == torch.float32                  # fp32 modes never use it
    with torch.no_grad():
        w.mul_(2.0)                                                                 # any in-place change invalidates it
    c = CF.weight_operand(w, CF.DT_BF16)
    assert c.data_ptr() != w._coocc_bf16.data_ptr()
    assert torch.equal(c, CF.weight_rows(w).detach().to(torch.bfloat16))
    # a relayout through `.data =` keeps the version counter: the storage address / strides invalidate it
    w._coocc_bf16_version = w._version
    assert CF.weight_operand(w, CF.DT_BF16).data_ptr() == w._coocc_bf16.data_ptr()
    w.data = w.data.clone()
    assert w._version == w._coocc_bf16_version
    assert CF.weight_operand(w, CF.DT_BF16).data_ptr() != w._coocc_bf16.data_ptr()


def test_graph_bucket_keys():
    g = GraphedStep(model=None, bucket=8192)
    assert g._round(1, 640000) == 8192 and g._round(8192, 640000) == 8192 and g._round(8193, 640000) == 16384
    assert g._round(639999, 640000) == 640000                                       # capped at the grid size
    assert coocc_b200.GraphedStep is GraphedStep
