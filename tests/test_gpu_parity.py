"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical seeded
inputs and against the reference-generated fixtures.

Tolerances.  Integer / index work: bit-exact.  Floating point: the north-star bound is 1e-3
relative for fp32 logits / renders.  The convolutions and linears run as TF32 tensor-core math
with fp32 accumulation (COOCC_DTYPE_TF32; the reference's own GPU default, torch 1.10
`allow_tf32=True`); measured per-layer error is ~3e-4 of the output range, which compounds
through the 30-layer stack -- the stack-level bounds below state what is asserted.
"""
import numpy as np
import pytest
import torch

import coocc_b200
from coocc_b200 import functional as CF
from coocc_b200 import synthetic as S
from oracle import oracle as O
from helpers import rel_err, rel_l2, sample, stats

pytestmark = pytest.mark.gpu
DEV = "cuda"
NC = dict(type="SyncBN", requires_grad=True)

# asserted bounds (max|a-b| / max|b| unless noted) per arithmetic mode of the tensor-core convs:
#   fp32 = COOCC_DTYPE_TF32X3 (3-pass split, fp32-accurate)  -> the north-star 1e-3 bound
#   tf32 = single pass                                        -> per-layer ~3e-4, stack ~1e-2
# Gradients of a ReLU network are discontinuous in the forward values: a forward error e flips
# ~e of the ReLU masks and perturbs gradients by ~sqrt(e); gradient bounds are therefore wider.
TOL = {
    "fp32": dict(fused=1e-3, fused_grad=1e-2, occ=1e-3, stack_grad=2e-2, render=1e-3, render_grad=1e-2),
    "tf32": dict(fused=3e-3, fused_grad=3e-1, occ=2e-2, stack_grad=5e-1, render=2e-3, render_grad=5e-1),
    # bf16 operands (8-bit mantissa), fp32 accumulation: per-layer ~3e-3, reported separately
    # (activations are also *stored* in bf16 between layers: one more rounding per layer)
    "bf16": dict(fused=2e-2, fused_grad=8e-1, occ=1.5e-1, stack_grad=1.0, render=2e-2, render_grad=1.0),
}


@pytest.fixture(params=["fp32", "tf32", "bf16"])
def precision(request):
    old = coocc_b200.get_precision()
    coocc_b200.set_precision(request.param)
    yield request.param
    coocc_b200.set_precision(old)


def _cuda_params(p):
    return {k: v.to(DEV) for k, v in p.items()}


def _report(name, **kw):
    print("[parity] %-28s " % name + "  ".join("%s=%.3e" % (k, v) for k, v in kw.items()), flush=True)


# ------------------------------------------------------------------------------- indices
@pytest.mark.parametrize("name", ["c1", "c1k1"])
def test_gsf_indices_bit_exact(name, golden):
    cfg = S.CONFIGS[name]
    K = cfg["K"]
    inp = S.make_inputs(name, with_render=False)
    img, pts = inp["img_voxel_feats"], inp["pts_voxel_feats"]
    cat, st = CF.gsf_index(img.to(DEV), pts.to(DEV), K, want_parts=True)
    torch.cuda.synchronize()
    ii, ip = O.occupied_indices(img), O.occupied_indices(pts)
    X, Y, Z = cfg["grid"]
    lin = lambda t: ((t[:, 1] * Y + t[:, 2]) * Z + t[:, 3]).int()
    assert st["n_img"] == len(ii) and st["n_pts"] == len(ip)
    assert torch.equal(st["lists"][0, :len(ii)].cpu(), lin(ii))
    assert torch.equal(st["lists"][1, :len(ip)].cpu(), lin(ip))
    # packed channels-last slices == the reference's permuted inputs, bit for bit
    C = cfg["C"]
    assert torch.equal(cat[:, :C].cpu(), img.permute(0, 2, 3, 4, 1).reshape(-1, C))
    assert torch.equal(cat[:, C:2 * C].cpu(), pts.permute(0, 2, 3, 4, 1).reshape(-1, C))
    for dname, q, k in (("A", ip, ii), ("B", ii, ip)):
        ref, parts = O.fps_nn_fast(q, k, K, tie="canonical", return_parts=True)
        ref = ref.reshape(K, -1)
        got = CF.gsf_nn_indices(st, dname).cpu()
        if parts:
            d = st[dname]
            assert np.array_equal(d["rep_idx"].cpu().numpy(), parts["rep_idx"]), "FPS order differs"
            assert np.array_equal(d["topk_idx"].cpu().numpy(), parts["topk_idx"]), "top-K differs"
            assert np.array_equal(d["group"].cpu().numpy(), parts["group"]), "ball query differs"
        assert torch.equal(got, ref), "assignment differs (%s)" % dname
    if K == 1:   # K=1 canonical == reference (fixture made by the unmodified reference)
        g = golden(name)
        assert np.array_equal(CF.gsf_nn_indices(st, "A")[0].cpu().numpy().astype(np.int32), g["nn_img"])
        assert np.array_equal(CF.gsf_nn_indices(st, "B")[0].cpu().numpy().astype(np.int32), g["nn_pts"])


def test_q1_raises_like_reference():
    inp = S.make_inputs("c1k1", with_render=False)
    with pytest.raises(IndexError):
        CF.gsf_index(inp["img_voxel_feats"].to(DEV), inp["pts_voxel_feats"].to(DEV), 2)


# ------------------------------------------------------------------------------- fuser
@pytest.mark.parametrize("name", ["c1", "c1k1"])
def test_fuser_forward_backward(name, golden, precision):
    tol = TOL[precision]
    cfg = S.CONFIGS[name]
    C, K = cfg["C"], cfg["K"]
    inp, P = S.make_inputs(name, with_render=False), S.make_params(name)
    # oracle (canonical tie rule), CPU fp32
    po = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P["occ_fuser"].items()}
    img_o = inp["img_voxel_feats"].clone().requires_grad_(True)
    pts_o = inp["pts_voxel_feats"].clone().requires_grad_(True)
    out_o, parts = O.bifuser_forward(po, img_o, pts_o, K, tie="canonical", return_parts=True)
    w = torch.linspace(-1, 1, out_o.numel()).reshape(out_o.shape)
    (out_o * w).sum().backward()
    # CUDA path
    m = coocc_b200.BiFuser_N(C, C, knum=K).to(DEV)
    m.load_state_dict(P["occ_fuser"])
    m.train()
    img = inp["img_voxel_feats"].to(DEV).requires_grad_(True)
    pts = inp["pts_voxel_feats"].to(DEV).requires_grad_(True)
    cat = CF.gsfusion_concat(img, pts, m.knn_enc[0].weight, m.knn_enc[0].bias, K)
    e_cat = rel_err(cat, parts["all_feats"].reshape(-1, 4 * C))
    out = m(img, pts)
    (out * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    # forward: max-norm; gradients: relative L2 (a single flipped ReLU mask moves individual
    # gradient entries at full scale, the L2 norm measures how many did)
    e_out, e_img, e_pts = rel_err(out, out_o), rel_l2(img.grad, img_o.grad), rel_l2(pts.grad, pts_o.grad)
    e_w = rel_l2(m.knn_enc[0].weight.grad, po["knn_enc.0.weight"].grad)
    e_b = rel_l2(m.knn_enc[0].bias.grad, po["knn_enc.0.bias"].grad)
    e_cw = rel_l2(m.con_enc[0].weight.grad, po["con_enc.0.weight"].grad)
    _report("fuser[%s,%s]" % (name, precision), cat=e_cat, out=e_out, dimg=e_img, dpts=e_pts, dknn_w=e_w, dknn_b=e_b, dconv_w=e_cw)
    assert e_cat < 1e-5          # gather + fp32 linear + modulate + scatter: fp32-exact
    assert e_out < tol["fused"]
    assert max(e_img, e_pts, e_w, e_b, e_cw) < tol["fused_grad"]
    if K == 1:                   # reference fixture (K=1: canonical == reference)
        g = golden(name)
        np.testing.assert_allclose(sample(out), g["fused_sample"], atol=tol["fused"] * np.abs(g["fused_sample"]).max())


# ------------------------------------------------------------------------------- conv stack
@pytest.mark.parametrize("name", ["c1", "c1k1"])
def test_conv_stack_forward_backward(name, golden, precision):
    tol = TOL[precision]
    g = golden(name)
    cfg = S.CONFIGS[name]
    C = cfg["C"]
    P = S.make_params(name)
    planes = [C, 2 * C, 4 * C, 8 * C]
    x0 = torch.randn(1, C, *cfg["grid"], generator=torch.Generator().manual_seed(1234)) * 0.5
    # oracle
    xo = x0.clone().requires_grad_(True)
    pe = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P["semantic_encoder"].items()}
    ph = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P["pts_bbox_head"].items()}
    mid_o = O.resnet3d_forward(pe, xo)
    nk_o = O.fpn3d_forward(P["semantic_neck"], mid_o)
    feats_o, occ_o = O.occhead_coarse_forward(ph, nk_o)
    w = torch.linspace(-1, 1, occ_o.numel()).reshape(occ_o.shape)
    (occ_o * w).sum().backward()
    # CUDA
    enc = coocc_b200.CustomResNet3D(depth=18, n_input_channels=C, block_inplanes=planes, out_indices=(0, 1, 2, 3), norm_cfg=NC).to(DEV)
    neck = coocc_b200.FPN3D(with_cp=True, in_channels=planes, out_channels=2 * C, norm_cfg=NC).to(DEV)
    head = coocc_b200.OccHead(norm_cfg=NC, soft_weights=True, num_level=4, in_channels=[2 * C] * 4, out_channel=17).to(DEV)
    enc.load_state_dict(P["semantic_encoder"]); neck.load_state_dict(P["semantic_neck"]); head.load_state_dict(P["pts_bbox_head"])
    for m in (enc, neck, head):
        m.train()
    x = x0.to(DEV).requires_grad_(True)
    mid = enc(x)
    nk = neck(mid)
    o = head.forward_coarse_voxel(nk)
    occ = o["occ"][0]
    (occ * w.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    errs = dict(mid0=rel_err(mid[0], mid_o[0]), mid3=rel_err(mid[3], mid_o[3]), neck0=rel_err(nk[0], nk_o[0]),
                feats=rel_err(o["out_voxel_feats"][0], feats_o), occ=rel_err(occ, occ_o), occ_l2=rel_l2(occ, occ_o),
                dx=rel_err(x.grad, xo.grad), dx_l2=rel_l2(x.grad, xo.grad),
                dw_proj=rel_err(enc.input_proj[0].weight.grad, pe["input_proj.0.weight"].grad),
                dw_pred=rel_err(head.occ_pred_conv[3].weight.grad, ph["occ_pred_conv.3.weight"].grad))
    _report("stack[%s,%s]" % (name, precision), **errs)
    assert errs["occ"] < tol["occ"] and errs["occ_l2"] < tol["occ"]
    assert errs["dx_l2"] < tol["stack_grad"] and errs["dw_pred"] < tol["stack_grad"] and errs["dw_proj"] < tol["stack_grad"]
    np.testing.assert_allclose(sample(occ), g["occ_sample"], atol=tol["occ"] * np.abs(g["occ_sample"]).max())
    # running statistics are updated like nn.BatchNorm3d
    assert int(enc.input_proj[1].num_batches_tracked) == 1


# ------------------------------------------------------------------------------- render
@pytest.mark.parametrize("name", ["c1", "c1k1"])
def test_render_forward_backward(name, golden, precision):
    tol = TOL[precision]
    g = golden(name)
    cfg = S.CONFIGS[name]
    C = cfg["C"]
    inp, P = S.make_inputs(name), S.make_params(name)
    vf0 = torch.randn(1, C, *cfg["grid"], generator=torch.Generator().manual_seed(4321)) * 0.5
    vo = vf0.clone().requires_grad_(True)
    pr = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P["render"].items()}
    ro = O.render_forward(pr, vo, inp["geom"], inp["gt_depth"], inp["gt_img"])
    (ro["loss_depth_render"] + ro["loss_rgb"]).backward()
    sig = coocc_b200.MLP(input_dim=C, output_dim=1, net_depth=1, skip_layer=None).to(DEV)
    rgb = coocc_b200.MLP(input_dim=C, output_dim=3, net_depth=3, skip_layer=None).to(DEV)
    sig.load_state_dict({k[11:]: v for k, v in P["render"].items() if k.startswith("sigma_head.")})
    rgb.load_state_dict({k[9:]: v for k, v in P["render"].items() if k.startswith("rgb_head.")})
    vf = vf0.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    rgbs, depths, losses = coocc_b200.render_fn(vf, inp["geom"].to(DEV), sig, rgb, inp["gt_depth"].to(DEV), inp["gt_img"].to(DEV))
    (losses["loss_depth_render"] + losses["loss_rgb"]).backward()
    torch.cuda.synchronize()
    errs = dict(rgbs=rel_err(rgbs, ro["rgbs"]), depths=rel_err(depths, ro["depths"]),
                l_depth=abs(losses["loss_depth_render"].item() - ro["loss_depth_render"].item()) / abs(ro["loss_depth_render"].item()),
                l_rgb=abs(losses["loss_rgb"].item() - ro["loss_rgb"].item()) / abs(ro["loss_rgb"].item()),
                dvf=rel_err(vf.grad, vo.grad), dvf_l2=rel_l2(vf.grad, vo.grad),
                dw_sig=rel_err(sig.output_layer.weight.grad, pr["sigma_head.output_layer.weight"].grad),
                dw_rgb=rel_err(rgb.hidden_layers[0].weight.grad, pr["rgb_head.hidden_layers.0.weight"].grad))
    _report("render[%s,%s]" % (name, precision), **errs)
    assert errs["rgbs"] < tol["render"] and errs["depths"] < tol["render"]
    assert errs["l_depth"] < tol["render"] and errs["l_rgb"] < tol["render"]
    assert errs["dvf_l2"] < tol["render_grad"] and errs["dw_sig"] < tol["render_grad"] and errs["dw_rgb"] < tol["render_grad"]
    np.testing.assert_allclose(sample(rgbs), g["render_rgbs_sample"], atol=tol["render"])
    assert abs(losses["loss_rgb"].item() - float(g["loss_rgb"])) < tol["render"] * float(g["loss_rgb"])


# ------------------------------------------------------------------------------- single convs
CONV_CASES = [
    # X, Y, Z, Cin, Cout, k, stride
    (12, 10, 4, 32, 32, 1, 1), (12, 10, 4, 64, 40, 3, 1), (13, 11, 4, 32, 64, 3, 2), (13, 13, 1, 64, 128, 3, 2),
    (12, 10, 4, 64, 96, 1, 2), (10, 10, 8, 128, 17, 1, 1), (9, 9, 2, 256, 3, 1, 1), (25, 25, 2, 96, 32, 3, 1),
    (64, 60, 10, 32, 64, 3, 1),      # > 148 double tiles: exercises the 256-row CTA tile (MT = 2) and its ragged tail
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3d_op_forward_backward(case, precision):
    """One convolution through the autograd wrapper (incl. the strided-dgrad glue) against
    torch's CPU fp32 conv3d on identical inputs."""
    import torch.nn.functional as F
    X, Y, Z, Cin, Cout, k, s = case
    gen = torch.Generator().manual_seed(sum(case))
    x0 = torch.randn(1, Cin, X, Y, Z, generator=gen)
    w0 = torch.randn(Cout, Cin, k, k, k, generator=gen) / (Cin * k ** 3) ** 0.5
    xr, wr = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, None, s, k // 2)
    gy = torch.randn(yr.shape, generator=gen)
    yr.backward(gy)
    x = x0.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    w = w0.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    x2d, dims = CF.to_cl2d(x)
    y2d = CF.conv3d(x2d, w, dims, k, s)
    odims = tuple(CF.out_dim(n, k, s) for n in dims)
    y = CF.to_5d(y2d, odims)
    y.backward(gy.to(DEV))
    torch.cuda.synchronize()
    e = dict(y=rel_err(y, yr), dx=rel_err(x.grad, xr.grad), dw=rel_err(w.grad, wr.grad))
    _report("conv%s[%s]" % (str(case), precision), **e)
    bound = {"fp32": 2e-5, "tf32": 2e-3, "bf16": 2e-2}[precision]
    assert max(e.values()) < bound


# ------------------------------------------------------------------------------- fused BN
@pytest.mark.parametrize("res", [False, True])
def test_conv_bn_relu_fused(res, precision):
    """conv -> BatchNorm3d(train) -> (+residual) -> ReLU through the fused kernels
    (conv-epilogue statistics + bn_act fwd/bwd) against torch CPU modules."""
    import torch.nn as nn
    from coocc_b200 import modules as M
    gen = torch.Generator().manual_seed(7)
    Cin, Cout, dims = 32, 64, (11, 9, 4)
    conv = nn.Conv3d(Cin, Cout, 3, 1, 1, bias=False)
    bn = nn.BatchNorm3d(Cout)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * 0.05)
        bn.weight.copy_(torch.rand(Cout, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(Cout, generator=gen) * 0.1)
    x0 = torch.randn(1, Cin, *dims, generator=gen)
    r0 = torch.randn(1, Cout, *dims, generator=gen)
    import copy
    conv_g, bn_g = copy.deepcopy(conv).to(DEV), copy.deepcopy(bn).to(DEV)
    # CPU reference
    xr, rr = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
    y = bn(conv(xr))
    if res:
        y = y + rr
    y = torch.relu(y)
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    # fused CUDA path
    x = x0.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    r = r0.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    x2d, d = CF.to_cl2d(x)
    r2d = CF.to_cl2d(r)[0] if res else None
    o2d, _ = M.conv_bn_act(x2d, d, conv_g, bn_g, relu=True, residual=r2d)
    o = CF.to_5d(o2d, d)
    o.backward(gy.to(DEV))
    torch.cuda.synchronize()
    e = dict(y=rel_err(o, y), dx=rel_l2(x.grad, xr.grad), dw=rel_l2(conv_g.weight.grad, conv.weight.grad),
             dgamma=rel_l2(bn_g.weight.grad, bn.weight.grad), dbeta=rel_l2(bn_g.bias.grad, bn.bias.grad),
             rmean=rel_err(bn_g.running_mean, bn.running_mean), rvar=rel_err(bn_g.running_var, bn.running_var))
    if res:
        e["dres"] = rel_l2(r.grad, rr.grad)
    _report("conv_bn_relu[res=%s,%s]" % (res, precision), **e)
    fwd_b, bwd_b = {"fp32": (2e-5, 2e-3), "tf32": (3e-3, 5e-2), "bf16": (3e-2, 3e-1)}[precision]
    assert e["y"] < fwd_b and e["rmean"] < fwd_b and e["rvar"] < fwd_b
    assert max(v for k, v in e.items() if k.startswith("d")) < bwd_b
    assert int(bn_g.num_batches_tracked) == 1


def test_eval_mode_matches_oracle_modules():
    """Inference (model.eval(), no_grad): BatchNorm uses the running statistics -- compared with the
    same torch modules run on the CPU."""
    import copy
    import torch.nn as nn
    from coocc_b200 import modules as M
    coocc_b200.set_precision("fp32")
    try:
        gen = torch.Generator().manual_seed(11)
        conv = nn.Conv3d(32, 64, 3, 1, 1, bias=False)
        bn = nn.BatchNorm3d(64)
        with torch.no_grad():
            bn.running_mean.copy_(torch.randn(64, generator=gen) * 0.1)
            bn.running_var.copy_(torch.rand(64, generator=gen) + 0.5)
            bn.weight.copy_(torch.rand(64, generator=gen) + 0.5)
            bn.bias.copy_(torch.randn(64, generator=gen) * 0.1)
        conv_g, bn_g = copy.deepcopy(conv).to(DEV).eval(), copy.deepcopy(bn).to(DEV).eval()
        conv.eval(); bn.eval()
        x0 = torch.randn(1, 32, 9, 8, 4, generator=gen)
        with torch.no_grad():
            ref = torch.relu(bn(conv(x0)))
            x2d, d = CF.to_cl2d(x0.to(DEV).contiguous(memory_format=torch.channels_last_3d))
            out, _ = M.conv_bn_act(x2d, d, conv_g, bn_g, relu=True)
        assert rel_err(CF.to_5d(out, d), ref) < 2e-5
        assert int(bn_g.num_batches_tracked) == 0
    finally:
        coocc_b200.set_precision("tf32")
