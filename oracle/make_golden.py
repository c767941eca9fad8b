"""make_golden -- TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the
UNMODIFIED reference Python (imported from /root/reference through oracle/refshim.py) on the
seeded synthetic inputs of coocc_b200.synthetic.  Run in the build container only:

    python -m oracle.make_golden

The reference has no tests of its own for BiFuser_N / CustomResNet3D / FPN3D / OccHead / the
render block, so these fixtures are what pins the oracle (and, through it, the
CUDA path) to the reference for those rows.  Large tensors are stored as strided samples
plus sums to keep the fixtures small.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import coocc_b200  # noqa: E402
from coocc_b200 import synthetic as S  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STRIDE = 97      # sample stride for large float tensors


def sample(t, stride=None):
    f = t.detach().reshape(-1)
    return f[::(stride or SAMPLE_STRIDE["cur"])].numpy().astype(np.float32)


SAMPLE_STRIDE = {"cur": STRIDE}
BIG_STRIDE = 1009       # r50-size tensors (10^7 elements): every 1009th element + the three sums


def stats(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()], dtype=np.float64)


def sub(prefix, p):
    return {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}


def golden_for(name, ns):
    """Fixture of one synthetic configuration.  The [K,N] neighbour tables are stored whole (also for the r50-size
    fixture, K=2 -- the benchmarked configuration); large float tensors as strided samples + sums, the stride is
    recorded under `stride`."""
    cfg = S.CONFIGS[name]
    X, Y, Z = cfg["grid"]
    SAMPLE_STRIDE["cur"] = BIG_STRIDE if X * Y * Z * cfg["C"] > (1 << 22) else STRIDE
    C, K = cfg["C"], cfg["K"]
    inp = S.make_inputs(name)
    P = S.make_params(name)
    torch.use_deterministic_algorithms(True)
    g = {"stride": np.int64(SAMPLE_STRIDE["cur"])}
    nc = dict(type="SyncBN", requires_grad=True)

    # ---- GSFusion (reference BiFuser_N, torch tie order) -----------------------------------
    fuser = ns.BiFuser_N(C, C, knum=K)
    fuser.load_state_dict(P["occ_fuser"], strict=True)
    fuser.train()
    img = inp["img_voxel_feats"].clone().requires_grad_(True)
    pts = inp["pts_voxel_feats"].clone().requires_grad_(True)
    inds_img = torch.nonzero(img.sum(1))
    inds_pts = torch.nonzero(pts.sum(1))
    g["n_img"], g["n_pts"] = np.int64(len(inds_img)), np.int64(len(inds_pts))
    nn_img = fuser.fps_NN_fast(inds_pts.contiguous(), inds_img.contiguous(), 2048, 6, 200, 13.3, K)
    nn_pts = fuser.fps_NN_fast(inds_img.contiguous(), inds_pts.contiguous(), 2048, 6, 200, 13.3, K)
    g["nn_img"] = nn_img.numpy().astype(np.int32)
    g["nn_pts"] = nn_pts.numpy().astype(np.int32)
    fused = fuser(img, pts)
    g["fused_sample"], g["fused_stats"] = sample(fused), stats(fused)
    wsum = torch.linspace(-1, 1, fused.numel()).reshape(fused.shape)
    (fused * wsum).sum().backward()
    g["fused_dimg_sample"], g["fused_dimg_stats"] = sample(img.grad), stats(img.grad)
    g["fused_dpts_sample"], g["fused_dpts_stats"] = sample(pts.grad), stats(pts.grad)
    g["fused_dknn_w"] = fuser.knn_enc[0].weight.grad.numpy().astype(np.float32)[:, ::5]
    g["fused_dknn_b"] = fuser.knn_enc[0].bias.grad.numpy().astype(np.float32)

    # ---- dense conv stack on a fixed seeded input -------------------------------------------
    gx = torch.Generator().manual_seed(1234)
    x = (torch.randn(1, C, *cfg["grid"], generator=gx) * 0.5).requires_grad_(True)
    planes = [C, 2 * C, 4 * C, 8 * C]
    enc = ns.CustomResNet3D(depth=18, n_input_channels=C, block_inplanes=planes,
                            out_indices=(0, 1, 2, 3), norm_cfg=nc)
    enc.load_state_dict(P["semantic_encoder"], strict=True)
    neck = ns.FPN3D(with_cp=True, in_channels=planes, out_channels=2 * C, norm_cfg=nc)
    neck.load_state_dict(P["semantic_neck"], strict=True)
    head = ns.OccHead(norm_cfg=nc, soft_weights=True, cascade_ratio=2, sample_from_voxel=False,
                      sample_from_img=False, final_occ_size=[2 * s for s in cfg["grid"]],
                      fine_topk=15000, empty_idx=0, num_level=4, in_channels=[2 * C] * 4,
                      out_channel=17, point_cloud_range=[-50, -50, -5.0, 50, 50, 3.0])
    head.load_state_dict(P["pts_bbox_head"], strict=True)
    for m in (enc, neck, head):
        m.train()
    mid = enc(x)
    nk = neck(mid)
    o = head.forward_coarse_voxel(nk)
    occ, feats = o["occ"][0], o["out_voxel_feats"][0]
    for i, t in enumerate(mid):
        g["mid%d_sample" % i], g["mid%d_stats" % i] = sample(t), stats(t)
    for i, t in enumerate(nk):
        g["neck%d_sample" % i], g["neck%d_stats" % i] = sample(t), stats(t)
    g["occ_sample"], g["occ_stats"] = sample(occ), stats(occ)
    g["occfeat_sample"], g["occfeat_stats"] = sample(feats), stats(feats)
    wocc = torch.linspace(-1, 1, occ.numel()).reshape(occ.shape)
    (occ * wocc).sum().backward()
    g["stack_dx_sample"], g["stack_dx_stats"] = sample(x.grad), stats(x.grad)
    g["stack_dw_proj"] = enc.input_proj[0].weight.grad.reshape(-1)[::13].numpy().astype(np.float32)
    g["stack_dw_pred"] = head.occ_pred_conv[3].weight.grad.reshape(-1).numpy().astype(np.float32)

    # ---- render block (reference lines executed as-is) -------------------------------------
    mlp_s = ns.MLP(input_dim=C, output_dim=1, net_depth=1, skip_layer=None)
    mlp_r = ns.MLP(input_dim=C, output_dim=3, net_depth=3, skip_layer=None)
    mlp_s.load_state_dict(sub("sigma_head.", P["render"]))
    mlp_r.load_state_dict(sub("rgb_head.", P["render"]))
    gv = torch.Generator().manual_seed(4321)
    vf = (torch.randn(1, C, *cfg["grid"], generator=gv) * 0.5).requires_grad_(True)
    rgbs, depths, losses = refshim.reference_render_block(
        vf, inp["geom"].clone(), mlp_s, mlp_r, inp["gt_depth"], inp["gt_img"])
    g["render_rgbs_sample"], g["render_rgbs_stats"] = sample(rgbs), stats(rgbs)
    g["render_depths_sample"], g["render_depths_stats"] = sample(depths), stats(depths)
    g["loss_depth_render"] = np.float64(losses["loss_depth_render"].item())
    g["loss_rgb"] = np.float64(losses["loss_rgb"].item())
    (losses["loss_depth_render"] + losses["loss_rgb"]).backward()
    g["render_dvf_sample"], g["render_dvf_stats"] = sample(vf.grad), stats(vf.grad)
    g["render_dw_sigma_out"] = mlp_s.output_layer.weight.grad.reshape(-1).numpy().astype(np.float32)
    g["render_dw_rgb_out"] = mlp_r.output_layer.weight.grad.reshape(-1).numpy().astype(np.float32)
    return g


def golden_losses(ns):
    """Reference OccHead.loss_voxel (occ_head.py:267-293) and its four loss functions, run as they are
    on seeded logits / labels of the c1 grid (labels at twice the resolution) and on a ratio-1 case."""
    import sys
    mod = sys.modules["projects.mmdet3d_plugin.coocc.dense_heads.occ_head"]
    cfg = S.CONFIGS["c1"]
    grid = cfg["grid"]
    head = ns.OccHead(norm_cfg=dict(type="SyncBN", requires_grad=True), soft_weights=True, cascade_ratio=2,
                      sample_from_voxel=False, sample_from_img=False, final_occ_size=[2 * s for s in grid],
                      fine_topk=15000, empty_idx=0, num_level=4, in_channels=[64] * 4, out_channel=17,
                      point_cloud_range=[-50, -50, -5.0, 50, 50, 3.0])
    g = {}
    gen = torch.Generator().manual_seed(2024)
    logits = (torch.randn(1, 17, *grid, generator=gen) * 2.0)
    logits[:, 0] += 1.5                                   # free space dominates, like a trained head
    gt = S.make_gt_occ(grid, 2, seed=0)
    seen = {}
    orig = mod.CE_ssc_loss

    def spy(pred, target, *a, **k):
        seen["tv"] = target.clone()
        return orig(pred, target, *a, **k)

    mod.CE_ssc_loss = spy
    try:
        x = logits.clone().requires_grad_(True)
        ld = head.loss_voxel(x, gt.clone(), tag="c_0")
    finally:
        mod.CE_ssc_loss = orig
    g["tv"] = seen["tv"].numpy().astype(np.uint8)
    names = ["loss_voxel_ce_c_0", "loss_voxel_sem_scal_c_0", "loss_voxel_geo_scal_c_0", "loss_voxel_lovasz_c_0"]
    g["losses"] = np.array([ld[k].item() for k in names], dtype=np.float64)
    coef = [1.0, 0.7, 1.3, 0.9]
    sum(c * ld[k] for c, k in zip(coef, names)).backward()
    g["coef"] = np.array(coef)
    g["dlogits_sample"], g["dlogits_stats"] = sample(x.grad), stats(x.grad)
    for i, k in enumerate(names):                          # per-loss gradients
        x2 = logits.clone().requires_grad_(True)
        head.loss_voxel(x2, gt.clone(), tag="c_0")[k].backward()
        g["d%d_sample" % i], g["d%d_stats" % i] = sample(x2.grad), stats(x2.grad)
    # ratio 1 (labels already at the output resolution), a class-poor label set
    gt1 = seen["tv"].clone()
    gt1[(gt1 > 5) & (gt1 != 255)] = 0
    ld1 = head.loss_voxel(logits.clone(), gt1.clone(), tag="c_0")
    g["losses_r1"] = np.array([ld1[k].item() for k in names], dtype=np.float64)
    g["class_weights"] = head.class_weights.numpy().astype(np.float64)
    return g


def golden_eval():
    """Reference evaluation_semantic / fast_hist (coocc_ray.py:659-684, 726-730) on seeded logits of the c1
    grid against labels at twice the resolution, with and without a visibility mask."""
    grid = S.CONFIGS["c1"]["grid"]
    gen = torch.Generator().manual_seed(77)
    pred = torch.randn(1, 17, *grid, generator=gen) * 2.0
    pred[:, 0] += 1.0
    gt = S.make_gt_occ(grid, 2, seed=1)
    vis = (torch.rand(gt.shape, generator=gen) < 0.6).to(torch.uint8)
    g = {}
    g["sc"], _ = refshim.reference_evaluation_semantic(pred, gt, "SC")
    g["ssc"], g["ssc_vis"] = refshim.reference_evaluation_semantic(pred, gt, "SSC", vis)
    # labels at the prediction's own resolution (identity up-sampling)
    gt1 = gt[:, ::2, ::2, ::2].contiguous()
    g["ssc_r1"], _ = refshim.reference_evaluation_semantic(pred, gt1, "SSC")
    return {k: np.asarray(v, dtype=np.int64) for k, v in g.items()}


LSS_GRID = dict(xbound=[-25.0, 25.0, 1.0], ybound=[-25.0, 25.0, 1.0], zbound=[-5.0, 3.0, 1.0], dbound=[2.0, 34.0, 0.5])
LSS_DATA = dict(input_size=(256, 704))


def lss_inputs(seed=0, C=32, cams=6):
    """Seeded inputs of the Lift-Splat golden case: a 50x50x8 grid, 6 cameras, 64 depth bins, 16x44 feature maps."""
    from oracle import lss as OLS
    rig = S.make_camera_rig(cams, seed)
    gen = torch.Generator().manual_seed(seed + 31)
    fH, fW = LSS_DATA["input_size"][0] // 16, LSS_DATA["input_size"][1] // 16
    ds = torch.arange(*LSS_GRID["dbound"], dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, LSS_DATA["input_size"][1] - 1, fW).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, LSS_DATA["input_size"][0] - 1, fH).view(1, fH, 1).expand(D, fH, fW)
    frustum = torch.stack((xs, ys, ds), -1)
    depth = torch.softmax(torch.randn(cams, D, fH, fW, generator=gen), 1)
    feat = torch.randn(cams, C, fH, fW, generator=gen)
    dx = torch.Tensor([LSS_GRID[k][2] for k in ("xbound", "ybound", "zbound")])
    bx = torch.Tensor([LSS_GRID[k][0] + LSS_GRID[k][2] / 2.0 for k in ("xbound", "ybound", "zbound")])
    nx = torch.Tensor([(LSS_GRID[k][1] - LSS_GRID[k][0]) / LSS_GRID[k][2] for k in ("xbound", "ybound", "zbound")])
    return dict(rig=rig, frustum=frustum, depth=depth, feat=feat, dx=dx, bx=bx, nx=nx)


def golden_lss():
    from oracle import lss as OLS
    i = lss_inputs()
    r = i["rig"]
    g = {}
    geom = refshim.reference_get_geometry(i["frustum"], r["rots"], r["trans"], r["intrins"], r["post_rots"],
                                          r["post_trans"], r["bda"])
    g["geom_sample"], g["geom_stats"] = sample(geom), stats(geom)
    vol = OLS.lift(i["depth"], i["feat"]).clone().requires_grad_(True)
    out = refshim.reference_voxel_pooling(geom, vol, i["bx"], i["dx"], i["nx"])      # [1,C,Z,X,Y] permuted to [1,C,X,Y,Z]
    g["out_shape"] = np.array(out.shape)
    g["out_sample"], g["out_stats"] = sample(out.contiguous()), stats(out)
    g["occupied"] = np.int64((out.abs().sum(1) != 0).sum())
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    (out * w).sum().backward()
    g["dvol_sample"], g["dvol_stats"] = sample(vol.grad), stats(vol.grad)
    idx = ((geom - (i["bx"] - i["dx"] / 2.)) / i["dx"]).long().reshape(-1, 3)
    g["idx_sample"] = idx[::101].numpy().astype(np.int32)
    return g


FINE_GRID = (20, 20, 4)


def fine_inputs(seed=1):
    """Seeded inputs of the fine-stage golden case (coarse grid 20x20x4, 6 cameras, 16x44 image features)."""
    gen = torch.Generator().manual_seed(seed)
    feats = torch.randn(1, 128, *FINE_GRID, generator=gen)
    occ = torch.randn(1, 17, *FINE_GRID, generator=gen)
    occ[:, 0] += 1.0
    rig = S.make_camera_rig(6, 0)
    img_feats = torch.randn(1, 6, 512, 16, 44, generator=gen) * 0.5
    transform = (rig["rots"], rig["trans"], rig["intrins"], rig["post_rots"], rig["post_trans"], rig["bda"],
                 None, None, None, None, None, None, (torch.tensor([256]), torch.tensor([704])))
    return feats, occ, img_feats, transform


fine_head_params = S.fine_head_params      # (moved to coocc_b200.synthetic; same seeded values)


def golden_fine(ns):
    """Reference OccHead fine / cascade stage (occ_head.py:182-237) + loss_point (:295-312), run as they are: the
    coarse half is fed in directly (forward_coarse_voxel replaced by the seeded tensors), fine_topk small enough
    to exercise the random-subset branch (coordinate_transform.py:19-21) under torch.manual_seed(123)."""
    feats, occ, img_feats, transform = fine_inputs()
    head = ns.OccHead(norm_cfg=dict(type="SyncBN", requires_grad=True), soft_weights=True, cascade_ratio=2,
                      sample_from_voxel=True, sample_from_img=True, final_occ_size=[2 * s for s in FINE_GRID],
                      fine_topk=150, empty_idx=0, num_level=4, in_channels=[256] * 4, out_channel=17,
                      point_cloud_range=[-10.0, -10.0, -5.0, 10.0, 10.0, 3.0])
    head.train()
    head.load_state_dict(fine_head_params(), strict=False)
    f = feats.clone().requires_grad_(True)
    head.forward_coarse_voxel = lambda vf: {"out_voxel_feats": [f], "occ": [occ]}
    torch.manual_seed(123)
    res = head(voxel_feats=[None] * 4, img_feats=[img_feats], transform=transform)
    fc, fo = res["output_coords_fine"][0], res["output_voxels_fine"][0]
    gt = S.make_gt_occ(FINE_GRID, 2, 3)
    ld = head.loss_point(fc, fo, gt, "fine")
    names = ["loss_voxel_ce_fine", "loss_voxel_sem_scal_fine", "loss_voxel_geo_scal_fine", "loss_voxel_lovasz_fine"]
    g = dict(fine_coord=fc.numpy().astype(np.int32), fine_output=fo.detach().numpy().astype(np.float32),
             losses=np.array([ld[k].item() for k in names], dtype=np.float64))
    sum(ld.values()).backward()
    g["dfeats_sample"], g["dfeats_stats"] = sample(f.grad), stats(f.grad)
    g["dw_fine3"] = head.fine_mlp[3].weight.grad.numpy().astype(np.float32)
    return g


GOLDEN_CONFIGS = ("c1", "c1k1", "r50")      # r50: 100x100x8, C=128, K=2 = the reference's own working grid and knum


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    if only:                                 # e.g. `python -m oracle.make_golden r50`
        ns = refshim.load_reference()
        for name in only:
            g = golden_for(name, ns)
            path = os.path.join(OUT, "reference_%s.npz" % name)
            np.savez_compressed(path, **g)
            print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
        return
    gl2 = golden_lss()
    path = os.path.join(OUT, "reference_lss.npz")
    np.savez_compressed(path, **gl2)
    print("lss ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    ge = golden_eval()
    path = os.path.join(OUT, "reference_eval.npz")
    np.savez_compressed(path, **ge)
    print("eval ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    ns = refshim.load_reference()
    gf = golden_fine(ns)
    path = os.path.join(OUT, "reference_fine.npz")
    np.savez_compressed(path, **gf)
    print("fine ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    gl = golden_losses(ns)
    path = os.path.join(OUT, "reference_losses.npz")
    np.savez_compressed(path, **gl)
    print("losses ->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    for name in GOLDEN_CONFIGS:
        g = golden_for(name, ns)
        path = os.path.join(OUT, "reference_%s.npz" % name)
        np.savez_compressed(path, **g)
        print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
