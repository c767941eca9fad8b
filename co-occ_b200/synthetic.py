"""Seeded synthetic inputs and parameters for the fused-voxel hot path.

Shapes follow what the kept upstream code hands to the path:
  img_voxel_feats / pts_voxel_feats  [1,C,X,Y,Z] fp32 *strided views*
      (img: ViewTransformerLSSVoxel.py:120-121, pts: sparse_lidar_enc.py:176),
  geom  [1,N_cam,D,fH,fW,3] ego-frame metres (ViewTransformerLSSBEVDepth.py:117-150),
  gt_img [1,N_cam,3,16fH,16fW] in [0,1], gt_depth [1,N_cam,16fH,16fW] metres, 0 = none
      (img_inputs[0], img_inputs[7]; P/datasets/pipelines/loading.py:129).

Everything is generated on the CPU from torch.Generator(seed) so the oracle and the CUDA
path see bit-identical inputs on any box.
"""
import math

import torch

CONFIGS = {
    # name: grid, C, K, cams, fH, fW, D, p_img, p_pts
    "c1":        dict(grid=(50, 50, 4),    C=32,  K=4, cams=1, fH=4,  fW=4,   D=112, p_img=0.6, p_pts=0.3),
    # K=1 only works in the reference when N_img, N_pts <= 2048 (brute-force branch,
    # bifuser_n.py:55-60); its K=1 FPS branch falls off the function without a return (Q11)
    "c1k1":      dict(grid=(24, 24, 4),    C=32,  K=1, cams=1, fH=4,  fW=4,   D=112, p_img=0.6, p_pts=0.3, min_pts=1),
    "r50":       dict(grid=(100, 100, 8),  C=128, K=2, cams=6, fH=16, fW=44,  D=112, p_img=0.6, p_pts=0.15),
    "r101":      dict(grid=(100, 100, 8),  C=128, K=2, cams=6, fH=56, fW=100, D=112, p_img=0.6, p_pts=0.15),
    "openocc":   dict(grid=(128, 128, 10), C=128, K=2, cams=6, fH=56, fW=100, D=112, p_img=0.6, p_pts=0.15),
    "northstar": dict(grid=(200, 200, 16), C=128, K=2, cams=6, fH=8,  fW=8,   D=96,  p_img=0.6, p_pts=0.15),
}

RENDER_BOX = (100, 100, 8)   # coocc_ray.py:372, hard-coded in the reference


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def make_voxel_feats(grid, C, p_img, p_pts, seed=0, min_pts=2049):
    """Two occupancy-masked feature grids with 2048 < N_pts <= N_img."""
    X, Y, Z = grid
    g = _gen(seed)
    for _ in range(64):
        m_img = torch.rand(Z, X, Y, generator=g) < p_img
        m_pts = torch.rand(Z, Y, X, generator=g) < p_pts
        n_i, n_p = int(m_img.sum()), int(m_pts.sum())
        if min_pts <= n_p <= n_i:
            break
    else:
        raise RuntimeError("could not draw masks with %d <= N_pts <= N_img" % min_pts)
    img = torch.randn(1, C, Z, X, Y, generator=g) * 0.3 * m_img[None, None]
    pts = torch.relu(torch.randn(1, C, Z, Y, X, generator=g) * 0.3 + 0.05) * m_pts[None, None]
    # a ReLU row can be all-zero with probability ~2^-C; force a positive entry so the mask
    # and the occupancy agree (the reference's mask is `sum != 0`, bifuser_n.py:130-131)
    pts[:, 0] = torch.where(m_pts[None] & (pts.sum(1) == 0), torch.full_like(pts[:, 0], 0.125), pts[:, 0])
    img[:, 0] = torch.where(m_img[None] & (img.sum(1) == 0), torch.full_like(img[:, 0], 0.125), img[:, 0])
    # views with the upstream strides: logical [1,C,X,Y,Z]
    return img.permute(0, 1, 3, 4, 2), pts.permute(0, 1, 4, 3, 2)


def make_geom(grid, cams, fH, fW, D, seed=0):
    """Pinhole rays, depths arange(2, 2+0.5*D, 0.5) (dbound, coocc_multi_r50_256x704.py:53).

    When the feature grid is smaller than the reference's hard-coded 100x100x8 render box,
    a sample that is 'inside' the box must also index inside the grid (the reference would
    raise otherwise), so the cameras sit inside the small grid and look towards -x/-y with
    non-positive pitch: rays leave the grid and the box together."""
    X, Y, Z = grid
    g = _gen(seed + 7919)
    depth = torch.arange(D, dtype=torch.float32) * 0.5 + 2.0
    small = X < RENDER_BOX[0] or Y < RENDER_BOX[1] or Z < RENDER_BOX[2]
    geoms = []
    for c in range(cams):
        if small:
            origin = torch.tensor([-50.0 + 0.55 * X, -50.0 + 0.55 * Y, -5.0 + 0.6 * Z])
            yaw0, fov, p_lo, p_hi = math.radians(225.0), math.radians(80.0), -0.35, 0.0
        else:
            origin = torch.tensor([0.3 * math.cos(c), 0.3 * math.sin(c), 0.6])
            yaw0, fov, p_lo, p_hi = math.radians(60.0 * c + 30.0), math.radians(70.0), -0.30, 0.12
        u = (torch.arange(fW, dtype=torch.float32) + 0.5) / fW - 0.5
        v = (torch.arange(fH, dtype=torch.float32) + 0.5) / fH
        yaw = yaw0 + u * fov + (torch.rand(1, generator=g).item() - 0.5) * 0.05
        pitch = p_hi + (p_lo - p_hi) * v
        dirs = torch.stack([torch.cos(pitch)[:, None] * torch.cos(yaw)[None, :],
                            torch.cos(pitch)[:, None] * torch.sin(yaw)[None, :],
                            torch.sin(pitch)[:, None].expand(fH, fW)], -1)       # [fH,fW,3]
        pts = origin[None, None, None, :] + depth[:, None, None, None] * dirs[None]  # [D,fH,fW,3]
        # sub-voxel jitter so that samples do not sit on voxel faces
        pts = pts + (torch.rand(pts.shape, generator=g) - 0.5) * 0.02
        geoms.append(pts)
    return torch.stack(geoms)[None].contiguous()


def make_render_targets(cams, fH, fW, seed=0):
    g = _gen(seed + 104729)
    H, W = 16 * fH, 16 * fW
    gt_img = torch.rand(1, cams, 3, H, W, generator=g)
    keep = torch.rand(1, cams, H, W, generator=g) < 0.1
    gt_depth = (torch.rand(1, cams, H, W, generator=g) * 56.0 + 2.0) * keep
    return gt_img, gt_depth


def make_gt_occ(grid, ratio=2, seed=0, num_cls=17):
    """gt_occ [1, X*r, Y*r, Z*r] int64 labels (0 = free, 1..16 classes, 255 = ignore) at `ratio` times the
    working grid, like the 200x200x16 nuScenes-occupancy labels over the 100x100x8 grid
    (coocc_multi_r50_256x704.py: occ_size vs. voxel grid).  About a quarter of the coarse cells hold
    an object; inside them roughly half of the sub-voxels carry its label, a few carry another class
    (so the reference's torch.mode vote, occ_head.py:269-280, sees majorities, ties and singletons),
    1 % of all sub-voxels are 255."""
    X, Y, Z = grid
    g = _gen(seed + 97)
    coarse = torch.randint(1, num_cls, (X, Y, Z), generator=g)
    occupied = torch.rand(X, Y, Z, generator=g) < 0.25
    up = lambda t: t.repeat_interleave(ratio, 0).repeat_interleave(ratio, 1).repeat_interleave(ratio, 2)
    fine = up(coarse)
    shape = fine.shape
    keep = up(occupied) & (torch.rand(shape, generator=g) < 0.5)
    fine = torch.where(keep, fine, torch.zeros_like(fine))
    other = torch.rand(shape, generator=g) < 0.03
    fine = torch.where(other, torch.randint(1, num_cls, shape, generator=g), fine)
    fine = torch.where(torch.rand(shape, generator=g) < 0.01, torch.full_like(fine, 255), fine)
    return fine[None].contiguous()


def make_camera_rig(cams=6, seed=0, in_h=256, in_w=704):
    """Plausible nuScenes-like calibration for N cameras (img_inputs[1:7] of P/datasets/pipelines/loading.py:129):
    rots [1,N,3,3] camera->ego, trans [1,N,3], intrins [1,N,3,3], post_rots / post_trans (image augmentation:
    resize + crop), bda [1,3,3] (BEV augmentation: small rotation, scale, flip)."""
    g = _gen(seed + 211)
    rots, trans, intr, prot, ptr = [], [], [], [], []
    for c in range(cams):
        yaw = math.radians(360.0 * c / cams + 5.0 * float(torch.rand(1, generator=g)))
        cy, sy = math.cos(yaw), math.sin(yaw)
        # camera axes (x right, y down, z forward) in the ego frame (x forward, y left, z up)
        fwd = torch.tensor([cy, sy, 0.0])
        right = torch.tensor([sy, -cy, 0.0])
        down = torch.tensor([0.0, 0.0, -1.0])
        rots.append(torch.stack([right, down, fwd], 1))
        trans.append(torch.tensor([1.5 * cy, 1.5 * sy, 1.6]) + 0.1 * torch.randn(3, generator=g))
        f = 1260.0 + 20.0 * float(torch.randn(1, generator=g))
        intr.append(torch.tensor([[f, 0.0, 800.0], [0.0, f, 450.0], [0.0, 0.0, 1.0]]))
        sc = in_w / 1600.0 * (1.0 + 0.05 * float(torch.randn(1, generator=g)))
        prot.append(torch.tensor([[sc, 0.0, 0.0], [0.0, sc, 0.0], [0.0, 0.0, 1.0]]))
        ptr.append(torch.tensor([-3.0 * float(torch.rand(1, generator=g)), -(900 * sc - in_h) * 0.9, 0.0]))
    a = math.radians(4.0)
    bda = torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]]) * 1.02
    bda[1] = -bda[1]                     # flip_dy
    st = lambda ts: torch.stack(ts)[None].float()
    return dict(rots=st(rots), trans=st(trans), intrins=st(intr), post_rots=st(prot), post_trans=st(ptr), bda=bda[None])


def make_img_feats(cams, fH, fW, seed=0, channels=512):
    """img_feats[0] of extract_img_feat (coocc_ray.py:164-197): the SECONDFPN output [1,N,512,fH,fW] the OccHead fine
    stage samples (occ_head.py:199-203)."""
    g = _gen(seed + 1299709)
    return torch.randn(1, cams, channels, fH, fW, generator=g) * 0.5


def make_transform(cams, fH, fW, seed=0):
    """`transform = img_inputs[1:]` (coocc_ray.py:348): (rots, trans, intrins, post_rots, post_trans, bda, gt_depths,
    sensor2sensors, denorm_imgs, aabb, intrin_nerf, c2ws, img_size) -- the head reads entries 0-5 and the last one
    (img_size = (H, W) per sample, occ_head.py:222-227); the others are placeholders here."""
    rig = make_camera_rig(cams, seed, in_h=16 * fH, in_w=16 * fW)
    return (rig["rots"], rig["trans"], rig["intrins"], rig["post_rots"], rig["post_trans"], rig["bda"],
            None, None, None, None, None, None, (torch.tensor([16 * fH]), torch.tensor([16 * fW])))


def make_inputs(name, seed=0, with_render=True):
    cfg = CONFIGS[name]
    img, pts = make_voxel_feats(cfg["grid"], cfg["C"], cfg["p_img"], cfg["p_pts"], seed,
                                min_pts=cfg.get("min_pts", 2049))
    out = dict(img_voxel_feats=img, pts_voxel_feats=pts)
    if with_render:
        out["geom"] = make_geom(cfg["grid"], cfg["cams"], cfg["fH"], cfg["fW"], cfg["D"], seed)
        out["gt_img"], out["gt_depth"] = make_render_targets(cfg["cams"], cfg["fH"], cfg["fW"], seed)
    return out


# ----------------------------------------------------------------------------------------
# Parameters, keyed like the reference modules' state_dicts
# ----------------------------------------------------------------------------------------
def _conv_w(g, cout, cin, k):
    fan_out = cout * k ** 3
    return torch.randn(cout, cin, k, k, k, generator=g) * math.sqrt(2.0 / fan_out)


def _bn(g, p, prefix, c):
    p[prefix + ".weight"] = torch.rand(c, generator=g) * 0.5 + 0.75
    p[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
    p[prefix + ".running_mean"] = torch.zeros(c)
    p[prefix + ".running_var"] = torch.ones(c)
    p[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _linear(g, p, prefix, cout, cin, bias_scale=0.05):
    a = math.sqrt(6.0 / (cin + cout))
    p[prefix + ".weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * a
    p[prefix + ".bias"] = torch.randn(cout, generator=g) * bias_scale


def fuser_params(C, K, seed=0):
    g = _gen(seed + 11)
    p = {}
    p["con_enc.0.weight"] = _conv_w(g, 2 * C, 4 * C, 3)
    _bn(g, p, "con_enc.1", 2 * C)
    p["con_enc.3.weight"] = _conv_w(g, C, 2 * C, 3)
    _bn(g, p, "con_enc.4", C)
    _linear(g, p, "knn_enc.0", C, C * K)
    return p


def resnet3d_params(C, planes, layers=(2, 2, 2, 2), strides=(1, 2, 2, 2), seed=0):
    g = _gen(seed + 23)
    p = {}
    p["input_proj.0.weight"] = _conv_w(g, planes[0], C, 1)
    _bn(g, p, "input_proj.1", planes[0])
    cin = planes[0]
    for s, n in enumerate(layers):
        for b in range(n):
            pre = "layers.%d.%d." % (s, b)
            stride = strides[s] if b == 0 else 1
            p[pre + "conv1.weight"] = _conv_w(g, planes[s], cin, 3)
            _bn(g, p, pre + "bn1", planes[s])
            p[pre + "conv2.weight"] = _conv_w(g, planes[s], planes[s], 3)
            _bn(g, p, pre + "bn2", planes[s])
            if b == 0 and (stride != 1 or cin != planes[s]):
                p[pre + "downsample.0.weight"] = _conv_w(g, planes[s], cin, 1)
                _bn(g, p, pre + "downsample.1", planes[s])
            cin = planes[s]
    return p


def fpn3d_params(in_channels, out_channels, seed=0):
    g = _gen(seed + 37)
    p = {}
    for i, c in enumerate(in_channels):
        p["lateral_convs.%d.0.conv.weight" % i] = _conv_w(g, out_channels, c, 1)
        _bn(g, p, "lateral_convs.%d.0.bn" % i, out_channels)
        p["fpn_convs.%d.0.conv.weight" % i] = _conv_w(g, out_channels, out_channels, 3)
        _bn(g, p, "fpn_convs.%d.0.bn" % i, out_channels)
    return p


def occhead_params(in_channels, num_cls=17, seed=0):
    g = _gen(seed + 41)
    p = {}
    mid = in_channels[0] // 2
    for i, c in enumerate(in_channels):
        p["occ_convs.%d.0.weight" % i] = _conv_w(g, c // 2, c, 3)
        _bn(g, p, "occ_convs.%d.1" % i, c // 2)
    p["occ_pred_conv.0.weight"] = _conv_w(g, mid // 2, mid, 1)
    _bn(g, p, "occ_pred_conv.1", mid // 2)
    p["occ_pred_conv.3.weight"] = _conv_w(g, num_cls, mid // 2, 1)
    p["voxel_soft_weights.0.weight"] = _conv_w(g, mid // 2, mid, 1)
    _bn(g, p, "voxel_soft_weights.1", mid // 2)
    p["voxel_soft_weights.3.weight"] = _conv_w(g, len(in_channels), mid // 2, 1)
    return p


def render_params(C, width=256, seed=0):
    """sigma_head = MLP(C,1,depth 1), rgb_head = MLP(C,3,depth 3) (coocc_ray.py:112-113)."""
    g = _gen(seed + 53)
    p = {}
    _linear(g, p, "sigma_head.hidden_layers.0", width, C)
    _linear(g, p, "sigma_head.output_layer", 1, width)
    # keep densities O(0.1) per metre so that transmittance is neither 0 nor 1 everywhere
    p["sigma_head.output_layer.weight"] *= 0.5
    p["sigma_head.output_layer.bias"] += 0.05
    cin = C
    for i in range(3):
        _linear(g, p, "rgb_head.hidden_layers.%d" % i, width, cin)
        cin = width
    _linear(g, p, "rgb_head.output_layer", 3, width)
    for h in ("sigma_head", "rgb_head"):
        p[h + ".posi_encoder.scales"] = torch.tensor([2 ** i for i in range(10)])   # nerf_mlp.py:188-190 (Q8)
    return p


def fine_head_params(seed=0):
    """State-dict entries of the OccHead fine / cascade stage (occ_head.py:58-82) with the reference's hard-wired
    layer shapes (512 -> 128 image conv, 128 -> 64, 192 -> 64 -> 17)."""
    g = torch.Generator().manual_seed(seed + 71)
    r = lambda *s: torch.randn(*s, generator=g)
    return {"img_mlp_0.0.weight": r(128, 512, 1, 1) * 0.04, "img_mlp_0.0.bias": r(128) * 0.1,
            "img_mlp_0.1.weight": 1 + 0.1 * r(128), "img_mlp_0.1.bias": 0.1 * r(128),
            "img_mlp.0.weight": r(64, 128) * 0.09, "img_mlp.0.bias": r(64) * 0.1,
            "img_mlp.1.weight": 1 + 0.1 * r(64), "img_mlp.1.bias": 0.1 * r(64),
            "fine_mlp.0.weight": r(64, 192) * 0.07, "fine_mlp.0.bias": r(64) * 0.1,
            "fine_mlp.1.weight": 1 + 0.1 * r(64), "fine_mlp.1.bias": 0.1 * r(64),
            "fine_mlp.3.weight": r(17, 64) * 0.12, "fine_mlp.3.bias": r(17) * 0.1}


def sparse_enc_params(input_channel=4, base=16, out_channel=128, seed=0):
    """State dict of SparseLiDAREnc8x (sparse_lidar_enc.py:125-160) with spconv's parameter names and weight layout
    [Cout, kz, ky, kx, Cin]; the GroupNorm biases are non-zero so that conv_input's degenerate GroupNorm(16, 16)
    (output = relu(bias), Q13) does not zero the whole encoder."""
    g = _gen(seed + 67)

    def conv(p, name, cin, cout, bias):
        p[name + ".weight"] = torch.randn(cout, 3, 3, 3, cin, generator=g) * math.sqrt(2.0 / (27 * cin))
        if bias:
            p[name + ".bias"] = torch.randn(cout, generator=g) * 0.1

    def bn(p, name, c):
        _bn(g, p, name, c)

    def gn(p, name, c):
        p[name + ".weight"] = torch.rand(c, generator=g) * 0.5 + 0.75
        p[name + ".bias"] = torch.rand(c, generator=g) * 0.5 + 0.1

    p = {}
    conv(p, "conv_input.0", input_channel, base, True)
    gn(p, "conv_input.1", base)
    cin = base
    for i, st in enumerate(("conv1", "conv2", "conv3")):
        cout = base * 2 ** (i + 1)
        conv(p, st + ".0.0", cin, cout, False)
        bn(p, st + ".0.1", cout)
        for b in (1, 2):
            conv(p, "%s.%d.net.0" % (st, b), cout, cout, False)
            bn(p, "%s.%d.net.1" % (st, b), cout)
            conv(p, "%s.%d.net.3" % (st, b), cout, cout, False)
            bn(p, "%s.%d.net.4" % (st, b), cout)
        cin = cout
    conv(p, "conv_out.0", cin, out_channel, True)
    gn(p, "conv_out.1", out_channel)
    return p


def make_lidar_voxels(sparse_shape_xyz, n, input_channel=4, seed=0):
    """HardSimpleVFE output: voxel_features [N, input_channel], coors [N,4] (batch, z, y, x) of N distinct voxels,
    in the lexicographic order the voxelizer's dense scan produces."""
    g = _gen(seed + 83)
    W, H, D = sparse_shape_xyz
    lin = torch.randperm(D * H * W, generator=g)[:n].sort().values
    z, y, x = lin // (H * W), (lin // W) % H, lin % W
    coors = torch.stack([torch.zeros_like(z), z, y, x], 1).int()
    feats = torch.randn(n, input_channel, generator=g)
    return feats, coors


def make_params(name, seed=0):
    cfg = CONFIGS[name]
    C, K = cfg["C"], cfg["K"]
    planes = [C, 2 * C, 4 * C, 8 * C]
    neck_out = 2 * C
    return dict(
        occ_fuser=fuser_params(C, K, seed),
        semantic_encoder=resnet3d_params(C, planes, seed=seed),
        semantic_neck=fpn3d_params(planes, neck_out, seed),
        pts_bbox_head=occhead_params([neck_out] * 4, 17, seed),
        render=render_params(C, 256, seed),
    )
