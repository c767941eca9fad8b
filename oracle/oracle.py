"""oracle.oracle -- TEST INFRASTRUCTURE ONLY (parity checker; never imported by the product).

Portable CPU restatement (numpy + torch-CPU fp32) of the reference's fused-voxel hot path.
It must travel to the GPU box, where /root/reference does not exist, so nothing here reads
the reference tree.  Every function cites the reference file:line it follows (paths relative
to the reference root, P/ = projects/mmdet3d_plugin/).

Parity pinning (see DESIGN.md "Oracle"):
  * FPS / ball-query: reference known-answer tests (tests/test_oracle_golden.py).
  * everything else: there are NO reference tests for this path; the restatement
    is pinned against outputs of the *real* reference Python run in the build container
    (oracle/refshim.py + oracle/make_golden.py -> tests/golden/*.npz, tests/test_oracle_vs_reference.py).

Parameters are passed as dicts keyed by the reference modules' state_dict names, so a
reference checkpoint (or `module.state_dict()`) can be fed in unchanged.

Tie rule (SURVEY §7 hard part 1): `tie="canonical"` orders equal distances by ascending key
index and resolves duplicate scatter targets as last writer in (representative, slot) order;
`tie="torch"` calls torch.topk like the reference does (order among ties unspecified).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

FPS_NUM = 2048          # P/coocc/fuser/bifuser_n.py:137
BALL_RADIUS = 6         # :137
BALL_SAMPLES = 200      # :137
DIST_THRESH = 13.3      # :137


# ----------------------------------------------------------------------------------------
# GSFusion index pipeline
# ----------------------------------------------------------------------------------------
def occupied_indices(feats):
    """P/coocc/fuser/bifuser_n.py:130-131: rows (b,x,y,z) of voxels whose fp32 channel sum is
    non-zero, lexicographic order (torch.nonzero)."""
    return torch.nonzero(feats.sum(1))


def _int_d2(a, b):
    """squared distances between integer coordinate sets a [M,3], b [N,3] -> int64 [M,N]."""
    a = a.astype(np.int32)          # voxel coordinates < 2^15: the squares fit int32, widened on return
    b = b.astype(np.int32)
    out = np.zeros((a.shape[0], b.shape[0]), dtype=np.int64)
    for c in range(a.shape[1]):
        d = a[:, None, c] - b[None, :, c]
        out += d * d
    return out


def topk_canonical(d2, k):
    """k smallest per row ordered by (d2 asc, column asc).  Returns (d2_sel, idx) [M,k]."""
    M, N = d2.shape
    keyed = d2 * np.int64(N) + np.arange(N, dtype=np.int64)[None, :]
    if k < N:
        part = np.partition(keyed, k - 1, axis=1)[:, :k]
    else:
        part = keyed
    part = np.sort(part, axis=1)
    return part // N, part % N


def fps_nn_fast(query, key, num, fps_num=FPS_NUM, radius=BALL_RADIUS,
                max_cluster_samples=BALL_SAMPLES, dist_thresh=DIST_THRESH, tie="canonical",
                return_parts=False):
    """P/coocc/fuser/bifuser_n.py:38-125.

    query [Nq,4], key [Nk,4] int64 rows (b,x,y,z).  Returns int64 [Nq] (num==1) or [num,Nq];
    -1 where no representative with a valid neighbour claimed the query (:48-50).
    """
    q4 = query.cpu().numpy()
    k4 = key.cpu().numpy()
    Nq = q4.shape[0]
    q = q4[:, 1:]            # :51  (batch column dropped)
    k = k4[:, 1:]            # :52
    out = np.full((num, Nq), -1, dtype=np.int64)
    parts = {}
    thr = np.float32(dist_thresh)

    if Nq <= fps_num:
        if num != 1:
            # :88-93 -- the 2-D boolean mask of topk(...) is applied to a 1-D row: the
            # reference raises here; the restatement raises the same type.
            raise IndexError("too many indices for tensor of dimension 1")
        d2 = _int_d2(q, k)                                   # :56
        sel, idx = topk_canonical(d2, 1)                     # :57 (min -> first minimum)
        valid = np.sqrt(sel[:, 0].astype(np.float32)) < thr  # :58
        out[0, valid] = idx[valid, 0]                        # :59
        res = torch.from_numpy(out[0])
        return (res, parts) if return_parts else res

    qf = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))[None]
    rep_idx = ops.furthest_point_sample(qf, fps_num)[0].numpy().astype(np.int64)   # :63 / :97
    rep = q[rep_idx]                                                               # :64 / :98
    # :67 / :101 -- the reference materialises the whole [2048, Nk] distance matrix; the restatement walks it in
    # row blocks (same values; ~64 M elements at a time so the north-star grid, Nk ~ 4e5, fits in memory)
    rows = max(1, min(len(rep), (64 << 20) // max(len(k), 1)))
    vals, idxs = [], []
    for r0 in range(0, len(rep), rows):
        d2 = _int_d2(rep[r0:r0 + rows], k)
        if tie == "canonical":
            sel, ix = topk_canonical(d2, num)
            vals.append(np.sqrt(sel.astype(np.float32)))
            idxs.append(ix)
        else:
            dist = torch.from_numpy(np.sqrt(d2.astype(np.float32)))
            if num == 1:
                v, i = dist.min(-1)                                                # :68
                vals.append(v[:, None].numpy())
                idxs.append(i[:, None].numpy())
            else:
                v, i = torch.topk(dist, num, dim=-1, largest=False)                # :103
                vals.append(v.numpy())
                idxs.append(i.numpy())
    val, idx = np.concatenate(vals), np.concatenate(idxs)
    valid = val < thr                                                              # :69 / :107
    repf = torch.from_numpy(np.ascontiguousarray(rep, dtype=np.float32))[None]
    group = ops.ball_query(0, radius, max_cluster_samples, qf, repf)[0].numpy().astype(np.int64)  # :71 / :109
    for i in range(num):
        # :77-85 / :115-123: duplicate targets -> last writer in (rep, slot) order wins
        row = out[i]
        for r in np.nonzero(valid[:, i])[0]:
            row[group[r]] = idx[r, i]
    parts = dict(rep_idx=rep_idx, topk_idx=np.where(valid, idx, -1), topk_d2=np.where(valid, val * val, -1),
                 valid=valid, group=group)
    res = torch.from_numpy(out[0] if num == 1 else out)
    return (res, parts) if return_parts else res


# ----------------------------------------------------------------------------------------
# GSFusion module forward
# ----------------------------------------------------------------------------------------
def _take(feats_cl, inds):
    return feats_cl[inds[:, 0], inds[:, 1], inds[:, 2], inds[:, 3]]


def _bn_train(x, p, prefix, eps=1e-5):
    return F.batch_norm(x, None, None, p[prefix + ".weight"], p[prefix + ".bias"], True, 0.1, eps)


def bifuser_forward(p, img_voxel_feats, pts_voxel_feats, knum, tie="canonical", return_parts=False):
    """P/coocc/fuser/bifuser_n.py:127-174 (BiFuser_N.forward), training-mode BatchNorm.

    p: state_dict-keyed params `con_enc.{0,3}.weight`, `con_enc.{1,4}.{weight,bias}`,
    `knn_enc.0.{weight,bias}`.
    """
    B, C, H, W, L = img_voxel_feats.shape
    inds_img = occupied_indices(img_voxel_feats)                       # :130
    inds_pts = occupied_indices(pts_voxel_feats)                       # :131
    img_cl = img_voxel_feats.permute(0, 2, 3, 4, 1)                    # :133
    pts_cl = pts_voxel_feats.permute(0, 2, 3, 4, 1)
    Wk, bk = p["knn_enc.0.weight"], p["knn_enc.0.bias"]

    def enc(rows):
        return F.relu(F.linear(rows, Wk, bk))                          # :32-36

    sel_pts = _take(pts_cl, inds_pts)                                  # :135
    nn_img, parts_i = fps_nn_fast(inds_pts, inds_img, knum, tie=tie, return_parts=True)   # :137
    if knum == 1:
        near_img = _take(img_cl, inds_img[nn_img])                     # :139-140 (-1 -> last row, Q3)
    else:
        near_img = torch.cat([_take(img_cl, inds_img[nn_img[i]]) for i in range(knum)], 1)  # :142-147
    near_img = enc(near_img) * sel_pts                                 # :148

    sel_img = _take(img_cl, inds_img)                                  # :150
    nn_pts, parts_p = fps_nn_fast(inds_img, inds_pts, knum, tie=tie, return_parts=True)   # :151
    if knum == 1:
        near_pts = _take(pts_cl, inds_pts[nn_pts])                     # :153-154
    else:
        # :156-161 -- indexes the *image* table with pts-table positions
        near_pts = torch.cat([_take(pts_cl, inds_img[nn_pts[i]]) for i in range(knum)], 1)
    near_pts = enc(near_pts) * sel_img                                 # :162

    fused_img = torch.zeros(B, H, W, L, C)                             # :164
    fused_img = fused_img.index_put((inds_pts[:, 0], inds_pts[:, 1], inds_pts[:, 2], inds_pts[:, 3]), near_img)  # :165
    fused_pts = torch.zeros(B, H, W, L, C)                             # :168
    fused_pts = fused_pts.index_put((inds_img[:, 0], inds_img[:, 1], inds_img[:, 2], inds_img[:, 3]), near_pts)  # :169
    all_feats = torch.cat([img_cl, pts_cl, fused_img, fused_pts], -1)  # :172
    x = all_feats.permute(0, 4, 1, 2, 3)
    x = F.conv3d(x, p["con_enc.0.weight"], None, 1, 1)                 # :24
    x = F.relu(_bn_train(x, p, "con_enc.1"))                           # :25-26
    x = F.conv3d(x, p["con_enc.3.weight"], None, 1, 1)                 # :27
    x = F.relu(_bn_train(x, p, "con_enc.4"))                           # :28-29
    if return_parts:
        return x, dict(inds_img=inds_img, inds_pts=inds_pts, nn_img=nn_img, nn_pts=nn_pts,
                       parts_img=parts_i, parts_pts=parts_p, all_feats=all_feats)
    return x


# ----------------------------------------------------------------------------------------
# Dense 3D conv decoder / head
# ----------------------------------------------------------------------------------------
def resnet3d_forward(p, x, layers=(2, 2, 2, 2), strides=(1, 2, 2, 2), out_indices=(0, 1, 2, 3)):
    """P/coocc/backbones/resnet3d.py:196-205 with BasicBlock :34-64 (depth 10/18/34)."""
    x = F.conv3d(x, p["input_proj.0.weight"])                          # :138-143
    x = F.relu(_bn_train(x, p, "input_proj.1"))
    res = []
    for s, nblk in enumerate(layers):
        for b in range(nblk):
            pre = "layers.%d.%d." % (s, b)
            stride = strides[s] if b == 0 else 1
            out = F.conv3d(x, p[pre + "conv1.weight"], None, stride, 1)        # :49
            out = F.relu(_bn_train(out, p, pre + "bn1"))                       # :50-51
            out = F.conv3d(out, p[pre + "conv2.weight"], None, 1, 1)           # :53
            out = _bn_train(out, p, pre + "bn2")                               # :54
            if (pre + "downsample.0.weight") in p:                             # :56-57
                idn = F.conv3d(x, p[pre + "downsample.0.weight"], None, stride, 0)
                idn = _bn_train(idn, p, pre + "downsample.1")
            else:
                idn = x
            x = F.relu(out + idn)                                              # :59-60
        if s in out_indices:
            res.append(x)
    return res


def fpn3d_forward(p, inputs):
    """P/coocc/necks/fpn3d.py:70-108 (checkpointing does not change values)."""
    n = len(inputs)
    lat = []
    for i in range(n):
        y = F.conv3d(inputs[i], p["lateral_convs.%d.0.conv.weight" % i])
        lat.append(F.relu(_bn_train(y, p, "lateral_convs.%d.0.bn" % i)))
    for i in range(n - 1, 0, -1):                                              # :91-94
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:],
                                                mode="trilinear", align_corners=False)
    outs = []
    for i in range(n):
        y = F.conv3d(lat[i], p["fpn_convs.%d.0.conv.weight" % i], None, 1, 1)
        outs.append(F.relu(_bn_train(y, p, "fpn_convs.%d.0.bn" % i)))
    return outs


def occhead_coarse_forward(p, voxel_feats, soft_weights=True):
    """P/coocc/dense_heads/occ_head.py:149-171 (forward_coarse_voxel)."""
    occs = []
    for i, f in enumerate(voxel_feats):
        y = F.conv3d(f, p["occ_convs.%d.0.weight" % i], None, 1, 1)            # :102-110
        occs.append(F.relu(_bn_train(y, p, "occ_convs.%d.1" % i)))
    nlev = len(voxel_feats)
    if soft_weights:
        w = F.conv3d(occs[0], p["voxel_soft_weights.0.weight"])                # :123-132
        w = F.relu(_bn_train(w, p, "voxel_soft_weights.1"))
        w = F.conv3d(w, p["voxel_soft_weights.3.weight"])
        w = torch.softmax(w, dim=1)                                            # :157
    else:
        w = torch.ones([occs[0].shape[0], nlev, 1, 1, 1]) / nlev               # :159
    H, W, D = occs[0].shape[2:]
    feats = 0
    for f, wi in zip(occs, torch.unbind(w, dim=1)):                            # :163-165
        f = F.interpolate(f, size=[H, W, D], mode="trilinear", align_corners=False)
        feats = feats + f * wi.unsqueeze(1)
    y = F.conv3d(feats, p["occ_pred_conv.0.weight"])                           # :113-119
    y = F.relu(_bn_train(y, p, "occ_pred_conv.1"))
    occ = F.conv3d(y, p["occ_pred_conv.3.weight"])
    return feats, occ


# ----------------------------------------------------------------------------------------
# Volume-render regulariser
# ----------------------------------------------------------------------------------------
def mlp_forward(p, prefix, x, depth):
    """P/utils/nerf_mlp.py:92-105 with skip_layer=None (as built at coocc_ray.py:112-113)."""
    for i in range(depth):
        x = F.relu(F.linear(x, p[prefix + "hidden_layers.%d.weight" % i],
                            p[prefix + "hidden_layers.%d.bias" % i]))
    return F.linear(x, p[prefix + "output_layer.weight"], p[prefix + "output_layer.bias"])


RENDER_NX = (100, 100, 8)          # coocc_ray.py:372 (hard-coded)
RENDER_ORIGIN = (-50.0, -50.0, -5.0)


def render_voxel_indices(geom):
    """coocc_ray.py:372-384: geom [D,H,W,3] ego metres -> (idx [H,W,D,3] int64, inside [H,W,D])."""
    dx = torch.tensor([1.0, 1.0, 1.0])
    bx = torch.tensor([-50.0 + 0.5, -50.0 + 0.5, -5.0 + 0.5])
    nx = torch.tensor([100.0, 100.0, 8.0])
    g = (geom - (bx - dx / 2.0)) / dx                                          # :377
    inside = ((g[..., 0] >= 0) & (g[..., 0] < nx[0]) & (g[..., 1] >= 0) & (g[..., 1] < nx[1])
              & (g[..., 2] >= 0) & (g[..., 2] < nx[2]))                        # :378-380
    g = torch.where(inside[..., None], g, torch.zeros_like(g))                 # :381
    return g.long().permute(1, 2, 0, 3), inside.permute(1, 2, 0)               # :384,387


def render_forward(p, voxel_feats, gemo, gt_depth, gt_img, per_voxel_heads=False):
    """coocc_ray.py:358-433.  voxel_feats [1,C,X,Y,Z]; gemo [1,N,D,H,W,3];
    gt_depth [1,N,16H,16W]; gt_img [1,N,3,16H,16W].  p holds `sigma_head.*` / `rgb_head.*`.
    Returns dict(rgbs, depths, rgb_map, depth_map, loss_depth_render, loss_rgb)."""
    B, N, D, H, W, _ = gemo.shape
    assert B == 1                                                              # :365
    vf = voxel_feats[0]
    rgbs, depths, lo_rgb, lo_depth = [], [], [], []
    for i in range(N):
        pts, mask = render_voxel_indices(gemo[0, i])
        feat = vf[:, pts[..., 0], pts[..., 1], pts[..., 2]].permute(1, 2, 3, 0)  # :385-386
        rgb = mlp_forward(p, "rgb_head.", feat, 3)                             # :389
        rgb = torch.sigmoid(rgb * mask[..., None])                             # :390-391
        sigma = F.relu(mlp_forward(p, "sigma_head.", feat, 1).squeeze(-1))     # :392-393
        ptsf = pts.float()
        dists = torch.norm(ptsf[:, :, 1:] - ptsf[:, :, :-1], dim=-1)           # :396
        dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)  # :397-400
        alpha = 1.0 - torch.exp(-F.relu(sigma * dists))                        # :401
        trans = torch.cumprod(torch.cat([torch.ones(H, W, 1), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
        weights = alpha * trans                                                # :402-406
        rgb_map = (weights[..., None] * rgb).sum(-2)                           # :407
        z_vals = torch.linspace(0, D, D).reshape(1, 1, D)                      # :409
        depth_map = (weights * z_vals).sum(-1)                                 # :410
        lo_rgb.append(rgb_map)
        lo_depth.append(depth_map)
        depths.append(F.interpolate(depth_map[None, None], scale_factor=16, mode="bilinear")[0, 0])  # :412-414
        rgbs.append(F.interpolate(rgb_map.permute(2, 0, 1)[None], scale_factor=16,
                                  mode="bilinear")[0].permute(1, 2, 0))        # :415-417
    rgbs = torch.stack(rgbs)
    depths = torch.stack(depths)
    dgt = (gt_depth[0] - (2.0 - 0.5 / 2.0)) / 0.5                              # :423-424
    dgt = dgt.clip(0, D)                                                       # :425
    fg = dgt > 0                                                               # :426
    loss_depth = F.mse_loss(depths[fg] / D, dgt[fg] / D)                       # :427-429
    loss_rgb = F.mse_loss(rgbs, gt_img[0].permute(0, 2, 3, 1))                 # :431
    return dict(rgbs=rgbs, depths=depths, rgb_map=torch.stack(lo_rgb), depth_map=torch.stack(lo_depth),
                loss_depth_render=loss_depth, loss_rgb=loss_rgb)


# ----------------------------------------------------------------------------------------
# Whole hot path (fuser -> encoder -> neck -> head, + render) used by the bench CPU legs
# ----------------------------------------------------------------------------------------
def hot_path_forward(params, inputs, knum, tie="canonical"):
    """Composition following COOCC_Ray.forward_train (coocc_ray.py:313-433) restricted to the
    hot-path modules: occ_fuser (:252-253) -> semantic_encoder (:328) -> semantic_neck (:329)
    -> pts_bbox_head coarse logits (:349) -> render block (:358-433)."""
    fused = bifuser_forward(params["occ_fuser"], inputs["img_voxel_feats"], inputs["pts_voxel_feats"],
                            knum, tie=tie)
    mid = resnet3d_forward(params["semantic_encoder"], fused)
    neck = fpn3d_forward(params["semantic_neck"], mid)
    feats, occ = occhead_coarse_forward(params["pts_bbox_head"], neck)
    out = dict(fused=fused, occ=occ, out_voxel_feats=feats)
    if "geom" in inputs:
        out.update(render_forward(params["render"], fused, inputs["geom"], inputs["gt_depth"],
                                  inputs["gt_img"]))
    return out
