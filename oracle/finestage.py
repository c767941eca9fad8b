"""oracle.finestage -- TEST INFRASTRUCTURE ONLY (parity checker; never imported by the product).

CPU restatement of the occupancy head's fine / cascade stage (second half; the product's
counterpart is csrc/fine_stage.cu + csrc/fine_select.cu behind modules.OccHead.forward_fine):
    OccHead.forward, fine branch      P/coocc/dense_heads/occ_head.py:182-237 (layers :58-82)
    coarse_to_fine_coordinates        P/utils/coordinate_transform.py:3-25
    project_points_on_img (nuScenes)  P/utils/coordinate_transform.py:29-70
    OccHead.loss_point                occ_head.py:295-312
Pinned against the unmodified reference `OccHead` run in the build container
(oracle/make_golden.py -> tests/golden/reference_fine.npz, tests/test_oracle_fine.py).

Parameters: the reference head's state_dict keys (`img_mlp_0.{0,1}.*`, `img_mlp.{0,1}.*`, `fine_mlp.{0,1,3}.*`).
Randomness: when more than `fine_topk` coarse voxels are occupied the reference keeps a random subset drawn
with torch.randperm on the default CPU generator (:20); the restatement draws from the same generator, so a
caller that seeds torch identically gets the same subset.
"""
import torch
import torch.nn.functional as F

from . import losses as OL


def coarse_to_fine_coordinates(coarse_cor, ratio, topk=30000):
    """coordinate_transform.py:3-25.  coarse_cor [3,N] int64 -> [3, ratio^3 * min(N, topk)]: every coarse voxel
    expands to its ratio^3 children, ordered child-major (all voxels for child offset 0, then offset 1, ...)."""
    n = coarse_cor.shape[1]
    r = torch.arange(ratio, device=coarse_cor.device)
    off = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), dim=3).reshape(-1, 3)       # [r^3, 3], x-major
    fine = coarse_cor[None] * ratio + off[:, :, None]                                       # [r^3, 3, N]
    if n >= topk:                                                                           # :19-21
        fine = fine[:, :, torch.randperm(n)[:topk]]
    return fine.permute(1, 0, 2).reshape(3, -1)


def project_points_on_img(points, rots, trans, intrins, post_rots, post_trans, bda_mat, pts_range, W_img, H_img,
                          W_occ, H_occ, D_occ):
    """coordinate_transform.py:29-70 (data_type == 'nus').  points [1,N,3] fine-voxel indices ->
    (uv [n_cam, N, 1, 2] in grid_sample's [-1,1] convention, mask [1?,N,n_cam] of points in front of / inside a camera)."""
    voxel_size = (pts_range[3:] - pts_range[:3]) / torch.tensor([W_occ - 1, H_occ - 1, D_occ - 1])
    p = points * voxel_size[None, None] + pts_range[:3][None, None]                          # :33-34 (metres)
    p = (bda_mat.inverse() @ p.unsqueeze(-1)).squeeze(-1)                                   # :38-39 undo BEV aug
    p = p.view(-1, 1, 3) - trans.view(1, -1, 3)                                             # :46-47
    p = rots.inverse().unsqueeze(0) @ p.unsqueeze(-1)                                       # :48-49 ego -> camera
    p = (intrins.unsqueeze(0) @ p).squeeze(-1)                                              # :53 camera -> raw pixel
    d = p[..., 2:3]
    uv = p[..., :2] / (d + 1e-5)                                                            # :58-59
    uv = (post_rots[..., :2, :2].unsqueeze(0) @ uv.unsqueeze(-1)).squeeze(-1) + post_trans[..., :2].unsqueeze(0)
    uv = torch.stack([(uv[..., 0] / (W_img - 1) - 0.5) * 2, (uv[..., 1] / (H_img - 1) - 0.5) * 2], -1)   # :65-66
    mask = (d[..., 0] > 1e-5) & (uv[..., 0] > -1) & (uv[..., 0] < 1) & (uv[..., 1] > -1) & (uv[..., 1] < 1)
    return uv.permute(2, 1, 0, 3), mask


def _gn_linear(x, p, pre_lin, pre_gn, groups=16):
    """Linear -> GroupNorm(16) -> ReLU on rows [N, C] (occ_head.py:66-70, 73-77)."""
    y = F.linear(x, p[pre_lin + ".weight"], p[pre_lin + ".bias"])
    return F.relu(F.group_norm(y, groups, p[pre_gn + ".weight"], p[pre_gn + ".bias"], 1e-5))


def fine_forward(p, out_voxel_feats, coarse_occ, img_feats, transform, final_occ_size, point_cloud_range,
                 cascade_ratio=2, fine_topk=15000, empty_idx=0, training=True):
    """occ_head.py:182-237 for B = 1.  out_voxel_feats [1,128,W,H,D], coarse_occ [1,17,W,H,D],
    img_feats [1,n_cam,512,Hf,Wf], transform = img_inputs[1:] (rots, trans, intrins, post_rots, post_trans,
    bda, ..., img_size at [-1]).  Returns (fine_coord [3,M] int64, fine_output [M,17])."""
    mask = coarse_occ.argmax(1) != empty_idx                                                # :183
    assert mask.sum() > 0, "no foreground in coarse voxel"
    _, W, H, D = mask.shape
    cx, cy, cz = torch.meshgrid(torch.arange(W), torch.arange(H), torch.arange(D), indexing="ij")
    Bi, Ni, Ci, Wi, Hi = img_feats.shape                                                    # :193-197
    f2d = F.conv2d(img_feats.reshape(-1, Ci, Wi, Hi), p["img_mlp_0.0.weight"], p["img_mlp_0.0.bias"])
    f2d = F.relu(F.group_norm(f2d, 16, p["img_mlp_0.1.weight"], p["img_mlp_0.1.bias"], 1e-5)).reshape(Bi, Ni, -1, Wi, Hi)
    coarse = torch.stack([cx[mask[0]], cy[mask[0]], cz[mask[0]]], 0)                         # :201-203
    fine = coarse_to_fine_coordinates(coarse, cascade_ratio, fine_topk if training else 30000)
    new_coord = fine[None].permute(0, 2, 1).float().contiguous()                            # [1,M,3]
    g = fine.float()                                                                        # :212-217, normalised to [-1,1]
    g = torch.stack([(g[i] / (final_occ_size[i] - 1) - 0.5) * 2 for i in range(3)], 0)
    grid = g[None, None, None].permute(0, 4, 1, 2, 3)                                       # [1,M,1,1,3]
    vox = F.grid_sample(out_voxel_feats.permute(0, 1, 4, 3, 2), grid, mode="bilinear", padding_mode="zeros",
                        align_corners=False)                                                # :219
    feats = [vox[0, :, :, 0, 0].permute(1, 0)]
    uv, m = project_points_on_img(new_coord, transform[0][0:1], transform[1][0:1], transform[2][0:1],
                                  transform[3][0:1], transform[4][0:1], transform[5][0:1], point_cloud_range,
                                  transform[-1][1][0:1], transform[-1][0][0:1], W * cascade_ratio, H * cascade_ratio,
                                  D * cascade_ratio)                                        # :226-230
    s = F.grid_sample(f2d[0].contiguous(), uv.contiguous(), align_corners=True, mode="bilinear", padding_mode="zeros")
    s = s * m.permute(2, 1, 0)[:, None]                                                     # :233
    feats.append(_gn_linear(s.sum(0)[:, :, 0].permute(1, 0), p, "img_mlp.0", "img_mlp.1"))  # :234
    x = _gn_linear(torch.cat(feats, 1), p, "fine_mlp.0", "fine_mlp.1")                      # :237 (layers :73-78)
    return fine, F.linear(x, p["fine_mlp.3.weight"], p["fine_mlp.3.bias"])


def loss_point(fine_coord, fine_output, target_voxels, tag="fine", weights=(1.0, 1.0, 1.0, 1.0), empty_idx=0):
    """occ_head.py:295-312: the four voxel losses on the sampled points (CE without class weights here)."""
    gt = target_voxels[:, fine_coord[0], fine_coord[1], fine_coord[2]].long()[0]
    pred = fine_output.t()[None, :, :, None, None]                       # [1,17,M,1,1] for the [B,C,...] loss helpers
    tv = gt[None, :, None, None]
    return {
        "loss_voxel_ce_%s" % tag: weights[0] * F.cross_entropy(fine_output, gt, ignore_index=255),
        "loss_voxel_sem_scal_%s" % tag: weights[1] * OL.sem_scal_loss(pred, tv, 255),
        "loss_voxel_geo_scal_%s" % tag: weights[2] * OL.geo_scal_loss(pred, tv, 255, empty_idx),
        "loss_voxel_lovasz_%s" % tag: weights[3] * OL.lovasz_softmax(pred, tv, 255),
    }
