"""The hot-path slice of COOCC_Ray.forward_train (P/coocc/detectors/coocc_ray.py:313-433), built
from config-style dicts through the registry exactly as the detector builds its sub-modules
(coocc_ray.py:80-83, 111-113): occ_fuser -> semantic_encoder -> semantic_neck -> pts_bbox_head
(coarse logits) and the render block on the fuser output.

Everything upstream of `img_voxel_feats` / `pts_voxel_feats` / `geom` (2D backbone, LSS view
transform, sparse LiDAR encoder, dataloader) stays the reference's code and is not part of this
package; this class is what tests, smoke() and bench.py drive.
"""
import torch
import torch.nn as nn

from . import registry
from .modules import MLP, render_fn


def model_cfg(C=128, K=2, num_cls=17):
    """The `model=dict(...)` entries of projects/configs/coocc_nusc/coocc_multi_r50_256x704.py
    (:136-190) that select hot-path modules, with numC_Trans = C."""
    planes = [C, 2 * C, 4 * C, 8 * C]
    nc = dict(type='SyncBN', requires_grad=True)
    return dict(
        occ_fuser=dict(type='BiFuser_N', knum=K, in_channels=C, out_channels=C),
        semantic_encoder=dict(type='CustomResNet3D', depth=18, n_input_channels=C, block_inplanes=planes,
                              out_indices=(0, 1, 2, 3), norm_cfg=nc),
        semantic_neck=dict(type='FPN3D', with_cp=True, in_channels=planes, out_channels=2 * C, norm_cfg=nc),
        pts_bbox_head=dict(type='OccHead', norm_cfg=nc, soft_weights=True, cascade_ratio=1,
                           sample_from_voxel=False, sample_from_img=False, num_level=4,
                           in_channels=[2 * C] * 4, out_channel=num_cls),
    )


class HotPath(nn.Module):
    def __init__(self, cfg, C, use_rendering=True):
        super().__init__()
        self.occ_fuser = registry.build_fusion_layer(cfg["occ_fuser"])
        self.semantic_encoder = registry.build_backbone(cfg["semantic_encoder"])
        self.semantic_neck = registry.build_neck(cfg["semantic_neck"])
        self.pts_bbox_head = registry.build_head(cfg["pts_bbox_head"])
        self.use_rendering = use_rendering
        if use_rendering:     # coocc_ray.py:111-113 (input_dim hard-wired to 128 there)
            self.sigma_head = MLP(input_dim=C, output_dim=1, net_depth=1, skip_layer=None)
            self.rgb_head = MLP(input_dim=C, output_dim=3, net_depth=3, skip_layer=None)

    def load_params(self, params):
        self.occ_fuser.load_state_dict(params["occ_fuser"])
        self.semantic_encoder.load_state_dict(params["semantic_encoder"])
        self.semantic_neck.load_state_dict(params["semantic_neck"])
        self.pts_bbox_head.load_state_dict(params["pts_bbox_head"])
        if self.use_rendering:
            r = params["render"]
            self.sigma_head.load_state_dict({k[11:]: v for k, v in r.items() if k.startswith("sigma_head.")})
            self.rgb_head.load_state_dict({k[9:]: v for k, v in r.items() if k.startswith("rgb_head.")})

    def forward_train(self, img_voxel_feats, pts_voxel_feats, geom=None, gt_depth=None, gt_img=None,
                      gt_occ=None):
        voxel_feats = self.occ_fuser(img_voxel_feats, pts_voxel_feats)          # coocc_ray.py:252-253
        mid_voxel = self.semantic_encoder(voxel_feats)                          # :328
        semantic_voxel = self.semantic_neck(mid_voxel)                          # :329
        outs = self.pts_bbox_head(voxel_feats=semantic_voxel)                   # :349 -> :282-290
        occ = outs["output_voxels"][0]
        losses = {}
        if gt_occ is not None:
            # coocc_ray.py:349-351 -> OccHead.loss: CE + sem_scal + geo_scal + Lovasz on the coarse logits
            # against gt_occ (label grid = integer multiple of the working grid, 255 = ignore)
            losses.update(self.pts_bbox_head.loss(output_voxels=outs["output_voxels"], target_voxels=gt_occ))
        if self.use_rendering and geom is not None:                             # :358-433
            _, _, rl = render_fn(voxel_feats, geom, self.sigma_head, self.rgb_head, gt_depth, gt_img)
            losses.update(rl)
        return losses, occ, voxel_feats
