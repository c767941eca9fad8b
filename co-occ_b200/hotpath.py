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

    # ------------------------------------------------------------------------------------
    def evaluation_semantic(self, pred, gt, eval_type, visible_mask=None):
        """COOCC_Ray.evaluation_semantic (coocc_ray.py:659-684): confusion matrices of the up-sampled
        argmax prediction against gt_occ; returns numpy arrays like the reference's fast_hist."""
        from . import functional as CF
        x2d, dims = CF.to_cl2d(pred)
        empty = self.pts_bbox_head.empty_idx
        h_ssc, h_vis, h_sc = CF.eval_confusion(x2d, dims, gt, visible_mask if eval_type == 'SSC' else None, empty, 255)
        if eval_type == 'SC':
            return h_sc.cpu().numpy(), None
        if eval_type == 'SSC':
            return h_ssc.cpu().numpy(), (h_vis.cpu().numpy() if h_vis is not None else None)
        raise ValueError(eval_type)

    @torch.no_grad()
    def simple_test(self, img_voxel_feats, pts_voxel_feats, gt_occ=None, visible_mask=None):
        """The hot-path slice of COOCC_Ray.simple_test (coocc_ray.py:520-575): fuser -> encoder -> neck ->
        head (eval-mode BatchNorm) and the SC / SSC confusion matrices of the coarse prediction."""
        voxel_feats = self.occ_fuser(img_voxel_feats, pts_voxel_feats)
        semantic_voxel = self.semantic_neck(self.semantic_encoder(voxel_feats))
        output = self.pts_bbox_head(voxel_feats=semantic_voxel)
        pred_c = output['output_voxels'][0]
        out = {'pred_c': pred_c, 'pred_f': None, 'output_voxels': pred_c, 'target_voxels': gt_occ}
        if gt_occ is not None:
            out['SC_metric'], _ = self.evaluation_semantic(pred_c, gt_occ, 'SC', visible_mask)
            out['SSC_metric'], out['SSC_occ_metric'] = self.evaluation_semantic(pred_c, gt_occ, 'SSC', visible_mask)
        return out
