"""The hot-path slice of COOCC_Ray.forward_train (P/coocc/detectors/coocc_ray.py:313-433), built from the config's
own dicts through the registry exactly as the detector builds its sub-modules (coocc_ray.py:80-83, 111-113):
occ_fuser -> semantic_encoder -> semantic_neck -> pts_bbox_head (coarse logits + fine / cascade stage + losses) and
the render block on the fuser output.

`HotPath` is `detector.COOCC_Ray` on the `UpstreamFeatures` base with a tensor-level calling convention: everything
upstream of `img_voxel_feats` / `pts_voxel_feats` / `img_feats` / `geom` (2D backbone, LSS view transform, sparse
LiDAR encoder, dataloader) stays the reference's code and is not part of this package; this class is what tests,
smoke() and bench.py drive.
"""
import torch

from .detector import CooccRayHotPath, UpstreamFeatures
from .modules import render_fn


def model_cfg(C=128, K=2, num_cls=17, fine=None, grid=None, fine_topk=15000):
    """The `model=dict(...)` entries of projects/configs/coocc_nusc/coocc_multi_r50_256x704.py (:136-178) that select
    hot-path modules, with numC_Trans = C.  `fine` = the config's cascade head (cascade_ratio=2, sample_from_voxel /
    sample_from_img, fine_topk=15000, :71-73,160-164); its layers are hard-wired to 128 / 512 input channels in the
    reference (occ_head.py:66-78), so it defaults to on only for C = 128.  `grid` = the working grid (final_occ_size
    = 2 x grid = the label grid); None keeps the config's [200, 200, 16]."""
    planes = [C, 2 * C, 4 * C, 8 * C]
    nc = dict(type='SyncBN', requires_grad=True)
    if fine is None:
        fine = C == 128
    occ_size = [2 * g for g in grid] if grid is not None else [200, 200, 16]
    head = dict(type='OccHead', norm_cfg=nc, soft_weights=True, cascade_ratio=2 if fine else 1,
                sample_from_voxel=bool(fine), sample_from_img=bool(fine), final_occ_size=occ_size,
                fine_topk=fine_topk, empty_idx=0, num_level=4, in_channels=[2 * C] * 4, out_channel=num_cls,
                point_cloud_range=[-50, -50, -5.0, 50, 50, 3.0],
                loss_weight_cfg=dict(loss_voxel_ce_weight=1.0, loss_voxel_sem_scal_weight=1.0,
                                     loss_voxel_geo_scal_weight=1.0, loss_voxel_lovasz_weight=1.0))
    return dict(
        occ_fuser=dict(type='BiFuser_N', knum=K, in_channels=C, out_channels=C),
        semantic_encoder=dict(type='CustomResNet3D', depth=18, n_input_channels=C, block_inplanes=planes,
                              out_indices=(0, 1, 2, 3), norm_cfg=nc),
        semantic_neck=dict(type='FPN3D', with_cp=True, in_channels=planes, out_channels=2 * C, norm_cfg=nc),
        pts_bbox_head=head,
    )


class HotPath(CooccRayHotPath, UpstreamFeatures):
    def __init__(self, cfg, C, use_rendering=True, loss_norm=False):
        # loss_norm=True in the configs (Q7): every head loss is divided by its own detached value, which makes the
        # loss values themselves uninformative (== 1); tests and the bench keep the raw losses unless asked
        super().__init__(occ_fuser=cfg["occ_fuser"], semantic_encoder=cfg["semantic_encoder"],
                         semantic_neck=cfg["semantic_neck"], pts_bbox_head=cfg["pts_bbox_head"], loss_norm=loss_norm,
                         use_rendering=use_rendering, render_input_dim=C)

    def load_params(self, params):
        self.occ_fuser.load_state_dict(params["occ_fuser"])
        self.semantic_encoder.load_state_dict(params["semantic_encoder"])
        self.semantic_neck.load_state_dict(params["semantic_neck"])
        # the fine-stage layers (img_mlp_0 / img_mlp / fine_mlp) keep their initialisation unless given
        self.pts_bbox_head.load_state_dict(params["pts_bbox_head"], strict=not self.pts_bbox_head.fine_stage)
        if self.use_rendering:
            r = params["render"]
            self.sigma_head.load_state_dict({k[11:]: v for k, v in r.items() if k.startswith("sigma_head.")})
            self.rgb_head.load_state_dict({k[9:]: v for k, v in r.items() if k.startswith("rgb_head.")})

    def forward_train(self, img_voxel_feats, pts_voxel_feats, geom=None, gt_depth=None, gt_img=None, gt_occ=None,
                      img_feats=None, transform=None):
        """Tensor-level form of COOCC_Ray.forward_train (same statements as CooccRayHotPath.forward_train, which
        tests/test_gpu_detector.py checks it against): the upstream tensors are given instead of being produced by
        extract_feat; gt_img / gt_depth = img_inputs[0] / [7], transform = img_inputs[1:]
        (P/datasets/pipelines/loading.py:129).  Returns (losses, coarse logits, fused grid)."""
        head = self.pts_bbox_head
        if head.fine_stage and head.sample_from_img and (img_feats is None or transform is None):
            raise ValueError("this head samples image features in its fine stage (sample_from_img=True): pass "
                             "img_feats [1,N,512,fH,fW] and transform = img_inputs[1:]")
        if img_feats is not None and not isinstance(img_feats, (list, tuple)):
            img_feats = [img_feats]
        voxel_feats = self.occ_fuser(img_voxel_feats, pts_voxel_feats)          # coocc_ray.py:252-253
        mid_voxel = self.semantic_encoder(voxel_feats)                          # :328
        semantic_voxel = self.semantic_neck(mid_voxel)                          # :329
        outs = self.pts_bbox_head(voxel_feats=semantic_voxel, img_feats=img_feats, transform=transform)   # :349 -> :282-290
        losses = {}
        if gt_occ is not None:                                                  # :291-299 -> OccHead.loss
            losses.update(head.loss(output_voxels=outs["output_voxels"], output_voxels_fine=outs["output_voxels_fine"],
                                    output_coords_fine=outs["output_coords_fine"], target_voxels=gt_occ))
        if self.loss_norm:                                                      # :353-356
            for k in list(losses.keys()):
                losses[k] = losses[k] / (losses[k].detach() + 1e-9)
        if self.use_rendering and geom is not None:                             # :358-433
            _, _, rl = render_fn(voxel_feats, geom, self.sigma_head, self.rgb_head, gt_depth, gt_img)
            losses.update(rl)
        self._last_outs = outs
        return losses, outs["output_voxels"][0], voxel_feats

    @torch.no_grad()
    def simple_test(self, img_voxel_feats, pts_voxel_feats, gt_occ=None, visible_mask=None, img_feats=None,
                    transform=None, geom=None, gt_img=None):
        """The hot-path slice of COOCC_Ray.simple_test (coocc_ray.py:520-656)."""
        if img_feats is not None and not isinstance(img_feats, (list, tuple)):
            img_feats = [img_feats]
        self.upstream = dict(img_voxel_feats=img_voxel_feats, pts_voxel_feats=pts_voxel_feats, img_feats=img_feats,
                             depth=None, geom=geom)
        try:
            img = ((gt_img,) + tuple(transform)) if transform is not None else None
            return CooccRayHotPath.simple_test(self, img=img, gt_occ=gt_occ, visible_mask=visible_mask)
        finally:
            self.upstream = None
