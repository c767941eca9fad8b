"""COOCC_Ray / COOCC_Ray_L: the reference detectors with the hot path running on the C ABI.

The reference has no module boundary around its volume renderer: the block is inline in
`COOCC_Ray.forward_train` (P/coocc/detectors/coocc_ray.py:358-433), copied into `simple_test` (:562-637) and, minus
the colour head, into the LiDAR-only detector (P/coocc/detectors/coocc_ray_lidar.py).  The drop-in therefore is a
detector of the same registry name whose `forward_train` / `simple_test` are the reference's, statement for
statement, except that

  * `occ_fuser`, `semantic_encoder`, `semantic_neck`, `pts_bbox_head` resolve (through the registry, as at
    coocc_ray.py:80-83) to this package's modules, and
  * the render block is one call of `modules.render_fn` (or `render_depth_fn` for the LiDAR-only branch, :435-494).

`CooccRayHotPath` holds those methods.  Where OpenMMLab and the reference plugin are importable the registered classes
are `class COOCC_Ray(CooccRayHotPath, <reference COOCC_Ray>)` -- everything upstream (`extract_img_feat`,
`extract_pts_feat`, `__init__`, `forward_test`, ...) stays the reference's own code; where they are not (this
repository's tests, bench.py) the base is `UpstreamFeatures`, which takes the upstream tensors as they are produced by
the kept code (`img_voxel_feats`, `pts_voxel_feats`, `img_feats`, `depth`, `geom`) from `img_inputs`' companion dict.
"""
import torch
import torch.nn as nn

from . import functional as CF
from . import registry
from .modules import MLP, render_depth_fn, render_fn


class CooccRayHotPath:
    """forward_pts_train / forward_train / simple_test / evaluation_semantic of COOCC_Ray on the C ABI."""

    lidar_only = False          # COOCC_Ray_L: no rgb_head, geometry from the calibration (coocc_ray.py:435-494)

    # -------------------------------------------------------------------------------- training
    def forward_pts_train(self, voxel_feats, gt_occ=None, points_occ=None, img_metas=None, transform=None,
                          img_feats=None, pts_feats=None, visible_mask=None):
        """coocc_ray.py:265-311."""
        outs = self.pts_bbox_head(voxel_feats=voxel_feats, points=points_occ, img_metas=img_metas, img_feats=img_feats,
                                  pts_feats=pts_feats, target_points=points_occ, transform=transform)
        self._last_outs = outs
        return self.pts_bbox_head.loss(output_voxels=outs['output_voxels'],
                                       output_voxels_fine=outs['output_voxels_fine'],
                                       output_coords_fine=outs['output_coords_fine'], target_voxels=gt_occ,
                                       target_points=points_occ, img_metas=img_metas, visible_mask=visible_mask)

    def forward_train(self, points=None, img_metas=None, img_inputs=None, gt_occ=None, points_occ=None,
                      visible_mask=None, gt_depths=None, **kwargs):
        """coocc_ray.py:313-509."""
        voxel_feats, img_feats, pts_feats, depth, gemo, img_voxel_feats = self.extract_feat(
            points, img=img_inputs, img_metas=img_metas)
        mid_voxel = self.semantic_encoder(voxel_feats)                                        # :328
        semantic_voxel = self.semantic_neck(mid_voxel)                                        # :329
        losses = dict()
        if not getattr(self, "disable_loss_depth", False) and depth is not None \
                and getattr(self, "img_view_transformer", None) is not None:                  # :339-340 (kept code)
            losses['loss_depth'] = self.img_view_transformer.get_depth_loss(img_inputs[7], depth)
        transform = img_inputs[1:] if img_inputs is not None else None                        # :348
        losses.update(self.forward_pts_train(semantic_voxel, gt_occ, points_occ, img_metas, img_feats=img_feats,
                                             pts_feats=pts_feats, transform=transform, visible_mask=visible_mask))
        if self.loss_norm:                                                                    # :353-356 (Q7)
            for k in list(losses.keys()):
                if k.startswith('loss'):
                    losses[k] = losses[k] / (losses[k].detach() + 1e-9)
        if self.use_rendering:                                                                # :358
            if img_feats is not None and not self.lidar_only:
                _, _, rl = render_fn(voxel_feats, gemo, self.sigma_head, self.rgb_head, img_inputs[7], img_inputs[0])
            else:                                                                             # :435-494
                rl = self._render_depth_only(voxel_feats, gt_depths)
            losses.update(rl)
        self._last_voxel_feats = voxel_feats
        return losses

    def _render_depth_only(self, voxel_feats, gt_depths):
        """LiDAR-only branch: frustum geometry from the calibration in `gt_depths` (get_frustum, coocc_ray.py:732-776),
        density head only, loss_depth_render only.  gt_depths = (rots, trans, intrins, post_rots, post_trans, bda,
        depth_gt [B,N,H,W], ..., input_size): indices 0-5, 6 (-2 in COOCC_Ray_L, coocc_ray_lidar.py:507) and -1."""
        from . import lss
        rots, trans, intrins, post_rots, post_trans, bda = gt_depths[:6]
        size = gt_depths[-1]
        ogfH, ogfW = int(size[0]), int(size[1])
        frustum = lss.create_frustum((ogfH, ogfW), (2.0, 58.0, 0.5), 16).to(voxel_feats.device)
        gemo = lss.get_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda)
        depth_gt = gt_depths[-2] if self.lidar_only else gt_depths[6]
        _, rl = render_depth_fn(voxel_feats, gemo, self.sigma_head, depth_gt)
        return rl

    # -------------------------------------------------------------------------------- inference
    def evaluation_semantic(self, pred, gt, eval_type, visible_mask=None):
        """coocc_ray.py:659-684: confusion matrices of the up-sampled argmax prediction against gt_occ, returned
        as numpy arrays like the reference's fast_hist; one kernel, only the counters leave the GPU."""
        x2d, dims = CF.to_cl2d(pred)
        empty = self.pts_bbox_head.empty_idx
        h_ssc, h_vis, h_sc = CF.eval_confusion(x2d, dims, gt, visible_mask if eval_type == 'SSC' else None, empty, 255)
        if eval_type == 'SC':
            return h_sc.cpu().numpy(), None
        if eval_type == 'SSC':
            return h_ssc.cpu().numpy(), (h_vis.cpu().numpy() if h_vis is not None else None)
        raise ValueError(eval_type)

    @torch.no_grad()
    def simple_test(self, img_metas=None, img=None, gt_depths=None, points=None, rescale=False, points_occ=None,
                    gt_occ=None, visible_mask=None):
        """coocc_ray.py:520-656 (lidarseg points and the PNG dump of the test-time render are not on the path; the
        rendered maps are returned under `render_rgbs` / `render_depths` instead of being written with cv2)."""
        voxel_feats, img_feats, pts_feats, depth, gemo, img_voxel_feats = self.extract_feat(
            points, img=img, img_metas=img_metas)
        semantic_voxel = self.semantic_neck(self.semantic_encoder(voxel_feats))
        transform = img[1:] if img is not None else None
        output = self.pts_bbox_head(voxel_feats=semantic_voxel, points=points_occ, img_metas=img_metas,
                                    img_feats=img_feats, pts_feats=pts_feats, target_points=points_occ,
                                    transform=transform)
        pred_c = output['output_voxels'][0]
        out = {'pred_c': pred_c, 'pred_f': None, 'output_voxels': pred_c, 'target_voxels': gt_occ}
        if gt_occ is not None:
            out['SC_metric'], _ = self.evaluation_semantic(pred_c, gt_occ, 'SC', visible_mask)
            out['SSC_metric'], out['SSC_occ_metric'] = self.evaluation_semantic(pred_c, gt_occ, 'SSC', visible_mask)
        if output['output_voxels_fine'] is not None and gt_occ is not None:                   # :544-554
            fine_pred, fine_coord = output['output_voxels_fine'][0], output['output_coords_fine'][0]
            pred_f = self.empty_idx * torch.ones_like(gt_occ)[:, None].repeat(1, fine_pred.shape[1], 1, 1, 1).float()
            fc = fine_coord.long()
            nsel = getattr(fine_coord, "_coocc_nsel", None)
            if nsel is not None:                  # device-side selection: drop the padding slots
                keep = (torch.arange(fc.shape[1], device=fc.device) % fine_coord._coocc_topk) < nsel[1]
                fc, fine_pred = fc[:, keep], fine_pred[keep]
            pred_f[:, :, fc[0], fc[1], fc[2]] = fine_pred.permute(1, 0)[None]
            out['pred_f'] = pred_f
            out['SC_metric'], _ = self.evaluation_semantic(pred_f, gt_occ, 'SC', visible_mask)
            out['SSC_metric_fine'], _ = self.evaluation_semantic(pred_f, gt_occ, 'SSC', visible_mask)
        if self.use_rendering and getattr(self, "test_rendering", False) and gemo is not None:    # :562-637
            rgbs, depths, _ = render_fn(voxel_feats, gemo, self.sigma_head, self.rgb_head, None, None)
            out['render_rgbs'], out['render_depths'] = rgbs, depths
        return out


class UpstreamFeatures(nn.Module):
    """Base used where OpenMMLab is absent: builds the four hot-path modules from the config's `model` dict exactly
    like COOCC_Ray.__init__ (coocc_ray.py:32-113) and takes the tensors the kept upstream code produces
    (extract_img_feat / extract_pts_feat, :164-234) from `self.upstream`, set by the caller before each step:
    dict(img_voxel_feats, pts_voxel_feats, img_feats=None, depth=None, geom=None)."""

    def __init__(self, occ_fuser=None, semantic_encoder=None, semantic_neck=None, pts_bbox_head=None, loss_norm=False,
                 use_rendering=False, test_rendering=False, empty_idx=0, disable_loss_depth=False, scale=16,
                 render_input_dim=128, **upstream_cfg):
        super().__init__()
        self.upstream_cfg = upstream_cfg      # img_backbone, img_neck, img_view_transformer, pts_* ...: kept code
        self.loss_norm, self.use_rendering, self.test_rendering = loss_norm, use_rendering, test_rendering
        self.empty_idx, self.disable_loss_depth, self.scale = empty_idx, disable_loss_depth, scale
        self.occ_fuser = registry.build_fusion_layer(occ_fuser) if occ_fuser is not None else None      # :80
        self.semantic_encoder = registry.build_backbone(semantic_encoder)                               # :82
        self.semantic_neck = registry.build_neck(semantic_neck)                                         # :83
        self.pts_bbox_head = registry.build_head(pts_bbox_head)            # MVXTwoStageDetector builds this one
        if use_rendering:                      # :111-113 (input_dim hard-wired to 128 there)
            self.sigma_head = MLP(input_dim=render_input_dim, output_dim=1, net_depth=1, skip_layer=None)
            if not self.lidar_only:
                self.rgb_head = MLP(input_dim=render_input_dim, output_dim=3, net_depth=3, skip_layer=None)
        self.upstream = None

    def extract_feat(self, points, img, img_metas):
        """coocc_ray.py:237-263 with the upstream halves supplied."""
        u = self.upstream
        if u is None:
            raise RuntimeError("set `.upstream = dict(img_voxel_feats=..., pts_voxel_feats=..., img_feats=..., depth=..., "
                               "geom=...)` (the outputs of the kept extract_img_feat / extract_pts_feat) before the step")
        img_v, pts_v = u.get("img_voxel_feats"), u.get("pts_voxel_feats")
        if self.occ_fuser is not None:
            voxel_feats = self.occ_fuser(img_v, pts_v)                                                  # :252-253
        else:
            assert (img_v is None) or (pts_v is None)
            voxel_feats = img_v if pts_v is None else pts_v
        return voxel_feats, u.get("img_feats"), u.get("pts_feats"), u.get("depth"), u.get("geom"), img_v


def _make_detectors():
    """Register COOCC_Ray / COOCC_Ray_L.  With OpenMMLab + the reference plugin importable the reference classes are
    the bases (their __init__, extract_* and test plumbing are kept); otherwise UpstreamFeatures."""
    base = base_l = UpstreamFeatures
    if registry.HAVE_MMDET3D:  # pragma: no cover - needs the OpenMMLab stack
        try:
            from projects.mmdet3d_plugin.coocc.detectors.coocc_ray import COOCC_Ray as base
            from projects.mmdet3d_plugin.coocc.detectors.coocc_ray_lidar import COOCC_Ray_L as base_l
        except Exception:  # noqa: BLE001
            base = base_l = UpstreamFeatures

    @registry.DETECTORS.register_module(force=True)
    class COOCC_Ray(CooccRayHotPath, base):
        pass

    @registry.DETECTORS.register_module(force=True)
    class COOCC_Ray_L(CooccRayHotPath, base_l):
        lidar_only = True

    return COOCC_Ray, COOCC_Ray_L


COOCC_Ray, COOCC_Ray_L = _make_detectors()


def build_detector(cfg):
    return registry.DETECTORS.build(cfg)
