"""The reference's plugin modules of the hot path, re-implemented on the C ABI.

Same registry names, constructor arguments, forward signatures and state_dict keys as
  BiFuser_N       P/coocc/fuser/bifuser_n.py:13-174
  CustomResNet3D  P/coocc/backbones/resnet3d.py:105-205
  FPN3D           P/coocc/necks/fpn3d.py:13-108
  OccHead         P/coocc/dense_heads/occ_head.py:15-171 (coarse path; fine stage/losses are §8f "next")
  MLP             P/utils/nerf_mlp.py:14-105 (the two render heads)
torch.nn layers are used only as *parameter containers* (identical names, shapes and default
initialisers); their forward() is never called -- every convolution / linear runs in
csrc/conv_tc.cu, the fusion in csrc/gsf_*.cu, the renderer in csrc/render.cu.
Tensors between modules are the reference's [1,C,X,Y,Z], physically channels_last_3d.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as CF
from .registry import BACKBONES, FUSION_LAYERS, HEADS, NECKS


# voxel counts per nuScenes-occupancy class (P/utils/nusc_param.py:10-12), class 0 = free
NUSC_CLASS_FREQUENCIES = [2242961742295, 25985376, 1561108, 28862014, 196106643, 15920504, 2158753, 26539491,
                          4004729, 34838681, 75173306, 2255027978, 50959399, 646022466, 869055679, 1446141335,
                          1724391378]


def _norm_layer(norm_cfg, c):
    cfg = dict(norm_cfg or dict(type="BN3d"))
    typ = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if typ not in ("BN3d", "SyncBN", "BN"):
        raise NotImplementedError("coocc_b200 hot path implements BatchNorm (BN3d/SyncBN) only, got %s" % typ)
    return nn.BatchNorm3d(c, **cfg)


def _cl3d_(conv):
    """Keep conv weights physically [Cout,kx,ky,kz,Cin] so the kernels read them in place."""
    w = conv.weight
    if w.dim() == 5 and not w.permute(0, 2, 3, 4, 1).is_contiguous():
        w.data = w.data.contiguous(memory_format=torch.channels_last_3d)
    return w


def _cl3d_all(module):
    """Relayout every Conv3d weight of `module` to [Cout,kx,ky,kz,Cin] at construction, so that optimizers and
    bf16 shadows built before the first forward see the final memory order (load_state_dict copies in place and
    .to(device) preserves strides)."""
    for m in module.modules():
        if isinstance(m, nn.Conv3d):
            _cl3d_(m)


def batch_norm2d(x2d, bn):
    """BatchNorm3d over a [V,C] view (N*D*H*W = V rows), same running-stat updates."""
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        CF.count_batch(bn.num_batches_tracked)
    mom = 0.0 if bn.momentum is None else bn.momentum
    use_batch = bn.training or bn.running_mean is None
    return F.batch_norm(x2d, bn.running_mean if not bn.training or bn.track_running_stats else None,
                        bn.running_var if not bn.training or bn.track_running_stats else None,
                        bn.weight, bn.bias, use_batch, mom, bn.eps)


def conv_bn_act(x2d, dims, conv, bn=None, relu=True, residual=None, skip=False):
    """conv -> BatchNorm (batch statistics from the conv epilogue) -> (+ residual) -> ReLU.
    skip=True: returns (y, odims, x_alias); hand x_alias to the OTHER consumer of x (the residual branch) and its gradient
    is added inside this convolution's data-gradient kernel (CF.conv3d) instead of in a separate pass."""
    k = conv.kernel_size[0]
    s = conv.stride[0]
    odims = tuple(CF.out_dim(n, k, s) for n in dims)
    fused = (bn is not None and bn.training and conv.bias is None and conv.out_channels % 4 == 0
             and bn.momentum is not None)
    if skip:
        if not (fused and torch.is_grad_enabled()):
            return conv_bn_act(x2d, dims, conv, bn, relu, residual) + (x2d,)
        y, stats, xs = CF.conv3d(x2d, _cl3d_(conv), dims, k, s, want_stats=True, out_bf16=CF.act_bf16(), skip=True)
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            CF.count_batch(bn.num_batches_tracked)
        track = bn.track_running_stats
        y = CF.bn_act(y, stats, bn.weight, bn.bias, residual, relu, bn.eps, bn.momentum,
                      bn.running_mean if track else None, bn.running_var if track else None)
        return y, odims, xs
    if fused:
        y, stats = CF.conv3d(x2d, _cl3d_(conv), dims, k, s, want_stats=True, out_bf16=CF.act_bf16())
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            CF.count_batch(bn.num_batches_tracked)
        track = bn.track_running_stats
        y = CF.bn_act(y, stats, bn.weight, bn.bias, residual, relu, bn.eps, bn.momentum,
                      bn.running_mean if track else None, bn.running_var if track else None)
        return y, odims
    y = CF.conv3d(x2d, _cl3d_(conv), dims, k, s, bias=conv.bias)
    if (bn is not None and not bn.training and bn.running_mean is not None and not torch.is_grad_enabled()
            and conv.out_channels % 4 == 0):
        # inference: running statistics, same fused kernel
        return CF.bn_act_eval(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.eps, residual, relu), odims
    if bn is not None:
        y = batch_norm2d(y, bn)
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    return y, odims


def resize_trilinear(x2d, dims, size):
    """F.interpolate(..., mode='trilinear', align_corners=False) on a [V,C] view."""
    if tuple(size) == tuple(dims):
        return x2d
    x5 = CF.to_5d(x2d, dims)
    y5 = F.interpolate(x5, size=list(size), mode="trilinear", align_corners=False)
    return CF.to_cl2d(y5)[0]


# ----------------------------------------------------------------------------------------
@FUSION_LAYERS.register_module(force=True)
class BiFuser_N(nn.Module):
    def __init__(self, in_channels, out_channels, knum=1, norm_cfg=None, fix_k1_fps=False):
        super().__init__()
        self.in_channels, self.out_channels, self.knum = in_channels, out_channels, knum
        self.fix_k1_fps = fix_k1_fps
        # container layers mirror bifuser_n.py:23-36 (plain BatchNorm3d defaults)
        self.con_enc = nn.Sequential(
            nn.Conv3d(in_channels * 4, out_channels * 2, 3, padding=1, bias=False),
            nn.BatchNorm3d(out_channels * 2), nn.ReLU(True),
            nn.Conv3d(in_channels * 2, out_channels, 3, padding=1, bias=False),
            nn.BatchNorm3d(out_channels), nn.ReLU(True))
        self.knn_enc = nn.Sequential(nn.Linear(in_channels * knum, out_channels), nn.ReLU())
        _cl3d_all(self)

    def forward(self, img_voxel_feats, pts_voxel_feats):
        dims = tuple(img_voxel_feats.shape[2:])
        cat = CF.gsfusion_concat(img_voxel_feats, pts_voxel_feats, self.knn_enc[0].weight,
                                 self.knn_enc[0].bias, self.knum, self.fix_k1_fps)
        CF.PRE_TAIL_MARK = True      # (this convolution's data gradient is the last convolution of the backward)
        y, _ = conv_bn_act(cat, dims, self.con_enc[0], self.con_enc[1])
        CF.PRE_TAIL_MARK = False
        y, _ = conv_bn_act(y, dims, self.con_enc[3], self.con_enc[4])
        if CF.PRE_TAIL_HOOK is not None and y.requires_grad:
            hook = CF.PRE_TAIL_HOOK              # (bound now: the global is cleared when the capture ends)
            y.register_hook(lambda g: hook("fuser"))          # runs when the fuser's backward starts
        return CF.to_5d(y, dims)


# ----------------------------------------------------------------------------------------
class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1, downsample=None, norm_cfg=None):
        super().__init__()
        self.conv1 = nn.Conv3d(in_planes, planes, 3, stride, 1, bias=False)
        self.bn1 = _norm_layer(norm_cfg, planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv3d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = _norm_layer(norm_cfg, planes)
        self.downsample = downsample
        self.stride = stride

    def forward2d(self, x, dims):
        if self.downsample is not None:
            out, odims = conv_bn_act(x, dims, self.conv1, self.bn1)
            idn, _ = conv_bn_act(x, dims, self.downsample[0], self.downsample[1], relu=False)
        else:
            # identity shortcut: the residual gradient is added in conv1's data-gradient epilogue
            out, odims, idn = conv_bn_act(x, dims, self.conv1, self.bn1, skip=True)
        out, _ = conv_bn_act(out, odims, self.conv2, self.bn2, relu=True, residual=idn)   # resnet3d.py:53-60
        return out, odims


@BACKBONES.register_module(force=True)
class CustomResNet3D(nn.Module):
    def __init__(self, depth, block_inplanes=[64, 128, 256, 512], block_strides=[1, 2, 2, 2],
                 out_indices=(0, 1, 2, 3), n_input_channels=3, shortcut_type='B',
                 norm_cfg=dict(type='BN3d', requires_grad=True), widen_factor=1.0):
        super().__init__()
        metas = {10: [1, 1, 1, 1], 18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}
        if depth not in metas:
            raise NotImplementedError("hot path covers the BasicBlock depths (10/18/34); configs use 18")
        if shortcut_type != 'B':
            raise NotImplementedError("shortcut_type 'B' (conv1x1x1 + norm) only")
        planes = [int(x * widen_factor) for x in block_inplanes]
        self.in_planes = planes[0]
        self.out_indices = out_indices
        self.input_proj = nn.Sequential(nn.Conv3d(n_input_channels, self.in_planes, 1, 1, bias=False),
                                        _norm_layer(norm_cfg, self.in_planes), nn.ReLU(inplace=True))
        self.layers = nn.ModuleList()
        for i in range(len(planes)):
            self.layers.append(self._make_layer(planes[i], metas[depth][i], block_strides[i], norm_cfg))
        for m in self.modules():            # resnet3d.py:150-158
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        _cl3d_all(self)

    def _make_layer(self, planes, blocks, stride, norm_cfg):
        downsample = None
        if stride != 1 or self.in_planes != planes:
            downsample = nn.Sequential(nn.Conv3d(self.in_planes, planes, 1, stride, bias=False),
                                       _norm_layer(norm_cfg, planes))
        layers = [BasicBlock(self.in_planes, planes, stride, downsample, norm_cfg)]
        self.in_planes = planes
        for _ in range(1, blocks):
            layers.append(BasicBlock(planes, planes, norm_cfg=norm_cfg))
        return nn.Sequential(*layers)

    def forward(self, x):
        x2d, dims = CF.to_cl2d(x)
        x2d, dims = conv_bn_act(x2d, dims, self.input_proj[0], self.input_proj[1])
        res = []
        for index, layer in enumerate(self.layers):
            for blk in layer:
                x2d, dims = blk.forward2d(x2d, dims)
            if index in self.out_indices:
                res.append(CF.to_5d(x2d, dims))
        return res


# ----------------------------------------------------------------------------------------
class _ConvModule(nn.Module):
    """Parameter container with mmcv ConvModule's child names (`conv`, `bn`)."""

    def __init__(self, cin, cout, k, norm_cfg):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, k, 1, k // 2, bias=False)
        self.bn = _norm_layer(norm_cfg, cout)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")


@NECKS.register_module(force=True)
class FPN3D(nn.Module):
    def __init__(self, in_channels=[80, 160, 320, 640], out_channels=256,
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), conv_cfg=dict(type='Conv3d'),
                 act_cfg=dict(type='ReLU'), with_cp=False, upsample_cfg=dict(mode='trilinear'), init_cfg=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.with_cp = with_cp      # activation checkpointing is a memory policy, not arithmetic
        self.upsample_cfg = upsample_cfg
        self.num_out = len(in_channels)
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.num_out):
            self.lateral_convs.append(nn.Sequential(_ConvModule(in_channels[i], out_channels, 1, norm_cfg)))
            self.fpn_convs.append(nn.Sequential(_ConvModule(out_channels, out_channels, 3, norm_cfg)))
        _cl3d_all(self)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        lat, dims = [], []
        for i, x in enumerate(inputs):
            x2d, d = CF.to_cl2d(x)
            m = self.lateral_convs[i][0]
            y, _ = conv_bn_act(x2d, d, m.conv, m.bn)
            lat.append(y)
            dims.append(d)
        for i in range(self.num_out - 1, 0, -1):       # fpn3d.py:91-94, resize fused with the add
            lat[i - 1] = CF.resize_mix([lat[i]], [dims[i]], dims[i - 1], base=lat[i - 1])
        outs = []
        for i in range(self.num_out):
            m = self.fpn_convs[i][0]
            y, _ = conv_bn_act(lat[i], dims[i], m.conv, m.bn)
            outs.append(CF.to_5d(y, dims[i]))
        return outs


# ----------------------------------------------------------------------------------------
@HEADS.register_module(force=True)
class OccHead(nn.Module):
    def __init__(self, in_channels, out_channel, num_level=1, num_img_level=1, soft_weights=False,
                 loss_weight_cfg=None, conv_cfg=dict(type='Conv3d', bias=False),
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), fine_topk=20000,
                 point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], final_occ_size=[256, 256, 20],
                 empty_idx=0, visible_loss=False, balance_cls_weight=True, cascade_ratio=1,
                 sample_from_voxel=False, sample_from_img=False, train_cfg=None, test_cfg=None,
                 padding_mode='border', data_type='nus'):
        super().__init__()
        if not isinstance(in_channels, list):
            in_channels = [in_channels]
        self.in_channels, self.out_channel, self.num_level = in_channels, out_channel, num_level
        self.soft_weights = soft_weights
        self.cascade_ratio = cascade_ratio
        self.sample_from_voxel, self.sample_from_img = sample_from_voxel, sample_from_img
        self.empty_idx = empty_idx
        self.final_occ_size = final_occ_size
        self.fine_topk = fine_topk
        # loss weights and class weights: occ_head.py:82-98, 134-145
        lw = loss_weight_cfg or {}
        self.loss_voxel_ce_weight = lw.get('loss_voxel_ce_weight', 1.0)
        self.loss_voxel_sem_scal_weight = lw.get('loss_voxel_sem_scal_weight', 1.0)
        self.loss_voxel_geo_scal_weight = lw.get('loss_voxel_geo_scal_weight', 1.0)
        self.loss_voxel_lovasz_weight = lw.get('loss_voxel_lovasz_weight', 1.0)
        if data_type != 'nus':
            raise NotImplementedError("hot path covers the nuScenes head (17 classes)")
        if balance_cls_weight:
            cw = 1.0 / torch.log(torch.tensor(NUSC_CLASS_FREQUENCIES, dtype=torch.float64) + 0.001)
        else:
            cw = torch.ones(17, dtype=torch.float64) / 17
        # a plain attribute in the reference (not in its state_dict): non-persistent buffer, so it follows
        # .to(device) and no host->device copy happens inside a captured step
        self.register_buffer("class_weights", cw.float(), persistent=False)
        bias = dict(conv_cfg).get("bias", True)
        self.occ_convs = nn.ModuleList()
        for i in range(num_level):
            mid = in_channels[i] // 2
            self.occ_convs.append(nn.Sequential(nn.Conv3d(in_channels[i], mid, 3, 1, 1, bias=bias),
                                                _norm_layer(norm_cfg, mid), nn.ReLU(inplace=True)))
        self.occ_pred_conv = nn.Sequential(nn.Conv3d(mid, mid // 2, 1, bias=bias), _norm_layer(norm_cfg, mid // 2),
                                           nn.ReLU(inplace=True), nn.Conv3d(mid // 2, out_channel, 1, bias=bias))
        self.num_point_sampling_feat = num_level
        # fine / cascade stage (occ_head.py:58-82): parameter containers with the reference's names and shapes
        self.point_cloud_range = torch.tensor(point_cloud_range).float()
        self.fine_stage = cascade_ratio != 1 and (sample_from_voxel or sample_from_img)
        self.fine_select = "auto"          # see forward_fine
        # (seed, draw counter) of the device-side subset selection; not part of the reference's state_dict
        self.register_buffer("fine_rng_state", torch.tensor([0x5EED5EED, 0], dtype=torch.int64), persistent=False)
        if self.fine_stage:
            fine_in = 128 if sample_from_voxel else 0
            if sample_from_img:
                self.img_mlp_0 = nn.Sequential(nn.Conv2d(512, 128, 1, 1, 0), nn.GroupNorm(16, 128), nn.ReLU(inplace=True))
                self.img_mlp = nn.Sequential(nn.Linear(128, 64), nn.GroupNorm(16, 64), nn.ReLU(inplace=True))
                fine_in += 64
            self.fine_mlp = nn.Sequential(nn.Linear(fine_in, 64), nn.GroupNorm(16, 64), nn.ReLU(inplace=True),
                                          nn.Linear(64, out_channel))
        if soft_weights:
            self.voxel_soft_weights = nn.Sequential(
                nn.Conv3d(mid, mid // 2, 1, bias=bias), _norm_layer(norm_cfg, mid // 2), nn.ReLU(inplace=True),
                nn.Conv3d(mid // 2, num_level, 1, bias=bias))
        _cl3d_all(self)

    def forward_coarse_voxel(self, voxel_feats):
        occs, dims = [], []
        for f, m in zip(voxel_feats, self.occ_convs):
            x2d, d = CF.to_cl2d(f)
            y, _ = conv_bn_act(x2d, d, m[0], m[1])
            occs.append(y)
            dims.append(d)
        d0 = dims[0]
        if self.soft_weights:
            w, _ = conv_bn_act(occs[0], d0, self.voxel_soft_weights[0], self.voxel_soft_weights[1])
            w, _ = conv_bn_act(w, d0, self.voxel_soft_weights[3], None, relu=False)
            w = torch.softmax(w, dim=1)
        else:
            w = torch.full((1, self.num_level), 1.0 / self.num_level, device=occs[0].device)
        if self.soft_weights:
            # occ_head.py:161-165: every level resized, weighted and summed in one pass
            feats = CF.resize_mix(occs, dims, d0, wts=w)
        else:
            feats = CF.resize_mix(occs, dims, d0) * (1.0 / self.num_level)
        # (out_voxel_feats is also sampled by the fine stage: that gradient joins occ_pred_conv's data gradient)
        y, _, feats = conv_bn_act(feats, d0, self.occ_pred_conv[0], self.occ_pred_conv[1], skip=True)
        occ, _ = conv_bn_act(y, d0, self.occ_pred_conv[3], None, relu=False)
        return {"out_voxel_feats": [CF.to_5d(feats, d0)], "occ": [CF.to_5d(occ, d0)]}

    def loss_voxel(self, output_voxels, target_voxels, tag):
        """occ_head.py:267-293: label vote to the output resolution, then CE / sem_scal / geo_scal /
        Lovasz-softmax, all in csrc/occ_loss.cu (one autograd node)."""
        x2d, dims = CF.to_cl2d(output_voxels)
        labels = CF.downsample_labels(target_voxels, dims, self.empty_idx)
        cw = self.class_weights
        if cw.device != x2d.device:
            cw = cw.to(x2d.device)
        l4 = CF.occ_voxel_losses(x2d, labels, cw, 255, self.empty_idx)
        return {'loss_voxel_ce_{}'.format(tag): self.loss_voxel_ce_weight * l4[0],
                'loss_voxel_sem_scal_{}'.format(tag): self.loss_voxel_sem_scal_weight * l4[1],
                'loss_voxel_geo_scal_{}'.format(tag): self.loss_voxel_geo_scal_weight * l4[2],
                'loss_voxel_lovasz_{}'.format(tag): self.loss_voxel_lovasz_weight * l4[3]}

    def loss_point(self, fine_coord, fine_output, target_voxels, tag):
        """occ_head.py:295-312: the four voxel losses on the sampled points (CE without class weights, :305).
        Coordinates from the device-side selection carry their padding count (`_coocc_nsel`): padding slots get
        label 255, which all four losses ignore."""
        nsel = getattr(fine_coord, "_coocc_nsel", None)
        if nsel is not None:
            gt = CF.fine_gather_labels(fine_coord, fine_coord._coocc_topk, nsel, target_voxels, 255)
        else:
            gt = target_voxels[:, fine_coord[0, :], fine_coord[1, :], fine_coord[2, :]].long()[0]
            gt = gt.to(torch.int32).contiguous()
        l4 = CF.occ_voxel_losses(fine_output, gt, None, 255, self.empty_idx)
        return {'loss_voxel_ce_{}'.format(tag): self.loss_voxel_ce_weight * l4[0],
                'loss_voxel_sem_scal_{}'.format(tag): self.loss_voxel_sem_scal_weight * l4[1],
                'loss_voxel_geo_scal_{}'.format(tag): self.loss_voxel_geo_scal_weight * l4[2],
                'loss_voxel_lovasz_{}'.format(tag): self.loss_voxel_lovasz_weight * l4[3]}

    def loss(self, output_voxels=None, output_coords_fine=None, output_voxels_fine=None, target_voxels=None,
             target_points=None, img_metas=None, visible_mask=None, **kwargs):
        """occ_head.py:314-337 (lidarseg is not on the path)."""
        loss_dict = {}
        for index, output_voxel in enumerate(output_voxels):
            loss_dict.update(self.loss_voxel(output_voxel, target_voxels, tag='c_{}'.format(index)))
        if self.cascade_ratio != 1 and output_voxels_fine is not None:          # :320-331
            acc = {}
            for fine_coord, fine_output in zip(output_coords_fine, output_voxels_fine):
                for k, v in self.loss_point(fine_coord, fine_output, target_voxels, tag='fine').items():
                    acc[k] = v if k not in acc else acc[k] + v
            for k, v in acc.items():
                loss_dict[k] = v / len(output_coords_fine)
        if target_points:
            raise NotImplementedError("OccHead lidarseg losses are outside the hot path (SURVEY §8f)")
        return loss_dict

    def _select_host(self, coarse_occ):
        """occ_head.py:183-205 as written: host synchronisation (nonzero) and torch's CPU generator (randperm)."""
        mask = coarse_occ.argmax(1) != self.empty_idx                                           # :183
        if int(mask.sum()) == 0:
            raise AssertionError('no foreground in coarse voxel')                               # :184
        coarse = torch.nonzero(mask[0]).t().contiguous()                                        # [3,N], (x,y,z) lexicographic
        r = self.cascade_ratio
        topk = self.fine_topk if self.training else 30000
        off = torch.stack(torch.meshgrid(*([torch.arange(r, device=coarse.device)] * 3), indexing='ij'), 3).reshape(-1, 3)
        fine = coarse[None] * r + off[:, :, None]                                               # [r^3,3,N]
        if fine.shape[-1] >= topk:
            fine = fine[:, :, torch.randperm(fine.shape[-1])[:topk].to(fine.device)]            # CPU generator like the reference
        return fine.permute(1, 0, 2).reshape(3, -1)

    def _select_device(self, coarse_occ):
        """Same selection without leaving the device (csrc/fine_select.cu): fixed-capacity buffers of fine_topk
        parents, a keyed pseudo-random subset instead of torch.randperm, padding slots flagged through `_coocc_nsel`."""
        x2d, dims = CF.to_cl2d(coarse_occ)
        topk = self.fine_topk if self.training else 30000
        topk = min(int(topk), dims[0] * dims[1] * dims[2])
        if self.fine_rng_state.device != x2d.device:
            self.fine_rng_state = self.fine_rng_state.to(x2d.device)
        coords, nsel = CF.fine_select(x2d, dims, self.empty_idx, self.cascade_ratio, topk, self.fine_rng_state)
        coords._coocc_nsel, coords._coocc_topk = nsel, topk
        return coords

    def forward_fine(self, out_voxel_feats, coarse_occ, img_feats, transform):
        """occ_head.py:182-237 for B = 1: occupied coarse voxels -> their cascade_ratio^3 children (a random subset of
        fine_topk parents in training, coordinate_transform.py:19-21) -> trilinear sample of out_voxel_feats,
        camera projection + bilinear sample of the image features (img_mlp_0, masked camera sum, img_mlp),
        fine_mlp -> per-point logits.  Sampling / projection / GroupNorm run in csrc/fine_stage.cu, the Linear
        layers on the tensor-core conv kernel.  Returns (fine_coord [3,M], fine_output [M, out_channel]).

        `self.fine_select`: "host" = the reference's own draw (nonzero + torch.randperm: one host synchronisation, the
        coordinates are bit-identical to the reference's under the same torch seed); "device" = csrc/fine_select.cu
        (no synchronisation, CUDA-graph capturable, a different but equally distributed subset); "auto" = device
        while a CUDA graph is being captured or replayed by GraphedStep, host otherwise."""
        x2d, dims = CF.to_cl2d(out_voxel_feats)
        r = self.cascade_ratio
        mode = self.fine_select
        if mode == "auto":
            mode = "device" if (CF.GSF_OVERRIDE is not None or torch.cuda.is_current_stream_capturing()) else "host"
        fine = self._select_device(coarse_occ) if mode == "device" else self._select_host(coarse_occ)
        feats = []
        if self.sample_from_voxel:
            feats.append(CF.fine_sample_voxels(x2d, dims, fine, self.final_occ_size))
        if self.sample_from_img and img_feats is not None:
            f = img_feats[0]                                                                    # [1,n,512,Hf,Wf]
            _, n, Ci, Hf, Wf = f.shape
            rows = f[0].permute(0, 2, 3, 1).reshape(n * Hf * Wf, Ci)
            c0 = self.img_mlp_0[0]
            rows = CF.linear(rows, c0.weight.reshape(c0.out_channels, Ci), c0.bias)
            rows = CF.group_norm_rows(rows, self.img_mlp_0[1], span=Hf * Wf, relu=True)
            X, Y, Z = dims
            W_img, H_img = transform[-1][1], transform[-1][0]        # img_size = (H, W), per-sample entries
            W_img = float(W_img[0]) if hasattr(W_img, "__getitem__") else float(W_img)
            H_img = float(H_img[0]) if hasattr(H_img, "__getitem__") else float(H_img)
            uv, m = CF.fine_project(fine, transform[0][0], transform[1][0], transform[2][0], transform[3][0],
                                    transform[4][0], transform[5][0], self.point_cloud_range,
                                    W_img, H_img, (X * r, Y * r, Z * r))
            s = CF.fine_sample_images(rows, n, Hf, Wf, uv, m)
            s = CF.linear(s, self.img_mlp[0].weight, self.img_mlp[0].bias)
            feats.append(CF.group_norm_rows(s, self.img_mlp[1], span=1, relu=True))
        h = torch.cat(feats, 1)
        h = CF.linear(h, self.fine_mlp[0].weight, self.fine_mlp[0].bias)
        h = CF.group_norm_rows(h, self.fine_mlp[1], span=1, relu=True)
        return fine, CF.linear(h, self.fine_mlp[3].weight, self.fine_mlp[3].bias)

    def forward(self, voxel_feats, img_feats=None, img_metas=None, pts_feats=None, target_points=None,
                transform=None, **kwargs):
        assert type(voxel_feats) is list and len(voxel_feats) == self.num_level
        out = self.forward_coarse_voxel(voxel_feats)
        fine_output = fine_coord = None
        if self.fine_stage:
            assert out["occ"][0].shape[0] == 1, "batch 1 like the rest of the path"
            c, o = self.forward_fine(out["out_voxel_feats"][0], out["occ"][0], img_feats, transform)
            fine_coord, fine_output = [c], [o]
        return {"output_voxels": out["occ"], "output_voxels_fine": fine_output, "output_coords_fine": fine_coord,
                "output_points": None}


# ----------------------------------------------------------------------------------------
class _Scales(nn.Module):
    """`posi_encoder.scales` buffer the reference MLP registers but never uses."""

    def __init__(self):
        super().__init__()
        self.register_buffer("scales", torch.tensor([2 ** i for i in range(0, 10)]))


class MLP(nn.Module):
    """nerf_mlp.MLP restricted to what coocc_ray.py:112-113 builds (skip_layer=None, ReLU hidden
    activations, identity output, xavier-uniform weights, zero biases)."""

    def __init__(self, input_dim, output_dim=None, net_depth=8, net_width=256, skip_layer=None, **kw):
        super().__init__()
        if skip_layer is not None:
            raise NotImplementedError("render heads are built with skip_layer=None")
        self.input_dim, self.output_dim, self.net_depth, self.net_width = input_dim, output_dim, net_depth, net_width
        self.hidden_layers = nn.ModuleList()
        self.posi_encoder = _Scales()
        cin = input_dim
        for _ in range(net_depth):
            self.hidden_layers.append(nn.Linear(cin, net_width))
            cin = net_width
        self.output_layer = nn.Linear(cin, output_dim)
        for m in list(self.hidden_layers) + [self.output_layer]:
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)

    def forward_rows(self, x2d, relu_out=False):
        for lin in self.hidden_layers:
            x2d = CF.linear(x2d, lin.weight, lin.bias, relu=True, out_bf16=True)
        return CF.linear(x2d, self.output_layer.weight, self.output_layer.bias, relu=relu_out)

    def forward(self, x):
        shp = x.shape
        return self.forward_rows(x.reshape(-1, shp[-1])).reshape(*shp[:-1], self.output_dim)


def render_fn(voxel_feats, gemo, sigma_head, rgb_head, gt_depth, gt_img):
    """The inline render block of COOCC_Ray.forward_train (coocc_ray.py:358-433) and its test-time copy (:562-637).

    voxel_feats [1,C,X,Y,Z] (fuser output), gemo [1,N,D,H,W,3], gt_depth = img_inputs[7] [1,N,16H,16W],
    gt_img = img_inputs[0] [1,N,3,16H,16W] (both None at test time: maps only, empty loss dict).
    Returns (rgbs [N,16H,16W,3], depths [N,16H,16W], losses dict).
    The heads run once per voxel of the render box, not once per sample (identical values).
    rgb_head None = the LiDAR-only detector (coocc_ray.py:435-494): density head only, colour columns zero.
    """
    B, N, D, H, W, _ = gemo.shape
    assert B == 1
    x2d, dims = CF.to_cl2d(voxel_feats)
    rows = CF.box_rows(x2d, dims)               # (bf16 activations: gathered as stored, no fp32 copy of the grid)
    sigma = sigma_head.forward_rows(rows, relu_out=True)        # [T,1] = relu(sigma_head(f))
    if rgb_head is not None:
        rgb_raw = rgb_head.forward_rows(rows)                   # [T,3]
    else:
        rgb_raw = torch.zeros(sigma.shape[0], 3, device=sigma.device, dtype=sigma.dtype)
    tab = torch.cat([rgb_raw, sigma], dim=1)                    # [T,4]
    rgb_map, depth_map = CF.composite(tab, gemo[0], dims)
    want_loss = gt_depth is not None
    if gt_depth is None:
        gt_depth = torch.zeros(1, N, 16 * H, 16 * W, device=tab.device)
    if gt_img is None:
        gt_img = torch.zeros(1, N, 3, 16 * H, 16 * W, device=tab.device)
    losses2, rgbs, depths = CF.upsample_losses(rgb_map, depth_map, gt_img[0], gt_depth[0], D)
    losses = {}
    if want_loss:
        losses["loss_depth_render"] = losses2[0]
        if rgb_head is not None:
            losses["loss_rgb"] = losses2[1]
    return rgbs, depths, losses


def render_depth_fn(voxel_feats, gemo, sigma_head, gt_depth):
    """LiDAR-only render branch (coocc_ray.py:435-494, COOCC_Ray_L): depth maps and loss_depth_render from the
    density head alone.  Returns (depths [N,16H,16W], {"loss_depth_render": ...})."""
    _, depths, losses = render_fn(voxel_feats, gemo, sigma_head, None, gt_depth, None)
    return depths, losses
