"""refshim -- TEST INFRASTRUCTURE ONLY.

Imports the *unmodified* reference Python files of the hot path from
``/root/reference`` behind minimal ``sys.modules`` stubs for the OpenMMLab
packages that are not installed here (mmcv / mmdet / mmdet3d).  It exists to

  * validate ``oracle/oracle.py`` (the portable restatement) against the real
    reference code, and
  * generate the golden fixtures under ``tests/golden`` (``oracle/make_golden.py``).

``/root/reference`` does not exist on the GPU box, so nothing under ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may import this module; ``available()`` is the
guard used by the CPU tests that do.

The two CUDA-only ops the reference calls (``furthest_point_sample``,
``ball_query``) are served by ``oracle/oracle_ops.c`` through
``oracle.ops``.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("COOCC_REFERENCE_ROOT", "/root/reference")
_PLUGIN = os.path.join(REF_ROOT, "projects", "mmdet3d_plugin")

_loaded = {}


def available() -> bool:
    return os.path.isfile(os.path.join(_PLUGIN, "coocc", "fuser", "bifuser_n.py"))


class _Registry:
    """Stand-in for mmcv.utils.Registry: decorator that records the class."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls

        if module is not None:
            return deco(module)
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg)


def _build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    typ = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if typ in ("BN3d", "SyncBN", "BN"):
        # SyncBN == BatchNorm3d at world size 1 (CPU has no SyncBN kernel)
        layer = nn.BatchNorm3d(num_features, **cfg)
        return "bn" + str(postfix), layer
    if typ == "GN":
        return "gn" + str(postfix), nn.GroupNorm(num_channels=num_features, **cfg)
    raise KeyError(typ)


def _build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type="Conv3d"))
    typ = cfg.pop("type")
    assert typ == "Conv3d", typ
    return nn.Conv3d(*args, **kwargs, **cfg)


class _ConvModule(nn.Module):
    """mmcv.cnn.ConvModule subset used by fpn3d.py: conv -> norm -> act, names conv/bn."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), bias="auto",
                 inplace=True):
        super().__init__()
        with_norm = norm_cfg is not None
        if bias == "auto":
            bias = not with_norm
        self.conv = _build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                      stride=stride, padding=padding, bias=bias)
        self.norm_name = None
        if with_norm:
            self.norm_name, norm = _build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        self.activate = nn.ReLU(inplace=inplace) if act_cfg is not None else None
        # mmcv ConvModule.init_weights(): kaiming_normal_(relu) for conv, constant 1/0 norm
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.norm_name:
            x = getattr(self, self.norm_name)(x)
        if self.activate is not None:
            x = self.activate(x)
        return x


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


def _passthrough_decorator(*dargs, **dkwargs):
    def deco(fn):
        return fn

    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return deco


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # mark as package
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _install_stubs():
    if "mmdet3d.models.builder" in sys.modules and getattr(
            sys.modules["mmdet3d.models.builder"], "_coocc_stub", False):
        return
    from . import ops  # C restatement of the CUDA-only ops

    MODELS = _Registry("models")
    regs = dict(FUSION_LAYERS=MODELS, BACKBONES=MODELS, NECKS=MODELS, HEADS=MODELS,
                DETECTORS=MODELS)
    _mod("mmcv")
    _mod("mmcv.cnn", build_norm_layer=_build_norm_layer, build_conv_layer=_build_conv_layer,
         build_upsample_layer=None, ConvModule=_ConvModule)
    _mod("mmcv.runner", BaseModule=_BaseModule, auto_fp16=_passthrough_decorator,
         force_fp32=_passthrough_decorator)
    _mod("mmdet")
    _mod("mmdet.core", reduce_mean=lambda x: x)
    _mod("mmdet.models", **regs)
    _mod("mmdet3d")
    _mod("mmdet3d.models", **regs)
    _mod("mmdet3d.models.builder", _coocc_stub=True, **regs)
    _mod("mmdet3d.ops", furthest_point_sample=ops.furthest_point_sample,
         ball_query=ops.ball_query, gather_points=None)


def _load(modname, relpath):
    if modname in _loaded:
        return _loaded[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(_PLUGIN, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    _loaded[modname] = mod
    return mod


def _install_plugin_namespace():
    """projects.mmdet3d_plugin.utils as a *synthetic* namespace: the real
    utils/__init__.py drags in torchmetrics/mmcv; occ_head.py only needs these."""
    _mod("projects")
    _mod("projects.mmdet3d_plugin")
    utils = _mod("projects.mmdet3d_plugin.utils")
    ct = _load("projects.mmdet3d_plugin.utils.coordinate_transform", "utils/coordinate_transform.py")
    _load("projects.mmdet3d_plugin.utils.nusc_param", "utils/nusc_param.py")
    _load("projects.mmdet3d_plugin.utils.semkitti", "utils/semkitti.py")
    utils.coarse_to_fine_coordinates = ct.coarse_to_fine_coordinates
    utils.project_points_on_img = ct.project_points_on_img
    utils.per_class_iu = None
    utils.fast_hist_crop = None
    _mod("projects.mmdet3d_plugin.coocc")
    _mod("projects.mmdet3d_plugin.coocc.dense_heads")
    _load("projects.mmdet3d_plugin.coocc.dense_heads.lovasz_softmax",
          "coocc/dense_heads/lovasz_softmax.py")


def load_reference():
    """Returns a namespace with the reference classes BiFuser_N, CustomResNet3D,
    FPN3D, OccHead, MLP (all the unmodified reference code)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    ns = types.SimpleNamespace()
    ns.BiFuser_N = _load("_ref_bifuser_n", "coocc/fuser/bifuser_n.py").BiFuser_N
    ns.CustomResNet3D = _load("_ref_resnet3d", "coocc/backbones/resnet3d.py").CustomResNet3D
    # reference calls torch.utils.checkpoint.checkpoint without use_reentrant (fpn3d.py:85,103)
    import torch.utils.checkpoint as _cp
    if not getattr(_cp, "_coocc_pinned", False):
        _orig = _cp.checkpoint

        def _pinned(fn, *a, **k):
            k.setdefault("use_reentrant", True)
            return _orig(fn, *a, **k)

        _cp.checkpoint = _pinned
        _cp._coocc_pinned = True
    ns.FPN3D = _load("_ref_fpn3d", "coocc/necks/fpn3d.py").FPN3D
    ns.MLP = _load("_ref_nerf_mlp", "utils/nerf_mlp.py").MLP
    _install_plugin_namespace()
    ns.OccHead = _load("projects.mmdet3d_plugin.coocc.dense_heads.occ_head",
                       "coocc/dense_heads/occ_head.py").OccHead
    return ns


_RENDER_FILE = os.path.join(_PLUGIN, "coocc", "detectors", "coocc_ray.py")
_RENDER_LINES = (358, 433)   # `if self.use_rendering:` ... `losses["loss_rgb"] = ...`
_render_code = None


def reference_render_block(voxel_feats, gemo, sigma_head, rgb_head, gt_depth, gt_img):
    """Runs the reference's inline render block (coocc_ray.py:358-433) *as is*.

    The detector file cannot be imported whole (it needs CenterPoint / matplotlib), and
    reference source must not be copied into this repo, so the block's own lines are
    read from the reference tree at run time, dedented, compiled and executed in a
    namespace that supplies the names the block reads: ``self.{use_rendering,sigma_head,
    rgb_head}``, ``voxel_feats``, ``gemo``, ``img_feats`` (non-None selects the camera
    branch), ``img_inputs`` (only [0] and [7] are read) and ``losses``.

    voxel_feats [1,C,X,Y,Z]; gemo [1,N,D,H,W,3]; gt_depth [1,N,16H,16W] (= img_inputs[7]);
    gt_img [1,N,3,16H,16W] (= img_inputs[0]).  Returns (rgbs, depths, losses dict).
    """
    global _render_code
    import textwrap
    import torch.nn.functional as F
    if _render_code is None:
        with open(_RENDER_FILE) as f:
            lines = f.readlines()
        block = "".join(lines[_RENDER_LINES[0] - 1:_RENDER_LINES[1]])
        assert block.lstrip().startswith("if self.use_rendering:"), "reference layout changed"
        _render_code = compile(textwrap.dedent(block), _RENDER_FILE, "exec")
    img_inputs = [None] * 14
    img_inputs[0] = gt_img
    img_inputs[7] = gt_depth
    ns = dict(torch=torch, F=F, losses={}, img_feats=True, gemo=gemo, voxel_feats=voxel_feats,
              img_inputs=tuple(img_inputs),
              self=types.SimpleNamespace(use_rendering=True, sigma_head=sigma_head,
                                         rgb_head=rgb_head))
    exec(_render_code, ns)
    return ns["rgbs"], ns["depths"], ns["losses"]


_EVAL_LINES = (659, 684)      # `def evaluation_semantic(self, pred, gt, eval_type, visible_mask=None):` ... `return hist, hist_occ`
_HIST_LINES = (726, 730)      # `def fast_hist(pred, label, max_label=18):`
_eval_ns = None


def reference_evaluation_semantic(pred, gt, eval_type, visible_mask=None, empty_idx=0):
    """Runs COOCC_Ray.evaluation_semantic and fast_hist (coocc_ray.py:659-684, 726-730) as they are: the
    two functions' own lines are read from the reference tree, dedented and exec'd (the detector file
    cannot be imported whole).  `np.int` (removed in numpy 2, used at :664) is supplied as `int`."""
    global _eval_ns
    import copy
    import textwrap
    import numpy
    import torch.nn.functional as F
    if _eval_ns is None:
        with open(_RENDER_FILE) as f:
            lines = f.readlines()
        src = textwrap.dedent("".join(lines[_EVAL_LINES[0] - 1:_EVAL_LINES[1]])) + "\n" + \
            textwrap.dedent("".join(lines[_HIST_LINES[0] - 1:_HIST_LINES[1]]))
        assert src.lstrip().startswith("def evaluation_semantic(self, pred, gt, eval_type"), "reference layout changed"

        class _NP:
            def __getattr__(self, k):
                return int if k == "int" else getattr(numpy, k)

        _eval_ns = dict(torch=torch, F=F, np=_NP(), copy=copy)
        exec(compile(src, _RENDER_FILE, "exec"), _eval_ns)
    return _eval_ns["evaluation_semantic"](types.SimpleNamespace(empty_idx=empty_idx), pred, gt, eval_type, visible_mask)


_LSSV_FILE = os.path.join(_PLUGIN, "coocc", "image2bev", "ViewTransformerLSSVoxel.py")
_LSSB_FILE = os.path.join(_PLUGIN, "coocc", "image2bev", "ViewTransformerLSSBEVDepth.py")
_VP_LINES = (100, 123)        # `def voxel_pooling(self, geom_feats, x):` ... `return final`
_GG_LINES = (117, 150)        # `def get_geometry(self, rots, ...):` ... `return points`
_lss_ns = None


def _cpu_bev_pool(feats, coords, B, D, H, W):
    """The CUDA-only op `bev_pool` (M/ops/bev_pool/bev_pool.py:80-97, src/bev_pool_cuda.cu:20-46) restated for
    the CPU: sort by voxel rank, sum every run, scatter to out[b][z][x][y][:], return [B,C,D,H,W]."""
    B, D, H, W = int(B), int(D), int(H), int(W)
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    idx = ranks.argsort()
    feats, coords = feats[idx], coords[idx]
    out = torch.zeros(B, D, H, W, feats.shape[1], dtype=feats.dtype)
    out.index_put_((coords[:, 3], coords[:, 2], coords[:, 0], coords[:, 1]), feats, accumulate=True)
    return out.permute(0, 4, 1, 2, 3).contiguous()


def _lss_namespace():
    global _lss_ns
    import textwrap
    if _lss_ns is None:
        ns = dict(torch=torch, bev_pool=_cpu_bev_pool)
        for path, (a, b), name in ((_LSSV_FILE, _VP_LINES, "voxel_pooling"), (_LSSB_FILE, _GG_LINES, "get_geometry")):
            with open(path) as f:
                lines = f.readlines()
            src = textwrap.dedent("".join(lines[a - 1:b]))
            assert src.lstrip().startswith("def %s(self" % name), "reference layout changed (%s)" % name
            exec(compile(src, path, "exec"), ns)
        _lss_ns = ns
    return _lss_ns


def reference_voxel_pooling(geom_feats, x, bx, dx, nx):
    """ViewTransformerLiftSplatShootVoxel.voxel_pooling (ViewTransformerLSSVoxel.py:100-123) run as it is, with
    `self` = (bx, dx, nx) and the CUDA op bev_pool replaced by its CPU restatement."""
    return _lss_namespace()["voxel_pooling"](types.SimpleNamespace(bx=bx, dx=dx, nx=nx), geom_feats, x)


def reference_get_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
    """ViewTransformerLiftSplatShoot.get_geometry (ViewTransformerLSSBEVDepth.py:117-150) run as it is."""
    return _lss_namespace()["get_geometry"](types.SimpleNamespace(frustum=frustum), rots, trans, intrins, post_rots,
                                            post_trans, bda)
