"""Registry glue: the reference resolves `type=` strings of projects/configs/coocc_nusc/*.py
through the OpenMMLab registries (mmdet3d/models/builder.py:10-14).  When mmdet3d is
importable the classes of this package are registered there with force=True (replacing the
reference plugin's classes of the same name); otherwise a minimal local registry with the same
`register_module` / `build` surface is used so configs can be exercised without OpenMMLab."""


class _LocalRegistry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError("%s is already registered in %s" % (key, self.name))
            self.module_dict[key] = cls
            return cls

        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        cfg = dict(cfg)
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


try:  # pragma: no cover - exercised only where OpenMMLab is installed
    from mmdet3d.models.builder import FUSION_LAYERS  # noqa: F401
    from mmdet.models import BACKBONES, DETECTORS, HEADS, NECKS  # noqa: F401
    try:
        from mmdet3d.models.builder import MIDDLE_ENCODERS  # noqa: F401
    except Exception:  # noqa: BLE001
        MIDDLE_ENCODERS = BACKBONES
    HAVE_MMDET3D = True
except Exception:  # noqa: BLE001
    MODELS = _LocalRegistry("models")
    FUSION_LAYERS = BACKBONES = NECKS = HEADS = DETECTORS = MIDDLE_ENCODERS = MODELS
    HAVE_MMDET3D = False


def build_fusion_layer(cfg):
    return FUSION_LAYERS.build(cfg)


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_middle_encoder(cfg):
    return MIDDLE_ENCODERS.build(cfg)
