"""coocc_b200 -- B200-native fused-voxel hot path of Co-Occ (GSFusion -> 3D conv decoder/head ->
volume-render regulariser) behind the reference's mmdet3d-plugin module API.

Host code is Python/PyTorch (device memory, streams, autograd plumbing); all arithmetic on the
path runs in hand-written sm_100a CUDA reached through the C-ABI in include/coocc_b200.h
(csrc/libcoocc_b200.so).  There is no CPU fallback: calling a module without the library or
without a CUDA device raises.
"""
__version__ = "0.1.0"

from . import synthetic  # noqa: F401
from . import _lib  # noqa: F401
from .functional import get_precision, set_precision  # noqa: F401
from .modules import BiFuser_N, CustomResNet3D, FPN3D, MLP, OccHead, render_fn  # noqa: F401
from .hotpath import HotPath, model_cfg  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from . import lss  # noqa: F401
from . import sparse_enc  # noqa: F401
from . import detector  # noqa: F401
