"""oracle.ops -- TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle_ops.c with the
call signatures of the reference's `mmdet3d.ops.furthest_point_sample` /
`mmdet3d.ops.ball_query` (furthest_point_sample.py:15-35, ball_query.py:14-40)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_ops.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle_ops.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_ops.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_fps.restype = ctypes.c_int
        _lib.oracle_ball_query.restype = ctypes.c_int
    return _lib


def _fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def furthest_point_sample(points_xyz: torch.Tensor, num_points: int) -> torch.Tensor:
    """(B,N,3) float32 -> (B,num_points) int32, start index 0."""
    assert points_xyz.is_contiguous()
    xyz = np.ascontiguousarray(points_xyz.detach().cpu().numpy(), dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.empty((B, num_points), dtype=np.int32)
    rc = lib().oracle_fps(ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(num_points),
                          _fptr(xyz), _fptr(out))
    assert rc == 0
    return torch.from_numpy(out)


def ball_query(min_radius, max_radius, sample_num, xyz: torch.Tensor,
               center_xyz: torch.Tensor) -> torch.Tensor:
    """xyz (B,N,3), center_xyz (B,M,3) -> (B,M,sample_num) int32."""
    assert center_xyz.is_contiguous() and xyz.is_contiguous()
    assert min_radius < max_radius
    p = np.ascontiguousarray(xyz.detach().cpu().numpy(), dtype=np.float32)
    c = np.ascontiguousarray(center_xyz.detach().cpu().numpy(), dtype=np.float32)
    B, N, _ = p.shape
    M = c.shape[1]
    out = np.empty((B, M, sample_num), dtype=np.int32)
    rc = lib().oracle_ball_query(ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M),
                                 ctypes.c_float(min_radius), ctypes.c_float(max_radius),
                                 ctypes.c_int(sample_num), _fptr(c), _fptr(p), _fptr(out))
    assert rc == 0
    return torch.from_numpy(out)
