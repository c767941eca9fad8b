"""Lift-Splat voxel pooling and frustum geometry on the C ABI (csrc/lss_pool.cu).

Mirrors the pieces of the reference view transformer that feed the hot path:
    ViewTransformerLiftSplatShootVoxel.voxel_pooling   P/coocc/image2bev/ViewTransformerLSSVoxel.py:100-123
    bev_pool                                            M/ops/bev_pool/bev_pool.py:80-97
    ViewTransformerLiftSplatShoot.get_geometry / create_frustum / gen_dx_bx
                                                        P/coocc/image2bev/ViewTransformerLSSBEVDepth.py:21-25, 103-150
The depth net, the 2D backbone and the depth losses stay the reference's code.  Batch 1 (like the rest of
the path).  Outputs are the reference's logical [1,C,X,Y,Z] tensors in channels-last memory.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import functional as CF
from .functional import _p, _stream


def gen_dx_bx(xbound, ybound, zbound):
    """ViewTransformerLSSBEVDepth.py:21-25 (fp32 tensors, like the reference's nn.Parameters)."""
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.Tensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


class GridSpec:
    """Host-side copy of the voxel grid constants (bx, dx, nx are tiny fp32 tensors in the reference): read
    once, so that pooling calls do not synchronise with the device."""

    def __init__(self, bx, dx, nx):
        bx, dx = bx.detach().float().cpu(), dx.detach().float().cpu()
        lo = bx - dx / 2.0                                        # (self.bx - self.dx / 2.) evaluated in fp32
        self.lo = (ctypes.c_float * 3)(*[float(v) for v in lo.tolist()])
        self.dx = (ctypes.c_float * 3)(*[float(v) for v in dx.tolist()])
        self.dims = tuple(int(v) for v in nx.detach().cpu().to(torch.long).tolist())


def _spec(bx, dx=None, nx=None):
    return bx if isinstance(bx, GridSpec) else GridSpec(bx, dx, nx)


def _sort_points(geom, spec):
    """-> (workspace, sorted_keys, sorted_vals, point_keys, segments pointers, npts, (X,Y,Z))."""
    L = _lib.lib()
    CF._require_cuda(geom)
    g = geom.reshape(-1, 3)
    if g.dtype != torch.float32 or not g.is_contiguous():
        g = g.float().contiguous()
    npts = g.shape[0]
    X, Y, Z = spec.dims
    nbytes = int(L.coocc_lss_workspace(npts, X * Y * Z))
    if nbytes < 0:
        raise RuntimeError("coocc_lss: unsupported size (%d points, %d voxels)" % (npts, X * Y * Z))
    ws = torch.empty(nbytes, device=g.device, dtype=torch.uint8)
    sk, sv, pk, sg = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.coocc_lss_sort(_p(g), npts, spec.lo, spec.dx, X, Y, Z, _p(ws), ctypes.byref(sk), ctypes.byref(sv),
                                ctypes.byref(pk), ctypes.byref(sg), _stream()), "lss_sort")
    return ws, sk, sv, pk, sg, npts, (X, Y, Z)


class _VoxelPoolFn(torch.autograd.Function):
    """plain mode: rows [npts, C] (the flattened volume) -> [V, C]."""

    @staticmethod
    def forward(ctx, rows, geom, spec):
        L = _lib.lib()
        ws, sk, sv, pk, sg, npts, dims = _sort_points(geom, spec)
        rows = CF._as_rows(rows.float() if rows.dtype != torch.float32 else rows)
        assert rows.shape[0] == npts
        C = rows.shape[1]
        V = dims[0] * dims[1] * dims[2]
        out = torch.empty(V, C, device=rows.device, dtype=torch.float32)
        _lib.check(L.coocc_lss_pool_fwd(sk, sv, sg, npts, V, C, _p(rows), rows.stride(0), None, 0, 0, _p(out), C,
                                        _stream()), "lss_pool_fwd")
        ctx.save_for_backward(ws)
        ctx.meta = (pk, npts, V, C)
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        (ws,) = ctx.saved_tensors
        pk, npts, V, C = ctx.meta
        gout = CF._as_rows(gout.float())
        d = torch.empty(npts, C, device=gout.device, dtype=torch.float32)
        _lib.check(L.coocc_lss_pool_bwd(pk, npts, V, C, _p(gout), gout.stride(0), None, 0, None, 0, 0, 0, _p(d), C,
                                        None, _stream()), "lss_pool_bwd")
        return d, None, None


class _LiftSplatFn(torch.autograd.Function):
    """fused mode: feat [ncam*HW, C] (NHWC image features), depth [ncam, D, HW] -> [V, C]."""

    @staticmethod
    def forward(ctx, feat, depth, geom, spec):
        L = _lib.lib()
        ws, sk, sv, pk, sg, npts, dims = _sort_points(geom, spec)
        feat = CF._as_rows(feat)
        depth = depth.contiguous()
        ncam, D, HW = depth.shape
        assert ncam * D * HW == npts and feat.shape[0] == ncam * HW
        C = feat.shape[1]
        V = dims[0] * dims[1] * dims[2]
        out = torch.empty(V, C, device=feat.device, dtype=torch.float32)
        _lib.check(L.coocc_lss_pool_fwd(sk, sv, sg, npts, V, C, _p(feat), feat.stride(0), _p(depth), D, HW, _p(out), C,
                                        _stream()), "lss_pool_fwd")
        ctx.save_for_backward(ws, feat, depth)
        ctx.meta = (pk, npts, V, C, ncam, D, HW)
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        ws, feat, depth = ctx.saved_tensors
        pk, npts, V, C, ncam, D, HW = ctx.meta
        gout = CF._as_rows(gout.float())
        dfeat = torch.empty(ncam * HW, C, device=gout.device, dtype=torch.float32)
        ddepth = torch.empty(ncam, D, HW, device=gout.device, dtype=torch.float32)
        _lib.check(L.coocc_lss_pool_bwd(pk, npts, V, C, _p(gout), gout.stride(0), _p(feat), feat.stride(0), _p(depth),
                                        ncam, D, HW, _p(dfeat), C, _p(ddepth), _stream()), "lss_pool_bwd")
        return dfeat, ddepth, None, None


def voxel_pooling(geom_feats, x, bx, dx=None, nx=None):
    """ViewTransformerLSSVoxel.py:100-123: geom_feats [1,N,D,H,W,3] ego metres, x [1,N,D,H,W,C] (the lifted
    volume) -> [1,C,X,Y,Z] (sum of the features of the frustum points falling into each voxel).
    bx, dx, nx: the reference's fp32 grid tensors, or a GridSpec built from them once."""
    B, N, D, H, W, C = x.shape
    assert B == 1, "the hot path is batch-1"
    spec = _spec(bx, dx, nx)
    out = _VoxelPoolFn.apply(x.reshape(-1, C), geom_feats, spec)
    return CF.to_5d(out, spec.dims)


def lift_splat(geom, depth_prob, img_feat, bx, dx=None, nx=None):
    """Lift + Splat of ViewTransformerLSSVoxel.forward (:137-145) without the [N,D,H,W,C] volume:
    depth_prob [N,D,H,W], img_feat [N,C,H,W], geom [1,N,D,H,W,3] -> [1,C,X,Y,Z]."""
    N, D, H, W = depth_prob.shape
    C = img_feat.shape[1]
    feat = img_feat.permute(0, 2, 3, 1).reshape(N * H * W, C)         # NHWC rows (a copy for NCHW inputs)
    spec = _spec(bx, dx, nx)
    out = _LiftSplatFn.apply(feat.contiguous(), depth_prob.reshape(N, D, H * W), geom, spec)
    return CF.to_5d(out, spec.dims)


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
    """ViewTransformerLSSBEVDepth.py:117-150 (3x3 intrinsics and 3x3 bda, the nuScenes case).
    frustum [D,H,W,3]; rots/intrins/post_rots [1,N,3,3]; trans/post_trans [1,N,3]; bda [1,3,3]
    -> [1,N,D,H,W,3].  The N tiny matrix inverses/products stay torch ops; the per-point transform is the kernel."""
    L = _lib.lib()
    CF._require_cuda(frustum, rots)
    B, N, _ = trans.shape
    assert B == 1 and intrins.shape[-1] == 3 and bda.shape[-1] == 3
    D, H, W, _ = frustum.shape
    inv_pr = torch.inverse(post_rots[0])                                  # :127
    combine = rots[0].matmul(torch.inverse(intrins[0]))                   # :138
    mats = torch.cat([post_trans[0].reshape(N, 3), inv_pr.reshape(N, 9), combine.reshape(N, 9),
                      trans[0].reshape(N, 3)], 1).float().contiguous()
    geom = torch.empty(1, N, D, H, W, 3, device=frustum.device, dtype=torch.float32)
    fr = frustum.float().contiguous()
    b9 = bda[0].reshape(9).float().contiguous()
    _lib.check(L.coocc_lss_geometry(_p(fr), N, D, H, W, _p(mats), _p(b9), _p(geom), _stream()), "lss_geometry")
    return geom


def create_frustum(input_size, dbound, downsample=16):
    """ViewTransformerLSSBEVDepth.py:103-115 / get_frustum of coocc_ray.py:738-748: [D, fH, fW, 3] (u, v, depth)."""
    ogfH, ogfW = input_size
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.arange(*dbound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


class LSSVoxelPool(nn.Module):
    """The geometry / splat half of ViewTransformerLiftSplatShootVoxel: same grid_config / data_config /
    downsample constructor fields, buffers dx / bx / nx / frustum, and the methods create_frustum,
    get_geometry, voxel_pooling (plus the fused lift_splat)."""

    def __init__(self, grid_config, data_config, downsample=16):
        super().__init__()
        self.grid_config, self.data_config, self.downsample = grid_config, data_config, downsample
        dx, bx, nx = gen_dx_bx(grid_config['xbound'], grid_config['ybound'], grid_config['zbound'])
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        self.spec = GridSpec(dx=dx, bx=bx, nx=nx)

    def create_frustum(self):          # ViewTransformerLSSBEVDepth.py:103-115
        return nn.Parameter(create_frustum(self.data_config['input_size'], self.grid_config['dbound'], self.downsample),
                            requires_grad=False)

    def get_geometry(self, rots, trans, intrins, post_rots, post_trans, bda):
        return get_geometry(self.frustum, rots, trans, intrins, post_rots, post_trans, bda)

    def voxel_pooling(self, geom_feats, x):
        return voxel_pooling(geom_feats, x, self.spec)

    def lift_splat(self, geom, depth_prob, img_feat):
        return lift_splat(geom, depth_prob, img_feat, self.spec)
