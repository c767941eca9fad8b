"""SparseLiDAREnc8x (P/coocc/voxel_encoder/sparse_lidar_enc.py:125-177) on the C ABI -- SURVEY §8f rank 3.

The reference builds this encoder from spconv 2.3.6 modules (SubMConv3d / SparseConv3d / SparseSequential; third
party, `import spconv.pytorch`, not vendored under /root/reference).  Here a sparse convolution over the N active
voxels is an explicit sparse im2col (csrc/sparse_conv.cu: neighbour table through a dense index grid, one gather)
followed by ONE GEMM on the tcgen05 conv kernel (csrc/conv_tc.cu as a 1x1x1 convolution over N rows, with the
BatchNorm-statistics epilogue); BatchNorm / GroupNorm / ReLU / residual run on the [N, C] rows with the same kernels
the dense path uses.  Same registry name, constructor arguments, forward signature, outputs and state_dict keys
(`conv_input.0.weight` [16,3,3,3,4] ... `conv3.2.net.4.running_var`, `conv_out.1.bias`; spconv stores its weights
[Cout, kz, ky, kx, Cin], which is exactly the [Cout, 27*Cin] GEMM operand).

Semantics reproduced (spconv's published algorithm, restated in oracle/sparse_enc.py as masked dense convolutions):
  SubMConv3d(k=3)                     output sites = input sites; out[o] = sum_k W_k in[o - 1 + k] over ACTIVE inputs
  SparseConv3d(k=3, stride 2, pad 1)  output sites = every site reached by an active input; same sum at o*2 - 1 + k
  norm layers / ReLU                  act on the [N, C] feature rows: BatchNorm statistics over the active voxels only;
                                      nn.GroupNorm(16, C) normalises every row on its own -- for conv_input (C = 16,
                                      one channel per group) that makes the output relu(bias) whatever the input
                                      (reference quirk Q13, reproduced)
  x.dense().permute(0,1,4,3,2)        [B, C, W, H, D] strided view of the dense [B, C, D, H, W] tensor
Batch size 1 like the rest of the path (sparse_lidar_enc.py:166 "bs=1 hardcode").
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from . import functional as CF
from .registry import MIDDLE_ENCODERS

_p, _stream = CF._p, CF._stream


# ------------------------------------------------------------------------------------------------------------
# index side
# ------------------------------------------------------------------------------------------------------------
class SpLevel:
    """Active sites of one resolution level: coords int32 [N,4] (b,z,y,x) in lexicographic order, the dense index
    grid, and (lazily) the SubM neighbour table shared by all SubMConv3d of the level (spconv's indice_key)."""

    def __init__(self, coords, dims):
        self.coords = coords.to(torch.int32).contiguous()
        self.dims = tuple(int(d) for d in dims)          # (D, H, W) = (z, y, x)
        self.n = self.coords.shape[0]
        D, H, W = self.dims
        c = self.coords.long()
        self.lin = (c[:, 1] * H + c[:, 2]) * W + c[:, 3]
        self.grid = torch.full((D * H * W,), -1, device=coords.device, dtype=torch.int32)
        self.grid[self.lin] = torch.arange(self.n, device=coords.device, dtype=torch.int32)
        self._subm = None

    def neighbors(self, out_coords, n_out, stride, pad):
        L = _lib.lib()
        D, H, W = self.dims
        nbr = torch.empty(max(n_out, 1), 27, device=self.grid.device, dtype=torch.int32)
        _lib.check(L.coocc_sp_neighbors(_p(out_coords), n_out, stride, pad, _p(self.grid), D, H, W, _p(nbr), _stream()),
                   "sp_neighbors")
        return nbr[:n_out]

    def subm_table(self):
        if self._subm is None:
            self._subm = self.neighbors(self.coords, self.n, 1, 1)
        return self._subm

    def downsample(self, stride=2, pad=1, ksize=3):
        """SparseConv3d(3, stride, padding=pad): (next level, neighbour table [N_next, 27] into this level)."""
        L = _lib.lib()
        od = tuple((d + 2 * pad - ksize) // stride + 1 for d in self.dims)
        flags = torch.zeros(od[0] * od[1] * od[2], device=self.grid.device, dtype=torch.uint8)
        _lib.check(L.coocc_sp_flag_outputs(_p(self.coords), self.n, stride, pad, od[0], od[1], od[2], _p(flags), _stream()),
                   "sp_flag_outputs")
        zyx = torch.nonzero(flags.view(od))                  # lexicographic (host sync: the count sizes the level)
        coords = torch.cat([torch.zeros(zyx.shape[0], 1, device=zyx.device, dtype=zyx.dtype), zyx], 1)
        nxt = SpLevel(coords, od)
        return nxt, self.neighbors(nxt.coords, nxt.n, stride, pad)


class _GatherColsFn(torch.autograd.Function):
    """feats [N_in, C] fp32, nbr [N_out, 27] -> cols [N_out, 27*C] (explicit sparse im2col); backward = scatter-add."""

    @staticmethod
    def forward(ctx, feats, nbr):
        L = _lib.lib()
        CF._require_cuda(feats, nbr)
        feats = CF._as_rows(feats.float() if feats.dtype != torch.float32 else feats)
        n_in, C = feats.shape
        n_out = nbr.shape[0]
        cols = torch.empty(n_out, 27 * C, device=feats.device, dtype=torch.float32)
        _lib.check(L.coocc_sp_gather_cols(_p(feats), feats.stride(0), C, _p(nbr), n_out, _p(cols), cols.stride(0),
                                          _stream()), "sp_gather_cols")
        ctx.save_for_backward(nbr)
        ctx.meta = (n_in, C)
        return cols

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        (nbr,) = ctx.saved_tensors
        n_in, C = ctx.meta
        g = g.float().contiguous()
        d = torch.zeros(n_in, C, device=g.device, dtype=torch.float32)
        _lib.check(L.coocc_sp_scatter_cols(_p(g), g.stride(0), C, _p(nbr), nbr.shape[0], _p(d), d.stride(0), _stream()),
                   "sp_scatter_cols")
        return d, None


def sparse_conv_rows(feats, nbr, conv, bn=None, relu=False, residual=None):
    """One SubMConv3d / SparseConv3d (+ BatchNorm over the active rows, + residual, + ReLU) on feature rows.
    conv: _SpConv parameter container (weight [Cout,3,3,3,Cin], optional bias)."""
    cols = _GatherColsFn.apply(feats, nbr)
    cout = conv.weight.shape[0]
    w5d = conv.weight.reshape(cout, -1, 1, 1, 1)
    n = cols.shape[0]
    if bn is None:
        y = CF.conv3d(cols, w5d, (n, 1, 1), 1, 1, bias=conv.bias)
        return torch.relu(y) if relu else y
    assert conv.bias is None
    if bn.training:
        y, stats = CF.conv3d(cols, w5d, (n, 1, 1), 1, 1, want_stats=True)
        if bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        track = bn.track_running_stats
        return CF.bn_act(y, stats, bn.weight, bn.bias, residual, relu, bn.eps, bn.momentum,
                         bn.running_mean if track else None, bn.running_var if track else None, True)
    y = CF.conv3d(cols, w5d, (n, 1, 1), 1, 1)
    return CF.bn_act_eval(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.eps, residual, relu)


# ------------------------------------------------------------------------------------------------------------
# parameter containers with spconv's names
# ------------------------------------------------------------------------------------------------------------
class _SpConv(nn.Module):
    """weight [Cout, kz, ky, kx, Cin] (+ bias) like spconv.SubMConv3d / SparseConv3d; spconv's default initialiser
    (kaiming_uniform_, a = sqrt(5), fan_in = 27 * Cin)."""

    def __init__(self, cin, cout, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, 3, 3, 3, cin))
        bound = 1.0 / math.sqrt(27 * cin)          # kaiming_uniform_(a = sqrt(5)): sqrt(6 / (6 * fan_in))
        nn.init.uniform_(self.weight, -bound, bound)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound)) if bias else None


def _bn1d(norm_cfg, c):
    cfg = dict(norm_cfg or dict(type="BN1d"))
    typ = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if typ not in ("BN1d", "SyncBN", "BN"):
        raise NotImplementedError("SparseLiDAREnc8x: BatchNorm (BN1d / SyncBN) norm_cfg only, got %s" % typ)
    return nn.BatchNorm1d(c, **cfg)


class _SparseBasicBlock(nn.Module):
    """sparse_lidar_enc.py:40-62: net = (SubMConv3d, norm, ReLU, SubMConv3d, norm); out = relu(net(x) + x)."""

    def __init__(self, c, norm_cfg):
        super().__init__()
        self.net = nn.Sequential(_SpConv(c, c, False), _bn1d(norm_cfg, c), nn.ReLU(inplace=True), _SpConv(c, c, False),
                                 _bn1d(norm_cfg, c))

    def forward_rows(self, x, nbr):
        y = sparse_conv_rows(x, nbr, self.net[0], self.net[1], relu=True)
        return sparse_conv_rows(y, nbr, self.net[3], self.net[4], relu=True, residual=x)


@MIDDLE_ENCODERS.register_module(force=True)
class SparseLiDAREnc8x(nn.Module):
    def __init__(self, input_channel, norm_cfg, base_channel, out_channel, sparse_shape_xyz, **kwargs):
        super().__init__()
        self.sparse_shape_xyz = sparse_shape_xyz
        b = base_channel
        self.conv_input = nn.Sequential(_SpConv(input_channel, b, True), nn.GroupNorm(16, b), nn.ReLU(inplace=True))

        def stage(cin, cout):          # post_act_block(spconv, stride 2) + two SparseBasicBlocks (:138-154)
            return nn.Sequential(nn.Sequential(_SpConv(cin, cout, False), _bn1d(norm_cfg, cout), nn.ReLU(inplace=True)),
                                 _SparseBasicBlock(cout, norm_cfg), _SparseBasicBlock(cout, norm_cfg))

        self.conv1, self.conv2, self.conv3 = stage(b, 2 * b), stage(2 * b, 4 * b), stage(4 * b, 8 * b)
        self.conv_out = nn.Sequential(_SpConv(8 * b, out_channel, True), nn.GroupNorm(16, out_channel),
                                      nn.ReLU(inplace=True))
        self._pad_in = (8 - input_channel % 8) % 8         # GEMM rows of 16-byte multiples in fp32 and in bf16

    def forward(self, voxel_features, coors, batch_size):
        """voxel_features [N, input_channel] fp32, coors [N,4] (batch, z, y, x), batch_size 1 ->
        {'x': [1, C, W, H, D] strided view (what BiFuser_N receives as pts_voxel_feats), 'pts_feats': [rows, coords]}"""
        CF._require_cuda(voxel_features, coors)
        if int(batch_size) != 1:
            raise NotImplementedError("batch size 1 like the rest of the path (sparse_lidar_enc.py:166)")
        dims = tuple(int(v) for v in self.sparse_shape_xyz[::-1])              # spconv spatial shape (z, y, x)
        lvl = SpLevel(coors.int(), dims)
        x = voxel_features.float()
        w_in = self.conv_input[0]
        if self._pad_in:               # zero channels (and zero weight columns) change nothing
            x = torch.nn.functional.pad(x, (0, self._pad_in))
        x = self._input_conv(x, lvl)
        x = CF.group_norm_rows(x, self.conv_input[1], span=1, relu=True)
        for st in (self.conv1, self.conv2, self.conv3):
            nxt, nbr_down = lvl.downsample()
            x = sparse_conv_rows(x, nbr_down, st[0][0], st[0][1], relu=True)
            lvl = nxt
            for blk in (st[1], st[2]):
                x = blk.forward_rows(x, lvl.subm_table())
        x = sparse_conv_rows(x, lvl.subm_table(), self.conv_out[0])
        x = CF.group_norm_rows(x, self.conv_out[1], span=1, relu=True)
        D, H, W = lvl.dims
        C = x.shape[1]
        dense = torch.zeros(D * H * W, C, device=x.device, dtype=x.dtype)
        dense = dense.index_put((lvl.lin,), x)                                  # x.dense(): rows -> grid
        dense = dense.t().contiguous().view(1, C, D, H, W)                      # [B, C, D, H, W] like spconv's dense()
        return {'x': dense.permute(0, 1, 4, 3, 2), 'pts_feats': [(x, lvl.coords)]}

    def _input_conv(self, x, lvl):
        conv = self.conv_input[0]
        if not self._pad_in:
            return sparse_conv_rows(x, lvl.subm_table(), conv)
        cout, cin = conv.weight.shape[0], conv.weight.shape[4]
        cols = _GatherColsFn.apply(x, lvl.subm_table())                         # [N, 27 * (cin + pad)]
        w = torch.nn.functional.pad(conv.weight, (0, self._pad_in)).reshape(cout, -1, 1, 1, 1)
        return CF.conv3d(cols, w, (cols.shape[0], 1, 1), 1, 1, bias=conv.bias)
