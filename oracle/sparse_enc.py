"""oracle.sparse_enc -- TEST INFRASTRUCTURE ONLY (parity checker; never imported by the product).

CPU restatement of SparseLiDAREnc8x (P/coocc/voxel_encoder/sparse_lidar_enc.py:125-177) as masked dense convolutions.

PARITY UNPINNED: the reference runs this encoder on spconv 2.3.6 (`import spconv.pytorch`, pinned in
docs/requirements_ref.txt:166-167), a third-party package that is neither vendored under /root/reference nor
installed here, and the reference has no test or fixture for the encoder.  What is restated is spconv's published
algorithm (SECOND, Yan et al. 2018; spconv docs):
  SubMConv3d(k=3)                     out[o] = sum_k W_k . in[o - 1 + k] at the ACTIVE INPUT sites o only (no dilation)
  SparseConv3d(k=3, stride 2, pad 1)  out[o] = sum_k W_k . in[2 o - 1 + k] at every site o reached by an active input
  weight layout                       [Cout, kz, ky, kx, Cin]; spatial shape (z, y, x) = sparse_shape_xyz[::-1]
  norm layers inside SparseSequential act on the [N, C] feature matrix (BatchNorm over active voxels, GroupNorm per row)
  dense()                             [B, C, D, H, W], zeros at inactive sites
`brute_force_conv` is an independent rulebook-style evaluation used by tests/test_oracle_sparse.py to check the masked
dense form against that definition.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _w5(w):
    """spconv weight [Cout, kz, ky, kx, Cin] -> torch conv3d weight [Cout, Cin, kz, ky, kx]"""
    return w.permute(0, 4, 1, 2, 3)


def rows_of(x, m):
    """dense [1,C,D,H,W], mask [1,1,D,H,W] bool -> [N,C] rows in lexicographic (z,y,x) order"""
    return x[0][:, m[0, 0]].t()


def put_rows(r, m):
    out = torch.zeros(1, r.shape[1], *m.shape[2:], dtype=r.dtype)
    out[0][:, m[0, 0]] = r.t()
    return out


def subm_conv(x, m, w, b=None):
    return F.conv3d(x, _w5(w), b, 1, 1) * m


def strided_conv(x, m, w):
    y = F.conv3d(x, _w5(w), None, 2, 1)
    m2 = F.conv3d(m.float(), torch.ones(1, 1, 3, 3, 3), None, 2, 1) > 0
    return y * m2, m2


def _bn_rows(x, m, p, pre, relu=True, residual=None, eps=1e-5):
    r = F.batch_norm(rows_of(x, m), None, None, p[pre + ".weight"], p[pre + ".bias"], True, 0.1, eps)
    if residual is not None:
        r = r + rows_of(residual, m)
    return put_rows(F.relu(r) if relu else r, m)


def _gn_rows(x, m, p, pre):
    r = F.group_norm(rows_of(x, m), 16, p[pre + ".weight"], p[pre + ".bias"], 1e-5)
    return put_rows(F.relu(r), m)


def _block(x, m, p, pre):
    """SparseBasicBlock (sparse_lidar_enc.py:40-62)"""
    y = _bn_rows(subm_conv(x, m, p[pre + ".net.0.weight"]), m, p, pre + ".net.1")
    y = subm_conv(y, m, p[pre + ".net.3.weight"])
    return _bn_rows(y, m, p, pre + ".net.4", relu=True, residual=x)


def sparse_encoder_forward(p, voxel_features, coors, sparse_shape_xyz):
    """sparse_lidar_enc.py:162-177 (training-mode BatchNorm).  voxel_features [N,Cin], coors [N,4] (b,z,y,x).
    Returns x.dense().permute(0,1,4,3,2): [1, C, W, H, D]."""
    D, H, W = sparse_shape_xyz[::-1]
    c = coors.long()
    m = torch.zeros(1, 1, D, H, W, dtype=torch.bool)
    m[0, 0, c[:, 1], c[:, 2], c[:, 3]] = True
    x = torch.zeros(1, voxel_features.shape[1], D, H, W)
    x[0][:, c[:, 1], c[:, 2], c[:, 3]] = voxel_features.t()
    x = subm_conv(x, m, p["conv_input.0.weight"], p["conv_input.0.bias"])           # :131-134
    x = _gn_rows(x, m, p, "conv_input.1")
    for st in ("conv1", "conv2", "conv3"):                                         # :136-154
        x, m = strided_conv(x, m, p[st + ".0.0.weight"])
        x = _bn_rows(x, m, p, st + ".0.1")
        x = _block(x, m, p, st + ".1")
        x = _block(x, m, p, st + ".2")
    x = subm_conv(x, m, p["conv_out.0.weight"], p["conv_out.0.bias"])              # :156-159
    x = _gn_rows(x, m, p, "conv_out.1")
    return x.permute(0, 1, 4, 3, 2), m


def brute_force_conv(feats, coords, dims, w, stride, pad, subm, bias=None):
    """Rulebook-style evaluation straight from the definition (numpy, tiny inputs): returns (out_coords [M,3] in
    lexicographic order, out_feats [M,Cout])."""
    feats = np.asarray(feats, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    index = {tuple(int(v) for v in c): i for i, c in enumerate(coords)}
    if subm:
        outs = sorted(index.keys())
    else:
        od = [(d + 2 * pad - 3) // stride + 1 for d in dims]
        s = set()
        for (z, y, x) in index.keys():
            for kz in range(3):
                for ky in range(3):
                    for kx in range(3):
                        t = (z + pad - kz, y + pad - ky, x + pad - kx)
                        if all(v >= 0 and v % stride == 0 for v in t):
                            o = tuple(v // stride for v in t)
                            if all(o[a] < od[a] for a in range(3)):
                                s.add(o)
        outs = sorted(s)
    res = np.zeros((len(outs), w.shape[0]))
    for oi, o in enumerate(outs):
        for kz in range(3):
            for ky in range(3):
                for kx in range(3):
                    i = (o[0] * stride - pad + kz, o[1] * stride - pad + ky, o[2] * stride - pad + kx)
                    j = index.get(i)
                    if j is not None:
                        res[oi] += w[:, kz, ky, kx, :] @ feats[j]
        if bias is not None:
            res[oi] += np.asarray(bias, dtype=np.float64)
    return np.array(outs, dtype=np.int64).reshape(-1, 3), res
