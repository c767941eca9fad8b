"""oracle.lss -- TEST INFRASTRUCTURE ONLY (parity checker; never imported by the product).

CPU restatement (torch fp32 / int64) of the Lift-Splat pieces that produce the hot path's inputs, SURVEY §8f rank 2:
    voxel_pooling      P/coocc/image2bev/ViewTransformerLSSVoxel.py:100-123
    bev_pool           M/ops/bev_pool/bev_pool.py:80-97 + src/bev_pool_cuda.cu:20-46 (interval sum)
    get_geometry       P/coocc/image2bev/ViewTransformerLSSBEVDepth.py:117-150
    lift               ViewTransformerLSSVoxel.py:137-140 (volume = depth_prob x img_feat)
Pinned against the reference's own Python lines executed in the build container (oracle/refshim.py
reference_voxel_pooling / reference_get_geometry -> tests/golden/reference_lss.npz); the CUDA-only bev_pool
op inside them is the restatement below (the reference cannot run it on a CPU).
"""
import torch


def voxel_indices(geom_feats, bx, dx, nx):
    """:107-117 -> (idx int64 [N',3], kept bool [N']).  `.long()` truncates toward zero, so coordinates in
    (-1, 0) land in cell 0 and are kept."""
    g = ((geom_feats - (bx - dx / 2.)) / dx).long().reshape(-1, 3)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    return g, kept


def voxel_pooling(geom_feats, x, bx, dx, nx):
    """:100-123.  geom_feats [1,N,D,H,W,3], x [1,N,D,H,W,C] -> [1,C,X,Y,Z] (sum per voxel)."""
    B, N, D, H, W, C = x.shape
    assert B == 1
    g, kept = voxel_indices(geom_feats, bx, dx, nx)
    X, Y, Z = (int(v) for v in nx.to(torch.long).tolist())
    rows = x.reshape(-1, C)[kept]
    g = g[kept]
    vox = (g[:, 0] * Y + g[:, 1]) * Z + g[:, 2]
    order = torch.argsort(vox, stable=True)              # bev_pool.py:89-90 (the reference's argsort is unstable:
    out = torch.zeros(X * Y * Z, C, dtype=x.dtype)       #  the fp32 summation order inside a voxel is unspecified)
    out.index_add_(0, vox[order], rows[order])
    return out.reshape(1, X, Y, Z, C).permute(0, 4, 1, 2, 3)


def lift(depth_prob, img_feat):
    """:137-140.  depth_prob [N,D,H,W], img_feat [N,C,H,W] -> volume [1,N,D,H,W,C]."""
    vol = depth_prob.unsqueeze(1) * img_feat.unsqueeze(2)          # [N,C,D,H,W]
    return vol.permute(0, 2, 3, 4, 1).unsqueeze(0)


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
    """ViewTransformerLSSBEVDepth.py:117-150 (3x3 intrinsics, 3x3 bda)."""
    B, N, _ = trans.shape
    p = frustum - post_trans.view(B, N, 1, 1, 1, 3)                                           # :126
    p = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(p.unsqueeze(-1))            # :127
    p = torch.cat((p[..., :2, :] * p[..., 2:3, :], p[..., 2:3, :]), 5)                         # :129-131
    combine = rots.matmul(torch.inverse(intrins))                                             # :138
    p = combine.view(B, N, 1, 1, 1, 3, 3).matmul(p).squeeze(-1) + trans.view(B, N, 1, 1, 1, 3)  # :139-140
    return bda.view(B, 1, 1, 1, 1, 3, 3).matmul(p.unsqueeze(-1)).squeeze(-1)                    # :148
