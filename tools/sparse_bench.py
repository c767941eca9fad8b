"""Times SparseLiDAREnc8x at the config's size (800x800x64 @ 0.125 m, ~90k voxels; coocc_multi_r50_256x704.py:127-134)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import synthetic as S
from coocc_b200.sparse_enc import SparseLiDAREnc8x
dev = "cuda"
shape = [800, 800, 64]
for prec in ("bf16", "tf32", "fp32"):
    coocc_b200.set_precision(prec)
    enc = SparseLiDAREnc8x(4, dict(type='SyncBN'), 16, 128, shape).to(dev).train()
    enc.load_state_dict(S.sparse_enc_params(), strict=True)
    for n in (90000,):
        # LiDAR-like occupancy: voxels concentrated in a slab (points cluster near the ground plane)
        g = torch.Generator().manual_seed(0)
        x = torch.randint(0, 800, (n * 2,), generator=g); y = torch.randint(0, 800, (n * 2,), generator=g)
        z = (torch.randn(n * 2, generator=g) * 4 + 20).clamp(0, 63).long()
        lin = torch.unique((z * 800 + y) * 800 + x)[:n]
        z, y, x = lin // 640000, (lin // 800) % 800, lin % 800
        coors = torch.stack([torch.zeros_like(z), z, y, x], 1).int().to(dev)
        feats = torch.randn(coors.shape[0], 4, generator=g).to(dev)

        def step():
            enc.zero_grad(set_to_none=True)
            out = enc(feats, coors, 1)['x']
            out.sum().backward()
            return out
        for _ in range(3):
            out = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record(); torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            enc(feats, coors, 1)
            e2.record()
            for _ in range(5):
                enc(feats, coors, 1)
            e3.record(); torch.cuda.synchronize()
        occ = int((out.sum(1) != 0).sum())
        print("%s: N=%d voxels -> %d of 80000 coarse sites; forward %.2f ms, forward+backward %.2f ms" % (
            prec, coors.shape[0], occ, e2.elapsed_time(e3) / 5, e0.elapsed_time(e1) / 5), flush=True)
