"""Lift-Splat (csrc/lss_pool.cu) timing at the r50 / r101 configurations: CUDA events, warm caches excluded by size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import lss as LSS

dev = "cuda"
for name, (N, D, H, W, C, size) in {"r50": (6, 112, 16, 44, 128, (256, 704)), "r101": (6, 112, 56, 100, 128, (896, 1600))}.items():
    grid = dict(xbound=[-50.0, 50.0, 1.0], ybound=[-50.0, 50.0, 1.0], zbound=[-5.0, 3.0, 1.0], dbound=[2.0, 58.0, 0.5])
    m = LSS.LSSVoxelPool(grid, dict(input_size=size), 16).to(dev)
    rig = {k: v.to(dev) for k, v in coocc_b200.synthetic.make_camera_rig(N, 1, *size).items()}
    depth = torch.softmax(torch.randn(N, D, H, W, device=dev), 1).requires_grad_(True)
    feat = torch.randn(N, C, H, W, device=dev).requires_grad_(True)

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    geom = m.get_geometry(rig["rots"], rig["trans"], rig["intrins"], rig["post_rots"], rig["post_trans"], rig["bda"])
    t_geo = timeit(lambda: m.get_geometry(rig["rots"], rig["trans"], rig["intrins"], rig["post_rots"], rig["post_trans"], rig["bda"]))
    t_fwd = timeit(lambda: m.lift_splat(geom, depth.detach(), feat.detach()))

    def fb():
        depth.grad = feat.grad = None
        m.lift_splat(geom, depth, feat).sum().backward()
    t_fb = timeit(fb)
    npts = N * D * H * W
    alg_fwd = npts * (12 + 4) + N * H * W * C * 4 + 80000 * C * 4          # geom + depth + features + output
    print("%s: %d points, C=%d | get_geometry %.3f ms | lift_splat fwd %.3f ms (%.0f GB/s algorithmic) | fwd+bwd %.3f ms"
          % (name, npts, C, t_geo, t_fwd, alg_fwd / t_fwd / 1e6, t_fb))
    if name == "r101":
        vol = (depth.detach().unsqueeze(1) * feat.detach().unsqueeze(2)).permute(0, 2, 3, 4, 1).unsqueeze(0).contiguous()
        t_mat = timeit(lambda: m.voxel_pooling(geom, vol), 5)
        print("r101: reference-signature voxel_pooling of the materialised %.2f GB volume: %.3f ms" % (vol.numel() * 4 / 1e9, t_mat))
    if name == "r101":
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fb()
            torch.cuda.synchronize()
        evs = [e for e in prof.key_averages() if e.device_time_total > 0]
        tot = sum(e.device_time_total for e in evs)
        print("r101 fwd+bwd kernel table: %.3f ms device time" % (tot / 1e3))
        for e in sorted(evs, key=lambda e: -e.device_time_total)[:14]:
            print("  %8.1f us %3d  %s" % (e.device_time_total, e.count, e.key[:90]))
