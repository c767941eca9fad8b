"""BASELINE config 5: the dense conv stack alone (con_enc of BiFuser_N, CustomResNet3D-18, FPN3D, OccHead
coarse convs) forward + backward on a 512x512x40 x C=128 grid -- 10.5 M voxels, the label resolution of
coocc_multi_r101_openoccupancy.py used as a working grid to stress the tensor-core path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import functional as CF, modules as M
grid = tuple(int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (512, 512, 40)))
C = 128
dev = "cuda"
torch.cuda.set_per_process_memory_fraction(0.92)        # an over-sized grid must raise, not take the box down
coocc_b200.set_precision("bf16")
torch.manual_seed(0)
cfg = coocc_b200.model_cfg(C, 2, fine=False, grid=grid)
model = coocc_b200.HotPath(cfg, C, use_rendering=False).to(dev).train()
X, Y, Z = grid
V = X * Y * Z
cat = (torch.randn(V, 4 * C, device=dev) * 0.3).to(torch.bfloat16)


def step():
    model.zero_grad(set_to_none=True)
    x = cat.requires_grad_(False)
    f = model.occ_fuser
    y, _ = M.conv_bn_act(x, grid, f.con_enc[0], f.con_enc[1])
    y, _ = M.conv_bn_act(y, grid, f.con_enc[3], f.con_enc[4])
    vf = CF.to_5d(y, grid)
    outs = model.pts_bbox_head.forward_coarse_voxel(model.semantic_neck(model.semantic_encoder(vf)))
    occ = outs["occ"][0]
    occ.float().square().mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
CF.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
prof, CF.PROFILE = CF.PROFILE, None
conv = [(a.elapsed_time(b), w) for a, b, w, t in prof if not t.startswith("hbm:")]
t = sum(c[0] for c in conv) / 1e3
f = sum(c[1] for c in conv)
ms = e0.elapsed_time(e1)
print(json.dumps(dict(workload="conv stack fwd+bwd, %dx%dx%d x C=%d, bf16" % (X, Y, Z, C), voxels=V, ms_per_step=ms,
                      voxels_per_s=V / ms * 1e3, conv_ms=t * 1e3, conv_launches=len(conv), conv_tflops=f / t / 1e12,
                      algorithmic_tflop=f / 1e12, peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)))
