"""Run-to-run difference of the r50 conv stack's input gradient (fp32 mode): how much of the error against the reference
fixture is ReLU-mask flips caused by the non-deterministic summation order of the BatchNorm statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import synthetic as S
DEV = "cuda"; NC = dict(type="BN3d", requires_grad=True)
coocc_b200.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
cfg = S.CONFIGS["r50"]; C = cfg["C"]; P = S.make_params("r50"); planes = [C, 2 * C, 4 * C, 8 * C]
x0 = torch.randn(1, C, *cfg["grid"], generator=torch.Generator().manual_seed(1234)) * 0.5
enc = coocc_b200.CustomResNet3D(depth=18, n_input_channels=C, block_inplanes=planes, out_indices=(0, 1, 2, 3), norm_cfg=NC).to(DEV)
neck = coocc_b200.FPN3D(with_cp=True, in_channels=planes, out_channels=2 * C, norm_cfg=NC).to(DEV)
head = coocc_b200.OccHead(norm_cfg=NC, soft_weights=True, num_level=4, in_channels=[2 * C] * 4, out_channel=17).to(DEV)
enc.load_state_dict(P["semantic_encoder"]); neck.load_state_dict(P["semantic_neck"]); head.load_state_dict(P["pts_bbox_head"])
for m in (enc, neck, head): m.train()
def run():
    x = x0.to(DEV).requires_grad_(True)
    o = head.forward_coarse_voxel(neck(enc(x)))
    occ = o["occ"][0]
    w = torch.linspace(-1, 1, occ.numel()).reshape(occ.shape)
    (occ * w.to(DEV)).sum().backward()
    return occ.detach().float(), x.grad.float()
o1, g1 = run(); o2, g2 = run(); o3, g3 = run()
for a, b, n in ((g1, g2, "dx run1 vs run2"), (g1, g3, "dx run1 vs run3"), (o1, o2, "occ run1 vs run2")):
    print("%s: rel L2 %.3e  max/max %.3e  differing elements %.3e" % (n, ((a - b).norm() / a.norm()).item(),
          ((a - b).abs().max() / a.abs().max()).item(), ((a != b).float().mean()).item()))
