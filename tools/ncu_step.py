"""One warmed-up forward+backward of the hot path between cudaProfilerStart/Stop, for
`ncu --profile-from-start off -k regex:<kernels>` captures of the non-conv kernels."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="northstar")
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
cfg = S.CONFIGS[a.workload]
C, K = cfg["C"], cfg["K"]
dev = "cuda"
coocc_b200.set_precision(a.precision)
torch.manual_seed(0)
model = coocc_b200.HotPath(coocc_b200.model_cfg(C, K), C).to(dev).train()
inp = S.make_inputs(a.workload)
d = {k: v.to(dev) for k, v in inp.items()}
X, Y, Z = cfg["grid"]
occ = S.make_gt_occ(cfg["grid"], 2, 0).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    losses, _, _ = model.forward_train(d["img_voxel_feats"], d["pts_voxel_feats"], d["geom"], d["gt_depth"], d["gt_img"], occ)
    sum(losses.values()).backward()


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
