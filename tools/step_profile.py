"""Kernel-time table of one hot-path training step (torch.profiler / CUPTI), warm caches."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="r50")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--top", type=int, default=45)
a = ap.parse_args()
cfg = S.CONFIGS[a.workload]
C, K = cfg["C"], cfg["K"]
dev = "cuda"
coocc_b200.set_precision(a.precision)
torch.manual_seed(0)
model = coocc_b200.HotPath(coocc_b200.model_cfg(C, K), C).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
inp = S.make_inputs(a.workload)
d = {k: v.to(dev) for k, v in inp.items()}
X, Y, Z = cfg["grid"]
occ = S.make_gt_occ(cfg["grid"], 2, 0).to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    losses, _, _ = model.forward_train(d["img_voxel_feats"], d["pts_voxel_feats"], d["geom"], d["gt_depth"], d["gt_img"], occ)
    sum(losses.values()).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(e.device_time_total for e in evs)
print("workload %s precision %s: total device time %.3f ms in %d kernel names" % (a.workload, a.precision, tot / 1e3, len(evs)))
for e in sorted(evs, key=lambda e: -e.device_time_total)[:a.top]:
    print("%9.1f us %5d  %5.1f%%  %s" % (e.device_time_total, e.count, 100 * e.device_time_total / tot, e.key[:100]))
