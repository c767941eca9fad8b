"""One warmed-up training step of the hot path (forward + backward + AdamW, eager launches) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures: the launch list of exactly one step
(`--metrics gpu__time_duration.sum`) or sections of selected kernels (`-k regex:<kernels>`)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import functional as CF
from coocc_b200 import synthetic as S
from coocc_b200.ddp import GradArena
from coocc_b200.optim import FusedAdamW, norm_decay_mults

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="northstar")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--no-fine", action="store_true")
ap.add_argument("--select", default="device", choices=["device", "host"])
ap.add_argument("--no-opt", action="store_true")
a = ap.parse_args()
cfg = S.CONFIGS[a.workload]
C, K = cfg["C"], cfg["K"]
dev = "cuda"
coocc_b200.set_precision(a.precision)
torch.manual_seed(0)
fine = not a.no_fine and C == 128
model = coocc_b200.HotPath(coocc_b200.model_cfg(C, K, fine=fine, grid=cfg["grid"]), C).to(dev).train()
model.pts_bbox_head.fine_select = a.select
inp = S.make_inputs(a.workload)
d = {k: v.to(dev) for k, v in inp.items()}
occ = S.make_gt_occ(cfg["grid"], 2, 0).to(dev)
img_feats = tr = None
if fine:
    img_feats = S.make_img_feats(cfg["cams"], cfg["fH"], cfg["fW"], 0).to(dev)
    tr = tuple(t.to(dev) if torch.is_tensor(t) else t for t in S.make_transform(cfg["cams"], cfg["fH"], cfg["fW"], 0))
params = [p for p in model.parameters() if p.requires_grad]
opt = None
if not a.no_opt:
    arena = GradArena(params)
    opt = FusedAdamW(params, lr=1e-4, weight_decay=0.01, shadow=(a.precision == "bf16"), arena=arena,
                     param_mults=norm_decay_mults(model, 0.0), max_norm=5.0)


def step():
    if opt is not None:
        opt.zero_grad()
    else:
        model.zero_grad(set_to_none=True)
    CF.begin_step()
    CF.zero_pool_begin(dev)
    losses, _, _ = model.forward_train(d["img_voxel_feats"], d["pts_voxel_feats"], d["geom"], d["gt_depth"], d["gt_img"], occ,
                                       img_feats, tr)
    sum(losses.values()).backward()
    CF.zero_pool_end()
    if opt is not None:
        opt.step()


step()
step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
