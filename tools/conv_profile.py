"""Per-launch table of the tensor-core conv kernel inside one training step (CUDA events)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import functional as CF, synthetic as S
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="northstar"); ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
cfg = S.CONFIGS[a.workload]; C, K = cfg["C"], cfg["K"]; dev = "cuda"
coocc_b200.set_precision(a.precision)
torch.manual_seed(0)
model = coocc_b200.HotPath(coocc_b200.model_cfg(C, K, fine=False), C).to(dev).train()
inp = S.make_inputs(a.workload); d = {k: v.to(dev) for k, v in inp.items()}
X, Y, Z = cfg["grid"]; occ = torch.randint(0, 17, (1, X, Y, Z), device=dev)
def step():
    model.zero_grad(set_to_none=True)
    losses, _, _ = model.forward_train(d["img_voxel_feats"], d["pts_voxel_feats"], d["geom"], d["gt_depth"], d["gt_img"], occ)
    sum(losses.values()).backward()
for _ in range(2): step()
torch.cuda.synchronize()
CF.PROFILE = []
# keep the GPU busy while the host issues the step: otherwise the event pairs around the small launches measure the
# host's launch overhead (~50 us floor), not the kernels
CF.PROFILE_AHEAD_MS = 40
step(); torch.cuda.synchronize()
rows = [(a0.elapsed_time(b0), f, t) for a0, b0, f, t in CF.PROFILE if not t.startswith("hbm:")]
CF.PROFILE = None
tot = sum(r[0] for r in rows); fl = sum(r[1] for r in rows)
print("%s %s: %d conv launches, %.2f ms, %.0f TFLOP/s" % (a.workload, a.precision, len(rows), tot, fl / tot / 1e9))
agg = {}
for ms, f, t in rows:
    m, ff, n = agg.get(t, (0, 0, 0)); agg[t] = (m + ms, ff + f, n + 1)
for t, (ms, f, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print("%8.3f ms %5.1f%% x%d %7.0f TF/s  %s" % (ms, 100 * ms / tot, n, f / ms / 1e9, t))
