"""2-rank check of the SyncBN path (run with torchrun --nproc-per-node 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.nn as nn
import coocc_b200
from coocc_b200 import functional as CF, modules as M
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
coocc_b200.set_precision("fp32")
torch.manual_seed(0)
conv = nn.Conv3d(32, 64, 3, 1, 1, bias=False).to(dev)
bn = nn.BatchNorm3d(64).to(dev)
with torch.no_grad():
    bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
dims = (12, 10, 4)
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(1, 32, *dims, generator=g).to(dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
gy = torch.randn(dims[0] * dims[1] * dims[2], 64, generator=g).to(dev)
x2d, d = CF.to_cl2d(x)
out, _ = M.conv_bn_act(x2d, d, conv, bn, relu=True)
(out * gy).sum().backward()
# reference: gather both ranks' inputs, run torch modules on the 2-sample batch (plain BN over both)
xs = [torch.empty_like(x) for _ in range(world)]; gys = [torch.empty_like(gy) for _ in range(world)]
dist.all_gather(xs, x.detach()); dist.all_gather(gys, gy)
conv_r = nn.Conv3d(32, 64, 3, 1, 1, bias=False).to(dev); conv_r.load_state_dict(conv.state_dict())
bn_r = nn.BatchNorm3d(64).to(dev)
with torch.no_grad():
    bn_r.weight.copy_(bn.weight); bn_r.bias.copy_(bn.bias)
torch.backends.cudnn.allow_tf32 = False
xb = torch.cat(xs, 0).contiguous().requires_grad_(True)
yb = torch.relu(bn_r(conv_r(xb)))
gyb = torch.stack([t.reshape(*dims, 64).permute(3, 0, 1, 2) for t in gys], 0)
(yb * gyb).sum().backward()
ref_out = yb[rank].permute(1, 2, 3, 0).reshape(-1, 64)
ref_dx = xb.grad[rank]
e_out = ((out - ref_out).abs().max() / ref_out.abs().max()).item()
e_dx = ((x.grad[0] - ref_dx).norm() / ref_dx.norm()).item()
e_rm = ((bn.running_mean - bn_r.running_mean).abs().max()).item()
# weight grads: DDP would average the per-rank grads; sum over ranks equals the big-batch grad
gw = conv.weight.grad.clone(); dist.all_reduce(gw)
e_dw = ((gw - conv_r.weight.grad).norm() / conv_r.weight.grad.norm()).item()
ggam = bn.weight.grad.clone(); dist.all_reduce(ggam)
e_dg = ((ggam - bn_r.weight.grad).norm() / bn_r.weight.grad.norm()).item()
print("rank %d syncbn: out %.2e dx %.2e running_mean %.2e dW(sum over ranks) %.2e dgamma %.2e" % (rank, e_out, e_dx, e_rm, e_dw, e_dg), flush=True)
assert max(e_out, e_dx, e_dw, e_dg) < 1e-3 and e_rm < 1e-5
dist.barrier(); dist.destroy_process_group()
