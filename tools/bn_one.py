"""A few launches of the three BatchNorm kernels at one shape (for ncu captures)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib()
dev = "cuda"
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
V, C, bf = 640000, 256, 1
x = torch.randn(V, C, device=dev).to(torch.bfloat16)
out = torch.empty_like(x); dout = torch.randn(V, C, device=dev).to(torch.bfloat16); dx = torch.empty_like(x)
mi = torch.stack([torch.zeros(C, device=dev), torch.ones(C, device=dev)]).contiguous()
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); sums = torch.zeros(2, C, device=dev)
for _ in range(2):
    L.coocc_bn_act_fwd(P(x), C, V, C, P(mi), P(g), P(b), None, 0, 1, P(out), C, bf, st())
    L.coocc_bn_act_bwd_reduce(P(dout), C, None, 0, P(x), C, V, C, P(mi), P(g), P(b), 1, P(sums), bf, st())
    L.coocc_bn_act_bwd_apply(P(dout), C, None, 0, P(x), C, V, C, P(mi), P(g), P(b), 1, P(sums), V, P(dx), C, bf, None, 0, st())
    out.copy_(x)
torch.cuda.synchronize()
