"""One coocc_gsf_fps launch at a given grid size (for ncu captures of fps_kernel)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib()
dev = "cuda"
X, Y, Z = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (200, 200, 16)))
V = X * Y * Z
g = torch.Generator().manual_seed(0)
lists, counts = [], []
for p in (0.6, 0.15):
    m = (torch.rand(V, generator=g) < p).nonzero().flatten().int().to(dev)
    lists.append(m); counts.append(torch.tensor([m.numel()], dtype=torch.int32, device=dev))
outs = [torch.empty(2048, dtype=torch.int32, device=dev) for _ in range(2)]
nmax = max(int(c.item()) for c in counts)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    L.coocc_gsf_fps(P(lists[0]), P(counts[0]), P(outs[0]), P(lists[1]), P(counts[1]), P(outs[1]), nmax, 2048, Y, Z,
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
