"""Phase timing of fps_kernel rounds (needs a library built with -DCOOCC_FPS_TRACE: tools/fps_trace.sh)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
so = sys.argv[1]
L = ctypes.CDLL(so)
X, Y, Z = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (200, 200, 16)))
V = X * Y * Z
dev = "cuda"
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
buf = np.zeros(16 * 64 * 8, dtype=np.int64)
L.coocc_gsf_fps_trace(buf.ctypes.data_as(ctypes.c_void_p))
T = buf.reshape(16, 64, 8)
print("grid", (X, Y, Z), "-- warp 0 of every CTA of job 0, cycles (median over 64 rounds)")
print(" cta   round  compute  barrier  cta-reduce  send  mbar wait  inbox-reduce  loop-tail")
for r in range(16):
    t = T[r]
    if t[0, 0] == 0:
        continue
    rnd = np.median(t[1:, 0] - t[:-1, 0])
    print(" %3d  %6.0f  %7.0f  %7.0f  %10.0f  %4.0f  %9.0f  %12.0f  %9.0f" % (
        r, rnd, np.median(t[:, 2] - t[:, 0]), np.median(t[:, 3] - t[:, 2]), np.median(t[:, 6] - t[:, 3]),
        np.median(t[:, 4] - t[:, 6]), np.median(t[:, 5] - t[:, 4]), np.median(t[:, 7] - t[:, 5]),
        np.median(t[1:, 0] - t[:-1, 7])))
w = T[:, :, 5] - T[:, :, 4]
ok = T[:, 0, 0] != 0
print("per round: min over CTAs of the mbarrier wait  median %.0f   (the critical CTA waits only for the exchange)" % np.median(w[ok].min(0)))
c = (T[:, :, 3] - T[:, :, 0])[ok]
print("per round: max over CTAs of compute+barrier     median %.0f" % np.median(c.max(0)))
