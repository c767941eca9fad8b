"""Times coocc_gsf_fps alone for cluster sizes / exchange variants."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coocc_b200
from coocc_b200 import _lib
L = _lib.lib()
dev = "cuda"
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for (X, Y, Z) in ((100, 100, 8), (200, 200, 16)):
    V = X * Y * Z
    g = torch.Generator().manual_seed(0)
    lists, counts = [], []
    for p in (0.6, 0.15):
        m = (torch.rand(V, generator=g) < p).nonzero().flatten().int().to(dev)
        lists.append(m); counts.append(torch.tensor([m.numel()], dtype=torch.int32, device=dev))
    outs = [torch.empty(2048, dtype=torch.int32, device=dev) for _ in range(2)]
    nmax = max(int(c.item()) for c in counts)
    ref = None
    for cs in (0, 4, 8, 16):
        for flags in (0, 1, 2, 3):
            L.coocc_gsf_fps_tune(cs, flags)
            rc = L.coocc_gsf_fps(P(lists[0]), P(counts[0]), P(outs[0]), P(lists[1]), P(counts[1]), P(outs[1]), nmax, 2048, Y, Z, st())
            torch.cuda.synchronize()
            if rc != 0:
                print("grid", (X, Y, Z), "cs", cs, "flags", flags, "rc", rc); continue
            if ref is None:
                ref = [o.clone() for o in outs]
            ok = all(torch.equal(a, b) for a, b in zip(ref, outs))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                L.coocc_gsf_fps(P(lists[0]), P(counts[0]), P(outs[0]), P(lists[1]), P(counts[1]), P(outs[1]), nmax, 2048, Y, Z, st())
            e1.record(); torch.cuda.synchronize()
            print("grid %s N=%d/%d cs=%d flags=%d: %.3f ms  same=%s" % ((X, Y, Z), int(counts[0]), int(counts[1]), cs, flags, e0.elapsed_time(e1) / 3, ok), flush=True)
L.coocc_gsf_fps_tune(0, 0)
