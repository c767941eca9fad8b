"""optim.FusedAdamW (csrc/adamw.cu) on the GPU: same update as torch.optim.AdamW, bf16 shadow == bf16(p) in the
parameter's own memory order, and correct under CUDA graphs whose gradients live at different addresses (each capture
gets its own pointer table; an eager step in between must not disturb the replays)."""
import pytest
import torch

from coocc_b200.optim import FusedAdamW

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = [(64, 32, 3, 3, 3), (17, 64, 1, 1, 1), (256,), (128, 256), (40000,), (3,)]
    ps = []
    for i, s in enumerate(shapes):
        t = torch.randn(*s, generator=g).to(DEV)
        if len(s) == 5:
            t = t.contiguous(memory_format=torch.channels_last_3d)
        ps.append(torch.nn.Parameter(t))
    return ps


def _grads(ps, seed):
    g = torch.Generator(device="cpu").manual_seed(100 + seed)
    out = []
    for p in ps:
        t = torch.randn(*p.shape, generator=g).to(DEV) * 0.1
        if p.dim() == 5:
            t = t.contiguous(memory_format=torch.channels_last_3d)
        out.append(t)
    return out


def test_fused_adamw_matches_torch_and_keeps_the_shadow():
    pa, pb = _params(), _params()
    oa = FusedAdamW(pa, lr=1e-2, weight_decay=0.05, shadow=True)
    ob = torch.optim.AdamW(pb, lr=1e-2, weight_decay=0.05)
    for it in range(5):
        for p, q, g in zip(pa, pb, _grads(pa, it)):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), (p - q).abs().max().item()
        flat = p.detach().permute(0, 2, 3, 4, 1).reshape(-1) if p.dim() == 5 else p.detach().reshape(-1)
        assert torch.equal(p._coocc_bf16.reshape(-1), flat.to(torch.bfloat16))
        assert p._coocc_bf16_version == p._version


def test_fused_adamw_two_graphs_and_an_eager_step_share_one_optimizer():
    """the pointer table uploaded inside a capture is re-read from pinned memory at every replay; a second
    capture / an eager step must not overwrite it."""
    pa, pb = _params(1), _params(1)
    oa = FusedAdamW(pa, lr=1e-2, weight_decay=0.0, shadow=True)
    ob = torch.optim.AdamW(pb, lr=1e-2, weight_decay=0.0)
    src = [torch.zeros_like(p) for p in pa]      # static gradient sources the graphs read

    def captured():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for p, t in zip(pa, src):
                p.grad = t * 1.0                  # graph-pool address, differs per capture
            oa.step()
        return g

    def feed(it):
        gs = _grads(pa, it)
        for t, g in zip(src, gs):
            t.copy_(g)
        for q, g in zip(pb, gs):
            q.grad = g.clone()
        ob.step()

    # warm (lazy init outside capture)
    feed(0)
    for p, t in zip(pa, src):
        p.grad = t.clone()
    oa.step()
    g1 = captured()           # capture executes nothing
    g2 = captured()
    seq = [g1, g2, None, g1, g2, g1]
    for it, g in enumerate(seq, start=1):
        feed(it)
        if g is None:
            for p, t in zip(pa, src):
                p.grad = t.clone()
            oa.step()
        else:
            g.replay()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=5e-6, atol=5e-7), (p - q).abs().max().item()


def test_arena_norm_decay_and_grad_clip_match_torch():
    """FusedAdamW on a ddp.GradArena (gradients accumulated in place, cleared by the kernel) with the reference's
    optimizer recipe: norm_decay_mult = 0 for normalisation parameters (coocc_multi_r50_256x704.py:276), grad_clip
    max_norm (:279), and an lr-schedule multiplier."""
    import torch.nn as nn
    from coocc_b200.ddp import GradArena
    from coocc_b200.optim import norm_decay_mults

    def build():
        torch.manual_seed(3)
        net = nn.Sequential(nn.Conv3d(4, 8, 3, padding=1, bias=False), nn.BatchNorm3d(8), nn.ReLU(),
                            nn.Conv3d(8, 4, 1)).to(DEV)
        net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last_3d)
        return net

    na, nb = build(), build()
    arena = GradArena(list(na.parameters()))
    oa = FusedAdamW(list(na.parameters()), lr=1e-2, weight_decay=0.1, shadow=False, arena=arena,
                    param_mults=norm_decay_mults(na, 0.0), max_norm=0.5)
    norm_p = [p for m in nb.modules() if isinstance(m, nn.BatchNorm3d) for p in m.parameters()]
    rest = [p for p in nb.parameters() if all(p is not q for q in norm_p)]
    ob = torch.optim.AdamW([dict(params=rest), dict(params=norm_p, weight_decay=0.0)], lr=1e-2, weight_decay=0.1)
    for it in range(4):
        if it == 2:
            oa.set_lr_scale(0.1)
            for g in ob.param_groups:
                g["lr"] = 1e-3
        x = torch.randn(2, 4, 6, 5, 4, generator=torch.Generator().manual_seed(it)).to(DEV)
        oa.zero_grad()
        ob.zero_grad()
        (na(x).square().mean() * 50).backward()
        (nb(x).square().mean() * 50).backward()
        assert all(p.grad.data_ptr() == p._coocc_grad.data_ptr() for p in na.parameters())
        total = torch.nn.utils.clip_grad_norm_(nb.parameters(), 0.5)
        oa.step()
        ob.step()
        assert abs(float(oa.grad_norm) - float(total)) < 1e-4 * float(total)
        assert float(arena.flat.abs().max()) == 0.0                   # cleared in the same pass
    for p, q in zip(na.parameters(), nb.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), (p - q).abs().max().item()
