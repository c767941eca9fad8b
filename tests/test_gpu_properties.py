"""Full-size checks through size-independent properties (no CPU oracle at these sizes):
FPS max-min monotonicity, exact top-K against a brute-force GPU search, compaction inverses,
conv adjointness (<conv(x), y> = <x, dgrad(y)> = <w, wgrad(x, y)>) and linearity, BatchNorm
moments, compositing bounds.  Sizes: the r50 working grid (100x100x8, C=128) and one
north-star layer (200x200x16)."""
import ctypes

import pytest
import torch

import coocc_b200
from coocc_b200 import _lib
from coocc_b200 import functional as CF
from coocc_b200 import synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _coords(list_, n, Y, Z):
    v = list_[:n].long()
    return torch.stack([v // (Y * Z), (v // Z) % Y, v % Z], 1)


@pytest.fixture(scope="module")
def r50_state():
    cfg = S.CONFIGS["r50"]
    img, pts = S.make_voxel_feats(cfg["grid"], cfg["C"], cfg["p_img"], cfg["p_pts"], seed=3)
    cat, st = CF.gsf_index(img.to(DEV), pts.to(DEV), cfg["K"], want_parts=True)
    torch.cuda.synchronize()
    return cfg, img, pts, cat, st


def test_compaction_is_sorted_and_inverse(r50_state):
    cfg, img, pts, cat, st = r50_state
    X, Y, Z = cfg["grid"]
    for i, t in enumerate((img, pts)):
        n = st["n_img"] if i == 0 else st["n_pts"]
        lst = st["lists"][i, :n].long()
        assert (lst[1:] > lst[:-1]).all()
        occ = (t.to(DEV).sum(1) != 0).reshape(-1)
        assert n == int(occ.sum())
        assert torch.equal(lst, occ.nonzero().flatten())
        rank = st["ranks"][i].long()
        assert torch.equal(rank[lst], torch.arange(n, device=DEV))
        assert (rank[~occ] == -1).all()


def test_fps_max_min_distance_is_monotone(r50_state):
    cfg, _, _, _, st = r50_state
    X, Y, Z = cfg["grid"]
    for name, qi in (("A", 1), ("B", 0)):
        n = st["n_pts"] if name == "A" else st["n_img"]
        c = _coords(st["lists"][qi], n, Y, Z).float()
        rep = st[name]["rep_idx"].long()
        assert rep[0] == 0 and rep.unique().numel() == rep.numel() and int(rep.max()) < n
        p = c[rep]
        d = torch.cdist(p, p).pow(2).round()
        # distance of sample j to the closest earlier sample
        mask = torch.tril(torch.ones_like(d), -1).bool()
        dmin = torch.where(mask, d, torch.full_like(d, 1e9)).min(1).values[1:]
        assert (dmin[1:] <= dmin[:-1]).all(), "FPS max-min distances must be non-increasing"
        # and each pick was a farthest point: no point is farther from the first j samples than sample j
        for j in (1, 7, 100, 1000, 2047):
            dj = torch.cdist(c, p[:j]).pow(2).round().min(1).values
            assert dj.max() == dmin[j - 1]


def test_topk_matches_brute_force(r50_state):
    cfg, _, _, _, st = r50_state
    X, Y, Z = cfg["grid"]
    K = cfg["K"]
    for name, qi, ki in (("A", 1, 0), ("B", 0, 1)):
        nq = st["n_pts"] if name == "A" else st["n_img"]
        nk = st["n_img"] if name == "A" else st["n_pts"]
        q = _coords(st["lists"][qi], nq, Y, Z)
        k = _coords(st["lists"][ki], nk, Y, Z)
        rep = q[st[name]["rep_idx"].long()]
        d2 = (rep[:, None, :] - k[None, :, :]).pow(2).sum(-1)            # [2048, nk] exact ints
        key = d2 * nk + torch.arange(nk, device=DEV)[None, :]             # (d2 asc, index asc)
        best = key.topk(K, dim=1, largest=False).values
        bd2, bidx = best // nk, best % nk
        exp_idx = torch.where(bd2 <= 176, bidx, torch.full_like(bidx, -1))
        assert torch.equal(st[name]["topk_idx"].long(), exp_idx)
        assert torch.equal(st[name]["topk_d2"].long(), torch.where(bd2 <= 176, bd2, torch.full_like(bd2, -1)))


def test_assignment_winner_is_valid_and_in_range(r50_state):
    cfg, _, _, _, st = r50_state
    X, Y, Z = cfg["grid"]
    for name, qi in (("A", 1), ("B", 0)):
        nq = st["n_pts"] if name == "A" else st["n_img"]
        q = _coords(st["lists"][qi], nq, Y, Z)
        d = st[name]
        rep = q[d["rep_idx"].long()]
        for k in range(cfg["K"]):
            w = d["winner"][k, :nq].long()
            ok = w >= 0
            assert (d["topk_idx"][w[ok], k] >= 0).all()
            dist2 = (q[ok] - rep[w[ok]]).pow(2).sum(1)
            assert (dist2 < 36).all()
            # a query inside the ball of a valid representative whose ball is not full cannot stay unassigned
            grp = d["group"].long()
            full = (grp[:, -1] != grp[:, 0]) | False
        # group rows are ascending until the padding starts and contain the representative itself
        g = d["group"].long()
        assert (g[:, 0] <= d["rep_idx"].long()).all()


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
def test_conv_adjoint_and_linearity_full_size(prec):
    """<conv(x; w), y> = <x, dgrad(y; w)> = <w, wgrad(x, y)> on a north-star layer."""
    L = _lib.lib()
    X, Y, Z, Cin, Cout = 200, 200, 16, 128, 128
    dt = 1 if prec == "bf16" else 0
    tdt = torch.bfloat16 if dt == 1 else torch.float32
    V = X * Y * Z
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(V, Cin, device=DEV, generator=g).to(tdt)
    x2 = torch.randn(V, Cin, device=DEV, generator=g).to(tdt)
    w = (torch.randn(Cout, 27 * Cin, device=DEV, generator=g) * 0.02).to(tdt)
    yb = torch.randn(V, Cout, device=DEV, generator=g).to(tdt)
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = _lib.ConvDesc(X, Y, Z, Cin, Cout, 3, 1, dt, Cin, Cout)

    def fwd(inp):
        out = torch.empty(V, Cout, device=DEV)
        _lib.check(L.coocc_conv3d_fwd(ctypes.byref(d), inp.data_ptr(), w.data_ptr(), out.data_ptr(), Cout, None, 0, None, st()), "fwd")
        return out

    y = fwd(x)
    dx = torch.empty(V, Cin, device=DEV)
    _lib.check(L.coocc_conv3d_dgrad(ctypes.byref(d), yb.data_ptr(), w.data_ptr(), dx.data_ptr(), Cin, st()), "dgrad")
    dw = torch.zeros(Cout, 27 * Cin, device=DEV)
    _lib.check(L.coocc_conv3d_wgrad(ctypes.byref(d), x.data_ptr(), yb.data_ptr(), dw.data_ptr(), st()), "wgrad")
    torch.cuda.synchronize()
    a = (y.double() * yb.double()).sum()
    b = (x.double() * dx.double()).sum()
    c = (w.double() * dw.double()).sum()
    scale = (y.double().norm() * yb.double().norm())
    tol = 2e-3
    assert abs(a - b) / scale < tol and abs(a - c) / scale < tol, (a.item(), b.item(), c.item())
    # linearity in x (exact inputs: x + x2 rounded once to the operand type)
    xs = (x.float() + x2.float()).to(tdt)
    ys = fwd(xs)
    y12 = fwd(x) + fwd(x2)
    rel = (ys - y12).norm() / y12.norm()
    assert rel < (2e-2 if prec == "bf16" else 2e-3), rel.item()


def test_batchnorm_moments_full_size():
    coocc_b200.set_precision("tf32")
    import torch.nn as nn
    from coocc_b200 import modules as M
    conv = nn.Conv3d(128, 128, 1, bias=False).to(DEV)
    bn = nn.BatchNorm3d(128).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    x = torch.randn(100 * 100 * 8, 128, device=DEV)
    y, _ = M.conv_bn_act(x, (100, 100, 8), conv, bn, relu=False)
    m, v = y.mean(0), y.var(0, unbiased=False)
    assert (m - bn.bias).abs().max() < 1e-3
    assert ((v.sqrt() - bn.weight).abs() / bn.weight).max() < 1e-3


def test_render_bounds_full_size():
    cfg = S.CONFIGS["r50"]
    geom = S.make_geom(cfg["grid"], cfg["cams"], cfg["fH"], cfg["fW"], cfg["D"]).to(DEV)
    T = 100 * 100 * 8
    g = torch.Generator(device=DEV).manual_seed(1)
    tab = torch.randn(T, 4, device=DEV, generator=g)
    tab[:, 3] = tab[:, 3].abs() * 0.2
    rgb_map, depth_map = CF.composite(tab, geom[0], cfg["grid"])
    torch.cuda.synchronize()
    assert torch.isfinite(rgb_map).all() and torch.isfinite(depth_map).all()
    assert (rgb_map >= 0).all() and (rgb_map <= 1.0 + 1e-5).all()      # sum of weights <= 1, rgb in (0,1)
    assert (depth_map >= 0).all() and (depth_map <= cfg["D"] * (1 + 1e-5)).all()
    # zero density everywhere except the last interval (dist = 1e10 but sigma = 0): no contribution
    tab0 = tab.clone()
    tab0[:, 3] = 0
    r0, d0 = CF.composite(tab0, geom[0], cfg["grid"])
    assert r0.abs().max() == 0 and d0.abs().max() == 0
