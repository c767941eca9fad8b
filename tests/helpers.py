"""Shared helpers for the parity tests."""
import numpy as np
import torch

STRIDE = 97   # must match oracle/make_golden.py


def sample(t, g=None):
    """strided sample of a tensor; `g` = the fixture it is compared with (its `stride` entry, default STRIDE)."""
    stride = int(g["stride"]) if (g is not None and "stride" in getattr(g, "files", g)) else STRIDE
    return t.detach().float().cpu().reshape(-1)[::stride].numpy()


def stats(t):
    t = t.detach().double().cpu()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'rel' of the north-star's 1e-3 bound)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
