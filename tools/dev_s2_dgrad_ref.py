"""Reference derivation (torch CPU) for the planned stride-2 data-gradient kernel.

Today the dgrad of a 3x3x3 stride-2 pad-1 convolution zero-inserts dy onto the input lattice (coocc_dilate2) and runs
a stride-1 dgrad over it: 8x the forward's FLOPs, 7/8 of them multiply inserted zeros.  Decomposing the input lattice
into its 8 parity classes p in {0,1}^3 removes them: with y[o] = sum_k W[k] x[2o + k - 1],
    dx[2o' + p] = sum over taps k with (p + 1 - k) even of  W[k]^T dy[o' + (p + 1 - k) / 2]
per axis:  p = 0 -> k = 1 (offset 0);   p = 1 -> k = 0 (offset +1) and k = 2 (offset 0),
so class p needs 2^(px+py+pz) taps (1, 2, 2, 4, 2, 4, 4, 8 = 27 in total = the forward's work), each class is a small
stride-1 correlation over the half-resolution dy grid whose result is written to the stride-2 sub-lattice of dx.
This script checks the identity against autograd for even and odd extents (the cases of the 200->100->50->25->13 pyramid).
"""
import itertools

import torch
import torch.nn.functional as F

TAPS = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}          # parity -> [(kernel index k, dy offset)]


def dgrad_s2_by_parity(dy, w, in_dims):
    """dy [1,Cout,oX,oY,oZ], w [Cout,Cin,3,3,3] -> dx [1,Cin,X,Y,Z] of conv3d(x, w, stride=2, padding=1)."""
    Cout, Cin = w.shape[:2]
    X, Y, Z = in_dims
    oX, oY, oZ = dy.shape[2:]
    dx = torch.zeros(1, Cin, X, Y, Z, dtype=dy.dtype)
    pad = F.pad(dy, (0, 1, 0, 1, 0, 1))                                  # offset +1 may run past the edge: zeros
    for px, py, pz in itertools.product((0, 1), repeat=3):
        nx, ny, nz = (X - px + 1) // 2, (Y - py + 1) // 2, (Z - pz + 1) // 2     # positions 2o'+p < extent
        acc = torch.zeros(1, Cin, nx, ny, nz, dtype=dy.dtype)
        for (kx, ox), (ky, oy), (kz, oz) in itertools.product(TAPS[px], TAPS[py], TAPS[pz]):
            g = pad[:, :, ox:ox + nx, oy:oy + ny, oz:oz + nz]                      # dy[o' + off]
            acc += torch.einsum("bodhw,oi->bidhw", g, w[:, :, kx, ky, kz])        # W[k]^T dy
        dx[:, :, px::2, py::2, pz::2] = acc
    return dx


def main():
    torch.manual_seed(0)
    for dims in [(8, 8, 4), (9, 7, 5), (25, 25, 2), (13, 13, 1)]:
        x = torch.randn(1, 6, *dims, dtype=torch.float64, requires_grad=True)
        w = torch.randn(10, 6, 3, 3, 3, dtype=torch.float64)
        y = F.conv3d(x, w, stride=2, padding=1)
        dy = torch.randn_like(y)
        y.backward(dy)
        mine = dgrad_s2_by_parity(dy, w, dims)
        err = (mine - x.grad).abs().max().item()
        taps = sum(2 ** sum(p) for p in itertools.product((0, 1), repeat=3))
        print("in %-12s out %-12s  max |dx - autograd| = %.2e   taps per 2x2x2 cell: %d" % (dims, tuple(y.shape[2:]), err, taps))
        assert err < 1e-10


if __name__ == "__main__":
    main()
