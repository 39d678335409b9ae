"""KAN classification head of the MedSSD_kan family -- host-side glue OUTSIDE the hot path (heads are plain
PyTorch), present so that BASELINE.json configs[3] (MedSSD_kan) can be built, loaded from a reference checkpoint and timed.

Mirrors the interface of the reference's `KANLayer` / `KansModule` (MedSSD_kan/MedSSD_kan.py:190-300, 475-501): the same
parameter names and shapes -- grid (in, G + 2k + 1) and mask (in, out) frozen, coef (in, out, G + k), scale_base, scale_sp
(in, out) -- and the same function
    y[b, o] = sum_i mask[i, o] * ( scale_base[i, o] * silu(x[b, i]) + scale_sp[i, o] * sum_g coef[i, o, g] * B_g^k(x[b, i]) ),
B_g^k the order-k B-spline basis on the (extended, uniform at init) knot vector grid[i].  Written from that definition: the
basis is built bottom-up by the Cox-de Boor recurrence, the initial coefficients by a least-squares fit of small noise."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def bspline_basis(x: torch.Tensor, grid: torch.Tensor, k: int) -> torch.Tensor:
    """x (batch, in), grid (in, T) knots -> (batch, in, T - k - 1) order-k B-spline basis values (Cox-de Boor)."""
    x = x.unsqueeze(-1)                                              # (batch, in, 1)
    g = grid.unsqueeze(0)                                            # (1, in, T)
    b = ((x >= g[..., :-1]) & (x < g[..., 1:])).to(x.dtype)          # order 0: indicator of the knot interval
    for p in range(1, k + 1):
        left = (x - g[..., :-(p + 1)]) / (g[..., p:-1] - g[..., :-(p + 1)])
        right = (g[..., p + 1:] - x) / (g[..., p + 1:] - g[..., 1:-p])
        b = left * b[..., :-1] + right * b[..., 1:]
    return torch.nan_to_num(b)                                       # degenerate (repeated) knots contribute nothing


class KANLayer(nn.Module):
    def __init__(self, in_dim=3, out_dim=2, num=5, k=3, noise_scale=0.1, scale_base_mu=0.0, scale_base_sigma=1.0, scale_sp=1.0,
                 base_fun=None, grid_range=(-1.0, 1.0), sp_trainable=True, sb_trainable=True, **kwargs):
        super().__init__()
        self.in_dim, self.out_dim, self.num, self.k = in_dim, out_dim, num, k
        h = (grid_range[1] - grid_range[0]) / num
        knots = torch.arange(-k, num + k + 1, dtype=torch.float32) * h + grid_range[0]           # uniform grid extended by k knots
        self.grid = nn.Parameter(knots[None, :].repeat(in_dim, 1), requires_grad=False)
        # initial spline = least-squares fit of small noise sampled at the interior knots
        xs = self.grid[:, k:-k].t().contiguous()                                                  # (num + 1, in)
        noise = (torch.rand(num + 1, in_dim, out_dim) - 0.5) * noise_scale / num
        basis = bspline_basis(xs, self.grid, k).permute(1, 0, 2)                                  # (in, num + 1, G + k)
        coef = torch.linalg.lstsq(basis, noise.permute(1, 0, 2)).solution                         # (in, G + k, out)
        self.coef = nn.Parameter(coef.permute(0, 2, 1).contiguous())                              # (in, out, G + k)
        self.mask = nn.Parameter(torch.ones(in_dim, out_dim), requires_grad=False)
        self.scale_base = nn.Parameter(scale_base_mu / math.sqrt(in_dim) +
                                       scale_base_sigma * (torch.rand(in_dim, out_dim) * 2 - 1) / math.sqrt(in_dim),
                                       requires_grad=sb_trainable)
        self.scale_sp = nn.Parameter(torch.ones(in_dim, out_dim) * scale_sp, requires_grad=sp_trainable)
        self.base_fun = base_fun if base_fun is not None else nn.SiLU()

    def forward(self, x):
        """x (batch, in) -> (y (batch, out), preacts, postacts, postspline) like the reference (the last three only feed its plots)."""
        spline = torch.einsum("big,iog->bio", bspline_basis(x, self.grid, self.k), self.coef)    # (batch, in, out)
        y = self.mask[None] * (self.scale_base[None] * self.base_fun(x)[:, :, None] + self.scale_sp[None] * spline)
        return y.sum(dim=1), x[:, None, :].expand(-1, self.out_dim, -1), y.permute(0, 2, 1), spline.permute(0, 2, 1)


class KansModule(nn.Module):
    """kan1 (C -> C) + BatchNorm1d + residual, then kan2 (C -> classes): MedSSD_kan/MedSSD_kan.py:475-501."""

    def __init__(self, in_channels, out_channels, num1=5, num2=5, num3=5, k1=3, k2=3, k3=3):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kan1 = KANLayer(in_dim=in_channels, out_dim=in_channels, num=num1, k=k1)
        self.kan2 = KANLayer(in_dim=in_channels, out_dim=out_channels, num=num2, k=k2)
        self.bn = nn.BatchNorm1d(in_channels)

    def forward(self, x):
        out = self.bn(self.kan1(x)[0]) + x
        return self.kan2(out)[0]
