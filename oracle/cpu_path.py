"""oracle/cpu_path.py -- TEST INFRASTRUCTURE: the reference's CPU data flow for the SS2D hot path.

Used only by bench.py's `cpu_baseline` leg / `--impl reference` arm and by tests.  It restates
`SS2D.forward_corev0` (reference MedMamba.py:386-424) and the merge at MedMamba.py:476-477 with
plain PyTorch CPU ops -- stack / transpose / flip / cat / einsum exactly as the reference does --
and binds `selective_scan_fn` to the C restatement of `selective_scan_ref` (oracle/sscan_oracle.c,
OpenMP over rows) instead of the reference's pure-PyTorch loop, whose autograd backward needs
minutes per image.  kind = "port".
"""
from __future__ import annotations

import numpy as np
import torch

import oracle


class OracleSelectiveScanFn(torch.autograd.Function):
    """selective_scan_fn on CPU tensors through the C oracle (fp32 instantiation, all host threads)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, delta_bias, delta_softplus):
        out, _ = oracle.sscan_fwd(u, delta, A, B, C, D=D, delta_bias=delta_bias,
                                  delta_softplus=delta_softplus, precision="f32")
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias)
        ctx.softplus = delta_softplus
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, D, delta_bias = ctx.saved_tensors
        g = oracle.sscan_bwd(u, delta, A, B, C, D=D, delta_bias=delta_bias, delta_softplus=ctx.softplus,
                             dout=dout.contiguous(), precision="f32")
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        return t(g["du"]), t(g["ddelta"]), t(g["dA"]), t(g["dB"]), t(g["dC"]), t(g["dD"]), t(g["ddelta_bias"]), None


def ss2d_core_cpu(m, x):
    """x (B, D, H, W) -> (B, H, W, D); `m` supplies the SS2D parameters (reference names)."""
    B, D, H, W = x.shape
    L = H * W
    K = 4
    x_hwwh = torch.stack([x.view(B, -1, L), x.transpose(2, 3).contiguous().view(B, -1, L)], dim=1).view(B, 2, -1, L)
    xs = torch.cat([x_hwwh, torch.flip(x_hwwh, dims=[-1])], dim=1)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs.view(B, K, -1, L), m.x_proj_weight)
    dts, Bs, Cs = torch.split(x_dbl, [m.dt_rank, m.d_state, m.d_state], dim=2)
    dts = torch.einsum("bkrl,kdr->bkdl", dts.view(B, K, -1, L), m.dt_projs_weight)
    out_y = OracleSelectiveScanFn.apply(
        xs.float().reshape(B, -1, L), dts.contiguous().float().view(B, -1, L),
        -torch.exp(m.A_logs.float()).view(-1, m.d_state), Bs.float().contiguous(), Cs.float().contiguous(),
        m.Ds.float().view(-1), m.dt_projs_bias.float().view(-1), True).view(B, K, -1, L)
    inv_y = torch.flip(out_y[:, 2:4], dims=[-1]).view(B, 2, -1, L)
    wh_y = torch.transpose(out_y[:, 1].view(B, -1, W, H), 2, 3).contiguous().view(B, -1, L)
    invwh_y = torch.transpose(inv_y[:, 1].view(B, -1, W, H), 2, 3).contiguous().view(B, -1, L)
    y = out_y[:, 0] + inv_y[:, 0] + wh_y + invwh_y
    return torch.transpose(y, 1, 2).contiguous().view(B, H, W, -1)


def bind_cpu_core(model):
    """Point every SS2D block of a (CPU) VSSM at the CPU data flow above."""
    from functools import partial
    n = 0
    for mod in model.modules():
        if hasattr(mod, "forward_core") and hasattr(mod, "x_proj_weight"):
            mod.forward_core = partial(ss2d_core_cpu, mod)
            n += 1
    return n
