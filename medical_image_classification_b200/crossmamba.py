"""CrossMamba two-branch SSD mixer -- B200 mirror of the reference module `CrossMamba`
(reference CrossMamba/CrossMamba_fusion_2b2.py:54-388): same constructor arguments,
parameter names and shapes (`in_proj, skip_in_proj, xs_in_proj, BCdts_in_proj, conv2d, xs_conv2d, BCdts_conv2d,
dt_bias (4, nheads), A_logs (4*nheads), Ds, norm.weight, out_proj`), so reference checkpoints load with strict=True,
and the same `forward(u1, u2, u2_cat_u1, u1_cat_u2) -> (out1, out2)`.

Each branch scans ITS OWN input sequence (xs, through `xs_in_proj` / `xs_conv2d`) with B, C and dt computed from the
mixed tensor of the OTHER ordering (`BCdts_in_proj` / `BCdts_conv2d`): two `mamba_chunk_scan_combined` calls per
forward on the four-direction cross-scan layout, served by libb200ssm's SSD kernels (ssd_combined.py).  `in_proj` and
`conv2d` exist in the reference's state_dict but are never used by its forward (:254-388); they are kept for
checkpoint compatibility only.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .cross import cross_scan4, ssd_merge4
from .ssd_combined import RMSNormGated, mamba_chunk_scan_combined


class CrossMamba(nn.Module):
    def __init__(self, d_model, d_state=128, d_conv=3, expand=2, headdim=64, d_ssm=None, ngroups=1,
                 A_init_range=(1, 16), D_has_hdim=False, rmsnorm=True, norm_before_gate=False, dt_rank="auto",
                 dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 dt_limit=(0.0, float("inf")), dropout=0.0, conv_bias=True, bias=False, chunk_size=256,
                 use_mem_eff_path=True, layer_idx=None, process_group=None, sequence_parallel=True,
                 device=None, dtype=None, **kwargs):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        if process_group is not None:
            raise NotImplementedError("CrossMamba: tensor parallelism is dead code in the reference (process_group=None)")
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.headdim = headdim
        self.d_ssm = self.d_inner if d_ssm is None else d_ssm
        self.ngroups = ngroups
        assert self.d_ssm % headdim == 0
        self.nheads = self.d_ssm // headdim
        self.D_has_hdim, self.rmsnorm, self.norm_before_gate = D_has_hdim, rmsnorm, norm_before_gate
        self.dt_limit, self.chunk_size = dt_limit, chunk_size
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank

        gn = ngroups * d_state
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner + 2 * gn + self.nheads, bias=bias, **fk)     # unused by forward (:120-121)
        self.skip_in_proj = nn.Linear(d_model, 2 * self.d_inner - self.d_ssm, bias=bias, **fk)         # (z0, x0, z)  (:128-129)
        self.xs_in_proj = nn.Linear(d_model, self.d_ssm, bias=bias, **fk)                               # (:131)
        self.BCdts_in_proj = nn.Linear(d_model, 2 * gn + self.nheads, bias=bias, **fk)                  # (:133-134)
        conv_dim = self.d_ssm + 2 * gn + self.nheads
        conv = lambda c: nn.Conv2d(c, c, groups=c, bias=conv_bias, kernel_size=d_conv, padding=(d_conv - 1) // 2, **fk)
        self.conv2d = conv(conv_dim)                   # unused by forward (:138-146)
        self.xs_conv2d = conv(self.d_ssm)
        self.BCdts_conv2d = conv(2 * gn + self.nheads)
        self.act = nn.SiLU()

        dt = torch.exp(torch.rand(self.nheads, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = torch.clamp(dt, min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))
        self.dt_bias = nn.Parameter(torch.stack([inv_dt] * 4, dim=0))                 # (4, nheads)
        self.dt_bias._no_weight_decay = True
        A = torch.empty(self.nheads, dtype=torch.float32, device=device).uniform_(*A_init_range)
        self.A_logs = nn.Parameter(torch.log(A).to(dtype=dtype).repeat(4))            # (4*nheads)
        self.A_logs._no_weight_decay = True
        self.Ds = nn.Parameter(torch.ones(4 * (self.d_ssm if D_has_hdim else self.nheads), device=device))
        self.Ds._no_weight_decay = True
        if rmsnorm:
            self.norm = RMSNormGated(self.d_ssm, eps=1e-5, norm_before_gate=norm_before_gate,
                                     group_size=self.d_ssm // ngroups, **fk)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else None

    def _scan_inputs(self, xs, bcdts, B, L):
        """Four-direction cross-scan of [x | B | C | dt] (reference :280-299) as (b, l, ...) views with L stride 1."""
        gn = self.ngroups * self.d_state
        # one pass per component (csrc/cross.cu::cross_scan4_kernel), channel slices read in place; CPU tensors raise
        x = cross_scan4(xs)
        Bm, Cm, dt = (cross_scan4(t) for t in torch.split(bcdts, [gn, gn, self.nheads], dim=1))
        x = x.float().reshape(B, -1, L).permute(0, 2, 1).unflatten(2, (-1, self.headdim))     # (B, L, 4*nheads, P)
        Bm = Bm.float().reshape(B, -1, L).permute(0, 2, 1).unflatten(2, (self.ngroups, -1))   # (B, L, G, 4*N)
        Cm = Cm.float().reshape(B, -1, L).permute(0, 2, 1).unflatten(2, (self.ngroups, -1))
        dt = dt.float().reshape(B, -1, L).permute(0, 2, 1)                                    # (B, L, 4*nheads)
        return x, dt, Bm, Cm

    def forward(self, u1, u2, u2_cat_u1, u1_cat_u2, seq_idx=None, cu_seqlens=None):
        B, H, W, C = u1.shape
        L, K = H * W, 4
        kw = {} if tuple(self.dt_limit) == (0.0, float("inf")) else dict(dt_limit=self.dt_limit)
        As = -torch.exp(self.A_logs.float())
        Ds = self.Ds.view(-1, self.headdim) if self.D_has_hdim else self.Ds
        d_mlp = (2 * self.d_inner - self.d_ssm - self.d_ssm) // 2

        def branch(u, mixed):
            z0, x0, z = torch.split(self.skip_in_proj(u), [d_mlp, d_mlp, self.d_ssm], dim=-1)
            xs = self.act(self.xs_conv2d(self.xs_in_proj(u).permute(0, 3, 1, 2).contiguous()))
            bcdts = self.act(self.BCdts_conv2d(self.BCdts_in_proj(mixed).permute(0, 3, 1, 2).contiguous()))
            x, dt, Bm, Cm = self._scan_inputs(xs, bcdts, B, L)
            y = mamba_chunk_scan_combined(x, dt, As, Bm, Cm, chunk_size=self.chunk_size, D=Ds, z=None,
                                          dt_bias=self.dt_bias.view(-1), dt_softplus=True, seq_idx=seq_idx,
                                          cu_seqlens=cu_seqlens, **kw)                         # (B, L, 4*nheads, P)
            y = y.reshape(B, L, K, -1)
            assert y.dtype == torch.float32
            out = ssd_merge4(y, H, W).view(B, H, W, -1)                                        # cross-merge (:344-356), one gather pass
            if self.rmsnorm:
                out = self.norm(out, z)
            if d_mlp > 0:
                out = torch.cat([F.silu(z0) * x0, out], dim=-1)
            out = self.out_proj(out.to(self.out_proj.weight.dtype) if not torch.is_autocast_enabled() else out)
            return self.dropout(out) if self.dropout is not None else out

        return branch(u1, u2_cat_u1), branch(u2, u1_cat_u2)
