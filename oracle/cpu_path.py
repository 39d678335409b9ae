"""oracle/cpu_path.py -- TEST INFRASTRUCTURE: the reference's CPU data flow as its own eager module tree.

Used only by bench.py's `cpu_baseline` leg / `--impl reference` arm and by tests.  The product modules
(medical_image_classification_b200/{ss2d,models}.py) have NO CPU path -- their forwards call libb200ssm and raise on CPU
tensors.  This file holds the eager restatement instead: sub-classes that keep the product constructors (hence parameter names and
shapes: a reference / product state_dict loads with strict=True) and replace every `forward` with the reference's own
PyTorch CPU ops, line for line:

  * CpuSS2D        -- SS2D.forward (reference MedMamba.py:466-483) and forward_corev0 (:386-424): stack / transpose / flip /
                      cat cross-scan, the two einsums, four-direction scan, flip / transpose / add cross-merge, out_norm, gate;
  * CpuSSConvSSM   -- SS_Conv_SSM.forward (:530-538) with channel_shuffle (:486-499);
  * CpuPatchEmbed2D / CpuPatchMerging2D -- :160-169 / :186-212.

`selective_scan_fn` is bound to the C restatement of `selective_scan_ref` (oracle/sscan_oracle.c, OpenMP over rows) instead
of the reference's pure-PyTorch loop, whose autograd backward needs minutes per image: kind = "port".
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

import oracle
from medical_image_classification_b200.models import PatchEmbed2D, PatchMerging2D, SS_Conv_SSD, SS_Conv_SSM, channel_shuffle
from medical_image_classification_b200.ss2d import SS2D
from medical_image_classification_b200.ss2d_ssd import SS2D_with_SSD


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
    """x (B, D, H, W) -> (B, H, W, D); `m` supplies the SS2D parameters (reference names).  MedMamba.py:386-424, 476-477."""
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


class CpuSS2D(SS2D):
    def forward(self, x, **kwargs):                       # MedMamba.py:466-483
        xz = self.in_proj(x)
        x, z = xz.chunk(2, dim=-1)
        x = x.permute(0, 3, 1, 2).contiguous()
        x = self.act(self.conv2d(x))
        y = ss2d_core_cpu(self, x)
        assert y.dtype == torch.float32
        y = self.out_norm(y)
        y = y * F.silu(z)
        out = self.out_proj(y)
        if self.dropout is not None:
            out = self.dropout(out)
        return out


class CpuSSConvSSM(SS_Conv_SSM):
    def forward(self, input):                             # MedMamba.py:530-538
        left, right = input.chunk(2, dim=-1)
        right = self.drop_path(self.self_attention(self.ln_1(right)))
        left = left.permute(0, 3, 1, 2).contiguous()
        left = self.conv33conv33conv11(left)
        left = left.permute(0, 2, 3, 1).contiguous()
        output = torch.cat((left, right), dim=-1)
        output = channel_shuffle(output, groups=2)
        return output + input


class CpuPatchEmbed2D(PatchEmbed2D):
    def forward(self, x):                                 # MedMamba.py:164-169
        x = self.proj(x).permute(0, 2, 3, 1)
        return self.norm(x) if self.norm is not None else x


class CpuPatchMerging2D(PatchMerging2D):
    def forward(self, x):                                 # MedMamba.py:186-212
        B, H, W, C = x.shape
        h2, w2 = H // 2, W // 2
        parts = [x[:, i::2, j::2, :][:, :h2, :w2, :] for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))]
        x = torch.cat(parts, dim=-1).view(B, h2, w2, 4 * C)
        return self.reduction(self.norm(x))


class OracleSsdFn(torch.autograd.Function):
    """mamba_chunk_scan_combined on CPU tensors through the from-definition fp64 recurrence (oracle/ssd_oracle.c)."""

    @staticmethod
    def forward(ctx, x, dt, A, B, C, D, dt_bias, dt_softplus):
        out, _ = oracle.ssd_fwd(x, dt, A, B, C, D=D, dt_bias=dt_bias, dt_softplus=dt_softplus)
        ctx.save_for_backward(x, dt, A, B, C, D, dt_bias)
        ctx.softplus = dt_softplus
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, dout):
        x, dt, A, B, C, D, dt_bias = ctx.saved_tensors
        g = oracle.ssd_bwd(x, dt, A, B, C, D=D, dt_bias=dt_bias, dt_softplus=ctx.softplus, dout=dout.contiguous())
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a))
        return t(g["dx"]), t(g["ddt"]), t(g["dA"]), t(g["dB"]), t(g["dC"]), t(g["dD"]), t(g["ddt_bias"]), None


class CpuSS2DWithSSD(SS2D_with_SSD):
    def forward(self, u, seqlen=None, seq_idx=None, cu_seqlens=None):      # SSD/MedSSD.py:310-402
        B, H, W, C = u.shape
        L, K = H * W, 4
        zxbcdt = self.in_proj(u)
        d_mlp = (zxbcdt.shape[-1] - 2 * self.d_ssm - 2 * self.ngroups * self.d_state - self.nheads) // 2
        z0, x0, z, xBCdt = torch.split(
            zxbcdt, [d_mlp, d_mlp, self.d_ssm, self.d_ssm + 2 * self.ngroups * self.d_state + self.nheads], dim=-1)
        xBCdt = self.act(self.conv2d(xBCdt.permute(0, 3, 1, 2).contiguous()))
        gn = self.ngroups * self.d_state
        hwwh = torch.stack([xBCdt.reshape(B, -1, L), xBCdt.transpose(2, 3).reshape(B, -1, L)], dim=1)
        xBCdts = torch.cat([hwwh, hwwh.flip(-1)], dim=1)
        xs, Bs, Cs, dts = torch.split(xBCdts, [self.d_ssm, gn, gn, self.nheads], dim=2)
        xs = xs.float().reshape(B, -1, L).permute(0, 2, 1).unflatten(2, (-1, self.headdim))
        Bs = Bs.float().reshape(B, -1, L).permute(0, 2, 1).unflatten(2, (self.ngroups, -1))
        Cs = Cs.float().reshape(B, -1, L).permute(0, 2, 1).unflatten(2, (self.ngroups, -1))
        dts = dts.float().reshape(B, -1, L).permute(0, 2, 1)
        As = -torch.exp(self.A_logs.float())
        y = OracleSsdFn.apply(xs.contiguous(), dts.contiguous(), As, Bs.contiguous(), Cs.contiguous(), self.Ds.float(),
                              self.dt_bias.view(-1).float(), True)
        y = y.reshape(B, L, K, -1)
        inv_y = y[:, :, 2:4].flip(1)
        wh_y = y[:, :, 1].view(B, W, H, -1).transpose(1, 2).reshape(B, L, -1)
        invwh_y = inv_y[:, :, 1].view(B, W, H, -1).transpose(1, 2).reshape(B, L, -1)
        out = (y[:, :, 0] + inv_y[:, :, 0] + wh_y + invwh_y).view(B, H, W, -1)
        if self.rmsnorm:                                                    # gated RMSNorm, norm_before_gate=False, one group
            v = out * F.silu(z)
            out = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + self.norm.eps) * self.norm.weight
        if d_mlp > 0:
            out = torch.cat([F.silu(z0) * x0, out], dim=-1)
        out = self.out_proj(out)
        return self.dropout(out) if self.dropout is not None else out


class CpuSSConvSSD(SS_Conv_SSD):
    forward = CpuSSConvSSM.forward


_CPU_CLASS = {SS2D: CpuSS2D, SS_Conv_SSM: CpuSSConvSSM, PatchEmbed2D: CpuPatchEmbed2D, PatchMerging2D: CpuPatchMerging2D,
              SS2D_with_SSD: CpuSS2DWithSSD, SS_Conv_SSD: CpuSSConvSSD}


def bind_cpu_core(model):
    """Turn a (CPU) product module tree into the eager CPU tree above: every module whose class has a Cpu twin is re-classed
    (the twins add no attributes, only replace `forward`), parameters and buffers stay where they are.
    Returns the number of SS2D blocks converted."""
    n = 0
    for mod in model.modules():
        twin = _CPU_CLASS.get(type(mod))
        if twin is not None:
            mod.__class__ = twin
            n += twin in (CpuSS2D, CpuSS2DWithSSD)
    return n
