"""oracle/ref_import.py -- TEST INFRASTRUCTURE.

Import the *unmodified* reference Python files from /root/reference inside THIS container, so
that the oracle can be pinned against the reference itself and golden vectors can be generated
(oracle/make_golden.py).  /root/reference does not exist on the GPU box: nothing under tests -m gpu,
smoke() or bench.py calls into this module at run time.

Loading trick: the vendored selective_scan_interface.py does
`import selective_scan_cuda` at module scope and MedMamba.py imports timm (absent here), so small
stub modules are planted in sys.modules before the files are exec'd by path.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _timm_shim():
    """Minimal stand-in for timm.models.layers.{DropPath,to_2tuple,trunc_normal_} (timm semantics)."""
    import torch
    import torch.nn as nn

    class DropPath(nn.Module):
        def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", models)
    sys.modules.setdefault("timm.models.layers", layers)


def load_selective_scan_interface():
    """The reference's vendored mamba_ssm 1.2.0 interface (selective_scan_ref lives here)."""
    if "ref_selective_scan_interface" in sys.modules:
        return sys.modules["ref_selective_scan_interface"]
    sys.modules.setdefault("selective_scan_cuda", types.ModuleType("selective_scan_cuda"))
    path = os.path.join(REF_ROOT, "CrossMamba/FusionMamba/mamba_ssm/ops/selective_scan_interface.py")
    return _load("ref_selective_scan_interface", path)


def load_medmamba(selective_scan_fn=None):
    """Reference MedMamba.py with `selective_scan_fn` bound to `selective_scan_fn`
    (default: the reference's own pure-PyTorch selective_scan_ref = the CPU path)."""
    iface = load_selective_scan_interface()
    _timm_shim()
    fn = selective_scan_fn or iface.selective_scan_ref
    pkg = types.ModuleType("mamba_ssm")
    ops = types.ModuleType("mamba_ssm.ops")
    ssi = types.ModuleType("mamba_ssm.ops.selective_scan_interface")
    ssi.selective_scan_fn = fn
    ssi.selective_scan_ref = iface.selective_scan_ref
    pkg.ops, ops.selective_scan_interface = ops, ssi
    saved = {k: sys.modules.get(k) for k in ("mamba_ssm", "mamba_ssm.ops", "mamba_ssm.ops.selective_scan_interface")}
    sys.modules.update({"mamba_ssm": pkg, "mamba_ssm.ops": ops, "mamba_ssm.ops.selective_scan_interface": ssi})
    try:
        mod = _load("ref_MedMamba", os.path.join(REF_ROOT, "MedMamba.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.selective_scan_fn = fn
    return mod


def _oracle_ssd_fn():
    """CPU stand-in for mamba_ssm's mamba_chunk_scan_combined: torch autograd over the from-definition oracle
    (oracle/ssd_oracle.c).  mamba_ssm==2.2.2 is not vendored in the reference (PARITY UNPINNED for the operator);
    what the golden vectors made with this stand-in DO pin is the reference module's data flow around the operator
    (SSD/MedSSD.py:310-402: projections, conv, cross-scan, the one-group / 4*d_state quirk, cross-merge, gated norm)."""
    import numpy as np
    import torch
    import oracle

    class Fn(torch.autograd.Function):
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

    def mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None,
                                  seq_idx=None, cu_seqlens=None, dt_softplus=False, dt_limit=(0.0, float("inf")),
                                  return_final_states=False):
        assert z is None and initial_states is None and seq_idx is None and cu_seqlens is None and not return_final_states
        return Fn.apply(x, dt, A, B, C, D, dt_bias, dt_softplus)

    return mamba_chunk_scan_combined


def load_medssd(rel_path: str = "SSD/MedSSD.py", name: str = "ref_MedSSD"):
    """Reference SSD/MedSSD.py with mamba_ssm's pieces replaced by CPU stand-ins: the SSD operator by the oracle,
    the gated RMSNorm by its published formula y = rmsnorm(x * silu(z)) * w (norm_before_gate=False)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    _timm_shim()

    class RMSNorm(nn.Module):
        def __init__(self, hidden_size, eps=1e-5, group_size=None, norm_before_gate=True, device=None, dtype=None):
            super().__init__()
            assert not norm_before_gate and (group_size is None or group_size == hidden_size)
            self.eps = eps
            self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
            self.register_parameter("bias", None)

        def forward(self, x, z=None):
            v = x * F.silu(z)
            return v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + self.eps) * self.weight

    names = ["mamba_ssm", "mamba_ssm.ops", "mamba_ssm.ops.triton", "mamba_ssm.ops.triton.layernorm_gated",
             "mamba_ssm.ops.triton.ssd_combined", "mamba_ssm.ops.triton.selective_state_update",
             "mamba_ssm.distributed", "mamba_ssm.distributed.tensor_parallel", "mamba_ssm.distributed.distributed_utils"]
    mods = {n: types.ModuleType(n) for n in names}
    mods["mamba_ssm.ops.triton.layernorm_gated"].RMSNorm = RMSNorm
    mods["mamba_ssm.ops.triton.ssd_combined"].mamba_chunk_scan_combined = _oracle_ssd_fn()
    mods["mamba_ssm.ops.triton.ssd_combined"].mamba_split_conv1d_scan_combined = None
    mods["mamba_ssm.ops.triton.selective_state_update"].selective_state_update = None
    mods["mamba_ssm.distributed.tensor_parallel"].ColumnParallelLinear = None
    mods["mamba_ssm.distributed.tensor_parallel"].RowParallelLinear = None
    mods["mamba_ssm.distributed.distributed_utils"].all_reduce = None
    mods["mamba_ssm.distributed.distributed_utils"].reduce_scatter = None
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(mods)
    try:
        import huggingface_hub  # noqa: F401
    except Exception:
        hub = types.ModuleType("huggingface_hub")
        hub.PyTorchModelHubMixin = type("PyTorchModelHubMixin", (), {})
        sys.modules["huggingface_hub"] = hub
    try:
        mod = _load(name, os.path.join(REF_ROOT, rel_path))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_crossmamba():
    """Reference CrossMamba/CrossMamba_fusion_2b2.py (two-branch CrossMamba module, :54-388) with the same CPU stand-ins
    as load_medssd (oracle SSD operator, formula RMSNorm)."""
    return load_medssd("CrossMamba/CrossMamba_fusion_2b2.py", "ref_CrossMamba")
