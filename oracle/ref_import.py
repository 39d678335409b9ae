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
