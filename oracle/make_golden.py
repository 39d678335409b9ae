#!/usr/bin/env python
"""oracle/make_golden.py -- TEST INFRASTRUCTURE: generate tests/golden/*.npz from the reference.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

It imports the UNMODIFIED reference Python (oracle/ref_import.py) and records, for seeded inputs:

* sscan_*.npz   -- selective_scan_ref (selective_scan_interface.py:92-158) forward + autograd grads.
                   Input distributions follow the reference's own test
                   (test_selective_scan.py:411-441, 474): A=-0.5*U, B,C,u,D~N(0,1), delta=0.5*U,
                   delta_bias=0.5*U, g~N(0,1); plus cases at the model's shapes (N=16, G=4,
                   L in {49,196}) and A=-(1..N) (MedMamba.py:357-372 init).
* ss2d_*.npz    -- reference SS2D module (MedMamba.py:253-483) forward/backward with its state_dict.
* vssm_tiny.npz -- reference VSSM (MedMamba.py:671-767) eval logits + train-mode grads, tiny dims.

* ss2d_ssd_*.npz / medssd_tiny.npz -- reference SS2D_with_SSD / VSSM (SSD/MedSSD.py) with mamba_ssm's operator
                   replaced by the oracle (module data flow pinned; operator itself PARITY UNPINNED).

The vectors travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return None if t is None else t.detach().cpu().numpy().copy()  # copy: BN buffers are updated in place later


def sscan_case(name, batch, dim, N, L, G, has_D, has_bias, softplus, has_z=False, model_A=False,
               bc_3d=False, with_grads=True, seed=0, itype=torch.float32):
    """itype: dtype of u / delta / B / C / z as in the reference's test grid (test_selective_scan.py:372-390: fp32, fp16, bf16;
    A, D, delta_bias stay fp32).  The arrays are stored as float32 holding the rounded values, `itype` names the dtype."""
    iface = ref_import.load_selective_scan_interface()
    torch.manual_seed(seed)
    if model_A:
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(dim, 1)
    else:
        A = -0.5 * torch.rand(dim, N)
    bshape = (batch, N, L) if bc_3d else (batch, G, N, L)
    Bm = torch.randn(*bshape)
    Cm = torch.randn(*bshape)
    D = torch.randn(dim) if has_D else None
    z = torch.randn(batch, dim, L) if has_z else None
    bias = 0.5 * torch.rand(dim) if has_bias else None
    u = torch.randn(batch, dim, L)
    delta = 0.5 * torch.rand(batch, dim, L)
    g = torch.randn(batch, dim, L)
    if itype != torch.float32:
        Bm, Cm, u, delta, g = (t.to(itype) for t in (Bm, Cm, u, delta, g))
        z = None if z is None else z.to(itype)
    leaves = [t for t in (u, delta, A, Bm, Cm, D, z, bias) if t is not None]
    for t in leaves:
        t.requires_grad_(with_grads)
    out, last = iface.selective_scan_ref(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias,
                                         delta_softplus=softplus, return_last_state=True)
    _np = lambda t: None if t is None else t.detach().float().cpu().numpy().copy()   # noqa: E731  (16-bit tensors -> float32 values)
    rec = dict(u=_np(u), delta=_np(delta), A=_np(A), B=_np(Bm), C=_np(Cm), g=_np(g),
               out=_np(out), last_state=_np(last), delta_softplus=np.array(int(softplus)))
    if itype != torch.float32:
        rec["itype"] = np.array(str(itype).replace("torch.", ""))
    if D is not None:
        rec["D"] = _np(D)
    if z is not None:
        rec["z"] = _np(z)
    if bias is not None:
        rec["delta_bias"] = _np(bias)
    if with_grads:
        out.backward(g)
        rec.update(du=_np(u.grad), ddelta=_np(delta.grad), dA=_np(A.grad), dB=_np(Bm.grad), dC=_np(Cm.grad))
        if D is not None:
            rec["dD"] = _np(D.grad)
        if z is not None:
            rec["dz"] = _np(z.grad)
        if bias is not None:
            rec["ddelta_bias"] = _np(bias.grad)
    np.savez_compressed(os.path.join(OUT, f"sscan_{name}.npz"), **rec)
    print("wrote", name, {k: v.shape for k, v in rec.items() if hasattr(v, "shape")})


def sscan_extra_cases():
    """More of the reference's own test grid (test_selective_scan.py:372-390: seqlen up to 4096, itype fp32 / fp16 / bf16)."""
    sscan_case("long_L4096", 1, 8, 16, 4096, 1, True, True, True, with_grads=False)
    sscan_case("grads_L1024", 1, 8, 16, 1024, 2, True, True, True, seed=3)
    sscan_case("fp16_L256", 2, 24, 16, 256, 2, True, True, True, seed=4, itype=torch.float16)
    sscan_case("bf16_L256", 2, 24, 16, 256, 2, True, True, True, seed=5, itype=torch.bfloat16)


def ss2d_case(name, d_model, H, W, batch, seed=0):
    mm = ref_import.load_medmamba()
    torch.manual_seed(seed)
    m = mm.SS2D(d_model=d_model, d_state=16)
    # move the parameters off their structured init so that every term matters
    with torch.no_grad():
        m.A_logs.add_(0.3 * torch.randn_like(m.A_logs))
        m.Ds.add_(0.5 * torch.randn_like(m.Ds))
        m.out_norm.weight.add_(0.2 * torch.randn_like(m.out_norm.weight))
        m.out_norm.bias.add_(0.2 * torch.randn_like(m.out_norm.bias))
    x = torch.randn(batch, H, W, d_model, requires_grad=True)
    g = torch.randn(batch, H, W, d_model)
    out = m(x)
    out.backward(g)
    rec = {"x": _np(x), "g": _np(g), "out": _np(out), "dx": _np(x.grad)}
    for k, v in m.state_dict().items():
        rec["sd." + k] = _np(v)
    for k, p in m.named_parameters():
        rec["grad." + k] = _np(p.grad)
    # the bare core: conv output -> 4 merged-order outputs (forward_corev0, MedMamba.py:386-424)
    with torch.no_grad():
        xc = torch.randn(batch, m.d_inner, H, W)
        ys = m.forward_corev0(xc)
        rec["core_x"] = _np(xc)
        rec["core_y"] = _np(torch.stack(ys, 0))
    np.savez_compressed(os.path.join(OUT, f"ss2d_{name}.npz"), **rec)
    print("wrote ss2d", name)


def vssm_case(seed=0):
    mm = ref_import.load_medmamba()
    torch.manual_seed(seed)
    kw = dict(num_classes=6, depths=[1, 1, 1, 1], dims=[8, 16, 32, 64], drop_path_rate=0.0)
    net = mm.VSSM(**kw)
    x = torch.randn(4, 3, 64, 64)
    y = torch.randint(0, 6, (4,))
    rec = {"x": _np(x), "y": _np(y)}
    for k, v in net.state_dict().items():
        rec["sd." + k] = _np(v)
    net.eval()
    with torch.no_grad():
        rec["logits_eval"] = _np(net(x))
    net.train()
    logits = net(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    rec["logits_train"] = _np(logits)
    rec["loss"] = _np(loss)
    for k, p in net.named_parameters():
        if p.grad is not None and p.numel() <= 4096:
            rec["grad." + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, "vssm_tiny.npz"), **rec)
    print("wrote vssm_tiny loss", float(loss))


def ss2d_ssd_case(name, d_model, d_state, headdim, H, W, batch, seed=0):
    """Reference SS2D_with_SSD (SSD/MedSSD.py:160-402) forward/backward with its state_dict; the SSD operator is the
    oracle stand-in (ref_import.load_medssd), so this pins the module's data flow, not the operator."""
    ms = ref_import.load_medssd()
    torch.manual_seed(seed)
    m = ms.SS2D_with_SSD(d_model=d_model, d_state=d_state, headdim=headdim, chunk_size=32)
    with torch.no_grad():
        m.Ds.add_(0.5 * torch.randn_like(m.Ds))
        m.A_logs.add_(0.3 * torch.randn_like(m.A_logs))
        m.dt_bias.add_(0.3 * torch.randn_like(m.dt_bias))
        m.norm.weight.add_(0.2 * torch.randn_like(m.norm.weight))
    x = torch.randn(batch, H, W, d_model, requires_grad=True)
    g = torch.randn(batch, H, W, d_model)
    out = m(x)
    out.backward(g)
    rec = {"x": _np(x), "g": _np(g), "out": _np(out), "dx": _np(x.grad),
           "cfg": np.array([d_model, d_state, headdim, H, W, batch])}
    for k, v in m.state_dict().items():
        rec["sd." + k] = _np(v)
    for k, p in m.named_parameters():
        rec["grad." + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, f"ss2d_ssd_{name}.npz"), **rec)
    print("wrote ss2d_ssd", name, float(out.abs().max()))


def medssd_case(seed=0):
    ms = ref_import.load_medssd()
    torch.manual_seed(seed)
    kw = dict(num_classes=6, depths=[1, 1], dims=[64, 128], d_state=8, drop_path_rate=0.0)
    net = ms.VSSM(**kw)
    x = torch.randn(2, 3, 64, 64)
    y = torch.randint(0, 6, (2,))
    rec = {"x": _np(x), "y": _np(y)}
    for k, v in net.state_dict().items():
        rec["sd." + k] = _np(v)
    net.eval()
    with torch.no_grad():
        rec["logits_eval"] = _np(net(x))
    net.train()
    logits = net(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    rec["logits_train"] = _np(logits)
    rec["loss"] = _np(loss)
    for k, p in net.named_parameters():
        if p.grad is not None and p.numel() <= 2048:
            rec["grad." + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, "medssd_tiny.npz"), **rec)
    print("wrote medssd_tiny loss", float(loss))


def crossmamba_case(name, d_model, d_state, headdim, H, W, batch, seed=0):
    """Reference CrossMamba (CrossMamba/CrossMamba_fusion_2b2.py:54-388): two inputs, x from each branch, B/C/dt from the
    mixed tensors; SSD operator = oracle stand-in (module data flow pinned, operator PARITY UNPINNED)."""
    cm = ref_import.load_crossmamba()
    torch.manual_seed(seed)
    m = cm.CrossMamba(d_model=d_model, d_state=d_state, headdim=headdim, chunk_size=32)
    with torch.no_grad():
        m.Ds.add_(0.5 * torch.randn_like(m.Ds))
        m.A_logs.add_(0.3 * torch.randn_like(m.A_logs))
        m.dt_bias.add_(0.3 * torch.randn_like(m.dt_bias))
        m.norm.weight.add_(0.2 * torch.randn_like(m.norm.weight))
    ins = [torch.randn(batch, H, W, d_model, requires_grad=True) for _ in range(4)]   # u1, u2, u2_cat_u1, u1_cat_u2
    g1, g2 = torch.randn(batch, H, W, d_model), torch.randn(batch, H, W, d_model)
    o1, o2 = m(*ins)
    (o1 * g1).sum().add((o2 * g2).sum()).backward()
    rec = {"g1": _np(g1), "g2": _np(g2), "out1": _np(o1), "out2": _np(o2),
           "cfg": np.array([d_model, d_state, headdim, H, W, batch])}
    for i, t in enumerate(ins):
        rec[f"in{i}"] = _np(t)
        rec[f"din{i}"] = _np(t.grad)
    for k, v in m.state_dict().items():
        rec["sd." + k] = _np(v)
    for k, p in m.named_parameters():
        if p.grad is not None:
            rec["grad." + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, f"crossmamba_{name}.npz"), **rec)
    print("wrote crossmamba", name, float(o1.abs().max()), float(o2.abs().max()))


def _fusionmamba_scan_merge():
    """EfficientMerge / EfficientScan of CrossMamba/FusionMamba/models/cross.py (:34-90, :139-190).  The module cannot be imported
    whole (mamba_ssm, timm and a compiled selective_scan_cuda at import time), so the two autograd classes -- pure tensor ops --
    are compiled from their source lines in place, unmodified."""
    import math
    import torch.nn.functional as F
    path = os.path.join(ref_import.REF_ROOT if hasattr(ref_import, "REF_ROOT") else "/root/reference", "CrossMamba/FusionMamba/models/cross.py")
    lines = open(path).read().split("\n")
    src = "\n".join(lines[33:90]) + "\n\n" + "\n".join(lines[138:190]) + "\n"
    ns = {"torch": torch, "F": F, "math": math}
    exec(compile(src, path, "exec"), ns)
    return ns["EfficientScan"], ns["EfficientMerge"]


def atrous_case(name, B, C, H, W, seed=0):
    """x -> xs = EfficientScan(x); ys -> y = EfficientMerge(ys), plus their autograd gradients, from the reference classes."""
    Scan, Merge = _fusionmamba_scan_merge()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g, requires_grad=True)
    xs = Scan.apply(x, 2)
    gxs = torch.randn(xs.shape, generator=g)
    xs.backward(gxs)
    ys = torch.randn(xs.shape, generator=g, requires_grad=True)
    y = Merge.apply(ys, H, W, 2)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    np.savez_compressed(os.path.join(OUT, f"atrous_{name}.npz"), x=_np(x), xs=_np(xs), gxs=_np(gxs), dx=_np(x.grad), ys=_np(ys), y=_np(y),
                        gy=_np(gy), dys=_np(ys.grad))
    print("wrote atrous", name, tuple(xs.shape))


def _grad_sample(t, n=4096):
    """<= n evenly strided elements of a gradient (every parameter stays comparable without storing 6 M floats)."""
    f = t.detach().flatten()
    if f.numel() <= n:
        return _np(f)
    return _np(f[torch.arange(0, f.numel(), f.numel() // n)[:n]])


def medmamba_real_dims_case(seed=0):
    """fp32 goldens for the configuration bench.py benchmarks under bf16 autocast: the UNMODIFIED reference VSSM at MedMamba-T's
    real widths (dims 96-192-384-768, d_state 16, 224 x 224 input; depths 1-1-1-1, batch 4 to keep the CPU run in minutes), with
    `selective_scan_fn` bound to the C restatement of selective_scan_ref (oracle/sscan_oracle.c, itself pinned on the reference's
    function by tests/test_oracle_golden.py; the reference's own autograd backward needs ~6 minutes per image at L = 3136).
    Weights are NOT stored (6 M floats): they are the product model's seeded CPU initialisation, which the test re-creates;
    `param_checksum` guards against RNG drift.  Gradients are stored as <= 4096 strided samples per parameter."""
    from medical_image_classification_b200.models import VSSM as ProductVSSM
    from oracle.cpu_path import OracleSelectiveScanFn

    def scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
        assert z is None and not return_last_state
        return OracleSelectiveScanFn.apply(u.contiguous(), delta.contiguous(), A.contiguous(), B.contiguous(), C.contiguous(), D, delta_bias,
                                           delta_softplus)

    mm = ref_import.load_medmamba(selective_scan_fn=scan_fn)
    kw = dict(num_classes=6, depths=[1, 1, 1, 1], dims=[96, 192, 384, 768], drop_path_rate=0.0)
    torch.manual_seed(seed)
    prod = ProductVSSM(**kw)                        # CPU construction only: nothing is launched
    net = mm.VSSM(**kw)
    net.load_state_dict(prod.state_dict(), strict=True)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(4, 3, 224, 224, generator=g)
    y = torch.randint(0, 6, (4,), generator=g)
    rec = {"y": _np(y), "seed": np.array(seed),
           "param_checksum": np.array([float(sum(p.double().sum() for p in prod.parameters())),
                                       float(sum(p.double().abs().sum() for p in prod.parameters()))])}
    net.eval()
    with torch.no_grad():
        rec["logits_eval"] = _np(net(x))
    net.train()
    logits = net(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    rec["logits_train"] = _np(logits)
    rec["loss"] = _np(loss)
    for k, p in net.named_parameters():
        if p.grad is not None:
            rec["grad." + k] = _grad_sample(p.grad)
            rec["gnorm." + k] = np.array(float(p.grad.double().norm()))
    np.savez_compressed(os.path.join(OUT, "medmamba_real_dims.npz"), **rec)
    print("wrote medmamba_real_dims: loss", float(loss), "logits_eval[0]", rec["logits_eval"][0])


def kan_head_case(seed=0):
    """Reference KansModule (MedSSD_kan/MedSSD_kan.py:475-501: two pykan-style KANLayers + BatchNorm1d) forward / backward with
    its state_dict, for the host-side head mirror medical_image_classification_b200/kan_head.py (outside the hot path)."""
    mod = ref_import.load_medssd("MedSSD_kan/MedSSD_kan.py", "ref_MedSSD_kan")
    torch.manual_seed(seed)
    m = mod.KansModule(24, 6)
    with torch.no_grad():                                  # move every knot vector a little so the grids are not all identical
        m.kan1.grid.add_(0.01 * torch.randn(24, 1))
    x = (1.2 * torch.randn(10, 24)).requires_grad_()
    g = torch.randn(10, 6)
    rec = {"x": _np(x), "g": _np(g)}
    for k, v in m.state_dict().items():
        rec["sd." + k] = _np(v)
    m.train()
    out = m(x)
    out.backward(g)
    rec["out_train"] = _np(out)
    rec["dx"] = _np(x.grad)
    for k, p in m.named_parameters():
        if p.grad is not None:
            rec["grad." + k] = _np(p.grad)
    m.eval()
    with torch.no_grad():
        rec["out_eval"] = _np(m(x))
    np.savez_compressed(os.path.join(OUT, "kan_head.npz"), **rec)
    print("wrote kan_head")


def main():
    assert ref_import.available(), "/root/reference is not mounted"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "--kan-only" in sys.argv:
        kan_head_case()
        return
    if "--real-dims-only" in sys.argv:
        medmamba_real_dims_case()
        return
    if "--atrous-only" in sys.argv:
        atrous_case("8x8", 2, 3, 8, 8)
        atrous_case("7x9", 1, 5, 7, 9, seed=1)
        atrous_case("14x14", 2, 2, 14, 14, seed=2)
        return
    if "--crossmamba-only" in sys.argv:
        crossmamba_case("d32_6x5", 32, 8, 16, 6, 5, 2)
        return
    if "--sscan-extra" in sys.argv:
        sscan_extra_cases()
        return
    if "--ssd-only" in sys.argv:
        ss2d_ssd_case("d32_7x5", 32, 8, 16, 7, 5, 2)
        ss2d_ssd_case("d64_6x6", 64, 16, 64, 6, 6, 1)
        medssd_case()
        return
    # reference test-grid style (dstate=1, G in {1 (3-D B/C), 2})
    sscan_case("grid_g1", 2, 24, 1, 64, 1, True, True, True, bc_3d=True)
    sscan_case("grid_g2_plain", 2, 24, 1, 64, 2, False, False, False)
    sscan_case("grid_g2_nosoftplus_bias", 2, 24, 1, 128, 2, True, True, False)
    # the model's shapes: N=16, 4 groups, L = 7*7 and 14*14
    sscan_case("model_L49", 2, 32, 16, 49, 4, True, True, True, model_A=True)
    sscan_case("model_L196", 1, 32, 16, 196, 4, True, True, True)
    # gate z, odd sizes
    sscan_case("z_L130", 2, 16, 8, 130, 1, True, True, True, has_z=True, bc_3d=True)
    # longer than the reference kernel's 2048-step chunk (selective_scan.cpp:307); forward only
    sscan_case("long_L2100", 1, 8, 16, 2100, 1, True, True, True, with_grads=False)
    sscan_extra_cases()
    ss2d_case("d8_7x5", 8, 7, 5, 2)
    ss2d_case("d16_4x6", 16, 4, 6, 1)
    vssm_case()
    ss2d_ssd_case("d32_7x5", 32, 8, 16, 7, 5, 2)
    ss2d_ssd_case("d64_6x6", 64, 16, 64, 6, 6, 1)
    medssd_case()
    medmamba_real_dims_case()
    kan_head_case()
    crossmamba_case("d32_6x5", 32, 8, 16, 6, 5, 2)
    atrous_case("8x8", 2, 3, 8, 8)
    atrous_case("7x9", 1, 5, 7, 9, seed=1)
    atrous_case("14x14", 2, 2, 14, 14, seed=2)


if __name__ == "__main__":
    main()
