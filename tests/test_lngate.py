"""LayerNorm * silu(z) output stage (csrc/lngate.cu) against the reference's two PyTorch ops
(MedMamba.py:478-479) in fp32 -- a floating-point kernel, so the checker is a plain torch fp32 evaluation.
Tolerances: fp32 I/O 2e-6 relative (max-norm) forward, 2e-5 backward (sums over up to 50 K rows);
bf16 z / bf16 out: the bf16 rounding of the output (4e-3) forward, 1e-2 backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref(y, z, w, b, eps):
    n = torch.nn.functional.layer_norm(y.float(), (y.shape[-1],), w.float(), b.float(), eps)
    return n * torch.nn.functional.silu(z.float())


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("D", [96, 192, 384, 768, 40, 1024])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ln_gate_matches_torch(D, mode):
    from medical_image_classification_b200.ss2d import LnGateFn
    dev = "cuda"
    torch.manual_seed(D)
    B, H, W = 3, 13, 11
    y = (2.0 * torch.randn(B, H, W, D, device=dev) + 0.5).requires_grad_()
    zdt = torch.float32 if mode == "fp32" else torch.bfloat16
    xz = torch.randn(B, H, W, 2 * D, device=dev, dtype=zdt).requires_grad_()
    z = xz.chunk(2, dim=-1)[1]                       # strided view, read in place
    w = (1.0 + 0.1 * torch.randn(D, device=dev)).requires_grad_()
    b = (0.1 * torch.randn(D, device=dev)).requires_grad_()
    odt = torch.float32 if mode == "fp32" else torch.bfloat16
    out = LnGateFn.apply(y, z, w, b, 1e-5, odt)
    assert out.dtype == odt and out.shape == y.shape
    g = torch.randn_like(out)
    out.backward(g)
    got = [t.grad.clone() for t in (y, xz, w, b)]
    for t in (y, xz, w, b):
        t.grad = None
    o2 = ref(y, xz.chunk(2, dim=-1)[1], w, b, 1e-5)
    o2.backward(g.float())
    want = [t.grad for t in (y, xz, w, b)]
    ftol, btol = (2e-6, 2e-5) if mode == "fp32" else (4e-3, 1e-2)
    assert relerr(out, o2) < ftol
    for name, a, c in zip(("dy", "dxz", "dw", "db"), got, want):
        assert relerr(a, c) < btol, name


def test_ln_gate_full_size_properties():
    """BASELINE config-2 size (stage 0: 64 x 56 x 56 rows of 96 channels): the output of a normalised row does not
    depend on the other rows (first rows equal a small run bit for bit) and is invariant to a shift of y by a
    per-row constant up to rounding (LayerNorm removes the mean)."""
    from medical_image_classification_b200.ss2d import LnGateFn
    dev = "cuda"
    torch.manual_seed(0)
    y = torch.randn(64, 56, 56, 96, device=dev)
    z = torch.randn(64, 56, 56, 96, device=dev)
    w, b = torch.rand(96, device=dev) + 0.5, torch.randn(96, device=dev)
    out = LnGateFn.apply(y, z, w, b, 1e-5, torch.float32)
    out2 = LnGateFn.apply(y[:1], z[:1], w, b, 1e-5, torch.float32)
    assert torch.equal(out[:1], out2)
    shifted = LnGateFn.apply(y + 3.0, z, w, b, 1e-5, torch.float32)
    assert relerr(shifted, out) < 1e-5


@pytest.mark.parametrize("D", [48, 96, 384])
@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_plain_layer_norm_on_strided_half(D, odt):
    """z = None: the block pre-norm (MedMamba.py:531) on the right half of the block input, read in place."""
    from medical_image_classification_b200.ss2d import LnGateFn
    dev = "cuda"
    torch.manual_seed(D + 1)
    inp = (1.5 * torch.randn(2, 9, 10, 2 * D, device=dev) - 0.3).requires_grad_()
    right = inp.chunk(2, dim=-1)[1]
    w = (1.0 + 0.1 * torch.randn(D, device=dev)).requires_grad_()
    b = (0.1 * torch.randn(D, device=dev)).requires_grad_()
    out = LnGateFn.apply(right, None, w, b, 1e-6, odt)
    g = torch.randn_like(out)
    out.backward(g)
    got = [t.grad.clone() for t in (inp, w, b)]
    for t in (inp, w, b):
        t.grad = None
    o2 = torch.nn.functional.layer_norm(inp.chunk(2, dim=-1)[1], (D,), w, b, 1e-6)
    o2.backward(g.float())
    ftol, btol = (2e-6, 2e-5) if odt == torch.float32 else (4e-3, 1e-2)
    assert relerr(out, o2) < ftol
    for name, a, c in zip(("dinput", "dw", "db"), got, (inp.grad, w.grad, b.grad)):
        assert relerr(a, c) < btol, name


@pytest.mark.parametrize("D", [96, 192, 384, 768])
@pytest.mark.parametrize("odt", [torch.float32, torch.bfloat16])
def test_plain_layer_norm_of_bf16_rows(D, odt):
    """bf16 y (the residual stream of an autocast model after PatchMerging's Linear), z = None: rows are read as bf16 and
    dy comes back in bf16 -- equal to F.layer_norm of the upcast rows (what autocast runs) up to the output rounding."""
    from medical_image_classification_b200.ss2d import LnGateFn
    dev = "cuda"
    torch.manual_seed(D + 2)
    inp = (1.5 * torch.randn(2, 9, 10, 2 * D, device=dev) - 0.3).to(torch.bfloat16).requires_grad_()
    right = inp.chunk(2, dim=-1)[1]
    w = (1.0 + 0.1 * torch.randn(D, device=dev)).requires_grad_()
    b = (0.1 * torch.randn(D, device=dev)).requires_grad_()
    out = LnGateFn.apply(right, None, w, b, 1e-6, odt)
    assert out.dtype == odt
    g = torch.randn_like(out)
    out.backward(g)
    got = [t.grad.clone() for t in (inp, w, b)]
    assert got[0].dtype == torch.bfloat16
    for t in (inp, w, b):
        t.grad = None
    o2 = torch.nn.functional.layer_norm(inp.chunk(2, dim=-1)[1].float(), (D,), w, b, 1e-6)
    o2.backward(g.float())
    assert relerr(out, o2) < (2e-6 if odt == torch.float32 else 4e-3)
    assert relerr(got[0][..., D:], inp.grad[..., D:]) < 8e-3          # bf16 rounding of dy
    assert float(got[0][..., :D].abs().max()) == 0.0
    assert relerr(got[1], w.grad) < (2e-5 if odt == torch.float32 else 1e-2)
    assert relerr(got[2], b.grad) < (2e-5 if odt == torch.float32 else 1e-2)
