"""Depth-wise conv3x3 + SiLU producer stage (csrc/dwconv.cu) against the reference's ops
(MedMamba.py:470-473: permute -> nn.Conv2d(D, D, 3, padding=1, groups=D) -> SiLU) evaluated in fp32 by PyTorch.
Floating-point kernel: tolerances 2e-6 (max-norm relative) forward, 2e-5 backward for fp32 input; with bf16 input the
kernel reads the same bf16 values the checker is given (converted exactly to fp32), so the same forward bound holds and
the input gradient is compared after rounding to bf16 (8e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.detach().float() - b.detach().float()).abs().max() / b.detach().float().abs().max().clamp_min(1e-30))


CASES = [(2, 14, 14, 96), (1, 56, 56, 96), (3, 7, 7, 768), (2, 13, 11, 24), (1, 5, 64, 40), (2, 3, 2, 16)]


@pytest.mark.parametrize("shape", CASES, ids=[str(c) for c in CASES])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dwconv_silu_matches_torch(shape, dtype):
    from medical_image_classification_b200.ss2d import DwConvSiluFn
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, D = shape
    dev = "cuda"
    torch.manual_seed(B * 1000 + D)
    xz = torch.randn(B, H, W, 2 * D, device=dev, dtype=dtype).requires_grad_()
    wgt = (0.3 * torch.randn(D, 1, 3, 3, device=dev)).requires_grad_()
    bias = (0.1 * torch.randn(D, device=dev)).requires_grad_()
    x = xz.chunk(2, dim=-1)[0]                      # the strided view SS2D.forward passes
    out = DwConvSiluFn.apply(x, wgt, bias)
    assert out.dtype == torch.float32 and out.shape == (B, D, H, W)
    g = torch.randn_like(out)
    out.backward(g)
    got = [t.grad.clone() for t in (xz, wgt, bias)]
    for t in (xz, wgt, bias):
        t.grad = None
    xr = xz.chunk(2, dim=-1)[0].float().permute(0, 3, 1, 2).contiguous()
    ref = F.silu(F.conv2d(xr, wgt, bias, padding=1, groups=D))
    ref.backward(g)
    assert relerr(out, ref) < 2e-6
    assert relerr(got[0], xz.grad) < (2e-5 if dtype == torch.float32 else 8e-3)
    assert relerr(got[1], wgt.grad) < 2e-5
    assert relerr(got[2], bias.grad) < 2e-5
    assert torch.count_nonzero(got[0][..., D:]) == 0   # the z half receives no gradient from this op


def test_dwconv_full_size_properties():
    """BASELINE config-2 size (stage 0: batch 64, 56 x 56, 96 channels): samples are independent (first sample equals a
    batch-1 run bit for bit) and a constant image gives a constant interior (stencil sums to the same value)."""
    from medical_image_classification_b200.ss2d import DwConvSiluFn
    dev = "cuda"
    torch.manual_seed(0)
    x = torch.randn(64, 56, 56, 96, device=dev)
    w, b = 0.3 * torch.randn(96, 1, 3, 3, device=dev), 0.1 * torch.randn(96, device=dev)
    out = DwConvSiluFn.apply(x, w, b)
    out1 = DwConvSiluFn.apply(x[:1], w, b)
    assert torch.equal(out[:1], out1)
    const = DwConvSiluFn.apply(torch.ones(1, 56, 56, 96, device=dev), w, b)
    inner = const[0, :, 1:-1, 1:-1]
    assert float((inner - inner[:, :1, :1]).abs().max()) < 1e-6
